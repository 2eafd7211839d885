/* Shared device-side descriptors. */
#ifndef VB200_K_COMMON_CUH
#define VB200_K_COMMON_CUH
#include <cuda_runtime.h>

/* How a ch-block finds `lastmdct`, the previous block's logmdct as aoTuV keeps it in
   private_state.mblock (psy.c:4462-4501).  The reference copies / 8x-expands / 8:1 min-pools the
   previous block's logmdct into a state buffer; here the previous block's logmdct is still in HBM, so
   the consumer applies the resampling rule on the fly and the state buffer never round-trips. */
enum { LAST_COPY = 0, LAST_EXPAND8 = 1, LAST_MINPOOL8 = 2, LAST_ZERO = 3, LAST_NONE = 4 };
#define LAST_SHORT_EXTENT 128   /* entries a writer whose next block is short leaves in lastmdct (block sizes 256 / 2048) */
struct LastRef {
  const float *base;        /* buffer the offsets index */
  const long long *off;     /* [items] element offset, or NULL: item*stride */
  const int *rule;          /* [items] LAST_*, or NULL: LAST_COPY */
  long long stride;
  __device__ __forceinline__ static float fetch(const float *p, int r, int i) {
    if (r == LAST_COPY) return __ldg(p + i);
    if (r == LAST_EXPAND8) return __ldg(p + (i >> 3));
    if (r == LAST_MINPOOL8) {
      float m = __ldg(p + i * 8);
#pragma unroll
      for (int j = 1; j < 8; j++) { float v = __ldg(p + i * 8 + j); if (m > v) m = v; }
      return m;
    }
    return 0.f;
  }
  __device__ __forceinline__ float at(int item, int i) const {
    /* base==NULL: off[] holds absolute device addresses (previous block / stream state live in
       different allocations) */
    const float *p = base ? base + (off ? off[item] : (long long)item * stride)
                          : reinterpret_cast<const float *>(static_cast<uintptr_t>(off[item]));
    return fetch(p, rule ? rule[item] : LAST_COPY, i);
  }
};
/* Managed mode.  A block whose nominal floor fit is NULL does not write lastmdct (mapping0.c:1097), so a long block can
   find the newest writer's 128 entries on top of an OLDER writer's 1024: bins >= LAST_SHORT_EXTENT then come from a
   second source (absolute addresses; LAST_NONE: one source is enough).  A type of its own, so that the VBR kernels
   compile exactly as they did without it. */
struct LastRef2 {
  LastRef first;
  const long long *off2;
  const int *rule2;
  __device__ __forceinline__ float at(int item, int i) const {
    if (i >= LAST_SHORT_EXTENT) {
      const int r2 = rule2[item];
      if (r2 != LAST_NONE) return LastRef::fetch(reinterpret_cast<const float *>(static_cast<uintptr_t>(off2[item])), r2, i);
    }
    return first.at(item, i);
  }
};
#endif
