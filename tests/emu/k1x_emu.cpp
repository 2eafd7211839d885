/* TEST INFRASTRUCTURE -- compiles the K1x kernel (vorbis_aotuv_lancer_b200/csrc/k_transform_exact.cu) as a serial host
 * program: ONE "thread" (THREADS = 1) runs every work item of every stage in order, __syncthreads() is a no-op, the
 * device intrinsics are their IEEE host equivalents (built with -ffp-contract=off).  This checks the kernel's index
 * arithmetic, stage order and buffer reuse on the CPU against the oracle and the reference; it says nothing about
 * races between threads (the stages' items are independent by construction) and it is not a way to run the product --
 * nothing in the package can reach it.  Built and used by tests/test_k1x_emulation.py only. */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector_types.h>          /* float2 / int2 / int4 of the setup structs */

#define VB200_DEV_UTIL_CUH          /* dev_util.cuh is replaced by the host equivalents below */
#define FULLMASK 0xffffffffu
#undef __device__
#undef __global__
#undef __forceinline__
#undef __shared__
#undef __align__
#undef __launch_bounds__
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
static struct { int x; } threadIdx, blockIdx;
static inline void __syncthreads() {}
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline double warp_sum_d(double v) { return v; }          /* one thread: the warp is that thread */
static inline float warp_max_f(float v) { return v; }
static inline float todB_raw(float x) {
  uint32_t i; memcpy(&i, &x, 4); i &= 0x7fffffffu;
  return (float)i * 7.17711438e-7f + -764.6161886f;
}
static inline float todB_p345(float x) { return (float)((double)todB_raw(x) + .345); }

#include "../../vorbis_aotuv_lancer_b200/csrc/k_transform_exact.cu"

/* one windowed block in (the window is applied by the kernel: windows of ones over a long-long-long block leave it
   as it is), the four outputs of the front end out */
extern "C" int k1x_emu(int n, const float *windowed, float *gmdct, float *logfft, float *logmdct, float *ampmax) {
  static DevSetup S;
  static float ones[1024], T[2][2048 + 512], wa[2][4096];
  static int bitrev[2][512];
  const int W = n == 2048 ? 1 : 0;
  if (n != 2048 && n != 256) return -1;
  memset(&S, 0, sizeof(S));
  for (int i = 0; i < 1024; i++) ones[i] = 1.f;
  S.bs[0] = n; S.bs[1] = n;                  /* lW = nW = W: whole-length windows, all ones */
  S.win[0] = ones; S.win[1] = ones;
  vb200_exact_tables_mdct(n, T[W], bitrev[W]);
  memset(wa[W], 0, sizeof(wa[W]));
  vb200_exact_tables_drft(n, wa[W]);
  S.xm_trig[W] = T[W]; S.xm_bitrev[W] = bitrev[W]; S.xf_wa[W] = wa[W];
  const int lW = W, nW = W;
  threadIdx.x = 0; blockIdx.x = 0;
  if (W) k_transform_exact<2048, 1>(&S, 1, windowed, nullptr, &lW, &nW, nullptr, nullptr, gmdct, logfft, logmdct, ampmax, nullptr);
  else   k_transform_exact<256, 1>(&S, 0, windowed, nullptr, &lW, &nW, nullptr, nullptr, gmdct, logfft, logmdct, ampmax, nullptr);
  return 0;
}
