/* Small device helpers shared by the kernels. */
#ifndef VB200_DEV_UTIL_CUH
#define VB200_DEV_UTIL_CUH
#include <cuda_runtime.h>
#include <stdint.h>

#define FULLMASK 0xffffffffu

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULLMASK, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULLMASK, v, o));
  return v;
}

/* todB of scales.h:43-51: a piecewise-linear log2 read off the IEEE bit pattern,
   (float)(bits&0x7fffffff) * 7.17711438e-7f - 764.6161886f, every op rounded to float, never fused. */
__device__ __forceinline__ float todB_raw(float x) {
  uint32_t i = __float_as_uint(x) & 0x7fffffffu;
  return __fadd_rn(__fmul_rn(__uint2float_rn(i), 7.17711438e-7f), -764.6161886f);
}
/* `todB(x) + .345` as written at every call site: the literal is a double, so the add is done in
   double and then narrowed. */
__device__ __forceinline__ float todB_p345(float x) {
  return (float)((double)todB_raw(x) + .345);
}

/* FLOOR1_fromdB_LOOKUP (psy.c:4517-4582 / floor1.c:301-366) regenerated: entry i is
   10^((i*0.5-127.5)/20)... kept as a table in constant memory, filled by the host at load time
   from the closed form and checked against the reference values in tests. */

#endif
