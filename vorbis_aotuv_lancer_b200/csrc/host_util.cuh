/* RAII device buffers for the stage calls (host pointers in, host pointers out). */
#ifndef VB200_HOST_UTIL_CUH
#define VB200_HOST_UTIL_CUH
#include <cuda_runtime.h>
#include "vb200_internal.h"

template <typename T>
struct DevArr {
  T *p = nullptr; size_t n = 0; cudaError_t err = cudaSuccess;
  DevArr() {}
  explicit DevArr(size_t count) { alloc(count); }
  DevArr(const T *host, size_t count, cudaStream_t st) { alloc(count); if (host && p && count) err = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st); }
  void alloc(size_t count) { n = count; err = cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T)); if (err != cudaSuccess) p = nullptr; }
  cudaError_t down(T *host, cudaStream_t st, size_t count = (size_t)-1) {
    if (count == (size_t)-1) count = n;
    if (!host || !count) return cudaSuccess;
    return cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, st);
  }
  ~DevArr() { if (p) cudaFree(p); }
  DevArr(const DevArr &) = delete; DevArr &operator=(const DevArr &) = delete;
  operator T *() { return p; }
};

#define STAGE_PROLOGUE(s) \
  if (!(s)) { vb200_set_error("%s: NULL setup", __func__); return VB200_EINVAL; } \
  if ((s)->device < 0 || !(s)->d) { vb200_set_error("%s: setup has no CUDA device (no CPU fallback)", __func__); return VB200_EFAULT; } \
  CUDA_TRY(cudaSetDevice((s)->device)); \
  cudaStream_t st = (cudaStream_t)(s)->stream;
#define CHECK_ALLOC(a) do{ if ((a).err != cudaSuccess) { vb200_set_error("%s: device allocation/copy failed: %s", __func__, cudaGetErrorString((a).err)); return VB200_EFAULT; } }while(0)
#endif
