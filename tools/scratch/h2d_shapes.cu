// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/scratch/h2d_shapes tools/scratch/h2d_shapes.cu
// scratch: H2D rate of the segmenter's upload pattern (one 2-row 2D copy per stream) against plain 1D copies
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
int main() {
  const int S = 64, ch = 2; const long n = 2646000, lead = 1023 + 64, T = lead + n + 8192;
  std::vector<float*> h(S);
  for (int i = 0; i < S; i++) { cudaMallocHost(&h[i], sizeof(float) * ch * n); for (long k = 0; k < ch * n; k += 1024) h[i][k] = 1.f; }
  float *d; cudaMalloc(&d, sizeof(float) * (size_t)S * ch * T);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0, st);
      for (int i = 0; i < S; i++) {
        float *dst = d + (size_t)i * ch * T + (mode == 3 ? 1088 : lead);
        if (mode == 0 || mode == 3) cudaMemcpy2DAsync(dst, sizeof(float) * T, h[i], sizeof(float) * n, sizeof(float) * n, ch, cudaMemcpyHostToDevice, st);
        else if (mode == 1) for (int c = 0; c < ch; c++) cudaMemcpyAsync(dst + (size_t)c * T, h[i] + (size_t)c * n, sizeof(float) * n, cudaMemcpyHostToDevice, st);
        else cudaMemcpyAsync(d + (size_t)i * ch * n, h[i], sizeof(float) * ch * n, cudaMemcpyHostToDevice, st);
      }
      cudaEventRecord(e1, st); cudaStreamSynchronize(st);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("mode %d (%s): %.1f ms, %.1f GB/s\n", mode, mode == 0 ? "2D, odd offset" : mode == 1 ? "1D per row, odd offset" : mode == 2 ? "1D whole stream" : "2D, 64B-aligned offset", ms, S * ch * n * 4.0 / 1e6 / ms);
    }
  }
  return 0;
}
