/* TEST INFRASTRUCTURE -- compiles the K1x kernel (vorbis_aotuv_lancer_b200/csrc/k_transform_exact.cu) as a host
 * program and runs one CTA of it with REAL threads: every CUDA thread of the block is an OS thread (threadIdx is
 * thread-local), __syncthreads()/__syncwarp() are a pthread barrier over the block, __shared__ arrays are statics, the
 * warp reductions go through a scratch array between barriers, and the device intrinsics are their IEEE host
 * equivalents (built with -ffp-contract=off).  This checks the kernel's index arithmetic, its thread-to-item mappings,
 * stage order, barrier placement and buffer reuse on the CPU against the oracle and the reference (a missing barrier
 * shows up as a mismatch sooner or later because the OS schedules the threads freely; the run under
 * -fsanitize=thread in tests/test_k1x_emulation.py reports it deterministically).  It is not a way to run the product:
 * nothing in the package can reach it.  Built and used by tests/test_k1x_emulation.py only. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include <vector_types.h>          /* float2 / int2 / int4 of the setup structs */

static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

#define VB200_DEV_UTIL_CUH          /* dev_util.cuh is replaced by the host equivalents below */
#define FULLMASK 0xffffffffu
#undef __device__
#undef __global__
#undef __host__
#undef __forceinline__
#undef __shared__
#undef __align__
#undef __launch_bounds__
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
struct Idx { int x; };
static thread_local Idx threadIdx;
static Idx blockIdx;
static pthread_barrier_t g_bar;
static int g_threads;
static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline void __syncwarp() { pthread_barrier_wait(&g_bar); }          /* a 32-thread block is one warp */
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
/* warp reductions: every thread of the block calls them together (they sit in block-uniform code) */
static double g_scr_d[1024];
static float g_scr_f[1024];
static inline double warp_sum_d(double v) {
  g_scr_d[threadIdx.x] = v;
  pthread_barrier_wait(&g_bar);
  const int w0 = threadIdx.x & ~31;
  double s = 0.;
  /* the device sums by xor-butterfly; the post-echo detector only compares magnitudes far from its thresholds on the
     test material, and the GPU test covers the real order */
  for (int i = w0; i < w0 + 32 && i < g_threads; i++) s += g_scr_d[i];
  pthread_barrier_wait(&g_bar);
  return s;
}
static inline float warp_max_f(float v) {
  g_scr_f[threadIdx.x] = v;
  pthread_barrier_wait(&g_bar);
  const int w0 = threadIdx.x & ~31;
  float m = v;
  for (int i = w0; i < w0 + 32 && i < g_threads; i++) m = fmaxf(m, g_scr_f[i]);
  pthread_barrier_wait(&g_bar);
  return m;
}
static inline float todB_raw(float x) {
  uint32_t i; memcpy(&i, &x, 4); i &= 0x7fffffffu;
  return (float)i * 7.17711438e-7f + -764.6161886f;
}
static inline float todB_p345(float x) { return (float)((double)todB_raw(x) + .345); }

#include "../../vorbis_aotuv_lancer_b200/csrc/k_transform_exact.cu"

struct Job { int n, tid; const DevSetup *S; const float *in; const int *lW, *nW; float *gmdct, *logfft, *logmdct, *ampmax; };
static void *run_thread(void *p) {
  const Job *j = (const Job *)p;
  threadIdx.x = j->tid;
  if (j->n == 2048) k_transform_exact<2048, 256>(j->S, 1, j->in, nullptr, j->lW, j->nW, nullptr, nullptr, j->gmdct, j->logfft, j->logmdct, j->ampmax, nullptr);
  else              k_transform_exact<256, 32>(j->S, 0, j->in, nullptr, j->lW, j->nW, nullptr, nullptr, j->gmdct, j->logfft, j->logmdct, j->ampmax, nullptr);
  return nullptr;
}

/* one windowed block in (the window is applied by the kernel: windows of ones over a long-long-long block leave it
   as it is), the four outputs of the front end out */
extern "C" int k1x_emu(int n, const float *windowed, float *gmdct, float *logfft, float *logmdct, float *ampmax) {
  static DevSetup S;
  static float ones[1024] __attribute__((aligned(16))), T[2][2048 + 512] __attribute__((aligned(16))), wa[2][4096] __attribute__((aligned(16)));
  static int bitrev[2][512];
  static float2 stw[2][512];
  static int2 rev[2][256];
  static float in[2048] __attribute__((aligned(16)));
  const int W = n == 2048 ? 1 : 0;
  if (n != 2048 && n != 256) return -1;
  memset(&S, 0, sizeof(S));
  for (int i = 0; i < 1024; i++) ones[i] = 1.f;
  memcpy(in, windowed, sizeof(float) * n);
  S.bs[0] = n; S.bs[1] = n;                  /* lW = nW = W: whole-length windows, all ones */
  S.win[0] = ones; S.win[1] = ones;
  vb200_exact_tables_mdct(n, T[W], bitrev[W]);
  memset(wa[W], 0, sizeof(wa[W]));
  vb200_exact_tables_drft(n, wa[W]);
  vb200_exact_tables_k1x(n, T[W], bitrev[W], stw[W], rev[W]);
  S.xm_trig[W] = T[W]; S.xm_bitrev[W] = bitrev[W]; S.xf_wa[W] = wa[W]; S.xm_stage[W] = stw[W]; S.xm_rev[W] = rev[W];
  const int lW = W, nW = W;
  const int threads = n / 8;
  g_threads = threads;
  blockIdx.x = 0;
  pthread_barrier_init(&g_bar, nullptr, threads);
  std::vector<pthread_t> th(threads);
  std::vector<Job> jobs(threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = Job{n, t, &S, in, &lW, &nW, gmdct, logfft, logmdct, ampmax};
    pthread_create(&th[t], nullptr, run_thread, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], nullptr);
  pthread_barrier_destroy(&g_bar);
  return 0;
}
