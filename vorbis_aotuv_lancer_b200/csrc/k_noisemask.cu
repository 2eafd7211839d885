/* K2 -- noise masking, one small CTA per ch-block.
 *
 * Replaces _vp_noisemask (psy.c:3770-4074): two bark-window regressions
 * (bark_noise_hybridmp, psy.c:3480-3638), the aoTuV M7 tone-peak correction (ntfix,
 * psy.c:3645-3768), the noise compander, M2 post-echo cut, M8 (npeak) and M9 (epeak).
 *
 * B200 mapping.  The regression is built on five float32 prefix sums whose left-to-right
 * summation order is part of the result, so each of them is an
 * inherently serial 1024-step chain.  The kernel therefore (a) computes the per-bin terms in
 * parallel, (b) lets five lanes walk the five chains out of shared memory, (c) evaluates the
 * per-bin window fits in parallel again -- and keeps the CTA tiny (64 threads, ~29 KB of
 * shared memory) so that 7-8 CTAs are resident per SM and the serial chains of different
 * ch-blocks overlap.  HBM traffic is the algorithmic minimum: logmdct (+lastmdct for M9) in,
 * noise + epeak + npeak out.
 */
#include "vb200_internal.h"
#include "dev_util.cuh"
#include "k_common.cuh"

namespace {

template <int N>
struct NoiseSmem {
  float sums[5][N + 4];   /* N, X, XX, Y, XY prefix sums; +4 keeps the rows 16-byte aligned and the five chain
                             lanes (one 128-bit access each) on different banks */
  float noise[N];
  __device__ float *tmp() { return sums[1]; }      /* scratch of the M7 tone fix; the prefix sums are dead by then */
  __device__ float *inm() { return sums[2]; }
};

/* A, B, D of the regression window of bin ii in pass `pass` (0: the bin's bark window, 1: the fixed window), from the
   prefix sums (the parallel part of bark_noise_hybridmp) */
template <int N>
__device__ __forceinline__ void fit_abd(const NoiseSmem<N> &sm, const int *__restrict__ bark, int fixed, int pass, int b1, int ii,
                                        float &A, float &B, float &D) {
  int lo, hi;
  if (pass == 0) { const int b = __ldg(bark + ii); lo = b >> 16; hi = b & 0xffff; }
  else { hi = ii + fixed / 2; lo = hi - fixed; }
  float tN, tX, tXX, tY, tXY;
  if (ii < b1) {
    tN = __fadd_rn(sm.sums[0][hi], sm.sums[0][-lo]); tX = __fsub_rn(sm.sums[1][hi], sm.sums[1][-lo]);
    tXX = __fadd_rn(sm.sums[2][hi], sm.sums[2][-lo]); tY = __fadd_rn(sm.sums[3][hi], sm.sums[3][-lo]);
    tXY = __fsub_rn(sm.sums[4][hi], sm.sums[4][-lo]);
  } else {
    tN = __fsub_rn(sm.sums[0][hi], sm.sums[0][lo]); tX = __fsub_rn(sm.sums[1][hi], sm.sums[1][lo]);
    tXX = __fsub_rn(sm.sums[2][hi], sm.sums[2][lo]); tY = __fsub_rn(sm.sums[3][hi], sm.sums[3][lo]);
    tXY = __fsub_rn(sm.sums[4][hi], sm.sums[4][lo]);
  }
  A = __fsub_rn(__fmul_rn(tY, tXX), __fmul_rn(tX, tXY));
  B = __fsub_rn(__fmul_rn(tN, tXY), __fmul_rn(tX, tY));
  D = __fsub_rn(__fmul_rn(tN, tXX), __fmul_rn(tX, tX));
}

/* One bark_noise_hybridmp call: f(q, i) -> out(q, i, noise).  Bins keep their owner thread (i = tid + q*T) through the
   whole kernel, so the only barriers are the ones around the serial prefix chains: a bin above the last full window
   (i >= b2) recomputes the fit of bin b2-1 itself instead of waiting for it to be published, and the second (fixed
   window) pass follows the first in the same loop. */
template <int N, int T, typename F, typename OUT>
__device__ __forceinline__ void noise_fit(NoiseSmem<N> &sm, const DevPsy &P, F f, float offset, int fixed, OUT out) {
  const int tid = threadIdx.x;
  __syncthreads();                     /* every reader of the previous prefix sums is done */
#pragma unroll
  for (int q = 0; q < N / T; q++) {
    const int i = tid + q * T;
    float y = __fadd_rn(f(q, i), offset);
    if (y < 1.f) y = 1.f;
    float w = __fmul_rn(y, y);
    if (i == 0) {
      w = __fmul_rn(w, .5f);
      sm.sums[0][0] = w; sm.sums[1][0] = w; sm.sums[2][0] = 0.f; sm.sums[3][0] = __fmul_rn(w, y); sm.sums[4][0] = 0.f;
    } else {
      float x = (float)i, wx = __fmul_rn(w, x);
      sm.sums[0][i] = w; sm.sums[1][i] = wx; sm.sums[2][i] = __fmul_rn(wx, x);
      sm.sums[3][i] = __fmul_rn(w, y); sm.sums[4][i] = __fmul_rn(wx, y);
    }
  }
  __syncthreads();
  if (tid < 5) {                       /* the five serial float32 chains */
    float4 *r = reinterpret_cast<float4 *>(sm.sums[tid]);
    float acc = 0.f;
    auto run4 = [&](float4 &v) {
      acc = __fadd_rn(acc, v.x); v.x = acc; acc = __fadd_rn(acc, v.y); v.y = acc;
      acc = __fadd_rn(acc, v.z); v.z = acc; acc = __fadd_rn(acc, v.w); v.w = acc;
    };
    float4 a0 = r[0], a1 = r[1];
#pragma unroll 2
    for (int i = 0; i < N / 4; i += 4) {      /* two register sets: the next 8 values arrive while 8 are summed */
      float4 b0 = r[i + 2], b1 = r[i + 3];
      run4(a0); r[i] = a0; run4(a1); r[i + 1] = a1;
      a0 = r[i + 4]; a1 = r[i + 5];            /* last pass: 8 floats past the row, still inside NoiseSmem, never used */
      run4(b0); r[i + 2] = b0; run4(b1); r[i + 3] = b1;
    }
  }
  __syncthreads();
  const int b1a = P.nf_b1[0], b2a = P.nf_b2[0], b1f = P.nf_b1[1], b2f = P.nf_b2[1];
#pragma unroll
  for (int q = 0; q < N / T; q++) {
    const int i = tid + q * T;
    float A = 0.f, B = 0.f, D = 1.f;
    if (b2a > 0) fit_abd<N>(sm, P.bark, fixed, 0, b1a, i < b2a ? i : b2a - 1, A, B, D);
    float R = __fdiv_rn(__fadd_rn(A, __fmul_rn((float)i, B)), D);
    if (R < 0.f) R = 0.f;
    float nv = __fsub_rn(R, offset);
    if (fixed > 0) {
      if (b2f > 0) fit_abd<N>(sm, P.bark, fixed, 1, b1f, i < b2f ? i : b2f - 1, A, B, D);
      else if (b2a > 0) fit_abd<N>(sm, P.bark, fixed, 0, b1a, b2a - 1, A, B, D);      /* the reference's A, B, D carry over */
      const float v = __fsub_rn(__fdiv_rn(__fadd_rn(A, __fmul_rn((float)i, B)), D), offset);
      if (v < nv) nv = v;
    }
    out(q, i, nv);
  }
}

/* aoTuV M7 for short blocks (block_mode <= 1, psy.c:3661-3731) by all threads of the CTA.  The reference walks the
   bins and jumps to the end of a peak's range after handling it -- but that range is non-increasing, so no strict local
   maximum is skipped: the peaks are exactly the strict local maxima of [4, nx), each a function of the spectrum
   alone.  temp[] starts at 0 and only ever takes max(ss, temp) clamped at 0, so the order of the peaks does not matter
   either: bin j ends with the largest ss of the <= 6 peaks (i in j-3 .. j+2) whose range ps..pe covers it.  Three
   steps: the compressed spectrum, one thread per candidate peak (ps, pe, ss), one thread per bin (max, cap, subtract).
   (Lane 0 used to walk all of it: two thirds of the short-block kernel's instructions ran on one lane.) */
template <int N, int T>
__device__ __forceinline__ void tone_fix_short(NoiseSmem<N> &sm, const DevPsy &P, const float *__restrict__ spec) {
  int nx = P.tonefix_end;
  if (!nx) return;
  const int tid = threadIdx.x;
  const float *off1 = P.noiseoffset + N;
  const float limit = fabsf(__ldg(off1));
  float *noise = sm.sums[3];             /* logmdct - noise fit, staged here by the kernel */
  float *ss_ = sm.tmp(), *inm_ = sm.inm();
  int *rng_ = reinterpret_cast<int *>(sm.sums[0]);       /* ps | pe << 16 of a peak */
  const int up = 3, un = 4;
  int nxp = nx + un;
  const float tol = (N == 256) ? 15.f : 9.f, strength = .6f;
  if (nxp > N) { nx = N; nxp = N - un; }
  for (int i = tid; i < N; i += T) {
    float v = 0.f;
    if (i < nxp) { const float s = spec[i]; v = (s < -70.f) ? (float)(-70. + (double)__fadd_rn(s, 70.f) * .1) : s; }
    inm_[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < N; i += T) {
    float ss = -1.f; int rng = 0;
    if (i >= un && i < nx && spec[i] > spec[i - 1] && spec[i] > spec[i + 1]) {
      int ps = i - 1, pe = i + 1;
      for (int j = ps; j > i - up; j--) { if (spec[j + 1] < spec[j]) break; ps = j; }
      for (int j = pe; j < i + un; j++) { if (spec[j - 1] < spec[j]) break; pe = j; }
      float v = fmaxf(__fsub_rn(inm_[i], inm_[ps]), __fsub_rn(inm_[i], inm_[pe]));
      if (v > tol) {
        if (spec[i] > noise[i]) { v = __fsub_rn(v, tol); v = __fmul_rn(v, strength); }
        ss = v; rng = ps | (pe << 16);
      }
    }
    ss_[i] = ss; rng_[i] = rng;
  }
  __syncthreads();
  float cut[N / T];
#pragma unroll
  for (int q = 0; q < N / T; q++) {
    const int j = tid + q * T;
    float t = 0.f;
    for (int i = (j - 3 < un ? un : j - 3); i <= j + 2 && i < nx; i++) {
      const float ss = ss_[i];
      const int rng = rng_[i];
      if (ss >= 0.f && j >= (rng & 0xffff) && j <= (rng >> 16)) { float m = fmaxf(ss, t); if (m < 0.f) m = 0.f; t = m; }
    }
    cut[q] = t;
  }
#pragma unroll
  for (int q = 0; q < N / T; q++) {
    const int j = tid + q * T;
    if (j >= un - 1 && j < nx) {
      const float cap = fminf(__ldg(P.ntfix + j), __fadd_rn(__ldg(off1 + j), limit));
      float t = cut[q];
      if (t > cap) t = cap;
      noise[j] = __fsub_rn(noise[j], t);
    }
  }
}

#ifndef NM_T
#define NM_T 256          /* threads of a long ch-block's CTA */
#endif
#ifndef NM_MINB
#define NM_MINB 8         /* resident CTAs per SM the compiler must leave room for */
#endif
template <int N, int T, typename LASTREF>
__global__ void __launch_bounds__(T, (T >= 128 ? NM_MINB : 1))
k_noisemask(const DevSetup *__restrict__ S, const int *__restrict__ psyv, const int *__restrict__ block_mode,
            const float *__restrict__ lowcomp, const float *__restrict__ postev,
            const float *__restrict__ logmdct, LASTREF last,
            float *__restrict__ noise_out, float *__restrict__ epeak_out, float *__restrict__ npeak_out, int npeak_stride) {
  extern __shared__ __align__(16) unsigned char smraw[];
  NoiseSmem<N> &sm = *reinterpret_cast<NoiseSmem<N> *>(smraw);
  const int item = blockIdx.x, tid = threadIdx.x;
  const DevPsy &P = S->psy[psyv[item]];
  const float *lm = logmdct + (size_t)item * N;
  const int mode = block_mode[item];
  const float level = lowcomp[item], poste = postev[item];
  const int part = P.partition;

  /* this thread's bins tid + q*T: logmdct (and, for M9, lastmdct) are read once and stay in registers */
  float lmr[N / T], wk[N / T];
#pragma unroll
  for (int q = 0; q < N / T; q++) lmr[q] = __ldg(lm + tid + q * T);
  const bool m9 = mode > 1;

  /* wk: logmdct minus the first fit, then logmdct minus that -- thread-owned like the bins, so it never needs shared
     memory (24.6 KB per CTA instead of 28.7: 8 CTAs per SM) */
  noise_fit<N, T>(sm, P, [&](int q, int) { return lmr[q]; }, 140.f, -1,
                  [&](int q, int, float nv) { wk[q] = __fsub_rn(lmr[q], nv); });
  noise_fit<N, T>(sm, P, [&](int q, int) { return wk[q]; }, 0.f, P.noisewindowfixed,
                  [&](int q, int i, float nv) { sm.noise[i] = nv; wk[q] = __fsub_rn(lmr[q], wk[q]); });
  if (mode <= 2 && P.tonefix_end) {        /* uniform per CTA; long blocks: transitions only (one block in ten) */
    __syncthreads();
#pragma unroll
    for (int q = 0; q < N / T; q++) sm.sums[3][tid + q * T] = wk[q];
    if constexpr (N >= 256) {
      /* a long block is a transition block here (block_mode = blocktype | W << 1, mapping0.c:774): psy.c:3735-3766.
         The group means come from the unmodified vector and every cut depends on the means only, so the walk over the
         peaks is three parallel steps: means (thread per group), thresholds (thread per candidate peak), and each bin
         takes the cuts of the <= 7 peaks whose range covers it, in peak order (the order of the float subtractions) */
      float *tmp_ = sm.sums[1], *th_ = sm.sums[2];
      int *cut0 = reinterpret_cast<int *>(sm.sums[2] + 256);
      const int nx = P.tonefix_end, ng = (nx + 7) >> 3, gp = nx >> 3;
      for (int k = tid; k < 256; k += T) tmp_[k] = 0.f;
      __syncthreads();
      for (int k = tid; k < ng; k += T) {
        double a = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) a += (double)sm.sums[3][k * 8 + j];
        tmp_[k] = (float)(a / 8);
      }
      __syncthreads();
      const float *off1 = P.noiseoffset + N;
      const float limit = fabsf(__ldg(off1));
      for (int i = tid; i < 256; i += T) {
        float th = 0.f; int a8 = -1;
        if (i >= 3 && i < gp && tmp_[i] > tmp_[i - 1] && tmp_[i] > tmp_[i + 1]) {
          float base; int a;
          if (tmp_[i - 1] > tmp_[i - 2]) { base = tmp_[i - 2]; a = i - 3; } else { base = tmp_[i - 1]; a = i - 2; }
          const float t = __fsub_rn(tmp_[i], base);
          if (t > 2.f) {
            const float cap = fminf(__ldg(P.ntfix + i * 8), __fadd_rn(__ldg(off1 + i * 8), limit));
            th = fminf(__fsub_rn(t, 2.f), cap); a8 = a * 8;
          }
        }
        th_[i] = th; cut0[i] = a8;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < N / T; q++) {
        const int j = tid + q * T, g = j >> 3;
        float v = wk[q];
        for (int i = g < 6 ? 3 : g - 3; i <= g + 3 && i < gp; i++) {
          const int a8 = cut0[i];
          if (a8 >= 0 && j >= a8 && j <= (i + 3) * 8) v = __fsub_rn(v, th_[i]);
        }
        wk[q] = v;
      }
    } else {
      __syncthreads();
      tone_fix_short<N, T>(sm, P, lm);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < N / T; q++) wk[q] = sm.sums[3][tid + q * T];
    }
  }
  __syncthreads();                     /* the prefix sums are dead: their rows are reused below */

  /* compander (psy.c:3878-4012); epeak staged in sums[0] for M9.  The companded noise and logmdct also go to two
     padded copies (index + index/32: a partition's bins then lie in different banks for every partition), which the
     per-partition pass below walks one thread per partition without bank conflicts */
  float *ep = sm.sums[0];
  float *npad = sm.sums[1], *lpad = sm.sums[3];          /* N + N/32 floats each: they run over into rows 2 and 4 */
#pragma unroll
  for (int q = 0; q < N / T; q++) {
    const int i = tid + q * T;
    int dB = (int)((double)sm.noise[i] + .5);
    if (dB >= VB200_COMPAND_LEVELS) dB = VB200_COMPAND_LEVELS - 1;
    if (dB < 0) dB = 0;
    const float w = wk[q], c = P.compand[dB];
    ep[i] = __fadd_rn(w, (float)S->stn_compand[dB]);
    float v = __fadd_rn(w, c);
    if (level > 0.f && i < P.n33p) v = __fsub_rn(v, __fmul_rn(__fsub_rn(c, P.compand_hi[dB]), level));
    sm.noise[i] = v;
    npad[i + (i >> 5)] = v; lpad[i + (i >> 5)] = lmr[q];
  }
  /* lastmdct for M9: requested here, consumed after the per-partition pass */
  float lastr[N / T];
#pragma unroll
  for (int q = 0; q < N / T; q++) lastr[q] = (m9 && tid + q * T < P.tonecomp_endp) ? last.at(item, tid + q * T) : 0.f;
  __syncthreads();

  /* per-partition work: M2 post-echo cut then M8 npeak (psy.c:4025-4063); one lane per partition */
  const int nparts = N / part;
  const float *off1 = P.noiseoffset + N;
  for (int k = tid; k < nparts; k += T) {
    const int i = k * part;
    float np = 0.f;
    if (i < P.min_nn_lp) {
      if (poste > 0.f) {
        float t = fminf(fminf(poste, 30.f), __fadd_rn(__ldg(off1 + i), 30.f));
        if (t > 0.f) {
          np = -1.f;
          for (int j = 0; j < part; j++) { const float v = __fsub_rn(sm.noise[i + j], t); sm.noise[i + j] = v; npad[i + j + ((i + j) >> 5)] = v; }
        }
      }
      const float nt = 4.f;
      const float o = __fadd_rn(__ldg(off1 + i + part - 1), 6.f);
      if (o > 0.f && !(np < -0.5f)) {
        float me = 0.f, av = 0.f;
#pragma unroll 8
        for (int j = 0; j < part; j++) {
          const int pi = i + j + ((i + j) >> 5);
          const float l = lpad[pi];
          const float t = __fsub_rn(l, npad[pi]);
          if (me < t) me = t;
          av = __fadd_rn(av, l);
        }
        if (!(av < (float)(-95 * part)) && me < nt) np = __fdiv_rn(fminf(o, __fsub_rn(nt, me)), nt);
      }
    }
    npeak_out[(size_t)item * npeak_stride + k] = np;
  }
  __syncthreads();

  /* M9 (psy.c:4065-4073) + stores */
  float *no = noise_out + (size_t)item * N, *eo = epeak_out + (size_t)item * N;
#pragma unroll
  for (int q = 0; q < N / T; q++) {
    const int i = tid + q * T;
    float e = 0.f;
    if (m9 && i < P.tonecomp_endp) {
      const float l = lmr[q];
      if (__fsub_rn(l, ep[i]) >= 12.f) { const float mi = __fsub_rn(l, lastr[q]); if (mi >= 1.f) e = mi; }
    }
    eo[i] = e;
    no[i] = sm.noise[i];
  }
}

}  // namespace

template <int N, int T, typename LASTREF>
static int launch_nm(const vb200_setup *s, cudaStream_t st, int items, const int *psy, const int *block_mode, const float *lowcomp,
                     const float *poste, const float *logmdct, LASTREF last, float *noise, float *epeak, float *npeak, int npeak_stride) {
  static bool attr = false;
  if (!attr && sizeof(NoiseSmem<N>) > 48 * 1024 - 64) {
    CUDA_TRY(cudaFuncSetAttribute(k_noisemask<N, T, LASTREF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NoiseSmem<N>)));
    attr = true;
  }
  k_noisemask<N, T, LASTREF><<<items, T, sizeof(NoiseSmem<N>), st>>>(s->d, psy, block_mode, lowcomp, poste, logmdct, last, noise, epeak, npeak, npeak_stride);
  LAUNCH_CHECK();
  return VB200_OK;
}

/* off2 / rule2: the second lastmdct source of managed mode (k_common.cuh), or NULL */
int vb200_launch_noisemask(const vb200_setup *s, cudaStream_t st, int W, int items, const int *psy,
                           const int *block_mode, const float *lowcomp, const float *poste,
                           const float *logmdct, LastRef last, const long long *off2, const int *rule2,
                           float *noise, float *epeak, float *npeak, int npeak_stride) {
  if (items <= 0) return VB200_OK;
  if (off2) {
    LastRef2 l2{last, off2, rule2};
    if (W) return launch_nm<1024, NM_T, LastRef2>(s, st, items, psy, block_mode, lowcomp, poste, logmdct, l2, noise, epeak, npeak, npeak_stride);
    return launch_nm<128, 32, LastRef2>(s, st, items, psy, block_mode, lowcomp, poste, logmdct, l2, noise, epeak, npeak, npeak_stride);
  }
  if (W) return launch_nm<1024, NM_T, LastRef>(s, st, items, psy, block_mode, lowcomp, poste, logmdct, last, noise, epeak, npeak, npeak_stride);
  return launch_nm<128, 32, LastRef>(s, st, items, psy, block_mode, lowcomp, poste, logmdct, last, noise, epeak, npeak, npeak_stride);
}
