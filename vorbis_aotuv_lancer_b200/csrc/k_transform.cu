/* K1 -- window + forward MDCT + real FFT + dB spectra, one CTA per ch-block.
 *
 * Replaces, for every channel of every block (mapping0.c:783-902, 933-951):
 *   _postnoise_detection   psy.c:619-648      (on the un-windowed PCM)
 *   _vorbis_apply_window   window.c:2137-2261
 *   mdct_forward           mdct.c:1799-1869   out[k]=(4/n) sum x[t] cos(pi/(2n)(2t+1+n/2)(2k+1))
 *   drft_forward           smallft.c:6770     + the |X|^2 -> dB loop mapping0.c:848-888
 *   logmdct                mapping0.c:933-951
 *
 * B200 mapping: the block is staged once in shared memory with float4 loads, both
 * transforms run out of shared memory (MDCT as an n/4-point complex FFT with
 * pre/post twiddles, the real FFT as an n/2-point complex FFT + split), and the
 * three result vectors leave with coalesced stores.  HBM traffic = the algorithmic
 * minimum: n*4 bytes in, 3*N2*4 bytes out (20480 / 2560 B).
 * Not a port of the reference's butterflies: Stockham autosort radix-2 stages
 * with twiddles precomputed in double (setup.cu).
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

__device__ __forceinline__ float2 cmulf(float2 v, float2 w) {
  return make_float2(__fsub_rn(__fmul_rn(v.x, w.x), __fmul_rn(v.y, w.y)), __fadd_rn(__fmul_rn(v.x, w.y), __fmul_rn(v.y, w.x)));
}
/* Index swizzles of the first two passes.  A radix-4 pass writes b[4(j-k) + k + q*Ns]: for Ns = 1 and Ns = 4 the 16
   lanes of a half-warp would hit only 4 of the 16 eight-byte bank pairs (4-way conflicts).  XOR-ing two index bits with
   bits of i >> 4 makes the 16 slots distinct, and the following pass, which reads 16 consecutive aligned elements per
   half-warp, sees a permutation inside the aligned 16 -- conflict-free both ways, no padding. */
template <int MODE> __device__ __forceinline__ int swz(int i) {
  if (MODE == 1) return i ^ ((i >> 4) & 3);
  if (MODE == 2) return i ^ (((i >> 4) & 3) << 2);
  return i;
}
template <int M, int THREADS, int Ns, int RD, int WR, bool SYNC = true>
__device__ __forceinline__ void fft_pass4(const float2 *a, float2 *b, const float2 *__restrict__ ptw) {
  for (int j = threadIdx.x; j < M / 4; j += THREADS) {
    const int k = j & (Ns - 1);
    float2 v0 = a[swz<RD>(j)], v1 = a[swz<RD>(j + M / 4)], v2 = a[swz<RD>(j + M / 2)], v3 = a[swz<RD>(j + 3 * (M / 4))];
    if (Ns > 1) {
      /* this pass's twiddles w^1, w^2, w^3 (w = exp(-2*pi*i*k/(4*Ns))) at unit stride: a table of 8 KB for all
         passes that stays in L1 (setup.cu: fft_pass_tw) */
      const float2 *tp = ptw + (Ns - 4) + k;
      const float2 w1 = __ldg(tp), w2 = __ldg(tp + Ns), w3 = __ldg(tp + 2 * Ns);
      v1 = cmulf(v1, w1); v2 = cmulf(v2, w2); v3 = cmulf(v3, w3);
    }
    const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
    const int o = ((j - k) << 2) + k;
    b[swz<WR>(o)] = make_float2(s02.x + s13.x, s02.y + s13.y);
    b[swz<WR>(o + Ns)] = make_float2(d02.x + d13.y, d02.y - d13.x);          /* d02 - i*d13 */
    b[swz<WR>(o + 2 * Ns)] = make_float2(s02.x - s13.x, s02.y - s13.y);
    b[swz<WR>(o + 3 * Ns)] = make_float2(d02.x - d13.y, d02.y + d13.x);      /* d02 + i*d13 */
  }
  if (SYNC) __syncthreads();
}
/* the radix-2 pass that ends a transform whose log2 is odd (Ns == M/2), no barrier */
template <int M, int THREADS>
__device__ __forceinline__ void fft_pass2(const float2 *a, float2 *b, const float2 *__restrict__ ptw, int r2off) {
  for (int j = threadIdx.x; j < M / 2; j += THREADS) {
    const float2 w = __ldg(ptw + r2off + j);      /* exp(-2*pi*i*j/M) */
    const float2 u = a[j], t = cmulf(a[j + M / 2], w);
    b[j] = make_float2(u.x + t.x, u.y + t.y);
    b[j + M / 2] = make_float2(u.x - t.x, u.y - t.y);
  }
}
/* Both transforms of a block, pass by pass between the same barriers: the N/4-point one of the MDCT (h0 <-> h1) and
   the N/2-point one of the real FFT (xa <-> xb).  They are independent, have the same number of passes for n = 2048
   (4 x radix-4 + radix-2, 5 x radix-4) and one pass apart for n = 256; sharing the barriers halves their number and
   gives the threads the smaller transform leaves idle something to do.  Results: MDCT in h1; real FFT in xb (n = 2048)
   or xa (n = 256). */
template <int N, int THREADS>
__device__ __forceinline__ void cta_fft_pair(float2 *h0, float2 *h1, float2 *xa, float2 *xb, const float2 *__restrict__ ptw, int r2off) {
  constexpr int M1 = N / 4, M2 = N / 2;
  static_assert(N == 2048 || N == 256, "block sizes of the built path");
  fft_pass4<M1, THREADS, 1, 0, 1, false>(h0, h1, ptw); fft_pass4<M2, THREADS, 1, 0, 1, false>(xa, xb, ptw); __syncthreads();
  fft_pass4<M1, THREADS, 4, 1, 2, false>(h1, h0, ptw); fft_pass4<M2, THREADS, 4, 1, 2, false>(xb, xa, ptw); __syncthreads();
  fft_pass4<M1, THREADS, 16, 2, 0, false>(h0, h1, ptw); fft_pass4<M2, THREADS, 16, 2, 0, false>(xa, xb, ptw); __syncthreads();
  if (N == 2048) {
    fft_pass4<M1, THREADS, 64, 0, 0, false>(h1, h0, ptw); fft_pass4<M2, THREADS, 64, 0, 0, false>(xb, xa, ptw); __syncthreads();
    fft_pass2<M1, THREADS>(h0, h1, ptw, r2off); fft_pass4<M2, THREADS, 256, 0, 0, false>(xa, xb, ptw); __syncthreads();
  } else {
    fft_pass2<M2, THREADS>(xb, xa, ptw, r2off); __syncthreads();
  }
}
/* Complex FFTs, forward (exp(-i...)): Stockham autosort, radix-4 passes (plus one radix-2 pass when log2 M is odd).  A
   radix-4 pass reads a[j + r*M/4], multiplies by w^r, w = exp(-2*pi*i*k/(4*Ns)), k = j mod Ns, and writes the 4-point
   DFT to b[4(j-k) + k + q*Ns]. */
template <int N, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_transform(const DevSetup *__restrict__ S, int W, const float *__restrict__ pcm, const long long *__restrict__ pcm_off,
            const int *__restrict__ lWv, const int *__restrict__ nWv,
            const int *__restrict__ block_mode, const int *__restrict__ lW_block_mode,
            float *__restrict__ gmdct, float *__restrict__ logfft, float *__restrict__ logmdct,
            float *__restrict__ local_ampmax, float *__restrict__ poste) {
  constexpr int N2 = N / 2, N4 = N / 4;
  __shared__ __align__(16) float x[N];
  __shared__ __align__(16) float2 bufA[N2];
  __shared__ __align__(16) float2 bufB[N2];
  __shared__ double red[2][THREADS / 32];
  __shared__ float redf[THREADS / 32];

  const int item = blockIdx.x, tid = threadIdx.x;
  const float *src = pcm + (pcm_off ? pcm_off[item] : (long long)item * N);

  const float2 *ptw = S->fft_pass_tw[W]; const int r2off = S->fft_r2_off[W];
  /* ---- window geometry (window.c:2137-2261): every boundary is a multiple of 4 samples ---- */
  const int lW = W ? lWv[item] : 0, nW = W ? nWv[item] : 0;
  const int ln = S->bs[lW], rn = S->bs[nW];
  const float *wl = S->win[lW], *wr = S->win[nW];
  const int leftbegin = N / 4 - ln / 4, leftend = leftbegin + ln / 2;
  const int rightbegin = N / 2 + N / 4 - rn / 4, rightend = rightbegin + rn / 2;

  /* ---- aoTuV M2 post-echo detector on the un-windowed PCM (psy.c:619-648), read straight from the timeline ---- */
  if (poste) {
    float ret = -1.f;
    const int mode = block_mode[item], lwm = lW_block_mode[item];
    if (mode == 2 && lwm == 0 && N >= 2048) {   /* uniform per CTA */
      const int sn = N >> 2, mn = sn + sn, en = sn + (N >> 1);
      double up = 0., un = 0.;
      for (int i = sn + tid; i < mn; i += THREADS) up += (double)fabsf(__ldg(src + i));
      for (int i = mn + tid; i < en; i += THREADS) un += (double)fabsf(__ldg(src + i));
      up = warp_sum_d(up); un = warp_sum_d(un);
      if ((tid & 31) == 0) { red[0][tid >> 5] = up; red[1][tid >> 5] = un; }
      __syncthreads();
      if (tid == 0) {
        double upt = 0., unt = 0.;
        for (int w = 0; w < THREADS / 32; w++) { upt += red[0][w]; unt += red[1][w]; }
        if (!(unt / sn > 0.01)) {
          upt *= upt; unt *= unt; unt *= 15;
          if (upt > unt) { ret = (float)(upt - unt); if (ret < 0.1) ret = -1.0f; }
        }
        poste[item] = ret;
      }
    } else if (tid == 0) poste[item] = ret;
  }

  /* ---- stage the block, windowed on the way in: float4 of PCM x float4 of the window table ---- */
  if ((((uintptr_t)src) & 15) == 0) {
    for (int c = tid; c < N / 4; c += THREADS) {
      const int i = c * 4;
      float4 v = __ldg(reinterpret_cast<const float4 *>(src) + c);
      if (i < leftbegin || i >= rightend) v = make_float4(0.f, 0.f, 0.f, 0.f);
      else if (i < leftend) {
        const float4 w = __ldg(reinterpret_cast<const float4 *>(wl + (i - leftbegin)));
        v.x = __fmul_rn(v.x, w.x); v.y = __fmul_rn(v.y, w.y); v.z = __fmul_rn(v.z, w.z); v.w = __fmul_rn(v.w, w.w);
      } else if (i >= rightbegin) {
        /* the falling half reads the table backwards: sample i+k takes wr[rn/2-1-(i-rightbegin)-k] */
        const float4 w = __ldg(reinterpret_cast<const float4 *>(wr + (rn / 2 - 4 - (i - rightbegin))));
        v.x = __fmul_rn(v.x, w.w); v.y = __fmul_rn(v.y, w.z); v.z = __fmul_rn(v.z, w.y); v.w = __fmul_rn(v.w, w.x);
      }
      reinterpret_cast<float4 *>(x)[c] = v;
    }
  } else {
    for (int i = tid; i < N; i += THREADS) {
      float v = __ldg(src + i);
      if (i < leftbegin) v = 0.f;
      else if (i < leftend) v = __fmul_rn(v, __ldg(wl + (i - leftbegin)));
      else if (i >= rightend) v = 0.f;
      else if (i >= rightbegin) v = __fmul_rn(v, __ldg(wr + (rn / 2 - 1 - (i - rightbegin))));
      x[i] = v;
    }
  }
  __syncthreads();

  /* ---- MDCT: fold to N2 reals, pre-twiddle to N4 complex (into the first half of bufB) ---- */
  float2 *h0 = bufB, *h1 = bufB + N4;
  float2 *xa = reinterpret_cast<float2 *>(x), *xb = bufA;      /* the windowed block read as N2 complex numbers IS the
                                                                  real FFT's input: no copy, x is its ping buffer */
  const float2 *tw = S->mdct_tw[W];
  for (int m = tid; m < N4; m += THREADS) {
    /* u[i] = -x[3N2/2-1-i]-x[3N2/2+i]  (i<N2/2);  x[i-N2/2]-x[3N2/2-1-i]  (i>=N2/2) */
    const int i0 = 2 * m, i1 = N2 - 1 - 2 * m;
    float u0, u1;
    if (i0 < N2 / 2) u0 = -x[3 * N2 / 2 - 1 - i0] - x[3 * N2 / 2 + i0];
    else u0 = x[i0 - N2 / 2] - x[3 * N2 / 2 - 1 - i0];
    if (i1 < N2 / 2) u1 = -x[3 * N2 / 2 - 1 - i1] - x[3 * N2 / 2 + i1];
    else u1 = x[i1 - N2 / 2] - x[3 * N2 / 2 - 1 - i1];
    float2 t = __ldg(tw + m);
    h0[m] = make_float2(u0 * t.x - u1 * t.y, u0 * t.y + u1 * t.x);
  }
  __syncthreads();
  /* ---- both transforms, pass by pass ---- */
  cta_fft_pair<N, THREADS>(h0, h1, xa, xb, ptw, r2off);
  const float scale = 4.f / N;
  /* ---- MDCT post-twiddle: h1 -> interleaved result staged in h0 for coalesced stores ---- */
  float *out = reinterpret_cast<float *>(h0);
  for (int k = tid; k < N4; k += THREADS) {
    float2 t = __ldg(tw + k), z = h1[k];
    float yr = z.x * t.x - z.y * t.y, yi = z.x * t.y + z.y * t.x;
    out[2 * k] = yr * scale;
    out[N2 - 1 - 2 * k] = -yi * scale;
  }
  /* ---- real FFT: split of the N2-point complex transform, dB spectrum, block maximum ---- */
  {
    const float2 *Z = (N == 2048) ? xb : xa;
    const float2 *rt = S->rfft_tw[W];
    const float scale_dB = todB_p345(scale);
    float *lf = logfft + (size_t)item * N2;
    float mx = -3.0e38f;
    for (int k = tid; k < N2; k += THREADS) {
      float v;
      if (k == 0) {
        float re0 = Z[0].x + Z[0].y;          /* X[0] = sum x[t] */
        v = (float)((double)__fadd_rn(scale_dB, todB_raw(re0)) + .345);
      } else {
        float2 a = Z[k], b = Z[N2 - k];
        /* X[k] = (a+conj(b))/2 - i*w*(a-conj(b))/2,  w = exp(-2*pi*i*k/N) */
        float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);
        float orr = 0.5f * (a.x - b.x), oi = 0.5f * (a.y + b.y);
        float2 w = __ldg(rt + k);
        /* -i*w*(orr + i*oi) = (w.y*orr + w.x*oi) + i*(w.y*oi - w.x*orr) */
        float xr = er + (w.y * orr + w.x * oi);
        float xi = ei + (w.y * oi - w.x * orr);
        float p = __fadd_rn(__fmul_rn(xr, xr), __fmul_rn(xi, xi));
        v = (float)((double)__fadd_rn(scale_dB, __fmul_rn(.5f, todB_raw(p))) + .345);
      }
      lf[k] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max_f(mx);
    if ((tid & 31) == 0) redf[tid >> 5] = mx;
  }
  __syncthreads();                     /* the staged MDCT result and the warp maxima are complete */
  {
    float *gm = gmdct + (size_t)item * N2, *lm = logmdct + (size_t)item * N2;
    for (int i = tid; i < N2; i += THREADS) {
      float v = out[i];
      gm[i] = v;
      lm[i] = todB_p345(v);
    }
    if (tid == 0) {
      float m = redf[0];
      for (int w = 1; w < THREADS / 32; w++) m = fmaxf(m, redf[w]);
      if (m > 0.f) m = 0.f;
      local_ampmax[item] = m;
    }
  }
}

}  // namespace

int vb200_launch_transform(const vb200_setup *s, cudaStream_t st, int W, int items,
                           const float *pcm, const long long *pcm_off, const int *lW, const int *nW,
                           const int *block_mode, const int *lW_block_mode,
                           float *gmdct, float *logfft, float *logmdct, float *local_ampmax, float *poste) {
  if (items <= 0) return VB200_OK;
  if (s->exact_transform)          /* opt-in: the reference's own float operations (k_transform_exact.cu) */
    return vb200_launch_transform_exact(s, st, W, items, pcm, pcm_off, lW, nW, block_mode, lW_block_mode, gmdct, logfft, logmdct, local_ampmax, poste);
  if (W) k_transform<2048, 256><<<items, 256, 0, st>>>(s->d, 1, pcm, pcm_off, lW, nW, block_mode, lW_block_mode, gmdct, logfft, logmdct, local_ampmax, poste);
  else   k_transform<256, 64><<<items, 64, 0, st>>>(s->d, 0, pcm, pcm_off, lW, nW, block_mode, lW_block_mode, gmdct, logfft, logmdct, local_ampmax, poste);
  LAUNCH_CHECK();
  return VB200_OK;
}
