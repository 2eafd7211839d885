/* K6 -- aoTuV channel coupling + quantisation + noise normalisation, one warp per block.
 *
 * Replaces _vp_couple_quantize_normalize (psy.c:4858-5142) with flag_lossless (:4584),
 * lossless_coupling[f] (:4626,:4643), min_indemnity_dipole_hypot (:4660), noise_normalize
 * (:4732) and ssort (:4709).
 *
 * B200 mapping.  The partition loop carries state from partition to partition (aoTuV M6's
 * side_resdef, the nepeak edits), so it is serial; inside a partition the <=32 bins map onto the
 * 32 lanes of one warp and all channels of the block live in that warp's shared-memory slice.
 * The two order-sensitive float accumulations (the threshold ramp of flag_lossless and M6's
 * residue deflection) are replayed in bin order with warp shuffles; integer votes use ballots.
 * One warp per block keeps the CTA at ~6 KB so many blocks are resident per SM.
 * HBM traffic: mdct + epeak + ilogmask in, quantised residue out (in place over ilogmask).
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {


template <int CH>
struct CoupleSmem {
  float raw[CH][32], quant[CH][32], floor[CH][32], res[CH][32];
  int flag[CH][32];
  int out[CH][32];
  int cand[32];
  __align__(16) float dsum[32];
};

/* the next partition's inputs are fetched with asynchronous global->shared copies while the current one is
   worked on: the partition loop is serial, so without this every partition starts with an exposed HBM round trip */
template <int CH>
struct CoupleStage { float md[2][CH][32], ep[2][CH][32]; int iw[2][CH][32]; };

__device__ __forceinline__ void cp_async4(void *smem, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int Pending> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(Pending) : "memory"); }

__device__ __forceinline__ bool opposite(float a, float b) { return (a > 0.f && b < 0.f) || (b > 0.f && a < 0.f); }
__device__ __forceinline__ bool big(float x) { return x < -0.5f || x >= 0.5f; }
__device__ __forceinline__ float unitnorm(float x) { return __uint_as_float((__float_as_uint(x) & 0x80000000u) | 0x3f800000u); }
__device__ __forceinline__ float sqrt_via_double(float x) { return (float)sqrt((double)x); }

/* min_indemnity_dipole_hypot, psy.c:4660-4674 */
__device__ __forceinline__ float dipole(float a, float b, float threv) {
  const float th = (float)0.94;
  const float a2 = fabsf(__fmul_rn(a, th)), b2 = fabsf(__fmul_rn(b, th));
  if (a > 0.f) {
    if (b > 0.f) return __fadd_rn(a2, b2);
    if (a > -b) return __fsub_rn(a2, __fmul_rn(b2, threv));
    return -__fsub_rn(b2, __fmul_rn(a2, threv));
  }
  if (b < 0.f) return -__fadd_rn(a2, b2);
  if (-a > b) return -__fsub_rn(a2, __fmul_rn(b2, threv));
  return __fsub_rn(b2, __fmul_rn(a2, threv));
}

/* noise_normalize (psy.c:4732-4854) for channel row k of the current partition; lane j = bin j */
template <int CH>
__device__ void normalize(CoupleSmem<CH> &sm, const DevPsy &P, int limit, int k, bool use_flags, float nepeak, int i, int jn, int lane) {
  int start = P.normal_p ? P.normal_start - i : jn;
  if (start > jn || nepeak < -0.5f) start = jn;
  const int j = lane;
  bool cand = false;
  float ve_c = 0.f;
  if (j < jn) {
    if (j < start) {
      if (!use_flags) sm.out[k][j] = (int)rint((double)sm.res[k][j]);
      else if (sm.flag[k][j] != 1) {
        const float ve = sqrt_via_double(__fdiv_rn(sm.quant[k][j], sm.floor[k][j]));
        const int r = (int)rint((double)ve);
        if (sm.raw[k][j] < 0.f) { sm.out[k][j] = -r; sm.res[k][j] = -ve; } else { sm.out[k][j] = r; sm.res[k][j] = ve; }
      }
    } else if (use_flags) {
      if (sm.flag[k][j] != 1) {
        float ve = __fdiv_rn(sm.quant[k][j], sm.floor[k][j]);
        if (ve < .25f && j >= limit - i) {
          cand = true; ve_c = ve;
          const float sq = sqrt_via_double(ve);
          sm.res[k][j] = (sm.raw[k][j] < 0.f) ? -sq : sq;
        } else {
          ve = sqrt_via_double(ve);
          const int r = (int)rint((double)ve);
          int o;
          if (sm.raw[k][j] < 0.f) { o = -r; sm.res[k][j] = -ve; } else { o = r; sm.res[k][j] = ve; }
          sm.out[k][j] = o;
          sm.quant[k][j] = __fmul_rn((float)(o * o), sm.floor[k][j]);
        }
      }
    } else {
      const float r = sm.res[k][j];
      const float ve = __fmul_rn(r, r);
      if (ve < .25f) { cand = true; ve_c = ve; }
      else { const int o = (int)rint((double)r); sm.out[k][j] = o; sm.quant[k][j] = __fmul_rn((float)(o * o), sm.floor[k][j]); }
    }
  }
  const unsigned cmask = __ballot_sync(FULLMASK, cand);
  if (cmask == 0) return;              /* uniform */
  /* energy of the candidates, summed in bin order */
  float acc = 0.f;
  int count = 0;
  for (unsigned m = cmask; m; m &= m - 1) {
    const int src = __ffs(m) - 1;
    acc = __fadd_rn(acc, __shfl_sync(FULLMASK, ve_c, src));
    if (lane == 0) sm.cand[count] = src;
    count++;
  }
  acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(acc, nepeak), nepeak));
  __syncwarp();
  if (lane == 0) {
    int top = ((int)acc) + 1;
    if (top > jn) top = jn;
    if (count < top) top = count;
    for (int a = 0; a < top; a++) {            /* partial selection sort, largest first, ties as ssort breaks them */
      int best = a;
      for (int b = a + 1; b < count; b++) if (sm.quant[k][sm.cand[best]] < sm.quant[k][sm.cand[b]]) best = b;
      const int t = sm.cand[a]; sm.cand[a] = sm.cand[best]; sm.cand[best] = t;
    }
    for (int c = 0; c < count; c++) {
      const int e = sm.cand[c];
      if ((double)acc >= P.normal_thresh) { sm.out[k][e] = (int)unitnorm(sm.raw[k][e]); acc = __fsub_rn(acc, 1.f); sm.quant[k][e] = sm.floor[k][e]; }
      else { sm.out[k][e] = 0; sm.quant[k][e] = 0.f; }
    }
  }
  __syncwarp();
}

/* CHC > 0 / STEPSC >= 0: channel count and coupling steps as compile-time facts (2, 1: stereo; 6, 0 and 6, 4: the
   uncoupled and the coupled 5.1 mappings) -- the channel and step loops unroll and index shared memory statically */
template <int CH, int CHC, int STEPSC>
__global__ void __launch_bounds__(32)
k_couple(const DevSetup *__restrict__ S, int W, const int *__restrict__ psyv, int blobno, int n, int ch_, int map_index,
         const float *__restrict__ mdct, const float *__restrict__ epeak, float *__restrict__ npeak, int npeak_stride,
         int *__restrict__ iwork, int *__restrict__ nonzero) {
  __shared__ CoupleSmem<CH> sm;
  __shared__ CoupleStage<CH> stg;
  __shared__ float s_np[CH];
  const int blk = blockIdx.x, lane = threadIdx.x;
  const DevPsy &P = S->psy[psyv[blk]];
  const DevMap &M = S->map[map_index];
  const int ch = CHC ? CHC : ch_;
  const int part = P.partition, steps = STEPSC >= 0 ? STEPSC : M.coupling_steps;
  const int limit = S->pointlimit[P.blockflag][blobno];
  const float pre = (float)S->stereo_th[S->prepointamp[blobno]], post = (float)S->stereo_th[S->postpointamp[blobno]];
  float pre_x = (float)S->stereo_thx[S->prepointamp[blobno]], post_x = (float)S->stereo_thx[S->postpointamp[blobno]];
  const float prae = (steps == 1) ? (float)0.34 : (float)0.825;
  const int sliding_lowpass = S->sliding_lowpass[W][blobno];
  int lowpassr = S->lowpassr[W];
  if (lowpassr % P.normal_partition) lowpassr = (lowpassr / P.normal_partition + 1) * P.normal_partition;   /* mapping0.c:778-781 */
  if (pre_x < pre) pre_x = pre;
  if (post_x < pre) post_x = pre;
  constexpr int NC = CH <= 2 ? 1 : VB200_MAX_COUPLING;     /* coupling steps a CH-channel mapping can hold */
  float resdef[NC];
#pragma unroll
  for (int k = 0; k < NC; k++) resdef[k] = -1.f;

  const float *md = mdct + (size_t)blk * ch * n, *ep = epeak + (size_t)blk * ch * n;
  float *npk = npeak + (size_t)blk * ch * npeak_stride;
  int *iw = iwork + (size_t)blk * ch * n;
  int *nzg = nonzero + (size_t)blk * ch;
  unsigned nzmask0 = 0;
  for (int k = 0; k < ch; k++) if (nzg[k]) nzmask0 |= 1u << k;

  auto prefetch = [&](int buf, int i) {
    if (i < lowpassr && lane < n - i && lane < part)
      for (int k = 0; k < ch; k++)
        if (nzmask0 & (1u << k)) {
          cp_async4(&stg.md[buf][k][lane], md + (size_t)k * n + i + lane);
          cp_async4(&stg.ep[buf][k][lane], ep + (size_t)k * n + i + lane);
          cp_async4(&stg.iw[buf][k][lane], iw + (size_t)k * n + i + lane);
        }
    cp_async_commit();
  };
  prefetch(0, 0);
  for (int i = 0, pi = 0; i < lowpassr; i += part, pi++) {
    const int jn = part > n - i ? n - i : part;
    const int j = lane;
    const int buf = pi & 1;
    unsigned nz = nzmask0;
    prefetch(buf ^ 1, i + part);
    cp_async_wait<1>();
    if (lane < ch) s_np[lane] = npk[lane * npeak_stride + pi];
    __syncwarp();
    for (int k = 0; k < ch; k++) {
      sm.flag[k][j] = 0;
      if (nz & (1u << k)) {
        /* flag_lossless: thresholds ramp across the partition holding the point limit; the ramp is a
           running float sum, replayed per lane in bin order */
        const int left = limit - i;
        float p1, p2, s1 = 0.f, s2 = 0.f; bool ramp = false;
        if (left > 0) {
          p1 = pre; p2 = pre_x;
          if (left - jn <= 0) { s1 = __fdiv_rn(__fsub_rn(post, pre), (float)jn); s2 = __fdiv_rn(__fsub_rn(post_x, pre_x), (float)jn); ramp = true; }
        } else { p1 = post; p2 = post_x; }
        if (ramp) for (int t = 0; t <= j && t < jn; t++) { p1 = __fadd_rn(p1, s1); p2 = __fadd_rn(p2, s2); }
        if (j < jn) {
          const float m = stg.md[buf][k][j];
          const float fl = S->fromdB[stg.iw[buf][k][j]];
          const float res = __fdiv_rn(m, fl);
          sm.res[k][j] = res;
          const float r = fabsf(res);
          float t = __fsub_rn(p1, stg.ep[buf][k][j]);
          if (t < pre) t = pre;
          sm.flag[k][j] = (r < t) ? ((r < p2) ? 0 : -1) : 1;
          float e = __fmul_rn(m, m);
          sm.quant[k][j] = e;
          sm.raw[k][j] = (m < 0.f) ? __fmul_rn(e, -1.f) : e;
          sm.floor[k][j] = __fmul_rn(fl, fl);
        }
        __syncwarp();
        normalize(sm, P, limit, k, false, s_np[k], i, jn, lane);
      } else if (j < jn) {
        sm.floor[k][j] = 1e-10f; sm.raw[k][j] = 0.f; sm.quant[k][j] = 0.f; sm.res[k][j] = 0.f; sm.out[k][j] = 0;
      }
    }
    __syncwarp();

    for (int step = 0; step < steps; step++) {
      const int Mi = M.mag[step], Ai = M.ang[step];
      if (!((nz >> Mi) & 1u) && !((nz >> Ai) & 1u)) continue;
      nz |= (1u << Mi) | (1u << Ai);
      const bool in = j < jn;
      if (P.tonefix_end > i) {                                 /* aoTuV M6, psy.c:5006-5048 */
        const bool act = in && (big(sm.res[Mi][j]) || big(sm.res[Ai][j]));
        const bool opp = in && opposite(sm.raw[Mi][j], sm.raw[Ai][j]);
        const unsigned am = __ballot_sync(FULLMASK, act);
        const int rp = __popc(__ballot_sync(FULLMASK, act && opp));
        const int ap = __popc(am);
        const float d = act ? fabsf(__fsub_rn(fabsf(sm.res[Mi][j]), fabsf(sm.res[Ai][j]))) : 0.f;
        /* sum of d over the active bins, in bin order (float: the order is part of the result).  d is +0 on the other
           lanes and the running sum is never negative, so with many active bins the sum simply runs over all 32
           lanes out of shared memory (8 vector loads + 32 adds) instead of one ffs/shuffle/add round per active bin */
        float def = 0.f;
        if (ap > 8) {
          sm.dsum[j] = d;
          __syncwarp();
          const float4 *q4 = reinterpret_cast<const float4 *>(sm.dsum);
#pragma unroll
          for (int t = 0; t < 8; t++) {
            const float4 v = q4[t];
            def = __fadd_rn(def, v.x); def = __fadd_rn(def, v.y); def = __fadd_rn(def, v.z); def = __fadd_rn(def, v.w);
          }
          __syncwarp();
        } else {
          for (unsigned m = am; m; m &= m - 1) def = __fadd_rn(def, __shfl_sync(FULLMASK, d, __ffs(m) - 1));
        }
        if (ap != 0) {
          const float cur = __fdiv_rn(def, (float)ap);
          def = cur;
          if (resdef[step] > 0.f) def = (float)((double)cur * 0.5 + (double)resdef[step] * 0.5);
          resdef[step] = cur;
          if (in) {
            const bool soft = (sm.flag[Mi][j] == -1 || sm.flag[Ai][j] == -1);
            if (def > 1.f && soft) sm.flag[Mi][j] = 1;
            if (__fdiv_rn((float)rp, (float)ap) >= prae && soft && opp) sm.flag[Mi][j] = 1;
          }
        } else resdef[step] = -1.f;
      }
      bool lossy = false;
      if (in) {
        if (j < sliding_lowpass - i) {
          if (sm.flag[Mi][j] == 1 || sm.flag[Ai][j] == 1) {      /* lossless coupling */
            sm.raw[Mi][j] = __fadd_rn(fabsf(sm.raw[Mi][j]), fabsf(sm.raw[Ai][j]));
            sm.quant[Mi][j] = __fadd_rn(sm.quant[Mi][j], sm.quant[Ai][j]);
            sm.flag[Mi][j] = 1; sm.flag[Ai][j] = 1;
            {
              const float fa = sm.res[Mi][j], fb = sm.res[Ai][j];
              float ra, rm = fa;
              if (fabsf(fa) > fabsf(fb)) ra = (fa > 0.f) ? __fsub_rn(fa, fb) : __fsub_rn(fb, fa);
              else { ra = (fb > 0.f) ? __fsub_rn(fa, fb) : __fsub_rn(fb, fa); rm = fb; }
              if ((double)ra >= (double)fabsf(rm) * 2) { ra = -ra; rm = -rm; }
              sm.res[Ai][j] = ra; sm.res[Mi][j] = rm;
            }
            {
              const int a = sm.out[Mi][j], b = sm.out[Ai][j];
              int ia, im = a;
              if (abs(a) > abs(b)) ia = (a > 0) ? a - b : b - a;
              else { ia = (b > 0) ? a - b : b - a; im = b; }
              if (ia >= abs(im) * 2) { ia = -ia; im = -im; }
              sm.out[Ai][j] = ia; sm.out[Mi][j] = im;
            }
          } else {                                                /* lossy point coupling */
            const float hL = .18f, hH = (steps == 1 || step == 3) ? .12f : .04f;
            const float r = dipole(sm.raw[Mi][j], sm.raw[Ai][j], (j < limit - i) ? hL : hH);
            sm.raw[Mi][j] = r;
            sm.quant[Mi][j] = fabsf(r);
            sm.raw[Ai][j] = 0.f; sm.quant[Ai][j] = 0.f;
            sm.flag[Ai][j] = 1;
            sm.out[Ai][j] = 0;
            sm.res[Ai][j] = 0.f;
            lossy = true;
          }
        }
        const float f2 = __fadd_rn(sm.floor[Mi][j], sm.floor[Ai][j]);
        sm.floor[Mi][j] = f2; sm.floor[Ai][j] = f2;
      }
      const bool point = __any_sync(FULLMASK, lossy);
      if (point) {
        if (lane == 0) {
          if (s_np[Mi] < -0.5f || s_np[Ai] < -0.5f) s_np[Mi] = -1.f;
          else s_np[Mi] = fminf(s_np[Mi], s_np[Ai]);
        }
        __syncwarp();
        normalize(sm, P, limit, Mi, true, s_np[Mi], i, jn, lane);
      }
      __syncwarp();
    }
    /* results of this partition */
    if (j < jn) for (int k = 0; k < ch; k++) iw[(size_t)k * n + i + j] = sm.out[k][j];
    if (lane < ch) npk[lane * npeak_stride + pi] = s_np[lane];
    __syncwarp();
  }
  if (lowpassr < n)
    for (int k = 0; k < ch; k++) for (int x = lowpassr + lane; x < n; x += 32) iw[(size_t)k * n + x] = 0;
  if (lane == 0) {
    /* psy.c:5133-5141: the reference updates nonzero[] in place, so earlier steps feed later ones */
    unsigned cur = nzmask0;
    for (int step = 0; step < steps; step++)
      if (((cur >> M.mag[step]) & 1u) || ((cur >> M.ang[step]) & 1u)) cur |= (1u << M.mag[step]) | (1u << M.ang[step]);
    for (int k = 0; k < ch; k++) nzg[k] = (cur >> k) & 1u;
  }
}

}  // namespace

int vb200_launch_couple(const vb200_setup *s, cudaStream_t st, int W, int blocks, const int *psy, int blobno,
                        const float *mdct, const float *epeak, float *npeak, int npeak_stride, int *iwork, int *nonzero) {
  if (blocks <= 0) return VB200_OK;
  if (blobno < 0 || blobno >= VB200_PACKETBLOBS) { vb200_set_error("blob number out of range"); return VB200_EINVAL; }
  const int n = s->h.bs[W] / 2;
  const int steps = s->h.map[s->h.mode_mapping[W]].coupling_steps;
#define KC_LAUNCH(CAP_, C_, S_) k_couple<CAP_, C_, S_><<<blocks, 32, 0, st>>>(s->d, W, psy, blobno, n, s->h.channels, s->h.mode_mapping[W], mdct, epeak, npeak, npeak_stride, iwork, nonzero)
  if (s->h.channels == 2 && steps == 1) KC_LAUNCH(2, 2, 1);
  else if (s->h.channels <= 2) KC_LAUNCH(2, 0, -1);
  else if (s->h.channels == 6 && steps == 0) KC_LAUNCH(VB200_MAX_CHANNELS, 6, 0);
  else if (s->h.channels == 6 && steps == 4) KC_LAUNCH(VB200_MAX_CHANNELS, 6, 4);
  else KC_LAUNCH(VB200_MAX_CHANNELS, 0, -1);
#undef KC_LAUNCH
  LAUNCH_CHECK();
  return VB200_OK;
}
