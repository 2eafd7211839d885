/* K3 -- tone masking, one warp per ch-block.
 *
 * Replaces _vp_tonemask (psy.c:4076-4142): ATH floor, seed_loop/seed_curve (:719,:652) splatting
 * the tone curves into the octave-spaced seed vector, seed_chase (:773) and max_seeds (:936).
 *
 * B200 mapping.  The curve splat is a max-scatter (order independent -> shared-memory atomic
 * max, bit-exact); the chase is a data-dependent stack walk over the ~780 seed lines that is
 * NOT a plain sliding maximum (a later, higher seed does not take over until the earlier
 * entry's span ends), so it stays serial: one lane walks it.  To keep the SM busy the CTA is a
 * single warp with ~8 KB of shared memory, so ~28 ch-blocks are resident per SM and their
 * serial walks overlap.  The final min-over-octave-span -> bin floor step runs over the
 * statically known groups (setup.cu) in parallel.  HBM traffic: logfft in, tone out.
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

#define TM_NEGINF (-9999.f)
constexpr int TM_MAXL = 1024;

struct ToneSmem {
  float seed[TM_MAXL];
  float amp[TM_MAXL];
  unsigned short pos[TM_MAXL];
};

__device__ __forceinline__ void smem_atomic_max(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

template <int N>
__global__ void __launch_bounds__(32)
k_tonemask(const DevSetup *__restrict__ S, const int *__restrict__ psyv, const float *__restrict__ logfft,
           const float *__restrict__ global_ampmax, const float *__restrict__ local_ampmax, float *__restrict__ tone) {
  __shared__ ToneSmem sm;
  const int item = blockIdx.x, lane = threadIdx.x;
  const DevPsy &P = S->psy[psyv[item]];
  const float *f = logfft + (size_t)item * N;
  const int L = P.total_octave_lines, per = P.eighth;
  float att = __fadd_rn(local_ampmax[item], P.ath_adjatt);
  if (att < P.ath_maxatt) att = P.ath_maxatt;
  const float dBoffset = __fsub_rn(P.max_curve_dB, global_ampmax[item]);

  for (int i = lane; i < L; i += 32) sm.seed[i] = TM_NEGINF;
  __syncwarp();

  /* seed_loop: the last bin of every run of bins that share an octave line owns the run */
  for (int i = lane; i < N; i += 32) {
    const int oc = __ldg(P.octave + i);
    if (i + 1 < N && __ldg(P.octave + i + 1) == oc) continue;
    float mx = __ldg(f + i);
    for (int j = i - 1; j >= 0 && __ldg(P.octave + j) == oc; j--) { float v = __ldg(f + j); if (v > mx) mx = v; }
    const float flr = __fadd_rn(__ldg(P.ath + i), att);
    if (__fadd_rn(mx, 6.f) > flr) {
      int band = oc >> P.shiftoc;
      if (band >= VB200_P_BANDS) band = VB200_P_BANDS - 1;
      if (band < 0) band = 0;
      int choice = (int)(((double)__fadd_rn(mx, dBoffset) - 30.) * (double).1f);
      if (choice < 0) choice = 0;
      if (choice > VB200_P_LEVELS - 1) choice = VB200_P_LEVELS - 1;
      const float *posts = P.tonecurves + (band * VB200_P_LEVELS + choice) * VB200_EHMER;
      const float p0 = __ldg(posts);
      const int k0 = (int)p0, k1 = (int)__ldg(posts + 1);
      int sp = (int)((float)(oc - P.firstoc) + (p0 - 16.f) * (float)per - (float)(per >> 1));
      for (int k = k0; k < k1; k++) {
        if (sp > 0) smem_atomic_max(&sm.seed[sp], __fadd_rn(mx, __ldg(posts + 2 + k)));
        sp += per;
        if (sp >= L) break;
      }
    }
  }
  __syncwarp();

  /* seed_chase: serial stack walk + left-to-right fill, lane 0 */
  if (lane == 0) {
    int top = 0;
    float a1 = 0.f;            /* amplitude on top of the stack */
    for (int i = 0; i < L; i++) {
      const float s = sm.seed[i];
      if (top >= 2) {
        while (!(s < a1)) {
          if (top > 1 && i < (int)sm.pos[top - 1] + per) {
            const float a2 = sm.amp[top - 2];
            if (a1 <= a2 && i < (int)sm.pos[top - 2] + per) { top--; a1 = a2; continue; }
          }
          break;
        }
      }
      sm.pos[top] = (unsigned short)i; sm.amp[top] = s; a1 = s; top++;
    }
    int w = 0;
    for (int i = 0; i < top; i++) {
      const float a = sm.amp[i];
      int end = (i < top - 1 && sm.amp[i + 1] > a) ? (int)sm.pos[i + 1] : (int)sm.pos[i] + per + 1;
      if (end > L) end = L;
      for (; w < end; w++) sm.seed[w] = a;
    }
  }
  __syncwarp();

  /* max_seeds: per statically known group, min over the octave span, then raise the bin floors */
  float *out = tone + (size_t)item * N;
  const float lim = P.tone_abs_limit;
  for (int g = lane; g < P.n_tone_groups; g += 32) {
    const int4 q = __ldg(P.tone_groups + g);
    float mv = sm.seed[q.x];
    if (mv > lim) mv = lim;
    for (int p = q.x + 1; p <= q.y; p++) { const float v = sm.seed[p]; if ((v > TM_NEGINF && v < mv) || mv == TM_NEGINF) mv = v; }
    for (int i = q.z; i < q.w; i++) { float fl = __fadd_rn(__ldg(P.ath + i), att); out[i] = fl < mv ? mv : fl; }
  }
  {
    const float mv = sm.seed[L - 1];
    for (int i = P.tone_tail + lane; i < N; i += 32) { float fl = __fadd_rn(__ldg(P.ath + i), att); out[i] = fl < mv ? mv : fl; }
  }
}

}  // namespace

int vb200_launch_tonemask(const vb200_setup *s, cudaStream_t st, int W, int items, const int *psy,
                          const float *logfft, const float *global_ampmax, const float *local_ampmax, float *tone) {
  if (items <= 0) return VB200_OK;
  if (W) k_tonemask<1024><<<items, 32, 0, st>>>(s->d, psy, logfft, global_ampmax, local_ampmax, tone);
  else   k_tonemask<128><<<items, 32, 0, st>>>(s->d, psy, logfft, global_ampmax, local_ampmax, tone);
  LAUNCH_CHECK();
  return VB200_OK;
}
