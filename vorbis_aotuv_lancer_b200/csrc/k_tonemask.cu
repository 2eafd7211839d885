/* K3 -- tone masking.
 *
 * Replaces _vp_tonemask (psy.c:4076-4142): ATH floor, seed_loop/seed_curve (:719,:652) splatting
 * the tone curves into the octave-spaced seed vector, seed_chase (:773) and max_seeds (:936).
 *
 * B200 mapping.  The curve splat is a max-scatter (order independent -> shared-memory atomic
 * max, bit-exact).  The chase is a data-dependent stack walk over the ~780 seed lines that is NOT
 * a plain sliding maximum (a later, higher seed does not take over until the earlier entry's span
 * ends), so it is serial per ch-block.  Round-1 first version gave every ch-block one warp and let
 * lane 0 walk the chase: ncu showed 4 active lanes per instruction and an issue-bound SM
 * (profiles/r01_ncu_tonemask_full.md).  This version keeps 16 ch-blocks per CTA in shared memory:
 * the splat and the final min-over-span step run warp-cooperatively per ch-block, while the chase
 * runs lane-per-ch-block (16 independent stack walks advance together in one half-warp, their seed
 * vectors laid out with an odd stride so that equal positions fall into different banks).
 * HBM traffic: logfft in, tone out.
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

#define TM_NEGINF (-9999.f)
constexpr int TM_ITEMS = 8;      /* ch-blocks per CTA: one warp each for the cooperative phases, 8 lanes for the chase */
constexpr int TM_THREADS = 256;
constexpr int TM_WARPS = TM_THREADS / 32;
constexpr int TM_QCAP = 64;      /* per-warp queue of curves to splat */

__device__ __forceinline__ void smem_atomic_max(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

struct SplatJob { float mx; int sp0; int curve; int k0, k1; };   /* amplitude, first seed position, curve offset in floats */

template <int N>
__global__ void __launch_bounds__(TM_THREADS)
k_tonemask(const DevSetup *__restrict__ S, int items, const int *__restrict__ psyv, const float *__restrict__ logfft,
           const float *__restrict__ global_ampmax, const float *__restrict__ local_ampmax, float *__restrict__ tone, int Lstride) {
  extern __shared__ __align__(16) unsigned char dyn[];
  float *seed_all = reinterpret_cast<float *>(dyn);                                  /* [TM_ITEMS][Lstride] */
  unsigned short *pos_all = reinterpret_cast<unsigned short *>(seed_all + TM_ITEMS * Lstride);   /* [TM_ITEMS][Lstride] */
  __shared__ SplatJob queue[TM_WARPS][TM_QCAP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int item0 = blockIdx.x * TM_ITEMS;
  const int nloc = min(TM_ITEMS, items - item0);

  for (int i = tid; i < TM_ITEMS * Lstride; i += TM_THREADS) seed_all[i] = TM_NEGINF;
  __syncthreads();

  /* ---- seed_loop: warps take ch-blocks in turn; lanes take the runs of bins that share an octave line ---- */
  for (int li = warp; li < nloc; li += TM_WARPS) {
    const int item = item0 + li;
    const DevPsy &P = S->psy[psyv[item]];
    const float *f = logfft + (size_t)item * N;
    float *seed = seed_all + li * Lstride;
    const int L = P.total_octave_lines, per = P.eighth;
    float att = __fadd_rn(local_ampmax[item], P.ath_adjatt);
    if (att < P.ath_maxatt) att = P.ath_maxatt;
    const float dBoffset = __fsub_rn(P.max_curve_dB, global_ampmax[item]);
    int nq = 0;                                          /* jobs waiting in this warp's queue */
    const int nruns = P.n_tone_runs;
    for (int r0 = 0; r0 < nruns || nq > 0; r0 += 32) {
      const int r = r0 + lane;
      bool hit = false;
      SplatJob job; job.mx = 0.f; job.sp0 = 0; job.curve = 0; job.k0 = 0; job.k1 = 0;
      if (r < nruns) {
        const int2 run = __ldg(P.tone_runs + r);          /* first and last bin of the run */
        float mx = __ldg(f + run.x);
        for (int j = run.x + 1; j <= run.y; j++) { const float v = __ldg(f + j); if (v > mx) mx = v; }
        const float flr = __fadd_rn(__ldg(P.ath + run.y), att);
        if (__fadd_rn(mx, 6.f) > flr) {
          const int oc = __ldg(P.octave + run.y);
          int band = oc >> P.shiftoc;
          if (band >= VB200_P_BANDS) band = VB200_P_BANDS - 1;
          if (band < 0) band = 0;
          int choice = (int)(((double)__fadd_rn(mx, dBoffset) - 30.) * (double).1f);
          if (choice < 0) choice = 0;
          if (choice > VB200_P_LEVELS - 1) choice = VB200_P_LEVELS - 1;
          job.curve = (band * VB200_P_LEVELS + choice) * VB200_EHMER;
          const float p0 = __ldg(P.tonecurves + job.curve);
          job.k0 = (int)p0; job.k1 = (int)__ldg(P.tonecurves + job.curve + 1);
          job.sp0 = (int)((float)(oc - P.firstoc) + (p0 - 16.f) * (float)per - (float)(per >> 1));
          job.mx = mx;
          hit = true;
        }
      }
      /* append this round's hits to the warp's queue; when 32 jobs wait (or the runs are exhausted) every
         lane splats one whole curve -- a max-scatter, so the order of the updates does not matter */
      const unsigned hm = __ballot_sync(FULLMASK, hit);
      if (hit) queue[warp][nq + __popc(hm & ((1u << lane) - 1u))] = job;
      nq += __popc(hm);
      __syncwarp();
      if (nq >= 32 || r0 + 32 >= nruns) {
        const int take = nq < 32 ? nq : 32;
        if (lane < take) {
          const SplatJob jb = queue[warp][lane];
          const float *curve = P.tonecurves + jb.curve + 2;
          int sp = jb.sp0;
          for (int k = jb.k0; k < jb.k1; k++) {
            if (sp >= L) break;
            if (sp > 0) smem_atomic_max(&seed[sp], __fadd_rn(jb.mx, __ldg(curve + k)));
            sp += per;
          }
        }
        __syncwarp();
        /* move the leftover (< 32 jobs) to the front */
        SplatJob keep; const bool mv = lane + take < nq;
        if (mv) keep = queue[warp][lane + take];
        __syncwarp();
        if (mv) queue[warp][lane] = keep;
        nq -= take;
        __syncwarp();
      }
    }
  }
  __syncthreads();

  /* ---- seed_chase: one lane per ch-block, amplitudes of the stack kept in place over the seeds ---- */
  if (warp == 0 && lane < nloc) {
    const int item = item0 + lane;
    const DevPsy &P = S->psy[psyv[item]];
    const int L = P.total_octave_lines, per = P.eighth;
    float *seed = seed_all + lane * Lstride;           /* seed[i] for i >= top: input; seed[i] for i < top: stack amplitude */
    unsigned short *pos = pos_all + lane * Lstride;
    /* The two entries on top of the stack live in registers (a1,p1 / a2,p2); shared memory is only read
       when a pop has to uncover a third one.  The seed two steps ahead is fetched early. */
    int top = 0, p1 = 0, p2 = 0;
    float a1 = 0.f, a2 = 0.f;
    float s_next = seed[0];
    for (int i = 0; i < L; i++) {
      const float s = s_next;
      s_next = (i + 1 < L) ? seed[i + 1] : 0.f;
      if (top >= 2) {
        while (!(s < a1)) {
          if (i < p1 + per && a1 <= a2 && i < p2 + per) {
            /* pop: the entry below becomes the top */
            top--; a1 = a2; p1 = p2;
            if (top >= 2) { a2 = seed[top - 2]; p2 = pos[top - 2]; }
            if (top < 2) break;
            continue;
          }
          break;
        }
      }
      pos[top] = (unsigned short)i; seed[top] = s;        /* top <= i: never clobbers unread input */
      a2 = a1; p2 = p1; a1 = s; p1 = i; top++;
    }
    /* turn (pos, amp) into the end of every entry's span: running maximum, clipped to L.  pos[i] becomes the
       first position NOT covered by entries 0..i, which is all the fill needs. */
    int run = 0;
    for (int i = 0; i < top; i++) {
      const int end = (i < top - 1 && seed[i + 1] > seed[i]) ? (int)pos[i + 1] : (int)pos[i] + per + 1;
      /* pos[i+1] is still the raw position here: entry i+1 is rewritten only in the next iteration */
      const int e2 = end > L ? L : end;
      if (e2 > run) run = e2;
      pos[i] = (unsigned short)run;
    }
    pos[L - 1 < Lstride - 1 ? Lstride - 1 : L - 1] = (unsigned short)top;   /* stash the stack height in the last slot */
  }
  __syncthreads();

  /* ---- fill + max_seeds: warps take ch-blocks in turn ---- */
  for (int li = warp; li < nloc; li += TM_WARPS) {
    const int item = item0 + li;
    const DevPsy &P = S->psy[psyv[item]];
    const int L = P.total_octave_lines;
    float *seed = seed_all + li * Lstride;
    unsigned short *pos = pos_all + li * Lstride;
    const int top = pos[Lstride - 1];
    /* entry i covers [pos[i-1], pos[i]); amplitudes sit in seed[0..top).  Expand from the back so that an
       amplitude is read before its slot is overwritten: positions written by entry i are >= pos[i-1] > i-1. */
    for (int i0 = ((top - 1) / 32) * 32; i0 >= 0; i0 -= 32) {
      const int i = i0 + lane;
      float a = 0.f; int lo = 0, hi = 0;
      if (i < top) { a = seed[i]; hi = pos[i]; lo = i > 0 ? (int)pos[i - 1] : 0; }
      __syncwarp();
      if (i < top) for (int w = lo; w < hi; w++) seed[w] = a;
      __syncwarp();
    }
    float att = __fadd_rn(local_ampmax[item], P.ath_adjatt);
    if (att < P.ath_maxatt) att = P.ath_maxatt;
    float *out = tone + (size_t)item * N;
    const float lim = P.tone_abs_limit;
    for (int g = lane; g < P.n_tone_groups; g += 32) {
      const int4 q = __ldg(P.tone_groups + g);
      float mv = seed[q.x];
      if (mv > lim) mv = lim;
      for (int p = q.x + 1; p <= q.y; p++) { const float v = seed[p]; if ((v > TM_NEGINF && v < mv) || mv == TM_NEGINF) mv = v; }
      for (int i = q.z; i < q.w; i++) { const float fl = __fadd_rn(__ldg(P.ath + i), att); out[i] = fl < mv ? mv : fl; }
    }
    {
      const float mv = seed[L - 1];
      for (int i = P.tone_tail + lane; i < N; i += 32) { const float fl = __fadd_rn(__ldg(P.ath + i), att); out[i] = fl < mv ? mv : fl; }
    }
  }
}

}  // namespace

int vb200_launch_tonemask(const vb200_setup *s, cudaStream_t st, int W, int items, const int *psy,
                          const float *logfft, const float *global_ampmax, const float *local_ampmax, float *tone) {
  if (items <= 0) return VB200_OK;
  int L = 0;
  for (int k = 0; k < s->h.psys; k++) if (s->h.psy[k].blockflag == W && s->h.psy[k].total_octave_lines > L) L = s->h.psy[k].total_octave_lines;
  const int Lstride = (L + 2) | 1;        /* odd: lanes walking equal positions hit different banks; +1 slot for the stack height */
  const size_t dyn = (size_t)TM_ITEMS * Lstride * (sizeof(float) + sizeof(unsigned short));
  const int grid = (items + TM_ITEMS - 1) / TM_ITEMS;
  if (W) {
    CUDA_TRY(cudaFuncSetAttribute(k_tonemask<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_tonemask<1024><<<grid, TM_THREADS, dyn, st>>>(s->d, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_tonemask<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_tonemask<128><<<grid, TM_THREADS, dyn, st>>>(s->d, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
  }
  LAUNCH_CHECK();
  return VB200_OK;
}
