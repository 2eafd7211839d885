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
 * (profiles/r01_ncu_tonemask_full.md).  Now the chase runs on bit masks (see tm_chase_step) and needs one
 * lane per ch-block and ~36 instructions per position, so it wants 32 ch-blocks per warp, while splat and fill want
 * a warp per ch-block.  Long blocks therefore take three kernels: k_tone_splat (warp per ch-block: splat + the
 * 14 comparison bits per position) -> k_tone_chase (lane per ch-block, comparison words transposed through shared
 * memory 32 positions at a time) -> k_tone_fill (warp per ch-block: spans, fill, min-over-span, output).  The seeds
 * (3 KB), comparison words (1.5 KB) and survivor bits (100 B) of a ch-block travel through rows of two per-batch
 * buffers that are dead at this point of the pipeline (logmask and iwork rows), +9.4 KB of HBM traffic per
 * ch-block.  Short blocks (rows too small for that) keep the fused kernel k_tonemask.
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

#define TM_NEGINF (-9999.f)
constexpr int TM_SVW = 33;       /* words of the survivor set of one ch-block (<= 1024 octave lines), odd stride */
constexpr int TM_QCAP = 64;      /* per-warp queue of curves to splat */
constexpr int TM_PER = 8;        /* eighth_octave_lines: every setup has 8 (vb200_setup_create refuses anything else), so the
                                    seed-position arithmetic divides and strides by a constant */

__device__ __forceinline__ void smem_atomic_max(float *addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

struct SplatJob { float mx; int sp0; int curve; int k0, k1; };   /* amplitude, first seed position, curve offset in floats */

/* seed_loop (psy.c:719-771) of one ch-block by one warp: lanes take the runs of bins that share an octave line,
   hits are queued, and when 32 wait every lane splats one whole curve (max-scatter: order does not matter) */
template <int N>
__device__ __forceinline__ void tm_splat(const DevSetup *__restrict__ S, const int *__restrict__ psyv, const float *__restrict__ logfft,
                                         const float *__restrict__ global_ampmax, const float *__restrict__ local_ampmax,
                                         int item, float *seed, SplatJob (*queue)[TM_QCAP], int warp, int lane) {
    const DevPsy &P = S->psy[psyv[item]];
    const float *f = logfft + (size_t)item * N;
    const int L = P.total_octave_lines; constexpr int per = TM_PER;
    float att = __fadd_rn(local_ampmax[item], P.ath_adjatt);
    if (att < P.ath_maxatt) att = P.ath_maxatt;
    const float dBoffset = __fsub_rn(P.max_curve_dB, global_ampmax[item]);
    int nq = 0;                                          /* jobs waiting in this warp's queue */
    const int nruns = P.n_tone_runs;
    for (int r0 = 0; r0 < nruns || nq > 0; r0 += 32) {
      const int r = r0 + lane;
      bool hit = false;
      SplatJob job; job.mx = 0.f; job.sp0 = 0; job.curve = 0; job.k0 = 0; job.k1 = 0;
      /* maximum of the run: runs are 1 bin long at low frequencies and up to ~11 at the top, so the warp walks as far
         as its longest run, four bins per lane and iteration with the four loads in flight together (a lane-private
         loop waited for every load in turn: 40 % of the kernel's stall samples were on this scoreboard) */
      int2 run = make_int2(0, -1);
      if (r < nruns) run = __ldg(P.tone_runs + r);        /* first and last bin of the run */
      float mx = r < nruns ? __ldg(f + run.x) : 0.f;
      {
        const int extra = __reduce_max_sync(FULLMASK, run.y - run.x);
        for (int t = 1; t <= extra; t += 4) {
          const int j = run.x + t;
          const float v0 = j <= run.y ? __ldg(f + j) : -INFINITY, v1 = j + 1 <= run.y ? __ldg(f + j + 1) : -INFINITY;
          const float v2 = j + 2 <= run.y ? __ldg(f + j + 2) : -INFINITY, v3 = j + 3 <= run.y ? __ldg(f + j + 3) : -INFINITY;
          if (v0 > mx) mx = v0;
          if (v1 > mx) mx = v1;
          if (v2 > mx) mx = v2;
          if (v3 > mx) mx = v3;
        }
      }
      if (r < nruns) {
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
          /* the reference walks k = k0..k1-1 with sp = sp0 + (k-k0)*per, skips sp <= 0 and stops at sp >= L
             (psy.c:672-706): clip the k range once instead of testing every step */
          int ka = jb.k0, kb = jb.k1;
          if (jb.sp0 <= 0) ka += (-jb.sp0) / per + 1;                      /* first k with sp > 0 */
          { const int room = (L - 1 - jb.sp0) / per; if (jb.sp0 >= L) kb = ka; else if (jb.k0 + room + 1 < kb) kb = jb.k0 + room + 1; }
          float *sd = seed + jb.sp0 + (ka - jb.k0) * per;
          for (int k = ka; k < kb; k++, sd += per) smem_atomic_max(sd, __fadd_rn(jb.mx, __ldg(curve + k)));
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

/* the 7+7 comparison bits of every position: bit k-1 = !(seed[p] < seed[p-k]), bit 8+d-1 = seed[p] <= seed[p-d] */
__device__ __forceinline__ unsigned tm_compare_word(const float *seed, int p) {
  const float s = seed[p];
  unsigned ge = 0, le = 0;
#pragma unroll
  for (int k = 1; k < 8; k++) {
    if (p - k >= 0) { const float o = seed[p - k]; if (!(s < o)) ge |= 1u << (k - 1); if (s <= o) le |= 1u << (k - 1); }
  }
  return ge | (le << 8);
}

/* One step of seed_chase (psy.c:773-830) on bit masks.  Every position is pushed once and can only be popped
   while it is "young" (less than 8 positions old), and a pop needs the entry below to be young too.  So the part of
   the stack that can still change is a 7-bit set of the last 7 positions (bit k-1 of M: position i-k is on the
   stack), an entry's amplitude is the seed at its position, and every comparison a pop makes is between two seeds
   less than 8 apart -- all precomputed in c.  Q bit k-1: that entry's amplitude is <= the amplitude of the entry it
   was pushed onto (which entry that is never changes while it stays on the stack).  The entries popped by seed[i]
   are the youngest ones up to the first entry that is not poppable -- seed[i] below it (GE clear), amplitude above
   its predecessor (Q clear), or no young predecessor left (the oldest young entry).  Returns the mask after the pops
   (its bit 6 is the final fate of position i-7) and advances M, Q. */
__device__ __forceinline__ unsigned tm_chase_step(unsigned c, unsigned &M, unsigned &Q) {
      const unsigned ge = c & 0x7fu, le = c >> 8;
      const unsigned oldest = M ? (0x80000000u >> __clz(M)) : 0u;
      const unsigned stop = (M & ~(ge & Q)) | oldest;
      const unsigned low = stop & (0u - stop);                     /* first entry that stays */
      const unsigned m = M & ~(low - 1u);                          /* M == 0: low == 0, m == 0 */
      const unsigned qbit = (le & low) ? 1u : 0u;                  /* new entry vs the entry it lands on, if young */
      M = ((m << 1) | 1u) & 0x7fu;
      Q = ((Q << 1) | qbit) & 0x7fu;
      return m;
}

/* spans of the survivors (running maximum of their ends = the fill pointer of the reference), fill, max_seeds
   (psy.c:936-992) and the ATH floor: one warp, one ch-block.  seed[] holds the raw seeds on entry. */
template <int N>
__device__ __forceinline__ void tm_fill(const DevSetup *__restrict__ S, const DevPsy &P, int item, float *seed, unsigned short *run16,
                                        const unsigned *sv, const float *__restrict__ local_ampmax, float *__restrict__ tone, int lane) {
    const int L = P.total_octave_lines; constexpr int per = TM_PER;
    const int nw = (L + 31) >> 5;
    const unsigned myw = lane < nw ? sv[lane] : 0u;
    /* first survivor after the word this lane holds */
    const unsigned nzm = __ballot_sync(FULLMASK, myw != 0u);
    int after = -1;
    {
      const unsigned hi = lane < 31 ? (nzm & ~((2u << lane) - 1u)) : 0u;
      const int w2 = hi ? __ffs(hi) - 1 : 0;
      const unsigned ww = __shfl_sync(FULLMASK, myw, w2);
      if (hi) after = w2 * 32 + __ffs(ww) - 1;
    }
    /* End of every survivor's span, running maximum in position order (the fill pointer of the reference).  Survivor p
       covers [end of the survivor before it, its own end); the spans tile [0, L).  Instead of filling the seed vector
       span by span (a serial loop per survivor), every survivor with a non-empty span leaves its position at the
       span's first slot, and a second sweep gives every position the last mark at or before it: run16[x] = the survivor
       whose amplitude the reference would have filled in at x.  The seeds themselves are never overwritten. */
    for (int x = lane; x < L; x += 32) run16[x] = 0xffff;
    __syncwarp();
    int carry = 0;
    for (int w = 0; w < nw; w++) {
      const unsigned word = __shfl_sync(FULLMASK, myw, w);
      const int aft = __shfl_sync(FULLMASK, after, w);
      const int p = w * 32 + lane;
      const bool sv_here = (word >> lane) & 1u;
      int end = 0;
      if (sv_here) {
        const unsigned hb = lane < 31 ? (word & ~((2u << lane) - 1u)) : 0u;
        const int nx = hb ? w * 32 + __ffs(hb) - 1 : aft;
        end = (nx >= 0 && seed[nx] > seed[p]) ? nx : p + per + 1;
        if (end > L) end = L;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULLMASK, end, o); if (lane >= o && y > end) end = y; }
      if (carry > end) end = carry;
      int lo = __shfl_up_sync(FULLMASK, end, 1);            /* running maximum before this position */
      if (lane == 0) lo = carry;
      if (sv_here && end > lo) run16[lo] = (unsigned short)p;       /* spans are disjoint: one writer per slot */
      carry = __shfl_sync(FULLMASK, end, 31);
    }
    __syncwarp();
    {
      int last = 0xffff;                                     /* slot 0 is always marked (the first survivor's span starts at 0) */
      for (int x0 = 0; x0 < L; x0 += 32) {
        const int x = x0 + lane;
        const int m = x < L ? (int)run16[x] : 0xffff;
        const unsigned has = __ballot_sync(FULLMASK, m != 0xffff);
        const unsigned below = has & ((2u << lane) - 1u);    /* lane 31: 2u << 31 == 0, mask == all ones */
        const int src = below ? 31 - __clz(below) : 0;
        const int got = __shfl_sync(FULLMASK, m, src);
        const int v = below ? got : last;
        if (x < L) run16[x] = (unsigned short)v;
        last = __shfl_sync(FULLMASK, v, 31);
      }
    }
    __syncwarp();
    float att = __fadd_rn(local_ampmax[item], P.ath_adjatt);
    if (att < P.ath_maxatt) att = P.ath_maxatt;
    float *out = tone + (size_t)item * N;
    const float lim = P.tone_abs_limit;
    for (int g = lane; g < P.n_tone_groups; g += 32) {
      const int4 q = __ldg(P.tone_groups + g);
      float mv = seed[run16[q.x]];
      if (mv > lim) mv = lim;
      for (int p = q.x + 1; p <= q.y; p++) { const float v = seed[run16[p]]; if ((v > TM_NEGINF && v < mv) || mv == TM_NEGINF) mv = v; }
      for (int i = q.z; i < q.w; i++) { const float fl = __fadd_rn(__ldg(P.ath + i), att); out[i] = fl < mv ? mv : fl; }
    }
    {
      const float mv = seed[run16[L - 1]];
      for (int i = P.tone_tail + lane; i < N; i += 32) { const float fl = __fadd_rn(__ldg(P.ath + i), att); out[i] = fl < mv ? mv : fl; }
    }
}

/* ---------------------------------------------------------------- fused kernel (short blocks) */
/* TM_ITEMS ch-blocks per CTA: one warp each for the cooperative phases, TM_ITEMS lanes for the chase */
template <int N, int TM_ITEMS>
__global__ void __launch_bounds__(TM_ITEMS * 32)
k_tonemask(const DevSetup *__restrict__ S, int items, const int *__restrict__ psyv, const float *__restrict__ logfft,
           const float *__restrict__ global_ampmax, const float *__restrict__ local_ampmax, float *__restrict__ tone, int Lstride) {
  constexpr int TM_THREADS = TM_ITEMS * 32, TM_WARPS = TM_ITEMS;
  extern __shared__ __align__(16) unsigned char dyn[];
  float *seed_all = reinterpret_cast<float *>(dyn);                                  /* [TM_ITEMS][Lstride] */
  unsigned short *pos_all = reinterpret_cast<unsigned short *>(seed_all + TM_ITEMS * Lstride);   /* [TM_ITEMS][Lstride] */
  __shared__ unsigned surv_all[TM_ITEMS * TM_SVW];                                   /* surviving positions, one bit each */
  __shared__ SplatJob queue[TM_WARPS][TM_QCAP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int item0 = blockIdx.x * TM_ITEMS;
  const int nloc = min(TM_ITEMS, items - item0);

  for (int i = tid; i < TM_ITEMS * Lstride; i += TM_THREADS) seed_all[i] = TM_NEGINF;
  __syncthreads();
  for (int li = warp; li < nloc; li += TM_WARPS) {
    float *seed = seed_all + li * Lstride;
    tm_splat<N>(S, psyv, logfft, global_ampmax, local_ampmax, item0 + li, seed, queue, warp, lane);
    __syncwarp();
    const int L = S->psy[psyv[item0 + li]].total_octave_lines;
    unsigned short *cmp = pos_all + li * Lstride;
    for (int p = lane; p < L; p += 32) cmp[p] = (unsigned short)tm_compare_word(seed, p);
    if (lane < Lstride - L) cmp[L + lane] = 0;
  }
  __syncthreads();
  if (warp == 0 && lane < nloc) {
    const int L = S->psy[psyv[item0 + lane]].total_octave_lines;
    const unsigned short *cm = pos_all + lane * Lstride;
    unsigned *sv = surv_all + lane * TM_SVW;
    unsigned M = 0, Q = 0, cur = 0;
    unsigned c_next = cm[0];
#pragma unroll 2
    for (int i = 0; i < L; i++) {
      const unsigned c = c_next;
      c_next = cm[i + 1];
      const unsigned m = tm_chase_step(c, M, Q);
      const int p = i - 7;                                         /* this position's fate is now sealed */
      if (p >= 0) { cur |= ((m >> 6) & 1u) << (p & 31); if ((p & 31) == 31) { sv[p >> 5] = cur; cur = 0; } }
    }
    for (int k = 7; k >= 1; k--) {
      const int p = L - k;
      if (p >= 0) { cur |= ((M >> (k - 1)) & 1u) << (p & 31); if ((p & 31) == 31) { sv[p >> 5] = cur; cur = 0; } }
    }
    if (L & 31) sv[L >> 5] = cur;
  }
  __syncthreads();
  for (int li = warp; li < nloc; li += TM_WARPS)
    tm_fill<N>(S, S->psy[psyv[item0 + li]], item0 + li, seed_all + li * Lstride, pos_all + li * Lstride, surv_all + li * TM_SVW, local_ampmax, tone, lane);
}

/* ---------------------------------------------------------------- three kernels (long blocks) */
constexpr int TS_WARPS = 4;
/* rows of the two scratch buffers: seeds in row `item` of rowA (N floats); comparison words at the start of row
   `item` of rowB (N ints = 2N u16), survivor words in its last TM_SVW ints */
template <int N>
__global__ void __launch_bounds__(TS_WARPS * 32)
k_tone_splat(const DevSetup *__restrict__ S, int items, const int *__restrict__ psyv, const float *__restrict__ logfft,
             const float *__restrict__ global_ampmax, const float *__restrict__ local_ampmax, float *__restrict__ rowA, int *__restrict__ rowB, int Lstride, int rowN) {
  extern __shared__ __align__(16) unsigned char dyn[];
  float *seed_all = reinterpret_cast<float *>(dyn);
  __shared__ SplatJob queue[TS_WARPS][TM_QCAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * TS_WARPS + warp;
  if (item >= items) return;
  float *seed = seed_all + warp * Lstride;
  for (int i = lane; i < Lstride; i += 32) seed[i] = TM_NEGINF;
  __syncwarp();
  tm_splat<N>(S, psyv, logfft, global_ampmax, local_ampmax, item, seed, queue, warp, lane);
  __syncwarp();
  const int L = S->psy[psyv[item]].total_octave_lines;
  float *sg = rowA + (size_t)item * rowN;
  unsigned short *cg = reinterpret_cast<unsigned short *>(rowB + (size_t)item * rowN);
  for (int p = lane; p < L; p += 32) { sg[p] = seed[p]; cg[p] = (unsigned short)tm_compare_word(seed, p); }
}

/* one lane per ch-block.  A lane reads ITS OWN row of comparison words, 32 positions (64 bytes, four 128-bit loads) at a
   time, the next 32 requested before the current ones are walked -- every 32-byte sector a lane touches is used in
   full, and nothing is transposed.  (The first version filled a 32 x 32 shared-memory tile row by row, one 64-byte
   load per row and iteration, each waiting for the one before it: 32 exposed DRAM latencies per 32 positions -- the
   kernel took 0.47 ms for 0.29 ms worth of instructions, and as long again for each small stream group of an e2e call.) */
__global__ void __launch_bounds__(TS_WARPS * 32)
k_tone_chase(int items, int L, int N, int *__restrict__ rowB) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item0 = (blockIdx.x * TS_WARPS + warp) * 32;
  if (item0 >= items) return;
  const bool mine = item0 + lane < items;
  const int item = mine ? item0 + lane : item0;
  const uint4 *cg = reinterpret_cast<const uint4 *>(rowB + (size_t)item * N);       /* 8 comparison words per uint4 */
  unsigned *sv = reinterpret_cast<unsigned *>(rowB + (size_t)item * N + (N - TM_SVW));
  unsigned M = 0, Q = 0, cur = 0;
  uint4 nx[4];
#pragma unroll
  for (int j = 0; j < 4; j++) nx[j] = cg[j];
  for (int p0 = 0; p0 < L; p0 += 32) {
    unsigned w[16];
#pragma unroll
    for (int j = 0; j < 4; j++) { w[4 * j] = nx[j].x; w[4 * j + 1] = nx[j].y; w[4 * j + 2] = nx[j].z; w[4 * j + 3] = nx[j].w; }
    if (p0 + 32 < L) {
#pragma unroll
      for (int j = 0; j < 4; j++) nx[j] = cg[(p0 + 32) / 8 + j];
    }
    const int n = min(32, L - p0);
#pragma unroll
    for (int k = 0; k < 32; k++) {
      if (k < n) {
        const unsigned c = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu);
        const unsigned m = tm_chase_step(c, M, Q);
        const int p = p0 + k - 7;
        if (p >= 0) { cur |= ((m >> 6) & 1u) << (p & 31); if ((p & 31) == 31) { if (mine) sv[p >> 5] = cur; cur = 0; } }
      }
    }
  }
  if (mine) {
    for (int k = 7; k >= 1; k--) {
      const int p = L - k;
      if (p >= 0) { cur |= ((M >> (k - 1)) & 1u) << (p & 31); if ((p & 31) == 31) { sv[p >> 5] = cur; cur = 0; } }
    }
    if (L & 31) sv[L >> 5] = cur;
  }
}

template <int N>
__global__ void __launch_bounds__(TS_WARPS * 32)
k_tone_fill(const DevSetup *__restrict__ S, int items, const int *__restrict__ psyv, const float *__restrict__ local_ampmax,
            const float *__restrict__ rowA, const int *__restrict__ rowB, float *__restrict__ tone, int Lstride, int rowN) {
  extern __shared__ __align__(16) unsigned char dyn[];
  float *seed_all = reinterpret_cast<float *>(dyn);
  unsigned short *run_all = reinterpret_cast<unsigned short *>(seed_all + TS_WARPS * Lstride);
  __shared__ unsigned surv_all[TS_WARPS * TM_SVW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * TS_WARPS + warp;
  if (item >= items) return;
  const DevPsy &P = S->psy[psyv[item]];
  const int L = P.total_octave_lines;
  float *seed = seed_all + warp * Lstride;
  const float *sg = rowA + (size_t)item * rowN;
  for (int p = lane; p < L; p += 32) seed[p] = sg[p];
  if (lane < Lstride - L) seed[L + lane] = TM_NEGINF;
  const unsigned *svg = reinterpret_cast<const unsigned *>(rowB + (size_t)item * rowN + (rowN - TM_SVW));
  unsigned *sv = surv_all + warp * TM_SVW;
  if (lane < ((L + 31) >> 5)) sv[lane] = svg[lane];
  __syncwarp();
  tm_fill<N>(S, P, item, seed, run_all + warp * Lstride, sv, local_ampmax, tone, lane);
}

}  // namespace

template <int N, int TI>
static int launch_tm(const vb200_setup *s, cudaStream_t st, int items, const int *psy, const float *logfft,
                     const float *global_ampmax, const float *local_ampmax, float *tone, int Lstride) {
  const size_t dyn = (size_t)TI * Lstride * (sizeof(float) + sizeof(unsigned short));
  const int grid = (items + TI - 1) / TI;
  CUDA_TRY(cudaFuncSetAttribute(k_tonemask<N, TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  k_tonemask<N, TI><<<grid, TI * 32, dyn, st>>>(s->d, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
  LAUNCH_CHECK();
  return VB200_OK;
}

/* rowA / rowB: two scratch buffers of `items` rows of rowN floats / ints that are dead while K3 runs (the pipeline
   lends the logmask and iwork rows of the LONG size class, to the short blocks as well: their own rows are too small
   for the ~590 seed lines); NULL selects the fused kernel */
int vb200_launch_tonemask(const vb200_setup *s, cudaStream_t st, int W, int items, const int *psy,
                          const float *logfft, const float *global_ampmax, const float *local_ampmax, float *tone,
                          float *rowA, int *rowB, int rowN) {
  if (items <= 0) return VB200_OK;
  int L = 0, Lmin = 1 << 30;
  for (int k = 0; k < s->h.psys; k++) if (s->h.psy[k].blockflag == W) {
    if (s->h.psy[k].total_octave_lines > L) L = s->h.psy[k].total_octave_lines;
    if (s->h.psy[k].total_octave_lines < Lmin) Lmin = s->h.psy[k].total_octave_lines;
  }
  const int Lstride = (L + 2) | 1;        /* odd: lanes walking equal positions hit different banks */
  const int N = s->h.bs[W] / 2;
  static const bool no_split = getenv("VB200_TM_FUSED") != nullptr;
  if (rowA && rowB && !no_split && L == Lmin && (N == 1024 || N == 128) && L <= rowN && 2 * L <= 4 * (rowN - TM_SVW)) {
    const int grid = (items + TS_WARPS - 1) / TS_WARPS;
    const size_t dynA = (size_t)TS_WARPS * Lstride * sizeof(float);
    const size_t dynC = (size_t)TS_WARPS * Lstride * (sizeof(float) + sizeof(unsigned short));
    if (N == 1024) k_tone_splat<1024><<<grid, TS_WARPS * 32, dynA, st>>>(s->d, items, psy, logfft, global_ampmax, local_ampmax, rowA, rowB, Lstride, rowN);
    else           k_tone_splat<128><<<grid, TS_WARPS * 32, dynA, st>>>(s->d, items, psy, logfft, global_ampmax, local_ampmax, rowA, rowB, Lstride, rowN);
    LAUNCH_CHECK();
    k_tone_chase<<<(items + TS_WARPS * 32 - 1) / (TS_WARPS * 32), TS_WARPS * 32, 0, st>>>(items, L, rowN, rowB);
    LAUNCH_CHECK();
    if (N == 1024) k_tone_fill<1024><<<grid, TS_WARPS * 32, dynC, st>>>(s->d, items, psy, local_ampmax, rowA, rowB, tone, Lstride, rowN);
    else           k_tone_fill<128><<<grid, TS_WARPS * 32, dynC, st>>>(s->d, items, psy, local_ampmax, rowA, rowB, tone, Lstride, rowN);
    LAUNCH_CHECK();
    return VB200_OK;
  }
  if (W) return launch_tm<1024, 4>(s, st, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
  /* short blocks: the chase of the fused kernel runs one lane per ch-block, so more ch-blocks per CTA = more chase lanes */
  static const int ti = getenv("VB200_TM_ITEMS") ? atoi(getenv("VB200_TM_ITEMS")) : 8;
  if (ti >= 16) return launch_tm<128, 16>(s, st, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
  if (ti >= 8) return launch_tm<128, 8>(s, st, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
  return launch_tm<128, 4>(s, st, items, psy, logfft, global_ampmax, local_ampmax, tone, Lstride);
}
