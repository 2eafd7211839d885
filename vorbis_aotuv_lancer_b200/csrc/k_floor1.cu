/* K5 -- floor-1: fit (K5a) and post quantisation + integer floor render (K5b), one warp per
 * ch-block.
 *
 * Replaces floor1_fit (floor1.c:597-750: accumulate_fit :427, fit_line :477, inspect_error :537),
 * floor1_interpolate_fit (:752) and the non-emitting parts of floor1_encode (:774-974).
 *
 * B200 mapping.  The greedy split is a data-dependent loop of at most posts-2 steps, so again a
 * small CTA (one warp, ~6 KB of shared memory) per ch-block and many resident CTAs per SM.
 * Inside a step the error inspection of a candidate line runs across the warp: the integer line is
 * evaluated in closed form per bin (base slope + accumulated carries) instead of being stepped,
 * and the early-exit tests become a warp vote.  Integer sums are exact in any order; the
 * least-squares solve stays in double on one lane in the reference's summation order, because
 * the rounded intercepts decide the posts.  HBM traffic: logmdct + logmask in, posts out (K5a);
 * posts in, ilogmask + coded values out (K5b).
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

__device__ __forceinline__ int dbq(float x) {      /* vorbis_dBquant, floor1.c:285-298 */
  int i = (int)__fadd_rn(__fmul_rn(x, 7.3142857f), 1023.5f);
  return i > 1023 ? 1023 : (i < 0 ? 0 : i);
}

/* num / den for 0 <= num < 2^21, den > 0: reciprocal multiply (num is exact in float) + correction */
__device__ __forceinline__ int div_small(int num, int den) {
  int q = (int)__fmul_rn((float)num, __frcp_rn((float)den));
  if (q * den > num) q--;
  if ((q + 1) * den <= num) q++;
  return q;
}
/* integer line (x0,y0)-(x1,y1) at x, as render_line0 / inspect_error step it (floor1.c:397-425): y0 + base * k +- (rem * k) / adx,
   with its per-line constants hoisted (they are warp-uniform during error inspection) and the
   per-bin division done as an exact reciprocal-multiply + correction: rem*k < 2^21 is exact in float */
struct LineK { int x0, y0, base, rem, adx, neg; float radx; };
__device__ __forceinline__ LineK line_setup(int x0, int x1, int y0, int y1) {
  LineK L; const int dy = y1 - y0; L.x0 = x0; L.y0 = y0; L.adx = x1 - x0;
  const int ab = div_small(abs(dy), L.adx);       /* |dy| < 2^11; C division truncates toward zero */
  L.base = dy < 0 ? -ab : ab;
  L.rem = abs(dy) - ab * L.adx; L.neg = dy < 0; L.radx = __frcp_rn((float)L.adx);
  return L;
}
__device__ __forceinline__ int line_eval(const LineK &L, int x) {
  const int k = x - L.x0, num = L.rem * k;
  int q = (int)((float)num * L.radx);
  if (q * L.adx > num) q--;
  if ((q + 1) * L.adx <= num) q++;
  return L.y0 + L.base * k + (L.neg ? -q : q);
}
/* render_point, floor1.c:260-283 */
__device__ __forceinline__ int point_at(int x0, int x1, int y0, int y1, int x) {
  y0 &= 0x7fff; y1 &= 0x7fff;
  const int dy = y1 - y0, adx = x1 - x0;
  const int off = div_small(abs(dy) * (x - x0), adx);     /* < 2^10 * 2^10 */
  return dy < 0 ? y0 - off : y0 + off;
}

struct Seg { int s[2][6]; };   /* s[class] = sum x, y, xx, yy, xy, count */

/* Per-warp shared memory, carved from the CTA's dynamic allocation with the post capacity of this setup's floors
   (q: 2 KB for a long block, T: 48 B per post interval, 10 B of bookkeeping per post => ~3.7 KB at 29 posts), so that
   ~60 warps are resident per SM; FIT_WARPS ch-blocks share a CTA (the per-SM CTA limit is 32). */
constexpr int FIT_WARPS = 4;
template <int N>
struct FitSmem {
  double (*T)[6];                 /* [pcap] weighted sums of a post interval */
  short *A, *B, *lo, *hi, *memo;  /* [pcap] each */
  unsigned short *q;              /* [QN] quantised mask, bit 15 = "loud" (mdct + twofitatten >= mask), bit 14 = loud and mask != 0 */
  unsigned *bm;                   /* [32] posts inside every lane's run of N/32 bins, one bit per bin */
  /* a lane walks N/32 consecutive bins in accumulate_fit: with 32-bin runs, one pad word per run puts the 32 lanes on
     32 banks (word 17 * lane + t / 2); the 32-bin slices of the error inspection stay conflict-free */
  static constexpr int CS = N / 32, QN = (CS == 32) ? N + 2 * (N / 32) : N;
  static __host__ __device__ constexpr int qi(int i) { return (CS == 32) ? i + ((i >> 5) << 1) : i; }
  static __host__ __device__ size_t bytes(int pcap) { return ((size_t)pcap * 48 + 128 + (size_t)pcap * 10 + (size_t)QN * 2 + 15) & ~(size_t)15; }
  __device__ FitSmem(unsigned char *base, int pcap) {
    T = reinterpret_cast<double (*)[6]>(base);
    bm = reinterpret_cast<unsigned *>(base + (size_t)pcap * 48);
    q = reinterpret_cast<unsigned short *>(base + (size_t)pcap * 48 + 128);
    A = reinterpret_cast<short *>(q + QN); B = A + pcap; lo = B + pcap; hi = lo + pcap; memo = hi + pcap;
  }
};

/* fit_line, floor1.c:477-535, split in three so that the warp shares it:
   (a) per-segment weighted terms, one lane per segment, once per ch-block;
   (b) the six running sums of a run of segments, one lane per sum, added in segment order exactly as
       the reference accumulates them (the rounded intercepts decide the posts, so the order stays);
   (c) pinned ends + the 2x2 solve. */
__device__ __forceinline__ void fit_terms(const Seg &a, float weight, double *T) {
  const int an = a.s[0][5], bnn = a.s[1][5];
  const double w = (double)__fdiv_rn(__fmul_rn((float)(bnn + an), weight), (float)(an + 1)) + 1.;
#pragma unroll
  for (int k = 0; k < 5; k++) T[k] = (double)a.s[1][k] + (double)a.s[0][k] * w;
  T[5] = (double)bnn + (double)an * w;
}

__device__ __forceinline__ int fit_solve(const double *sum, int x0, int x1, int *y0, int *y1) {
  double xb = sum[0], yb = sum[1], x2b = sum[2], y2b = sum[3], xyb = sum[4], bn = sum[5];
  if (*y0 >= 0) { xb += x0; yb += *y0; x2b += x0 * x0; y2b += *y0 * *y0; xyb += *y0 * x0; bn++; }
  if (*y1 >= 0) { xb += x1; yb += *y1; x2b += x1 * x1; y2b += *y1 * *y1; xyb += *y1 * x1; bn++; }
  (void)y2b;
  const double den = bn * x2b - xb * xb;
  if (den > 0.) {
    const double ic = (yb * x2b - xyb * xb) / den, sl = (bn * xyb - xb * yb) / den;
    int v0 = (int)rint(ic + sl * x0), v1 = (int)rint(ic + sl * x1);
    if (v0 > 1023) v0 = 1023;
    if (v1 > 1023) v1 = 1023;
    if (v0 < 0) v0 = 0;
    if (v1 < 0) v1 = 0;
    *y0 = v0; *y1 = v1;
    return 0;
  }
  *y0 = 0; *y1 = 0;
  return 1;
}

/* lanes 0-5 sum run A = segments [a0,a0+na), lanes 6-11 run B = [b0,b0+nb), each in segment order; lane 0 then holds
   the six sums of run A and lane 6 those of run B (the other lanes' copies are meaningless), so that the two solves
   of a split run as ONE pass of fit_solve on lanes 0 and 6 */
__device__ __forceinline__ void fit_two(const double (*T)[6], int a0, int na, int b0, int nb, int lane, double *s) {
  double acc = 0.;
  const int k = lane < 6 ? lane : lane - 6;
  if (lane < 12) {
    const int s0 = lane < 6 ? a0 : b0, n = lane < 6 ? na : nb;
    for (int i = 0; i < n; i++) acc += T[s0 + i][k];
  }
#pragma unroll
  for (int q = 0; q < 6; q++) s[q] = __shfl_down_sync(FULLMASK, acc, q);
}

__device__ __forceinline__ int pick(const short *A, const short *B, int i) {
  if (A[i] < 0) return B[i];
  if (B[i] < 0) return A[i];
  return (A[i] + B[i]) >> 1;
}

template <int N, int MINB>
__global__ void __launch_bounds__(FIT_WARPS * 32, MINB)
k_floor_fit(const DevSetup *__restrict__ S, int floor_index, int W, int ch, const float *__restrict__ logmdct, const float *__restrict__ logmask,
            int mdct_stride, int mask_stride, const int *__restrict__ gate, int *__restrict__ posts_out, int *__restrict__ fitted, int items, int pcap) {
  extern __shared__ __align__(16) unsigned char fit_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * FIT_WARPS + warp;
  if (item >= items) return;                      /* warps are independent: only __syncwarp below */
  FitSmem<N> sm(fit_dyn + (size_t)warp * FitSmem<N>::bytes(pcap), pcap);
  const DevFloor &F = S->flr[floor_index >= 0 ? floor_index : S->floor_of_ch[W][item % ch]];
  const int n = F.n, posts = F.posts;
  int *out = posts_out + (size_t)item * VB200_MAX_POSTS;
  if (gate && !gate[item]) { if (lane == 0) fitted[item] = 0; return; }
  const float *lm = logmdct + (size_t)item * mdct_stride, *mk = logmask + (size_t)item * mask_stride;

  for (int i = lane; i < n; i += 32) {
    const float m = __ldg(mk + i);
    const int qm = dbq(m);
    const bool loud = __fadd_rn(__ldg(lm + i), F.twofitatten) >= m;
    sm.q[FitSmem<N>::qi(i)] = (unsigned short)(qm | (loud ? (qm ? 0xc000 : 0x8000) : 0));
  }
  /* accumulate_fit (floor1.c:427-475) for all post intervals in ONE walk over the bins.  Interval j holds the bins
     sorted[j] .. min(sorted[j+1], n-1), both ends included: a bin at an inner post belongs to two intervals.  Lane L
     walks bins L*CS .. L*CS+CS-1 with twelve running integer sums (class a = loud bins, class o = all bins with a
     quantised mask != 0; class b of the reference is o - a) for the interval it is in; at a post it adds them to the
     interval's row (shared-memory atomics: integer sums are exact in any order), moves on to the next interval and
     starts it with the post's own bin; what is left at the end of its run goes to the row of the interval it ends in.
     The rows are the T rows of the next step, read back as twelve ints and rewritten as six doubles by one lane each.
     (Was: the whole warp over one interval after the other, 12 warp reductions each -- 17 of the 28 intervals of a
     long block are narrower than a warp.) */
  constexpr int CS = FitSmem<N>::CS;
  int *Ti = reinterpret_cast<int *>(sm.T);
  for (int i = lane; i < (posts - 1) * 12; i += 32) Ti[i] = 0;
  sm.bm[lane] = 0;
  __syncwarp();
  for (int k = 1 + lane; k < posts - 1; k += 32) { const int x = F.sorted[k]; atomicOr(&sm.bm[x / CS], 1u << (x % CS)); }
  __syncwarp();
  int any;
  {
    const unsigned mask = sm.bm[lane];
    int cur = __popc(mask);                         /* -> number of inner posts below this lane's first bin */
    {
      int incl = cur;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += y; }
      cur = incl - cur;
    }
    int ax = 0, ay = 0, ax2 = 0, ay2 = 0, axy = 0, an = 0, ox = 0, oy = 0, ox2 = 0, oy2 = 0, oxy = 0, on_ = 0, loudbins = 0;
    const int i0 = lane * CS;
    auto flush = [&]() {
      int *row = Ti + cur * 12;
      atomicAdd(row + 0, ax); atomicAdd(row + 1, ay); atomicAdd(row + 2, ax2); atomicAdd(row + 3, ay2); atomicAdd(row + 4, axy); atomicAdd(row + 5, an);
      atomicAdd(row + 6, ox); atomicAdd(row + 7, oy); atomicAdd(row + 8, ox2); atomicAdd(row + 9, oy2); atomicAdd(row + 10, oxy); atomicAdd(row + 11, on_);
    };
#pragma unroll 4
    for (int t = 0; t < CS; t++) {
      const int i = i0 + t;
      int la = 0, on = 0, q = 0;
      if (i < n) { const int qv = sm.q[FitSmem<N>::qi(i)]; q = qv & 0x3fff; on = q != 0; la = (qv >> 14) & 1; }
      const int ii = i * i, qq = q * q, iq = i * q;
      const int cx = on ? i : 0, cx2 = on ? ii : 0, dx = la ? i : 0, dy = la ? q : 0, dx2 = la ? ii : 0, dy2 = la ? qq : 0, dxy = la ? iq : 0;
      ax += dx; ay += dy; ax2 += dx2; ay2 += dy2; axy += dxy; an += la;
      ox += cx; oy += q; ox2 += cx2; oy2 += qq; oxy += iq; on_ += on;        /* q == 0 adds nothing to the y sums */
      loudbins += la;
      if ((mask >> t) & 1u) {                       /* bin i is the post that ends interval cur and starts cur + 1 */
        flush();
        cur++;
        ax = dx; ay = dy; ax2 = dx2; ay2 = dy2; axy = dxy; an = la;
        ox = cx; oy = q; ox2 = cx2; oy2 = qq; oxy = iq; on_ = on;
      }
    }
    if (i0 < n) flush();
    any = __any_sync(FULLMASK, loudbins > 0);
  }
  __syncwarp();
  for (int j = lane; j < posts - 1; j += 32) {
    Seg keep;
    const int *row = Ti + j * 12;
#pragma unroll
    for (int k = 0; k < 6; k++) { keep.s[0][k] = row[k]; keep.s[1][k] = row[6 + k] - row[k]; }
    fit_terms(keep, F.twofitweight, sm.T[j]);       /* in place: the row has been read */
  }
  __syncwarp();

  if (!any) { if (lane == 0) fitted[item] = 0; return; }
  for (int i = lane; i < posts; i += 32) { sm.A[i] = -200; sm.B[i] = -200; sm.lo[i] = 0; sm.hi[i] = 1; sm.memo[i] = -1; }
  __syncwarp();
  {
    double sa[6];
    fit_two(sm.T, 0, posts - 1, 0, 0, lane, sa);
    if (lane == 0) {
      int y0 = -200, y1 = -200;
      fit_solve(sa, F.sorted[0], F.sorted[posts - 1], &y0, &y1);
      sm.A[0] = sm.B[0] = y0; sm.A[1] = sm.B[1] = y1;
    }
  }
  __syncwarp();

  const int imaxover = (int)F.maxover, imaxunder = (int)F.maxunder;
  const bool int_limits = (float)imaxover == F.maxover && (float)imaxunder == F.maxunder && imaxover >= 0 && imaxunder >= 0 &&
                          imaxover < 4096 && imaxunder < 4096;
  for (int i = 2; i < posts; i++) {
    const int sp = F.reverse[i], ln = sm.lo[sp], hn = sm.hi[sp];
    if (sm.memo[ln] == hn) continue;            /* uniform across the warp */
    __syncwarp();
    if (lane == 0) sm.memo[ln] = hn;
    const int lsp = F.reverse[ln], hsp = F.reverse[hn];
    const int lx = F.postlist[ln], hx = F.postlist[hn];
    const int ly = pick(sm.A, sm.B, ln), hy = pick(sm.A, sm.B, hn);
    /* inspect_error across the warp */
    int viol = 0, mse = 0;
    const LineK LK = line_setup(lx, hx, ly, hy);
    if (int_limits) {
      /* y, v < 2^10 and the limits are whole numbers: the reference's float comparisons are exact integer ones.
         A lane evaluates the line once, at its first bin, and then steps it 32 bins at a time: with
         rem * k = q * adx + r and rem * 32 = a * adx + b, the carry count at k + 32 is q + a + (r + b >= adx).
         A bin can violate only when it is loud and its mask is not zero (flag bit 14) -- or when it is the first
         bin of the line, loud, with a zero mask (floor1.c:550-553), which lane 0 tests on its own.  A violation
         settles the outcome (floor1.c:556,562), so the warp stops at the first 32-bin slice that has one. */
      const int a32 = div_small(LK.rem * 32, LK.adx), b32 = LK.rem * 32 - a32 * LK.adx;
      const int ystep = LK.base * 32 + (LK.neg ? -a32 : a32), ycarry = LK.neg ? -1 : 1;
      const unsigned span = (unsigned)(imaxover + imaxunder);
      int x = lx + lane, y = 0, r = 0;
      int xi = FitSmem<N>::qi(x);                    /* x += 32 is one run further: 32 + its pad */
      constexpr int XSTEP = FitSmem<N>::qi(32);
      if (x < hx) {
        const int k = lane, num = LK.rem * k;
        int q = (int)((float)num * LK.radx);
        if (q * LK.adx > num) q--;
        if ((q + 1) * LK.adx <= num) q++;
        r = num - q * LK.adx;
        y = ly + LK.base * k + (LK.neg ? -q : q);
      }
      if (lane == 0) {
        const int qv = sm.q[FitSmem<N>::qi(lx)];
        if ((qv & 0xc000) == 0x8000 && (unsigned)(ly + imaxover) > span) viol = 1;
      }
      for (int x0 = lx; x0 < hx; x0 += 32) {         /* uniform trip count */
        if (x < hx) {
          const int qv = sm.q[xi], d = y - (qv & 0x3fff);
          mse += d * d;
          if ((qv & 0x4000) && (unsigned)(d + imaxover) > span) viol = 1;
          x += 32; xi += XSTEP; y += ystep; r += b32;
          if (r >= LK.adx) { r -= LK.adx; y += ycarry; }
        }
        if (__any_sync(FULLMASK, viol)) { viol = 1; break; }
      }
    } else {
      for (int x = lx + lane; x < hx; x += 32) {
        const int qv = sm.q[FitSmem<N>::qi(x)];
        const int y = line_eval(LK, x), v = qv & 0x3fff;
        mse += (y - v) * (y - v);
        if ((qv >> 15) && (x == lx || v)) {
          if (__fadd_rn((float)y, F.maxover) < (float)v) viol = 1;
          if (__fsub_rn((float)y, F.maxunder) > (float)v) viol = 1;
        }
      }
      viol = __any_sync(FULLMASK, viol);
    }
    if (!viol) mse = warp_sum_i(mse);
    int bad;
    {
      const int cnt = hx - lx;
      if (viol) bad = 1;
      else if (__fdiv_rn(__fmul_rn(F.maxover, F.maxover), (float)cnt) > F.maxerr) bad = 0;
      else if (__fdiv_rn(__fmul_rn(F.maxunder, F.maxunder), (float)cnt) > F.maxerr) bad = 0;
      else bad = ((float)(mse / cnt) > F.maxerr) ? 1 : 0;
    }
    __syncwarp();
    int r0 = 0, r1 = 0, ly0 = -200, ly1 = -200, hy0 = -200, hy1 = -200;
    if (bad) {                                             /* `bad` is warp-uniform */
      double ss[6];
      fit_two(sm.T, lsp, sp - lsp, sp, hsp - sp, lane, ss);
      /* the two solves run side by side on lanes 0 and 6, one pass through the code */
      int ya = -200, yb = -200, rr = 0;
      if (lane == 0 || lane == 6) rr = fit_solve(ss, F.sorted[lane ? sp : lsp], F.sorted[lane ? hsp : sp], &ya, &yb);
      r0 = __shfl_sync(FULLMASK, rr, 0); ly0 = __shfl_sync(FULLMASK, ya, 0); ly1 = __shfl_sync(FULLMASK, yb, 0);
      r1 = __shfl_sync(FULLMASK, rr, 6); hy0 = __shfl_sync(FULLMASK, ya, 6); hy1 = __shfl_sync(FULLMASK, yb, 6);
    }
    /* every value below is warp-uniform (the solves were broadcast), so the bookkeeping needs no single-lane section */
    bool relink = false;
    if (bad) {
      if (r0) { ly0 = ly; ly1 = hy0; }
      if (r1) { hy0 = ly1; hy1 = hy; }
      relink = !(r0 && r1) && (ly1 >= 0 || hy0 >= 0);
    }
    if (lane == 0) {
      if (bad && !(r0 && r1)) {
        sm.B[ln] = ly0; if (ln == 0) sm.A[ln] = ly0;
        sm.A[i] = ly1; sm.B[i] = hy0;
        sm.A[hn] = hy1; if (hn == 1) sm.B[hn] = hy1;
      } else { sm.A[i] = -200; sm.B[i] = -200; }
    }
    if (relink) {
      /* the new post becomes the neighbour of the run of sorted positions next to it that pointed past it
         (floor1.c:703-710: two loops that stop at the first position pointing elsewhere): lane t looks at the t-th
         position below / above sp, the run is the leading ones of the vote */
      for (int j0 = sp - 1; j0 >= 0; j0 -= 32) {
        const int j = j0 - lane;
        const unsigned m = __ballot_sync(FULLMASK, j >= 0 && sm.hi[j] == hn);
        const int run = __ffs(~m) - 1;             /* 32 set bits: ffs(0) - 1 = -1 */
        if (run < 0 || lane < run) { if (j >= 0) sm.hi[j] = i; }
        if (run >= 0) break;
      }
      for (int j0 = sp + 1; j0 < posts; j0 += 32) {
        const int j = j0 + lane;
        const unsigned m = __ballot_sync(FULLMASK, j < posts && sm.lo[j] == ln);
        const int run = __ffs(~m) - 1;
        if (run < 0 || lane < run) { if (j < posts) sm.lo[j] = i; }
        if (run >= 0) break;
      }
    }
    __syncwarp();
  }
  __syncwarp();
  /* final posts (floor1.c:728-747): predictions level by level over the neighbour tree, as in k_floor_encode */
  {
    unsigned short *o16 = reinterpret_cast<unsigned short *>(sm.lo);      /* lo/hi are dead now */
    if (lane < 2) { const int v = pick(sm.A, sm.B, lane); o16[lane] = (unsigned short)v; out[lane] = v; }
    __syncwarp();
    for (int l = 1; l <= F.nlevels; l++) {
      for (int k = F.lv_start[l] + lane; k < F.lv_start[l + 1]; k += 32) {
        const int i = F.lv_order[k];
        const int ln = F.lonb[i - 2], hn = F.hinb[i - 2];
        const int pred = point_at(F.postlist[ln], F.postlist[hn], o16[ln], o16[hn], F.postlist[i]);
        const int vx = pick(sm.A, sm.B, i);
        const int v = (vx >= 0 && pred != vx) ? vx : (pred | 0x8000);
        o16[i] = (unsigned short)v; out[i] = v;
      }
      __syncwarp();
    }
    if (lane == 0) fitted[item] = 1;
  }
}

/* floor1_interpolate_fit (floor1.c:752-771) for the managed-mode blob ladder: out = blend(A,B,del) */
__global__ void k_floor_interp(const DevSetup *__restrict__ S, int floor_index, int W, int ch, const int *__restrict__ A, const int *__restrict__ B,
                               const int *__restrict__ fa, const int *__restrict__ fb, int del, int *__restrict__ out, int *__restrict__ fo, int items) {
  const int item = blockIdx.x * blockDim.y + threadIdx.y, i = threadIdx.x;
  if (item >= items) return;
  const int posts = S->flr[floor_index >= 0 ? floor_index : S->floor_of_ch[W][item % ch]].posts;
  const int ok = fa[item] && fb[item];
  if (i == 0) fo[item] = ok;
  if (!ok) return;
  for (int k = i; k < posts; k += blockDim.x) {
    const int a = A[(size_t)item * VB200_MAX_POSTS + k], b = B[(size_t)item * VB200_MAX_POSTS + k];
    int v = ((65536 - del) * (a & 0x7fff) + del * (b & 0x7fff) + 32768) >> 16;
    if ((a & 0x8000) && (b & 0x8000)) v |= 0x8000;
    out[(size_t)item * VB200_MAX_POSTS + k] = v;
  }
}

template <int N>
__global__ void __launch_bounds__(32)
k_floor_encode(const DevSetup *__restrict__ S, int floor_index, int W, int ch, int *__restrict__ posts_io, const int *__restrict__ fitted,
               int *__restrict__ vals_out, int *__restrict__ ilogmask, int *__restrict__ nonzero) {
  __shared__ int post[VB200_MAX_POSTS];
  __shared__ int segx[VB200_MAX_POSTS], segy[VB200_MAX_POSTS];
  __shared__ int nseg;
  const int item = blockIdx.x, lane = threadIdx.x;
  const DevFloor &F = S->flr[floor_index >= 0 ? floor_index : S->floor_of_ch[W][item % ch]];
  const int posts = F.posts;
  int *ilog = ilogmask + (size_t)item * N;
  if (!fitted[item]) {
    for (int i = lane; i < N; i += 32) ilog[i] = 0;
    if (lane == 0) nonzero[item] = 0;
    return;
  }
  int *pio = posts_io + (size_t)item * VB200_MAX_POSTS, *vals = vals_out + (size_t)item * VB200_MAX_POSTS;
  for (int i = lane; i < posts; i += 32) {
    const int p = pio[i];
    int v = p & 0x7fff;
    switch (F.mult) { case 1: v >>= 2; break; case 2: v >>= 3; break; case 3: v /= 12; break; case 4: v >>= 4; break; }
    post[i] = v | (p & 0x8000);
  }
  __syncwarp();
  /* Post prediction (floor1.c:880-920).  The reference walks the posts in index order; a post's prediction reads its
     two neighbours, which lie on lower levels of the neighbour tree, and only posts on higher levels touch its flag --
     so the posts of one level go side by side on the lanes, level after level. */
  if (lane == 0) { vals[0] = post[0]; vals[1] = post[1]; }
  for (int l = 1; l <= F.nlevels; l++) {
    for (int k = F.lv_start[l] + lane; k < F.lv_start[l + 1]; k += 32) {
      const int i = F.lv_order[k];
      const int ln = F.lonb[i - 2], hn = F.hinb[i - 2];
      const int pred = point_at(F.postlist[ln], F.postlist[hn], post[ln], post[hn], F.postlist[i]);
      const int p = post[i];
      if ((p & 0x8000) || pred == p) { post[i] = pred | 0x8000; vals[i] = 0; }
      else {
        const int room = (F.quant_q - pred < pred) ? F.quant_q - pred : pred;
        int d = p - pred;
        if (d < 0) d = (d < -room) ? room - d - 1 : -1 - (d << 1);
        else d = (d >= room) ? d + room : d << 1;
        vals[i] = d;
        atomicAnd(&post[ln], 0x7fff); atomicAnd(&post[hn], 0x7fff);
      }
    }
    __syncwarp();
  }
  /* the posts that take part in the rendered curve, in x order: a compaction over the sorted posts */
  {
    int k = 1;
    if (lane == 0) { segx[0] = 0; segy[0] = post[0] * F.mult; }
    for (int j0 = 1; j0 < posts; j0 += 32) {
      const int j = j0 + lane;
      int cur = 0, hy = 0; bool in = false;
      if (j < posts) { cur = F.forward[j]; const int pv = post[cur]; hy = pv & 0x7fff; in = (hy == pv); }
      const unsigned m = __ballot_sync(FULLMASK, in);
      if (in) { const int o = k + __popc(m & ((1u << lane) - 1u)); segx[o] = F.postlist[cur]; segy[o] = hy * F.mult; }
      k += __popc(m);
    }
    if (lane == 0) { nseg = k; nonzero[item] = 1; }
  }
  __syncwarp();
  for (int i = lane; i < posts; i += 32) pio[i] = post[i];
  const int ns = nseg;
  for (int s = 0; s + 1 < ns; s++) {
    const int lx = segx[s], hx = segx[s + 1], ly = segy[s], hy = segy[s + 1];
    const int lim = N < hx ? N : hx;
    const LineK LK = line_setup(lx, hx, ly, hy);         /* warp-uniform per segment; one exact reciprocal division per bin */
    for (int x = lx + lane; x < lim; x += 32) ilog[x] = line_eval(LK, x);
  }
  { const int hx = ns > 1 ? segx[ns - 1] : 0, ly = segy[ns - 1]; for (int x = hx + lane; x < N; x += 32) ilog[x] = ly; }
}

}  // namespace

int vb200_launch_floor_fit(const vb200_setup *s, cudaStream_t st, int W, int items, int floor_index,
                           const float *logmdct, const float *logmask, int mdct_stride, int mask_stride,
                           const int *gate, int *posts, int *fitted) {
  if (items <= 0) return VB200_OK;
  /* look->n = postlist[1] may be smaller than N2 (the 5.1 LFE floor covers 10 bins) */
  if (floor_index >= 0 && (s->h.flr[floor_index].type != 1 || s->h.flr[floor_index].n > s->h.bs[W] / 2 || s->h.flr[floor_index].posts < 2)) { vb200_set_error("floor %d is not a floor-1 of this block size", floor_index); return VB200_EINVAL; }
  int pcap = 2;
  for (int k = 0; k < s->h.floors; k++) if (s->h.flr[k].type == 1 && s->h.flr[k].posts > pcap) pcap = s->h.flr[k].posts;
  pcap = (pcap + 3) & ~3;
  const int grid = (items + FIT_WARPS - 1) / FIT_WARPS;
  static const int occ = getenv("VB200_FIT_OCC") ? atoi(getenv("VB200_FIT_OCC")) : 8;    /* resident CTAs per SM the register budget aims at */
#define FIT_ARGS s->d, floor_index, W, s->h.channels, logmdct, logmask, mdct_stride, mask_stride, gate, posts, fitted, items, pcap
  if (W) {
    const size_t sh = FIT_WARPS * FitSmem<1024>::bytes(pcap);
    if (occ <= 8) k_floor_fit<1024, 8><<<grid, FIT_WARPS * 32, sh, st>>>(FIT_ARGS);
    else if (occ <= 10) k_floor_fit<1024, 10><<<grid, FIT_WARPS * 32, sh, st>>>(FIT_ARGS);
    else k_floor_fit<1024, 12><<<grid, FIT_WARPS * 32, sh, st>>>(FIT_ARGS);
  } else k_floor_fit<128, 8><<<grid, FIT_WARPS * 32, FIT_WARPS * FitSmem<128>::bytes(pcap), st>>>(FIT_ARGS);
#undef FIT_ARGS
  LAUNCH_CHECK();
  return VB200_OK;
}

int vb200_launch_floor_interp(const vb200_setup *s, cudaStream_t st, int W, int items, int floor_index, const int *A, const int *B,
                              const int *fa, const int *fb, int del, int *out, int *fo) {
  if (items <= 0) return VB200_OK;
  dim3 blk(32, 4);
  k_floor_interp<<<(items + 3) / 4, blk, 0, st>>>(s->d, floor_index, W, s->h.channels, A, B, fa, fb, del, out, fo, items);
  LAUNCH_CHECK();
  return VB200_OK;
}

int vb200_launch_floor_encode(const vb200_setup *s, cudaStream_t st, int W, int items, int floor_index,
                              int *posts, const int *fitted, int *vals, int *ilogmask, int *nonzero) {
  if (items <= 0) return VB200_OK;
  if (W) k_floor_encode<1024><<<items, 32, 0, st>>>(s->d, floor_index, W, s->h.channels, posts, fitted, vals, ilogmask, nonzero);
  else   k_floor_encode<128><<<items, 32, 0, st>>>(s->d, floor_index, W, s->h.channels, posts, fitted, vals, ilogmask, nonzero);
  LAUNCH_CHECK();
  return VB200_OK;
}
