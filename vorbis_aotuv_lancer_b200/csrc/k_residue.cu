/* K7 -- residue partition classification (K7a) and cascade lattice VQ search (K7b).
 *
 * Replaces res1_class/_01class (res0.c:735,406-468), res2_class/_2class (:759,473-526) and
 * res1_forward/res2_forward -> _01forward -> _encodepart -> local_book_besterror
 * (:715,773-799,528-640,384-404,316-378).  Codeword emission (vorbis_book_encode,
 * codebook.c:402-410) is not done here: the kernel returns the (book,entry) symbol list in the
 * reference's emission order and the bit packer consumes it.
 *
 * B200 mapping.  Partitions are independent; the cascade (stage s quantises what stage s-1 left)
 * is serial per partition.  One CTA per block and submap, four warps, each warp owning every
 * fourth partition in shared memory.  Inside a stage the dim-sized vectors of the partition map to
 * lanes; the lattice index is arithmetic, and when it lands on an entry without a codeword the
 * whole warp searches the book's compacted used-entry table (first minimum wins, as the
 * reference's linear scan).  Entries land in a shared-memory table; a warp scan over the per-
 * (stage, partition, row) symbol counts then turns the table into the emission-ordered list with
 * coalesced stores.  HBM traffic: the quantised residue in, ~2 KB of symbols out.
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

/* ------------------------------------------------------------------ K7a */
template <int CHC>                /* CHC > 0: residue type 2 over CHC channels as a compile-time fact, as in k_residue_forward */
__global__ void __launch_bounds__(64)
k_residue_class(const DevSetup *__restrict__ S, int resnum, int n, int ch_, const int *__restrict__ chlist, int tot_ch,
                const int *__restrict__ in, const int *__restrict__ nonzero, int *__restrict__ partword, int *__restrict__ classified) {
  const int blk = blockIdx.x, tid = threadIdx.x;
  const DevResidue &R = S->res[resnum];
  constexpr bool STEREO2 = CHC == 2;
  const int ch = CHC ? CHC : ch_;
  const int partvals = (R.end - R.begin) / R.grouping;
  const int *base = in + (size_t)blk * tot_ch * n;
  const int *nz = nonzero + (size_t)blk * tot_ch;
  int *pw = partword + (size_t)blk * ch * partvals;
  int used = 0;
  for (int k = 0; k < ch; k++) if (nz[chlist[k]]) used++;
  if (tid == 0) classified[blk] = used ? ((CHC || R.type == 2) ? 1 : used) : 0;
  if (!used) return;
  if (CHC || R.type == 2) {
    const int per = (R.grouping + ch - 1) / ch;   /* positions per channel per partition (res0.c:498-503) */
    if (STEREO2 && per == 16 && (R.grouping & 1) == 0 && ((R.begin >> 1) & 3) == 0 && (n & 3) == 0) {
      /* a partition is 16 consecutive values of each channel: four 128-bit loads per channel, all eight requested
         before the first is used (the scalar walk read a 64-byte run four bytes at a time, lane by lane) */
      const int4 *r0 = reinterpret_cast<const int4 *>(base + (size_t)chlist[0] * n + (R.begin >> 1));
      const int4 *r1 = reinterpret_cast<const int4 *>(base + (size_t)chlist[1] * n + (R.begin >> 1));
      for (int p = tid; p < partvals; p += blockDim.x) {
        int4 a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { a[q] = __ldg(r0 + p * 4 + q); b[q] = __ldg(r1 + p * 4 + q); }
        int mag = 0, ang = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          mag = max(max(max(mag, abs(a[q].x)), max(abs(a[q].y), abs(a[q].z))), abs(a[q].w));
          ang = max(max(max(ang, abs(b[q].x)), max(abs(b[q].y), abs(b[q].z))), abs(b[q].w));
        }
        int c = 0;
        for (; c < R.partitions - 1; c++) if (mag <= R.classmetric1[c] && ang <= R.classmetric2[c]) break;
        pw[p] = c;
      }
      return;
    }
    for (int p = tid; p < partvals; p += blockDim.x) {
      int l = R.begin / ch + p * per, mag = 0, ang = 0;
      for (int j = 0; j < R.grouping; j += ch, l++) {
        const int a = abs(__ldg(base + (size_t)chlist[0] * n + l));
        if (a > mag) mag = a;
        for (int k = 1; k < ch; k++) { const int b = abs(__ldg(base + (size_t)chlist[k] * n + l)); if (b > ang) ang = b; }
      }
      int c = 0;
      for (; c < R.partitions - 1; c++) if (mag <= R.classmetric1[c] && ang <= R.classmetric2[c]) break;
      pw[p] = c;
    }
  } else {
    const float scale = (float)(100. / (double)R.grouping);
    int row = 0;
    for (int k = 0; k < ch; k++) {
      if (!nz[chlist[k]]) continue;
      const int *v = base + (size_t)chlist[k] * n + R.begin;
      for (int p = tid; p < partvals; p += blockDim.x) {
        int mx = 0, ent = 0;
        for (int t = 0; t < R.grouping; t++) { const int a = abs(__ldg(v + p * R.grouping + t)); if (a > mx) mx = a; ent += a; }
        ent = (int)__fmul_rn((float)ent, scale);
        int c = 0;
        for (; c < R.partitions - 1; c++) if (mx <= R.classmetric1[c] && (R.classmetric2[c] < 0 || ent < R.classmetric2[c])) break;
        pw[row * partvals + p] = c;
      }
      row++;
    }
  }
}

/* ------------------------------------------------------------------ K7b */
constexpr int RF_WARPS = 4;
constexpr int RF_MAXG = 64;     /* largest partition size handled (grouping) */

constexpr int RF_SLOTS = 4;     /* partitions a warp works on at once */
constexpr int RF_LP = 32 / RF_SLOTS;

constexpr int RF_SKEW = 9;      /* the four slots of a warp start 9 banks apart: distinct mod 4 (8 lanes at stride dim = 4
                                   words cover every fourth bank) and nearly disjoint runs at stride 2 -- at a stride of
                                   RF_MAXG all four slots sat on the same banks (4-way conflicts on every access) */
struct RfShared {
  int vec[RF_WARPS][RF_SLOTS][RF_MAXG + RF_SKEW];
  int red_d[RF_WARPS], red_i[RF_WARPS];
};

#ifndef RF_MINB
#define RF_MINB 10         /* 48 registers, 10 CTAs per SM (specialised kernel: 8 / 10 / 12 CTAs: 2.65 / 2.57 / 2.64 ms) */
#endif
/* CHC > 0: residue type 2 over CHC channels as compile-time facts (2: stereo, interleave by a shift; 5: the five
   full-band channels of a 5.1 stream, interleave by a constant division): one row, the row / channel divisions and the
   type switches fold away (the generic kernel is 3.9 k instructions and spent a fifth of its stall samples waiting for
   instruction fetch) */
template <int CHC>
__global__ void __launch_bounds__(RF_WARPS * 32, RF_MINB)
k_residue_forward(const DevSetup *__restrict__ S, int resnum, int n, int ch_, const int *__restrict__ chlist, int tot_ch,
                  int *__restrict__ in, const int *__restrict__ nonzero, const int *__restrict__ partword,
                  int maxvec, int *__restrict__ syms, int max_syms, int *__restrict__ nsyms) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ RfShared sh;
  const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DevResidue &R = S->res[resnum];
  const DevBook *books = S->book;
  constexpr bool STEREO2 = CHC == 2;
  const int ch = CHC ? CHC : ch_;
  const int rtype = CHC ? 2 : R.type;
  const int partvals = (R.end - R.begin) / R.grouping, G = R.grouping, stages = R.stages;
  int *base = in + (size_t)blk * tot_ch * n;
  const int *nz = nonzero + (size_t)blk * tot_ch;
  int used = 0;
  for (int k = 0; k < ch; k++) if (nz[chlist[k]]) used++;
  if (!used) { if (tid == 0) nsyms[blk] = 0; return; }
  const int rows = CHC ? 1 : ((rtype == 2) ? 1 : used);
  const int *pw = partword + (size_t)blk * ch * partvals;
  /* dynamic shared memory: entry table [stages][partvals][rows][maxvec] (int16) + scan offsets + row->channel map */
  short *ent = reinterpret_cast<short *>(dyn);
  const size_t ent_elems = (size_t)stages * partvals * rows * maxvec;
  int *offs = reinterpret_cast<int *>(dyn + ((ent_elems * 2 + 15) & ~(size_t)15));   /* [stages*partvals*rows + 1] */
  __shared__ int rowch[VB200_MAX_CHANNELS];
  __shared__ short tab_bk[64 * 8], tab_nvec[64 * 8];      /* per (class, stage): book and vectors per partition */
  __shared__ int next_pr;
  if (tid == 0) { next_pr = 0; int r = 0; for (int k = 0; k < ch; k++) if (rtype == 2 || nz[chlist[k]]) rowch[r++] = chlist[k]; }
  for (int i = tid; i < R.partitions * 8; i += RF_WARPS * 32) {
    const int cls = i >> 3, st2 = i & 7;
    int bk = -1, nv = 0;
    if (st2 < stages && (R.secondstages[cls] & (1 << st2))) { bk = R.partbooks[i]; if (bk >= 0) nv = G / books[bk].dim; }
    tab_bk[i] = (short)bk; tab_nvec[i] = (short)nv;
  }
  const int ch_shift = STEREO2 ? 1 : (CHC ? -1 : ((ch & (ch - 1)) == 0 ? __ffs(ch) - 1 : -1));
  const int pd_mask = (R.phrase_dim & (R.phrase_dim - 1)) == 0 ? R.phrase_dim - 1 : -1;
  __syncthreads();

  /* ---- emission order: symbols per (stage, partition, row) slot and their scan.  The counts depend only on the
     partition classes, so warp 0 lays the offsets out BEFORE the cascade and then joins it (the partitions are drawn
     from a counter, so its late start balances out); slots are walked stage by stage so that no index has to be split
     by division. ---- */
  const int pd = R.phrase_dim;
  const int prow = partvals * rows;
  const int nslots = stages * prow;
  const int phrase_entries = books[R.groupbook].entries;
  auto phrase_val = [&](int r2, int p) {
    int val = pw[r2 * partvals + p];
    for (int k = 1; k < pd; k++) { val *= R.partitions; if (p + k < partvals) val += pw[r2 * partvals + p + k]; }
    return val;
  };
  if (warp == 0) {
    int carry = 0;
    for (int s = 0; s < stages; s++)
      for (int b0 = 0; b0 < prow; b0 += 32) {
        const int pr = b0 + lane;
        int tot = 0;
        if (pr < prow) {
          const int p = rows == 1 ? pr : pr / rows, row = rows == 1 ? 0 : pr % rows;
          tot = tab_nvec[pw[row * partvals + p] * 8 + s];
          if (s == 0 && row == 0 && (pd_mask >= 0 ? (p & pd_mask) == 0 : (p % pd) == 0))
            for (int r2 = 0; r2 < rows; r2++) if (phrase_val(r2, p) < phrase_entries) tot++;   /* phrase words lead the group */
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += y; }
        if (pr < prow) offs[s * prow + pr] = carry + incl - tot;
        carry += __shfl_sync(FULLMASK, incl, 31);
      }
    if (lane == 0) { offs[nslots] = carry; nsyms[blk] = carry < max_syms ? carry : max_syms; }
  }

  /* ---- pass 1: cascade VQ per (partition,row).  A partition holds G/dim <= 16 vectors (often 4 or 8), so a warp
     works on RF_SLOTS partitions at once, RF_LP lanes each; the stage loop is shared, the books differ per slot ---- */
  const int slot = lane / RF_LP, l = lane % RF_LP;
  for (;;) {
    /* partitions cost very different amounts (stages of their class, exhaustive searches): the warps draw the next
       RF_SLOTS partitions from a counter instead of owning a fixed quarter */
    int pr0 = 0;
    if (lane == 0) pr0 = atomicAdd(&next_pr, RF_SLOTS);
    pr0 = __shfl_sync(FULLMASK, pr0, 0);
    if (pr0 >= partvals * rows) break;
    const int pr = pr0 + slot;
    const bool valid = pr < partvals * rows;
    const int p = valid ? (rows == 1 ? pr : pr / rows) : 0, row = valid ? (rows == 1 ? 0 : pr % rows) : 0;
    const int cls = valid ? pw[row * partvals + p] : 0;
    int *v = sh.vec[warp][slot];
    const int off = p * G + R.begin;
    /* stage the partition: res2 reads the channel-interleaved vector work[j*ch+c]=in[c][j] (res0.c:782-787) */
    if (valid)
      for (int t0 = l; t0 < G; t0 += 4 * RF_LP) {          /* four loads in flight per lane, then the stores */
        int tmp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = t0 + u * RF_LP;
          tmp[u] = 0;
          if (t < G) {
            if (rtype == 2) {
              const int w = off + t;
              const int wc = ch_shift >= 0 ? (w & (ch - 1)) : w % ch, wj = ch_shift >= 0 ? (w >> ch_shift) : w / ch;
              tmp[u] = base[(size_t)chlist[wc] * n + wj];
            }
            else tmp[u] = base[(size_t)rowch[row] * n + off + t];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int t = t0 + u * RF_LP; if (t < G) v[t] = tmp[u]; }
      }
    __syncwarp();
    for (int s = 0; s < stages; s++) {
      const int bk = valid ? tab_bk[cls * 8 + s] : -1;
      if (!__any_sync(FULLMASK, bk >= 0)) continue;
      const DevBook &B = books[bk >= 0 ? bk : 0];
      const int dim = B.dim, nvec = bk >= 0 ? tab_nvec[cls * 8 + s] : 0, qv = B.quantvals, ze = qv >> 1;
      const int bdelta = B.delta, bmin = B.minval;
      const unsigned brcp = B.delta_rcp;
      short *e_out = ent + (((size_t)s * partvals + p) * rows + row) * maxvec;
      for (int t = l; __any_sync(FULLMASK, t < nvec); t += RF_LP) {
        bool fb = false;
        int index = 0;
        if (t < nvec) {
          int *a = v + t * dim;
          int pt[8];
          bool on_lattice = true;          /* the input is exactly a lattice point inside the book's range */
          /* components from the last to the first (the index is built most-significant first); the loop is unrolled
             to the largest dimension so that pt[] is indexed by constants and stays in registers */
#pragma unroll
          for (int i = 0; i < 8; i++) {
            if (i >= dim) break;
            const int o = dim - 1 - i;
            int x = a[o] - bmin;
            if (bdelta != 1) {
              /* C division (truncating) by the book's step: exact reciprocal multiply while the value is small */
              const int num = x + (bdelta >> 1), an = abs(num);
              if (brcp && an < (1 << 21)) { const int q = (int)__umulhi((unsigned)an, brcp); x = num < 0 ? -q : q; }
              else x = num / bdelta;
            }
            const int m = (x < ze) ? ((ze - x) << 1) - 1 : ((x - ze) << 1);
            index = index * qv + (m < 0 ? 0 : (m >= qv ? qv - 1 : m));
            pt[i] = x * bdelta + bmin;
            if (m < 0 || m >= qv || pt[i] != a[o]) on_lattice = false;
          }
          if (__ldg(B.len + index) == 0) {
            if (B.nearest && on_lattice) {
              /* the exhaustive search of res0.c:343-370 depends only on the lattice point: table lookup */
              const int u = __ldg(B.nearest + index);
              const int *ev = B.used_vals + (size_t)u * 8;
              for (int i = 0; i < dim; i++) a[i] -= __ldg(ev + i);
              e_out[t] = (short)__ldg(B.used_list + u);
            } else fb = true;
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++) { if (i >= dim) break; a[dim - 1 - i] -= pt[i]; }
            e_out[t] = (short)index;
          }
        }
        /* entries without a codeword: warp-wide exhaustive search over the used entries, one vector at a time */
        unsigned fm = __ballot_sync(FULLMASK, fb);
        while (fm) {
          const int src = __ffs(fm) - 1; fm &= fm - 1;
          const int tv = __shfl_sync(FULLMASK, t, src), sbk = __shfl_sync(FULLMASK, bk, src);
          const DevBook &SB = books[sbk];
          const int sdim = SB.dim;
          int *a = sh.vec[warp][src / RF_LP] + tv * sdim;
          int bestd = 0x7fffffff, besti = 0x7fffffff;
          for (int u = lane; u < SB.used; u += 32) {
            const int *ev = SB.used_vals + (size_t)u * 8;
            int d = 0;
            for (int j = 0; j < sdim; j++) { const int df = __ldg(ev + j) - a[j]; d += df * df; }
            if (d < bestd) { bestd = d; besti = u; }      /* u ascending per lane: first minimum kept */
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const int od = __shfl_xor_sync(FULLMASK, bestd, o), oi = __shfl_xor_sync(FULLMASK, besti, o);
            if (od < bestd || (od == bestd && oi < besti)) { bestd = od; besti = oi; }
          }
          __syncwarp();
          const int ps = __shfl_sync(FULLMASK, p, src), rs2 = __shfl_sync(FULLMASK, row, src);
          short *eo = ent + (((size_t)s * partvals + ps) * rows + rs2) * maxvec;
          if (besti != 0x7fffffff) {
            if (lane < sdim) a[lane] -= __ldg(SB.used_vals + (size_t)besti * 8 + lane);
            if (lane == 0) eo[tv] = (short)__ldg(SB.used_list + besti);
          } else if (lane == 0) eo[tv] = -1;
          __syncwarp();
        }
      }
      __syncwarp();
    }
    /* res1 leaves the cascade residual in place (res2 worked on a private interleaved copy) */
    if (rtype != 2 && valid) for (int t = l; t < G; t += RF_LP) base[(size_t)rowch[row] * n + off + t] = v[t];
    __syncwarp();
  }
  __syncthreads();

  /* ---- pass 2: scatter in emission order, one thread per (stage, partition, row) slot -- most slots hold nothing (the class has no book at that
     stage) and the others 4..16 symbols, too few to give a warp */
  int *out = syms + (size_t)blk * max_syms * 2;
  for (int sl = tid; sl < nslots; sl += RF_WARPS * 32) {
    const int s = sl / prow, pr = sl - s * prow;
    const int p = rows == 1 ? pr : pr / rows, row = rows == 1 ? 0 : pr % rows;
    int o = offs[sl];
    if (s == 0 && row == 0 && (pd_mask >= 0 ? (p & pd_mask) == 0 : (p % pd) == 0)) {
      for (int r2 = 0; r2 < rows; r2++) {
        const int val = phrase_val(r2, p);
        if (val < phrase_entries) { if (o < max_syms) { out[o * 2] = R.groupbook; out[o * 2 + 1] = val; } o++; }
      }
    }
    const int cls = pw[row * partvals + p];
    const int bk = tab_bk[cls * 8 + s];
    if (bk < 0) continue;
    const int nvec = tab_nvec[cls * 8 + s];
    const short *e_in = ent + (((size_t)s * partvals + p) * rows + row) * maxvec;
    for (int t = 0; t < nvec; t++) if (o + t < max_syms) *reinterpret_cast<int2 *>(out + (size_t)(o + t) * 2) = make_int2(bk, e_in[t]);
  }
}

}  // namespace

int vb200_launch_residue_class(const vb200_setup *s, cudaStream_t st, int W, int blocks, int resnum, int ch, const int *chlist,
                               int tot_ch, const int *in, const int *nonzero, int *partword, int *classified) {
  if (blocks <= 0) return VB200_OK;
  const int n = s->h.bs[W] / 2;
  if (s->h.res[resnum].type == 2 && ch == 2) k_residue_class<2><<<blocks, 64, 0, st>>>(s->d, resnum, n, ch, chlist, tot_ch, in, nonzero, partword, classified);
  else if (s->h.res[resnum].type == 2 && ch == 5) k_residue_class<5><<<blocks, 64, 0, st>>>(s->d, resnum, n, ch, chlist, tot_ch, in, nonzero, partword, classified);
  else k_residue_class<0><<<blocks, 64, 0, st>>>(s->d, resnum, n, ch, chlist, tot_ch, in, nonzero, partword, classified);
  LAUNCH_CHECK();
  return VB200_OK;
}

int vb200_residue_maxvec(const vb200_setup *s, int resnum) {
  const DevResidue &R = s->h.res[resnum];
  int mindim = 1 << 30;
  for (int c = 0; c < R.partitions; c++)
    for (int k = 0; k < 8; k++) { const int b = R.partbooks[c * 8 + k]; if (b >= 0 && s->book[b].dim < mindim) mindim = s->book[b].dim; }
  if (mindim == (1 << 30)) return 1;
  return R.grouping / mindim;
}

int vb200_launch_residue_forward(const vb200_setup *s, cudaStream_t st, int W, int blocks, int resnum, int ch, const int *chlist,
                                 int tot_ch, int *in, const int *nonzero, const int *partword, int *syms, int max_syms, int *nsyms) {
  if (blocks <= 0) return VB200_OK;
  const int n = s->h.bs[W] / 2;
  const DevResidue &R = s->h.res[resnum];
  if (R.grouping > RF_MAXG) { vb200_set_error("residue %d: partition size %d not built", resnum, R.grouping); return VB200_EIMPL; }
  const int partvals = (R.end - R.begin) / R.grouping;
  const int rows = R.type == 2 ? 1 : ch;
  const int maxvec = vb200_residue_maxvec(s, resnum);
  const size_t ent_bytes = ((size_t)R.stages * partvals * rows * maxvec * 2 + 15) & ~(size_t)15;
  const size_t dyn = ent_bytes + sizeof(int) * ((size_t)R.stages * partvals * rows + 1);
  if (dyn > 200 * 1024) { vb200_set_error("residue %d: symbol table does not fit shared memory", resnum); return VB200_EIMPL; }
#define RF_LAUNCH(C_) do { CUDA_TRY(cudaFuncSetAttribute(k_residue_forward<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
    k_residue_forward<C_><<<blocks, RF_WARPS * 32, dyn, st>>>(s->d, resnum, n, ch, chlist, tot_ch, in, nonzero, partword, maxvec, syms, max_syms, nsyms); } while (0)
  if (R.type == 2 && ch == 2) RF_LAUNCH(2);
  else if (R.type == 2 && ch == 5) RF_LAUNCH(5);
  else RF_LAUNCH(0);
#undef RF_LAUNCH
  LAUNCH_CHECK();
  return VB200_OK;
}
