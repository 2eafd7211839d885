/* K8 -- device bit packing: per block, turn floor values + residue symbols into the final packet.
 *
 * Replaces the emitting side of mapping0_forward (mapping0.c:1209-1218), floor1_encode
 * (floor1.c:850-942) and vorbis_book_encode / vorbis_oggpack_write (codebook.c:80-110,402-410):
 * SURVEY.md section 8(f1), the first "next" row.  The host path in emit.cu stays as the reference-order
 * implementation the parity tests compare this kernel against.
 *
 * B200 mapping.  One CTA per block.  Codeword lengths are looked up per symbol, an exclusive
 * CTA-wide scan gives every symbol its bit offset, and symbols OR themselves into the packet
 * image in shared memory (LSb first, a symbol touches at most two 32-bit words).  The finished
 * image is copied out with coalesced word stores into a packet pool; the only D2H traffic of the
 * whole path is then the packets themselves (~17 KB per stereo audio-second at q4).
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

constexpr int PK_T = 128;
constexpr int PK_WORDS_BIG = 4096;    /* 16 KB packet image: far above any block the setups produce */
constexpr int PK_WORDS_SMALL = 1024;  /* 4 KB image of the first pass over mono/stereo blocks (16 CTAs per SM instead of 11: the kernel is
                                         latency-bound); a block that does not fit is marked and packed again with the big image */
constexpr int PK_ZERO0 = 512;         /* words cleared up front; more are cleared on demand as the packet grows */
constexpr long long PK_RETRY = -2;    /* pk_off of a block the small image could not hold */
constexpr int PK_PARTS = 31;          /* floor1 partitions (5 bits in the header) */
constexpr int PK_PSYM = 9;            /* symbols of one floor partition: class word + up to 8 values */
constexpr int PK_PER = 4;             /* residue symbols per thread and pass */

template <int CH, int PK_WORDS>
struct PackSmem {
  unsigned words[PK_WORDS];
  unsigned pcode[CH][PK_PARTS][PK_PSYM];
  unsigned char plen[CH][PK_PARTS][PK_PSYM + 3];
  int pbits[CH][PK_PARTS], pbase[CH][PK_PARTS];
  int warp_tot[2][PK_T / 32];
  int base, overflow, limit_bits;      /* limit_bits: PK_WORDS * 32, or less when a test forces the second pass */
};

template <int CH, int PK_WORDS>
__device__ __forceinline__ void put_bits(PackSmem<CH, PK_WORDS> &sm, int bitpos, unsigned value, int len) {
  if (len <= 0) return;
  if (bitpos + len > sm.limit_bits) { sm.overflow = 1; return; }
  if (len < 32) value &= (1u << len) - 1u;
  const int w = bitpos >> 5, sh = bitpos & 31;
  atomicOr(&sm.words[w], value << sh);
  if (sh + len > 32) atomicOr(&sm.words[w + 1], value >> (32 - sh));
}

__device__ __forceinline__ int ilog_(unsigned v) { return 32 - __clz(v); }

template <int CH, int PK_WORDS, bool RETRY, int CHC, int SUBC>
__device__ __forceinline__ void
pack_block(const int blk, const DevSetup *__restrict__ S, int W, int ch_, const int *__restrict__ lWv, const int *__restrict__ nWv,
           const int *__restrict__ vals, const int *__restrict__ floor_nonzero,
           int submaps_, const int *__restrict__ syms, const int *__restrict__ nsyms, int cap_blocks, int max_syms,
           unsigned char *__restrict__ pool, unsigned long long pool_cap, unsigned long long *__restrict__ pool_top,
           long long *__restrict__ pk_off, int *__restrict__ pk_bytes, int limit_words) {
  __shared__ PackSmem<CH, PK_WORDS> sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = CHC ? CHC : ch_, submaps = SUBC ? SUBC : submaps_;      /* channels and submaps as compile-time facts (2, 1: stereo; 6, 2: 5.1): loops unroll */
  const DevBook *books = S->book;
  for (int i = tid; i < PK_ZERO0; i += PK_T) sm.words[i] = 0;
  int zeroed = PK_ZERO0;               /* uniform: words [0, zeroed) are clear or already carry bits */
  if (tid == 0) { sm.overflow = 0; sm.limit_bits = (RETRY && limit_words > 0 && limit_words < PK_WORDS ? limit_words : PK_WORDS) * 32; }
  /* make sure the image is clear up to bit `endbit` before anything is OR-ed there; callers are uniform */
  auto ensure = [&](int endbit) {
    int need = (endbit >> 5) + 2;
    if (need > PK_WORDS) need = PK_WORDS;
    if (need > zeroed) {
      const int upto = min(PK_WORDS, max(need, zeroed + PK_ZERO0));
      for (int i = zeroed + tid; i < upto; i += PK_T) sm.words[i] = 0;
      zeroed = upto;
      __syncthreads();
    }
  };

  /* floor symbols (floor1.c:850-942): one thread per (channel, partition) looks its <= 9 codewords up; the
     bit offsets of the partitions follow from a short serial prefix */
  for (int idx = tid; idx < ch * PK_PARTS; idx += PK_T) {
    const int c = idx / PK_PARTS, i = idx - c * PK_PARTS, it = blk * ch + c;
    const DevFloor &F = S->flr[S->floor_of_ch[W][c]];
    int bits = 0, ns = 0;
    if (i < F.partitions && floor_nonzero[it]) {
      const int *v = vals + (size_t)it * VB200_MAX_POSTS;
      int j = 2;
      for (int q = 0; q < i; q++) j += F.class_dim[F.partclass[q]];
      auto add_book = [&](int book, int entry) {
        const DevBook &B = books[book];
        if (entry >= 0 && entry < B.entries) { const int len = __ldg(B.len + entry); sm.pcode[c][i][ns] = __ldg(B.code + entry); sm.plen[c][i][ns] = (unsigned char)len; ns++; bits += len; }
      };
      const int cls = F.partclass[i], cdim = F.class_dim[cls], csubbits = F.class_subs[cls], csub = 1 << csubbits;
      int bookas[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cval = 0, cshift = 0;
      if (csubbits) {
        for (int k = 0; k < cdim; k++) {
          for (int l = 0; l < csub; l++) {
            const int bn = F.class_subbook[cls * 8 + l];
            const int mv = bn < 0 ? 1 : books[bn].entries;
            if (v[j + k] < mv) { bookas[k] = l; break; }
          }
          cval |= bookas[k] << cshift;
          cshift += csubbits;
        }
        add_book(F.class_book[cls], cval);
      }
      for (int k = 0; k < cdim; k++) {
        const int book = F.class_subbook[cls * 8 + bookas[k]];
        if (book >= 0 && v[j + k] < books[book].entries) add_book(book, v[j + k]);
      }
    }
    for (int k = ns; k < PK_PSYM; k++) sm.plen[c][i][k] = 0;
    sm.pbits[c][i] = bits;
  }
  __syncthreads();
  if (warp == 0) {
    /* bit offsets: header, then per channel [nonzero flag, two first posts, the partitions].  Lane i owns partition i
       (PK_PARTS = 31 <= 32): a warp scan replaces the serial walk over channels x partitions. */
    int pos = 1 + S->modebits + (W ? 2 : 0);
    for (int c = 0; c < ch; c++) {
      const int it = blk * ch + c;
      const DevFloor &F = S->flr[S->floor_of_ch[W][c]];
      const int nzf = floor_nonzero[it];
      const int qb = ilog_((unsigned)(F.quant_q - 1));
      const int hdr = nzf ? 1 + 2 * qb : 1;
      const int mine = lane < PK_PARTS ? sm.pbits[c][lane] : 0;
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += y; }
      if (lane < PK_PARTS) sm.pbase[c][lane] = pos + hdr + incl - mine;
      if (lane == 0) {
        if (c == 0) {
          put_bits(sm, 0, 0, 1);                                          /* packet type: audio */
          put_bits(sm, 1, (unsigned)W, S->modebits);                      /* mode number */
          if (W) { put_bits(sm, 1 + S->modebits, (unsigned)lWv[blk * ch], 1); put_bits(sm, 2 + S->modebits, (unsigned)nWv[blk * ch], 1); }
        }
        if (!nzf) put_bits(sm, pos, 0, 1);
        else {
          const int *v = vals + (size_t)it * VB200_MAX_POSTS;
          put_bits(sm, pos, 1, 1);
          put_bits(sm, pos + 1, (unsigned)v[0], qb);
          put_bits(sm, pos + 1 + qb, (unsigned)v[1], qb);
        }
      }
      pos += hdr + __shfl_sync(FULLMASK, incl, 31);
    }
    if (lane == 0) sm.base = pos;
  }
  __syncthreads();
  int base = sm.base;
  ensure(base);
  for (int idx = tid; idx < ch * PK_PARTS; idx += PK_T) {
    const int c = idx / PK_PARTS, i = idx - c * PK_PARTS;
    int pos = sm.pbase[c][i];
#pragma unroll
    for (int k = 0; k < PK_PSYM; k++) { const int len = sm.plen[c][i][k]; put_bits(sm, pos, sm.pcode[c][i][k], len); pos += len; }
  }

  /* residue symbols, submap by submap, in the emission order K7 produced; PK_PER consecutive symbols per
     thread, one barrier per pass (the warp totals are double-buffered, the running base is uniform) */
  int flip = 0;
  for (int smi = 0; smi < submaps; smi++) {
    const int n = nsyms[(size_t)smi * cap_blocks + blk];
    const int2 *sy = reinterpret_cast<const int2 *>(syms + ((size_t)smi * cap_blocks + blk) * max_syms * 2);
    for (int c0 = 0; c0 < n; c0 += PK_T * PK_PER) {
      const int i0 = c0 + tid * PK_PER;
      unsigned code[PK_PER]; int len[PK_PER];
      int mine = 0;
#pragma unroll
      for (int q = 0; q < PK_PER; q++) {
        code[q] = 0; len[q] = 0;
        if (i0 + q < n) {
          const int2 be = __ldg(sy + i0 + q);
          const DevBook &B = books[be.x];
          if (be.y >= 0 && be.y < B.entries) { code[q] = __ldg(B.code + be.y); len[q] = __ldg(B.len + be.y); }
        }
        mine += len[q];
      }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += y; }
      if (lane == 31) sm.warp_tot[flip][warp] = incl;
      __syncthreads();
      int wbase = 0, tot = 0;
#pragma unroll
      for (int w2 = 0; w2 < PK_T / 32; w2++) { const int t = sm.warp_tot[flip][w2]; if (w2 < warp) wbase += t; tot += t; }
      flip ^= 1;
      ensure(base + tot);
      int pos = base + wbase + incl - mine;
#pragma unroll
      for (int q = 0; q < PK_PER; q++) { put_bits(sm, pos, code[q], len[q]); pos += len[q]; }
      base += tot;
    }
  }
  __syncthreads();

  /* publish */
  __shared__ unsigned long long s_off;
  const int bits = base, bytes = (bits + 7) >> 3, words = (bytes + 3) >> 2;
  if (tid == 0) {
    unsigned long long off = sm.overflow ? ~0ull : atomicAdd(pool_top, (unsigned long long)words * 4ull);
    if (off + (unsigned long long)words * 4ull > pool_cap || sm.overflow) { off = ~0ull; }
    s_off = off;
    pk_off[blk] = (off == ~0ull) ? ((RETRY && sm.overflow) ? PK_RETRY : -1ll) : (long long)off;
    pk_bytes[blk] = bytes;
  }
  __syncthreads();
  if (s_off != ~0ull) {
    unsigned *dst = reinterpret_cast<unsigned *>(pool + s_off);
    for (int i = tid; i < words; i += PK_T) dst[i] = sm.words[i];
  }
}

#define PK_PARAMS const DevSetup *__restrict__ S, int W, int ch, const int *__restrict__ lWv, const int *__restrict__ nWv, \
       const int *__restrict__ vals, const int *__restrict__ floor_nonzero, \
       int submaps, const int *__restrict__ syms, const int *__restrict__ nsyms, int cap_blocks, int max_syms, \
       unsigned char *__restrict__ pool, unsigned long long pool_cap, unsigned long long *__restrict__ pool_top, \
       long long *__restrict__ pk_off, int *__restrict__ pk_bytes, int limit_words
#define PK_ARGS S, W, ch, lWv, nWv, vals, floor_nonzero, submaps, syms, nsyms, cap_blocks, max_syms, pool, pool_cap, pool_top, pk_off, pk_bytes, limit_words

template <int CH, int PK_WORDS, bool RETRY, int CHC, int SUBC>
__global__ void __launch_bounds__(PK_T)
k_pack(PK_PARAMS) {
  pack_block<CH, PK_WORDS, RETRY, CHC, SUBC>(blockIdx.x, PK_ARGS);
}

/* second pass: a CTA looks at 32 blocks and packs the ones the small image could not hold (normally none) */
template <int CH>
__global__ void __launch_bounds__(PK_T)
k_pack_retry(int blocks, PK_PARAMS) {
  const int b0 = blockIdx.x * 32;
  for (int k = 0; k < 32 && b0 + k < blocks; k++) {
    if (pk_off[b0 + k] != PK_RETRY) continue;             /* uniform across the CTA */
    __syncthreads();                                      /* the previous block's image is no longer read */
    pack_block<CH, PK_WORDS_BIG, false, 0, 0>(b0 + k, PK_ARGS);
  }
}

}  // namespace

int vb200_launch_pack(const vb200_setup *s, cudaStream_t st, int W, int blocks, const int *lW, const int *nW,
                      const int *vals, const int *floor_nonzero, int submaps, const int *syms, const int *nsyms,
                      int cap_blocks, int max_syms, unsigned char *pool, unsigned long long pool_cap,
                      unsigned long long *pool_top, long long *pk_off, int *pk_bytes) {
  if (blocks <= 0) return VB200_OK;
#define PK_CALL s->d, W, s->h.channels, lW, nW, vals, floor_nonzero, submaps, syms, nsyms, cap_blocks, max_syms, pool, pool_cap, pool_top, pk_off, pk_bytes, limit
  static const bool big_only = getenv("VB200_PACK_BIG") != nullptr;
  static const int limit = getenv("VB200_PACK_LIMIT") ? atoi(getenv("VB200_PACK_LIMIT")) : 0;    /* tests: words the first pass may fill */
  if (s->h.channels == 2 && submaps == 1 && !big_only) {
    k_pack<2, PK_WORDS_SMALL, true, 2, 1><<<blocks, PK_T, 0, st>>>(PK_CALL);
    LAUNCH_CHECK();
    k_pack_retry<2><<<(blocks + 31) / 32, PK_T, 0, st>>>(blocks, PK_CALL);
  } else if (s->h.channels <= 2 && !big_only) {
    k_pack<2, PK_WORDS_SMALL, true, 0, 0><<<blocks, PK_T, 0, st>>>(PK_CALL);
    LAUNCH_CHECK();
    k_pack_retry<2><<<(blocks + 31) / 32, PK_T, 0, st>>>(blocks, PK_CALL);
  } else if (s->h.channels <= 2)
    k_pack<2, PK_WORDS_BIG, false, 0, 0><<<blocks, PK_T, 0, st>>>(PK_CALL);
  else if (s->h.channels == 6 && submaps == 2)
    k_pack<VB200_MAX_CHANNELS, PK_WORDS_BIG, false, 6, 2><<<blocks, PK_T, 0, st>>>(PK_CALL);
  else
    k_pack<VB200_MAX_CHANNELS, PK_WORDS_BIG, false, 0, 0><<<blocks, PK_T, 0, st>>>(PK_CALL);
#undef PK_CALL
  LAUNCH_CHECK();
  return VB200_OK;
}
