/* The batched whole path: vb200_stream / vb200_batch.
 *
 * Replaces `vorbis_analysis` -> `_mapping_P[0]->forward` (analysis.c:29-63,
 * mapping0.c:738-1322), `vorbis_bitrate_addblock` (bitrate.c:73-227) and
 * `vorbis_bitrate_flushpacket` (bitrate.c:229-252) for many blocks of many independent
 * streams per call.
 *
 * Data layout in HBM.  Blocks are split by size class (W); inside a class every per-bin
 * vector is a dense [block][channel][N2] array (one cudaMalloc'd arena per batch, carved once),
 * so that every kernel reads and writes unit-stride rows.  Cross-block state:
 *   lastmdct  -- never materialised: a consumer gets the address of the previous block's
 *                logmdct row (or of the stream's saved state) plus a resampling rule;
 *   tempmdct  -- lives in the per-stream state slot; runs of consecutive impulse blocks form
 *                chains that one CTA walks in order;
 *   ampmax / lowcomp -- a per-stream scan kernel between K1 and K2/K3;
 *   lW_block_mode, lW_no, impadnum -- pure functions of the block-type sequence: host.
 * The host then emits the codewords (emit.cu) and runs the bitrate manager.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <chrono>
#include <algorithm>
#include <atomic>
#include <thread>
#include <mutex>
#include "vb200_internal.h"
#include "dev_util.cuh"
#include "emit.h"

#define PIPE_MAX_STREAM_SLOTS 4096
#define PIPE_SYMS_PER_ITEM_LONG 1536   /* worst case q10: 64 partitions x 16 vectors x 3 stages / 2 ch */
#define PIPE_SYMS_PER_ITEM_SHORT 256

/* ------------------------------------------------------------------ per-setup stream state pool */
struct StatePool {
  float *last;    /* [slots][ch][2048]  private_state.mblock  */
  float *temp;    /* [slots][ch][256]   private_state.tblock  */
  float *temp2;   /* managed mode: where K4 leaves tblock until the floor-fit gates of the batch have settled */
  float *scal;    /* [slots][2+ch]      g->ampmax, vbi->ampmax, lownoise_compand_level[ch] */
  unsigned char used[PIPE_MAX_STREAM_SLOTS];
};

/* the reservoirs of the bitrate manager (bitrate.h:26-38): plain data, so that the same code walks it on the
   host (host-emission debug path) and on the device (k_bitrate) */
struct BitrateState { long avg_reservoir, minmax_reservoir, avg_bitsper, min_bitsper, max_bitsper, short_per_long; double avgfloat; };
struct BitrateParams { long reservoir_bits, rate; double bias, slew_damp; int bs[2]; };

struct vb200_stream {
  vb200_setup *s;
  int slot;
  int lW_block_mode, lW_no, impadnum;      /* codec_internal.h:89-91 */
  bool started;
  /* lastmdct source for the next block: in the state slot until a batch leaves a newer one there */
  /* bitrate manager (bitrate.h:26-38) */
  int managed;
  BitrateState br;
  int64_t sequence;
};

static StatePool *pool_of(vb200_setup *s);

/* ------------------------------------------------------------------ small kernels */
namespace {

struct ScanBlock { int W, first_item, block_mode, lW_block_mode, psy; };

/* per-stream sequential pass between K1 and K2/K3: the ampmax max-plus recurrence
   (block.c:649-651, psy.c:4504-4515, mapping0.c:887-888,1184) and lb_loudnoise_fix (psy.c:5152-5180).
   One warp per stream: lanes fetch 32 blocks' inputs in parallel (and evaluate the rare loud-noise
   mean for their own block), then the recurrence itself is replayed in block order through shuffles --
   the float adds of the decay must happen one after the other to round like the reference. */
__global__ void __launch_bounds__(128)
k_stream_scan(const DevSetup *__restrict__ S, int nstreams, const int *__restrict__ slot,
              const int *__restrict__ blk_start, const ScanBlock *__restrict__ blks,
              float *__restrict__ scal, const float *const *__restrict__ local_ampmax,
              float *const *__restrict__ global_ampmax, const float *const *__restrict__ logmdct,
              float *const *__restrict__ lowcomp) {
  const int lane = threadIdx.x & 31;
  const int sidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sidx >= nstreams) return;
  const int ch = S->channels;
  float *st = scal + (size_t)slot[sidx] * (2 + VB200_MAX_CHANNELS);
  float g = st[0], vbi = st[1];
  float lv[VB200_MAX_CHANNELS];
#pragma unroll
  for (int c = 0; c < VB200_MAX_CHANNELS; c++) lv[c] = (c < ch) ? st[2 + c] : 0.f;
  const int b_end = blk_start[sidx + 1];
  for (int b0 = blk_start[sidx]; b0 < b_end; b0 += 32) {
    const int b = b0 + lane;
    const bool have = b < b_end;
    ScanBlock B; B.W = 0; B.first_item = 0; B.block_mode = 0; B.lW_block_mode = 0; B.psy = 0;
    if (have) B = blks[b];
    const int N2 = S->bs[B.W] / 2;
    float lmax = -3.0e38f;                 /* max over the block's channels of local_ampmax */
    float decay = 0.f;
    /* per-channel loud-noise decision of this lane's block: 0 carry, 1 force -1, 2 new value in nv[] */
    int kind[VB200_MAX_CHANNELS]; float nv[VB200_MAX_CHANNELS];
#pragma unroll
    for (int c = 0; c < VB200_MAX_CHANNELS; c++) { kind[c] = 0; nv[c] = 0.f; }
    if (have) {
      for (int c = 0; c < ch; c++) { const float la = local_ampmax[B.W][B.first_item + c]; if (la > lmax) lmax = la; }
      decay = __fmul_rn(__fdiv_rn((float)N2, (float)S->rate), S->ampmax_att_per_sec);
      const DevPsy &P = S->psy[B.psy];
      const bool off = (P.m_val < 0.5f) || (P.normal_thresh > .45);
      const bool recompute = (B.block_mode == 2 && B.lW_block_mode == 3) || (B.block_mode == 3 && B.lW_block_mode == 2);
#pragma unroll
      for (int c = 0; c < VB200_MAX_CHANNELS; c++) {
        if (c >= ch) continue;
        if (off) kind[c] = 1;
        else if (recompute) {
          const float *lm = logmdct[B.W] + (size_t)(B.first_item + c) * N2;
          double acc = 0;
          for (int i = P.n25p; i < P.n75p; i++) { const float v = lm[i]; acc += (v > -130.f) ? (double)v : -130.; }
          acc /= P.n;
          kind[c] = 2;
          nv[c] = (acc > -40.) ? -1.f : ((acc < -50.) ? 1.f : (float)(1. - ((acc + 50) / 10)));
        }
      }
    }
    float my_gl = 0.f;
    float my_lv[VB200_MAX_CHANNELS];
#pragma unroll
    for (int c = 0; c < VB200_MAX_CHANNELS; c++) my_lv[c] = 0.f;
    const int cnt = min(32, b_end - b0);
    /* loud-noise level: a block either carries the previous value, forces -1, or (rare: a long<->transition
       change at q<=2) sets a new one.  Without a "new value" in these 32 blocks the carried level is known from two
       ballots; only otherwise is it replayed block by block. */
    bool slow = false;
#pragma unroll
    for (int c = 0; c < VB200_MAX_CHANNELS; c++) {
      if (c >= ch) continue;
      const unsigned m1 = __ballot_sync(FULLMASK, kind[c] == 1), m2 = __ballot_sync(FULLMASK, kind[c] == 2);
      if (m2) { slow = true; continue; }
      my_lv[c] = (m1 & (0xffffffffu >> (31 - lane))) ? -1.f : lv[c];
      if (m1) lv[c] = -1.f;
    }
    for (int t = 0; t < cnt; t++) {
      const float lm_t = __shfl_sync(FULLMASK, lmax, t), d_t = __shfl_sync(FULLMASK, decay, t);
      if (vbi > g) g = vbi;
      g = __fadd_rn(g, d_t);
      if (g < -9999.f) g = -9999.f;
      float gl = g;
      if (lm_t > gl) gl = lm_t;
      vbi = gl;
      if (lane == t) my_gl = gl;
    }
    if (slow) {                                /* uniform across the warp */
      for (int t = 0; t < cnt; t++) {
#pragma unroll
        for (int c = 0; c < VB200_MAX_CHANNELS; c++) {
          if (c >= ch) continue;
          if (!__ballot_sync(FULLMASK, kind[c] == 2)) continue;
          const int k_t = __shfl_sync(FULLMASK, kind[c], t);
          const float n_t = __shfl_sync(FULLMASK, nv[c], t);
          if (k_t == 1) lv[c] = -1.f; else if (k_t == 2) lv[c] = n_t;
          if (lane == t) my_lv[c] = lv[c];
        }
      }
    }
    if (have) {
#pragma unroll
      for (int c = 0; c < VB200_MAX_CHANNELS; c++) {
        if (c >= ch) continue;
        global_ampmax[B.W][B.first_item + c] = my_gl;
        lowcomp[B.W][B.first_item + c] = my_lv[c];
      }
    }
  }
  if (lane == 0) {
    st[0] = g; st[1] = vbi;
    for (int c = 0; c < ch; c++) st[2 + c] = lv[c];
  }
}

/* The same pass for the setups whose lb_loudnoise_fix is switched off for every psy (normal_thresh > .45 or no M
   value, psy.c:5162: every quality >= q3, i.e. all BASELINE configurations): lownoise_compand_level is -1 throughout,
   which leaves the ampmax recurrence.  k_stream_scan pays one chain of dependent global loads per 32 blocks (blocks ->
   local_ampmax), ~2.8 us, 96 times in a 60 s stream.  Here one warp takes a stream in chunks of SC_CHUNK blocks:
   (A) all lanes fetch the chunk's block maxima and decays into shared memory, every load of the chunk in flight at
   once; (B) the recurrence walks the chunk out of shared memory, four blocks per 128-bit access (every lane runs it:
   the values are warp-uniform); (C) all lanes write the results.  0.27 ms -> ~0.05 ms per launch, which an e2e call
   pays once per stream group. */
constexpr int SC_CHUNK = 1024, SC_T = 128;
__global__ void __launch_bounds__(SC_T)
k_stream_scan_ampmax(const DevSetup *__restrict__ S, int nstreams, const int *__restrict__ slot,
                     const int *__restrict__ blk_start, const ScanBlock *__restrict__ blks,
                     float *__restrict__ scal, const float *const *__restrict__ local_ampmax,
                     float *const *__restrict__ global_ampmax, float *const *__restrict__ lowcomp) {
  __shared__ __align__(16) float s_lm[SC_CHUNK], s_d[SC_CHUNK], s_gl[SC_CHUNK];
  const int tid = threadIdx.x, sidx = blockIdx.x;
  if (sidx >= nstreams) return;
  const int ch = S->channels;
  float *st = scal + (size_t)slot[sidx] * (2 + VB200_MAX_CHANNELS);
  float g = st[0], vbi = st[1];
  const float dec0 = __fmul_rn(__fdiv_rn((float)(S->bs[0] / 2), (float)S->rate), S->ampmax_att_per_sec);
  const float dec1 = __fmul_rn(__fdiv_rn((float)(S->bs[1] / 2), (float)S->rate), S->ampmax_att_per_sec);
  const float *la[2] = {local_ampmax[0], local_ampmax[1]};
  float *ga[2] = {global_ampmax[0], global_ampmax[1]}, *lc[2] = {lowcomp[0], lowcomp[1]};
  const int b_end = blk_start[sidx + 1];
  for (int c0 = blk_start[sidx]; c0 < b_end; c0 += SC_CHUNK) {
    const int cnt = min(SC_CHUNK, b_end - c0);
#pragma unroll 8
    for (int j = tid; j < cnt; j += SC_T) {
      const ScanBlock B = blks[c0 + j];
      const float *p = la[B.W] + B.first_item;
      float lmax = -3.0e38f;
      for (int c = 0; c < ch; c++) { const float v = p[c]; if (v > lmax) lmax = v; }
      s_lm[j] = lmax; s_d[j] = B.W ? dec1 : dec0;
    }
    for (int j = cnt + tid; j < ((cnt + 3) & ~3); j += SC_T) { s_lm[j] = 0.f; s_d[j] = 0.f; }
    __syncthreads();
    if (tid < 32) {
      /* block.c:649-651 / mapping0.c:1184 in the reference's order of float operations; fmaxf(vbi, g) is
         `if (vbi > g) g = vbi` (the sign of a zero does not survive the add that follows) */
      auto step = [&](float lm_t, float d_t) {
        g = fmaxf(vbi, g);
        g = __fadd_rn(g, d_t);
        g = fmaxf(g, -9999.f);
        vbi = fmaxf(g, lm_t);
        return vbi;
      };
      const int full = cnt & ~3;
      for (int t = 0; t < full; t += 4) {
        const float4 lm4 = *reinterpret_cast<const float4 *>(s_lm + t), d4 = *reinterpret_cast<const float4 *>(s_d + t);
        float4 o;
        o.x = step(lm4.x, d4.x); o.y = step(lm4.y, d4.y); o.z = step(lm4.z, d4.z); o.w = step(lm4.w, d4.w);
        if (tid == 0) *reinterpret_cast<float4 *>(s_gl + t) = o;
      }
      for (int t = full; t < cnt; t++) { const float o = step(s_lm[t], s_d[t]); if (tid == 0) s_gl[t] = o; }
    }
    __syncthreads();
#pragma unroll 8
    for (int j = tid; j < cnt; j += SC_T) {
      const ScanBlock B = blks[c0 + j];
      const float gl = s_gl[j];
      float *go = ga[B.W] + B.first_item, *lo = lc[B.W] + B.first_item;
      for (int c = 0; c < ch; c++) { go[c] = gl; lo[c] = -1.f; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    st[0] = g; st[1] = vbi;
    for (int c = 0; c < ch; c++) st[2 + c] = -1.f;
  }
}

/* end of batch: leave lastmdct in the stream's state slot exactly as the reference leaves
   private_state.mblock (psy.c:4462-4501) */
__global__ void k_save_last(int n, const float *const *__restrict__ src, const int *__restrict__ rule,
                            const int *__restrict__ srcN, float *const *__restrict__ dst) {
  const int j = blockIdx.x;
  if (j >= n) return;
  const float *s = src[j]; float *d = dst[j];
  const int r = rule[j], N = srcN[j];
  if (r == LAST_COPY) { for (int i = threadIdx.x; i < N; i += blockDim.x) d[i] = s[i]; }
  else if (r == LAST_EXPAND8) { for (int i = threadIdx.x; i < N * 8; i += blockDim.x) d[i] = s[i >> 3]; }
  else if (r == LAST_MINPOOL8) {
    for (int i = threadIdx.x; i < (N >> 3); i += blockDim.x) {
      float m = s[i * 8];
      for (int k = 1; k < 8; k++) { const float v = s[i * 8 + k]; if (m > v) m = v; }
      d[i] = m;
    }
  }
}

/* What the host says about one block (32 bytes; every per-ch-block descriptor array of the kernels is derived from it
   on the device by k_expand_desc -- the host used to fill and ship 92 bytes per ch-block) */
struct BlockDesc {
  long long pcm_off;            /* channel 0 of the block in the PCM timeline */
  int pcm_stride;               /* channel stride */
  int slot;                     /* state slot of the stream */
  int lW_no;                    /* aoTuV history as the block finds it */
  int prev;                     /* previous block of the stream in this batch: (W << 30) | index in its size class; -1: none */
  unsigned char lW, nW, blocktype, lW_block_mode, impadnum, prev_rule, pad0, pad1;
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc layout");

struct ExpandArgs {
  int *psy_item, *psy_blk, *block_mode, *lW_block_mode, *lW, *nW, *params;
  long long *pcm_off, *last_addr, *last_addr2, *temp_off; int *last_rule, *last_rule2;
  const float *logmdct[2]; int N2[2];
  float *last_pool;
  int W, ch, managed, hs, flr_end, nblocks;
};
__global__ void __launch_bounds__(256) k_expand_desc(const BlockDesc *__restrict__ desc, const ExpandArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nblocks * a.ch) return;
  const int idx = t / a.ch, k = t - idx * a.ch;
  const BlockDesc d = desc[idx];
  const int block_mode = d.blocktype | (a.W << 1), psy = d.blocktype + (a.W ? 2 : 0);
  if (k == 0) a.psy_blk[idx] = psy;
  a.psy_item[t] = psy; a.block_mode[t] = block_mode; a.lW_block_mode[t] = d.lW_block_mode; a.lW[t] = d.lW; a.nW[t] = d.nW;
  int4 *pr = reinterpret_cast<int4 *>(a.params + (size_t)t * 8);
  pr[0] = make_int4(1, a.managed, a.flr_end, block_mode);
  pr[1] = make_int4(d.nW, d.lW_block_mode, d.lW_no, d.impadnum);
  a.pcm_off[t] = d.pcm_off + (long long)k * d.pcm_stride;
  a.temp_off[t] = ((long long)d.slot * a.ch + k) * 256;
  float *slot = a.last_pool + ((size_t)d.slot * a.ch + k) * 2048;
  if (!a.hs) { a.last_addr[t] = (long long)(uintptr_t)slot; a.last_rule[t] = LAST_ZERO; }
  else if (d.prev < 0) { a.last_addr[t] = (long long)(uintptr_t)slot; a.last_rule[t] = LAST_COPY; }
  else {
    const int pW = d.prev >> 30, pidx = d.prev & 0x3fffffff;
    a.last_addr[t] = (long long)(uintptr_t)(a.logmdct[pW] + ((size_t)pidx * a.ch + k) * a.N2[pW]);
    a.last_rule[t] = d.prev_rule;
  }
  if (a.last_addr2) { a.last_addr2[t] = 0; a.last_rule2[t] = LAST_NONE; }
}

/* managed mode: copy the tempmdct slots K4 left in the staging pool into the state pool */
__global__ void k_commit_temp(const long long *__restrict__ off, const float *__restrict__ src, float *__restrict__ dst) {
  const long long o = off[blockIdx.x];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) dst[o + i] = src[o + i];
}

__global__ void __launch_bounds__(256) k_copy_i32(const int *__restrict__ src, int *__restrict__ dst, size_t n) {
  /* 16 bytes per access (both pointers are 256-byte aligned arena slices), the odd tail by single ints */
  const size_t n4 = n >> 2, stride = (size_t)gridDim.x * blockDim.x;
  const int4 *s4 = reinterpret_cast<const int4 *>(src);
  int4 *d4 = reinterpret_cast<int4 *>(dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) d4[i] = s4[i];
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
static inline int copy_grid(size_t n) { const size_t g = (n / 4 + 255) / 256; return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g)); }
/* Descriptor upload without the copy engine: the arrays of a batch's host pre-pass lie in page-locked (device-visible)
   memory, and a kernel fetches them all in one launch.  A cudaMemcpyAsync would queue behind the PCM uploads of the
   stream groups that follow (the H2D engine is FIFO), holding back this batch's first kernel by milliseconds. */
constexpr int FETCH_MAX = 48;
struct FetchTable { void *dst[FETCH_MAX]; const void *src[FETCH_MAX]; unsigned bytes[FETCH_MAX]; int n; };
__global__ void __launch_bounds__(256) k_fetch(const FetchTable t) {
  const int e = blockIdx.y;
  const unsigned bytes = t.bytes[e];
  const unsigned n16 = bytes >> 4;
  const int4 *s16 = reinterpret_cast<const int4 *>(t.src[e]); int4 *d16 = reinterpret_cast<int4 *>(t.dst[e]);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) d16[i] = s16[i];
  const int *s4 = reinterpret_cast<const int *>(t.src[e]); int *d4 = reinterpret_cast<int *>(t.dst[e]);
  if (blockIdx.x == 0) for (unsigned i = (n16 << 2) + threadIdx.x; i < (bytes >> 2); i += blockDim.x) d4[i] = s4[i];
}

}  // namespace

/* ------------------------------------------------------------------ arena */
struct Arena {
  unsigned char *base = nullptr; size_t cap = 0, top = 0;
  template <typename T> T *take(size_t n) {
    top = (top + 255) & ~(size_t)255;
    T *p = reinterpret_cast<T *>(base + top);
    top += n * sizeof(T);
    return top <= cap ? p : nullptr;
  }
};

struct ClassBufs {           /* one size class (W) */
  int cap_items = 0, cap_blocks = 0, N2 = 0, nps = 0;
  float *gmdct, *logfft, *logmdct, *noise, *tone, *epeak, *logmask, *npeak, *local_ampmax, *global_ampmax, *lowcomp, *poste;
  float *npeak_work;
  int *posts[3], *fitted[3], *posts_k, *fitted_k, *vals, *iwork, *nonzero, *fit_gate;
  int *psy_item, *psy_blk, *block_mode, *lW_block_mode, *lW, *nW, *params;
  long long *pcm_off, *last_addr, *last_addr2; int *last_rule, *last_rule2;
  BlockDesc *desc;
  int *chain_start, *chain_len, *chain_items; long long *temp_off;
  int *partword, *classified, *syms, *nsyms, *floor_nz;
  int max_syms;
};

struct HostBlock { int stream, W, lW, nW, blocktype, block_mode, psy, idx; int64_t granulepos; int eof; int hist[3]; };

struct vb200_batch {
  vb200_setup *s;
  int max_blocks;
  Arena dev, pin;
  ClassBufs c[2];
  float *d_pcm; size_t pcm_cap;
  ScanBlock *d_scan; int *d_scan_start, *d_slot; float **d_ptrs;
  const float **d_save_src; float **d_save_dst; int *d_save_rule, *d_save_N; int *d_chl;
  long long *d_commit; int link_rounds;      /* managed mode: tempmdct slots to commit; K2/K4 rounds the last run took */
  /* host staging (pinned) mirrors of the descriptor arrays are carved from `pin` per run */
  std::vector<HostBlock> hb;
  /* results */
  BitWriter bw;
  const unsigned char *pk_base;
  int host_emit, host_emit_ready;
  cudaStream_t st;                 /* this batch's own stream: batches and segmenters overlap */
  cudaEvent_t ev_done; bool ev_done_set; vb200_batch *after;   /* device work ordered behind another batch's (vb200_batch_after) */
  bool pending; int run_nblocks, run_nb[2];
  std::vector<vb200_stream *> run_streams; std::vector<long> blob_sizes;
  std::vector<BitWriter> bws; std::vector<std::vector<size_t>> boff;
  unsigned char *d_pool, *h_pool; size_t pool_cap; unsigned long long *d_pool_top, *h_pool_top;
  long long *d_pk_off, *h_pk_off; int *d_pk_bytes, *h_pk_bytes;
  /* managed mode: the device bitrate manager's state, decisions and the pool of chosen packets */
  BitrateState *d_br, *h_br; int *d_choice, *h_choice, *d_final, *h_final; long long *d_out_off, *h_out_off;
  unsigned long long *d_out_top, *h_out_top; unsigned char *d_out_pool; size_t out_cap;
  std::vector<size_t> pk_off, pk_bytes; std::vector<int64_t> pk_gp; std::vector<int> pk_eos, pk_stream;
  /* timing */
  cudaEvent_t ev[24]; int nev; const char *evname[24]; float ev_ms[24]; float total_ms;
  cudaEvent_t ev_mark;             /* vb200_batch_mark: a point on this batch's stream for the caller's own timing */
  /* host copies for vb200_batch_get */
  int n_items[2], n_blocks[2];
};

/* ------------------------------------------------------------------ streams */
/* the pool lives in the setup it belongs to (created with the first stream, freed by vb200_setup_destroy) */
static std::mutex g_pool_mutex;
static StatePool *pool_of(vb200_setup *s) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (s->state_pool) return (StatePool *)s->state_pool;
  StatePool *p = (StatePool *)calloc(1, sizeof(StatePool));
  if (!p) return nullptr;
  const size_t ch = s->h.channels;
  if (cudaMalloc(&p->last, sizeof(float) * PIPE_MAX_STREAM_SLOTS * ch * 2048) != cudaSuccess ||
      cudaMalloc(&p->temp, sizeof(float) * PIPE_MAX_STREAM_SLOTS * ch * 256) != cudaSuccess ||
      cudaMalloc(&p->temp2, sizeof(float) * PIPE_MAX_STREAM_SLOTS * ch * 256) != cudaSuccess ||
      cudaMalloc(&p->scal, sizeof(float) * PIPE_MAX_STREAM_SLOTS * (2 + VB200_MAX_CHANNELS)) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p->last); cudaFree(p->temp); cudaFree(p->temp2); cudaFree(p->scal); free(p);
    return nullptr;
  }
  s->state_pool = p;
  return p;
}
extern "C" void vb200_state_pool_free(void *pool) {
  StatePool *p = (StatePool *)pool;
  if (!p) return;
  cudaFree(p->last); cudaFree(p->temp); cudaFree(p->temp2); cudaFree(p->scal);
  free(p);
}

extern "C" int vb200_stream_create(vb200_setup *s, vb200_stream **out) {
  if (!s || !out) { vb200_set_error("vb200_stream_create: bad arguments"); return VB200_EINVAL; }
  if (s->device < 0 || !s->d) { vb200_set_error("vb200_stream_create: setup has no CUDA device (no CPU fallback)"); return VB200_EFAULT; }
  CUDA_TRY(cudaSetDevice(s->device));
  StatePool *p = pool_of(s);
  if (!p) { vb200_set_error("vb200_stream_create: state pool allocation failed"); return VB200_EFAULT; }
  int slot = -1;
  for (int i = 0; i < PIPE_MAX_STREAM_SLOTS; i++) if (!p->used[i]) { slot = i; break; }
  if (slot < 0) { vb200_set_error("vb200_stream_create: more than %d live streams per setup", PIPE_MAX_STREAM_SLOTS); return VB200_EIMPL; }
  p->used[slot] = 1;
  const size_t ch = s->h.channels;
  cudaStream_t st = (cudaStream_t)s->stream;
  /* everything calloc-zero, ampmax -9999 (block.c:93,196,314-317; psy.c:128) */
  CUDA_TRY(cudaMemsetAsync(p->last + (size_t)slot * ch * 2048, 0, sizeof(float) * ch * 2048, st));
  CUDA_TRY(cudaMemsetAsync(p->temp + (size_t)slot * ch * 256, 0, sizeof(float) * ch * 256, st));
  float init[2 + VB200_MAX_CHANNELS]; init[0] = -9999.f; init[1] = -9999.f;
  for (int c = 0; c < VB200_MAX_CHANNELS; c++) init[2 + c] = 0.f;
  CUDA_TRY(cudaMemcpyAsync(p->scal + (size_t)slot * (2 + VB200_MAX_CHANNELS), init, sizeof(init), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  vb200_stream *x = new vb200_stream();
  memset(x, 0, sizeof(*x));
  x->s = s; x->slot = slot;
  /* vorbis_bitrate_init, bitrate.c:28-61 */
  if (s->bi_i32[3] > 0) {
    const long rate = s->h.rate; const int half = s->h.bs[0] >> 1;
    x->br.short_per_long = s->h.bs[1] / s->h.bs[0];
    x->managed = 1;
    x->br.avg_bitsper = (long)rint(1. * s->bi_i32[0] * half / rate);
    x->br.min_bitsper = (long)rint(1. * s->bi_i32[1] * half / rate);
    x->br.max_bitsper = (long)rint(1. * s->bi_i32[2] * half / rate);
    x->br.avgfloat = VB200_PACKETBLOBS / 2;
    const long fill = (long)(s->bi_i32[3] * s->bi_f64[0]);
    x->br.minmax_reservoir = fill; x->br.avg_reservoir = fill;
  }
  *out = x;
  return VB200_OK;
}

extern "C" void vb200_stream_destroy(vb200_stream *x) {
  if (!x) return;
  StatePool *p = (StatePool *)x->s->state_pool;      /* streams go before their setup (as vorbis_dsp_clear goes before vorbis_info_clear) */
  if (p) p->used[x->slot] = 0;
  delete x;
}

/* ------------------------------------------------------------------ batch */
static bool carve_class(vb200_batch *b, int W, int cap_blocks) {
  const vb200_setup *s = b->s;
  ClassBufs &c = b->c[W];
  const int ch = s->h.channels, N2 = s->h.bs[W] / 2;
  int part = 16;
  for (int k = 0; k < s->h.psys; k++) if (s->h.psy[k].blockflag == W) part = s->h.psy[k].partition;
  c.cap_blocks = cap_blocks; c.cap_items = cap_blocks * ch; c.N2 = N2; c.nps = N2 / part;
  const size_t I = c.cap_items, V = I * N2;
  const int nsel = s->h.managed ? 3 : 1;
  Arena &a = b->dev;
  c.gmdct = a.take<float>(V); c.logfft = a.take<float>(V); c.logmdct = a.take<float>(V); c.noise = a.take<float>(V);
  c.tone = a.take<float>(V); c.epeak = a.take<float>(V); c.logmask = a.take<float>(V * nsel);
  c.npeak = a.take<float>(I * c.nps); c.npeak_work = a.take<float>(I * c.nps);
  c.local_ampmax = a.take<float>(I); c.global_ampmax = a.take<float>(I); c.lowcomp = a.take<float>(I); c.poste = a.take<float>(I);
  for (int k = 0; k < 3; k++) { c.posts[k] = a.take<int>(I * VB200_MAX_POSTS); c.fitted[k] = a.take<int>(I); }
  c.posts_k = a.take<int>(I * VB200_MAX_POSTS); c.fitted_k = a.take<int>(I);
  c.vals = a.take<int>(I * VB200_MAX_POSTS); c.iwork = a.take<int>(V); c.nonzero = a.take<int>(I); c.fit_gate = a.take<int>(I);
  c.psy_item = a.take<int>(I); c.psy_blk = a.take<int>(cap_blocks); c.block_mode = a.take<int>(I); c.lW_block_mode = a.take<int>(I);
  c.lW = a.take<int>(I); c.nW = a.take<int>(I); c.params = a.take<int>(I * 8);
  c.pcm_off = a.take<long long>(I); c.last_addr = a.take<long long>(I); c.last_rule = a.take<int>(I);
  c.last_addr2 = a.take<long long>(I); c.last_rule2 = a.take<int>(I);
  c.desc = a.take<BlockDesc>(cap_blocks);
  c.chain_start = a.take<int>(I); c.chain_len = a.take<int>(I); c.chain_items = a.take<int>(I); c.temp_off = a.take<long long>(I);
  int maxpv = 1;
  for (int r = 0; r < s->h.residues; r++) { const DevResidue &R = s->h.res[r]; const int pv = (R.end - R.begin) / R.grouping; if (pv > maxpv) maxpv = pv; }
  const int nsub = s->h.map[s->h.mode_mapping[W]].submaps;
  c.partword = a.take<int>((size_t)cap_blocks * nsub * ch * maxpv);
  c.classified = a.take<int>((size_t)cap_blocks * nsub);
  c.max_syms = (W ? PIPE_SYMS_PER_ITEM_LONG : PIPE_SYMS_PER_ITEM_SHORT) * ch;
  c.syms = a.take<int>((size_t)cap_blocks * nsub * c.max_syms * 2);
  c.nsyms = a.take<int>((size_t)cap_blocks * nsub);
  c.floor_nz = a.take<int>(I);
  return c.nsyms != nullptr;
}

static size_t class_bytes(const vb200_setup *s, int W, int cap_blocks) {
  const size_t ch = s->h.channels, N2 = s->h.bs[W] / 2, I = (size_t)cap_blocks * ch, V = I * N2;
  const size_t nsel = s->h.managed ? 3 : 1;
  size_t maxpv = 1;
  for (int r = 0; r < s->h.residues; r++) { const DevResidue &R = s->h.res[r]; const size_t pv = (R.end - R.begin) / R.grouping; if (pv > maxpv) maxpv = pv; }
  const size_t max_syms = (size_t)(W ? PIPE_SYMS_PER_ITEM_LONG : PIPE_SYMS_PER_ITEM_SHORT) * ch;
  size_t b = 4 * (V * (7 + nsel) + I * (N2 / 8) * 2 + I * 8 + I * VB200_MAX_POSTS * 5 + I * 32 + cap_blocks * 16)
           + 8 * I * 4 + 4 * ((size_t)cap_blocks * s->h.map[s->h.mode_mapping[W]].submaps * (ch * maxpv + 2 + max_syms * 2));
  return b + 64 * 256;
}

extern "C" int vb200_batch_create(vb200_setup *s, int max_blocks, vb200_batch **out) {
  if (!s || !out || max_blocks < 1) { vb200_set_error("vb200_batch_create: bad arguments"); return VB200_EINVAL; }
  if (s->device < 0 || !s->d) { vb200_set_error("vb200_batch_create: setup has no CUDA device (no CPU fallback)"); return VB200_EFAULT; }
  CUDA_TRY(cudaSetDevice(s->device));
  vb200_batch *b = new vb200_batch();
  b->s = s; b->max_blocks = max_blocks; b->d_pcm = nullptr; b->pcm_cap = 0; b->nev = 0; b->total_ms = 0;
  const size_t bytes = class_bytes(s, 0, max_blocks) + class_bytes(s, 1, max_blocks) + sizeof(ScanBlock) * (size_t)max_blocks
                     + (size_t)PIPE_MAX_STREAM_SLOTS * (8 + VB200_MAX_CHANNELS * 64) + 65536;
  if (cudaMalloc(&b->dev.base, bytes) != cudaSuccess) { cudaGetLastError(); vb200_set_error("vb200_batch_create: cannot allocate %zu bytes of HBM", bytes); delete b; return VB200_EFAULT; }
  b->dev.cap = bytes;
  b->d_scan = b->dev.take<ScanBlock>(max_blocks); b->d_scan_start = b->dev.take<int>(PIPE_MAX_STREAM_SLOTS + 1); b->d_slot = b->dev.take<int>(PIPE_MAX_STREAM_SLOTS);
  b->d_ptrs = b->dev.take<float *>(8);
  /* two groups of state writes per (stream, channel): a 1024-entry writer, then a 128-entry writer on top of it */
  b->d_save_src = b->dev.take<const float *>((size_t)2 * PIPE_MAX_STREAM_SLOTS * VB200_MAX_CHANNELS); b->d_save_dst = b->dev.take<float *>((size_t)2 * PIPE_MAX_STREAM_SLOTS * VB200_MAX_CHANNELS);
  b->d_save_rule = b->dev.take<int>((size_t)2 * PIPE_MAX_STREAM_SLOTS * VB200_MAX_CHANNELS); b->d_save_N = b->dev.take<int>((size_t)2 * PIPE_MAX_STREAM_SLOTS * VB200_MAX_CHANNELS);
  b->d_commit = b->dev.take<long long>((size_t)PIPE_MAX_STREAM_SLOTS * VB200_MAX_CHANNELS); b->link_rounds = 0;
  b->d_chl = b->dev.take<int>(2 * VB200_MAX_SUBMAPS * VB200_MAX_CHANNELS);
  if (!b->d_chl || !carve_class(b, 0, max_blocks) || !carve_class(b, 1, max_blocks)) { vb200_set_error("vb200_batch_create: arena sizing bug"); cudaFree(b->dev.base); delete b; return VB200_EFAULT; }
  /* pinned staging for descriptors and results */
  const size_t ch = s->h.channels;
  /* descriptors only; the (large) symbol staging of the host-emission debug path is added on demand */
  const size_t pin_bytes = (size_t)max_blocks * ch * (64 * 4) * 2 + (size_t)PIPE_MAX_STREAM_SLOTS * 64 + (1 << 20);
  if (cudaMallocHost(&b->pin.base, pin_bytes) != cudaSuccess) { cudaGetLastError(); vb200_set_error("vb200_batch_create: cannot allocate pinned staging"); cudaFree(b->dev.base); delete b; return VB200_EFAULT; }
  b->pin.cap = pin_bytes;
  {
    const int nbl = s->h.managed ? VB200_PACKETBLOBS : 1;
    b->pool_cap = (size_t)max_blocks * 3072 * nbl + (1 << 16);
    b->host_emit = 0; b->host_emit_ready = 0; b->pk_base = nullptr;
    const size_t nent = (size_t)nbl * 2 * max_blocks;
    if (cudaMalloc(&b->d_pool, b->pool_cap) != cudaSuccess || cudaMalloc(&b->d_pool_top, 8) != cudaSuccess ||
        cudaMalloc(&b->d_pk_off, sizeof(long long) * nent) != cudaSuccess || cudaMalloc(&b->d_pk_bytes, sizeof(int) * nent) != cudaSuccess ||
        cudaMallocHost(&b->h_pool, b->pool_cap) != cudaSuccess || cudaMallocHost(&b->h_pool_top, 8) != cudaSuccess ||
        cudaMallocHost(&b->h_pk_off, sizeof(long long) * nent) != cudaSuccess || cudaMallocHost(&b->h_pk_bytes, sizeof(int) * nent) != cudaSuccess) {
      cudaGetLastError(); vb200_set_error("vb200_batch_create: cannot allocate the packet pool"); return VB200_EFAULT;
    }
  }
  b->d_br = nullptr; b->h_br = nullptr; b->d_choice = b->h_choice = b->d_final = b->h_final = nullptr; b->d_out_off = b->h_out_off = nullptr;
  b->d_out_top = b->h_out_top = nullptr; b->d_out_pool = nullptr; b->out_cap = 0;
  if (s->h.managed) {
    const size_t n2 = 2 * (size_t)max_blocks;
    b->out_cap = (size_t)max_blocks * 4096 + (1 << 16);
    if (cudaMalloc(&b->d_br, sizeof(BitrateState) * PIPE_MAX_STREAM_SLOTS) != cudaSuccess || cudaMallocHost(&b->h_br, sizeof(BitrateState) * PIPE_MAX_STREAM_SLOTS) != cudaSuccess ||
        cudaMalloc(&b->d_choice, sizeof(int) * n2) != cudaSuccess || cudaMallocHost(&b->h_choice, sizeof(int) * n2) != cudaSuccess ||
        cudaMalloc(&b->d_final, sizeof(int) * n2) != cudaSuccess || cudaMallocHost(&b->h_final, sizeof(int) * n2) != cudaSuccess ||
        cudaMalloc(&b->d_out_off, sizeof(long long) * n2) != cudaSuccess || cudaMallocHost(&b->h_out_off, sizeof(long long) * n2) != cudaSuccess ||
        cudaMalloc(&b->d_out_top, 8) != cudaSuccess || cudaMallocHost(&b->h_out_top, 8) != cudaSuccess || cudaMalloc(&b->d_out_pool, b->out_cap) != cudaSuccess) {
      cudaGetLastError(); vb200_set_error("vb200_batch_create: cannot allocate the bitrate manager's buffers"); return VB200_EFAULT;
    }
  }
  for (int i = 0; i < 24; i++) cudaEventCreate(&b->ev[i]);
  cudaEventCreate(&b->ev_mark);
  cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming); b->ev_done_set = false; b->after = nullptr;
  b->pending = false;
  if (cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking) != cudaSuccess) { vb200_set_error("vb200_batch_create: cannot create a stream"); return VB200_EFAULT; }
  *out = b;
  return VB200_OK;
}

extern "C" void vb200_batch_destroy(vb200_batch *b) {
  if (!b) return;
  cudaSetDevice(b->s->device);
  for (int i = 0; i < 24; i++) cudaEventDestroy(b->ev[i]);
  cudaEventDestroy(b->ev_mark);
  cudaEventDestroy(b->ev_done);
  cudaStreamDestroy(b->st);
  if (b->d_pcm) cudaFree(b->d_pcm);
  cudaFree(b->dev.base);
  cudaFreeHost(b->pin.base);
  cudaFree(b->d_pool); cudaFree(b->d_pool_top); cudaFree(b->d_pk_off); cudaFree(b->d_pk_bytes);
  cudaFreeHost(b->h_pool); cudaFreeHost(b->h_pool_top); cudaFreeHost(b->h_pk_off); cudaFreeHost(b->h_pk_bytes);
  cudaFree(b->d_br); cudaFreeHost(b->h_br); cudaFree(b->d_choice); cudaFreeHost(b->h_choice); cudaFree(b->d_final); cudaFreeHost(b->h_final);
  cudaFree(b->d_out_off); cudaFreeHost(b->h_out_off); cudaFree(b->d_out_top); cudaFreeHost(b->h_out_top); cudaFree(b->d_out_pool);
  delete b;
}

/* ------------------------------------------------------------------ bitrate manager (host) */
/* vorbis_bitrate_addblock, bitrate.c:73-227, restated over the 15 blob sizes of one block.
   Returns the chosen blob and the final byte size (after truncation / zero padding). */
static BitrateParams bitrate_params(const vb200_setup *s) {
  BitrateParams p;
  p.reservoir_bits = s->bi_i32[3]; p.rate = s->h.rate; p.bias = s->bi_f64[0]; p.slew_damp = s->bi_f64[1]; p.bs[0] = s->h.bs[0]; p.bs[1] = s->h.bs[1];
  return p;
}

__host__ __device__ static int bitrate_choose(BitrateState *x, const BitrateParams &P, int W, const long *blob_bytes, long *final_bytes) {
  const long reservoir_bits = P.reservoir_bits; const double bias = P.bias, slew_damp = P.slew_damp;
  int choice = (int)rint(x->avgfloat);
  long this_bits = blob_bytes[choice] * 8;
  const long min_target = W ? x->min_bitsper * x->short_per_long : x->min_bitsper;
  const long max_target = W ? x->max_bitsper * x->short_per_long : x->max_bitsper;
  const int samples = P.bs[W] >> 1;
  const long desired_fill = (long)(reservoir_bits * bias);
  const long rate = P.rate;
  if (x->avg_bitsper > 0) {
    double slew = 0.;
    const long avg_target = W ? x->avg_bitsper * x->short_per_long : x->avg_bitsper;
    const double slewlimit = 15. / slew_damp;
    if (x->avg_reservoir + (this_bits - avg_target) > desired_fill) {
      while (choice > 0 && this_bits > avg_target && x->avg_reservoir + (this_bits - avg_target) > desired_fill) { choice--; this_bits = blob_bytes[choice] * 8; }
    } else if (x->avg_reservoir + (this_bits - avg_target) < desired_fill) {
      while (choice + 1 < VB200_PACKETBLOBS && this_bits < avg_target && x->avg_reservoir + (this_bits - avg_target) < desired_fill) { choice++; this_bits = blob_bytes[choice] * 8; }
    }
    slew = rint(choice - x->avgfloat) / samples * rate;
    if (slew < -slewlimit) slew = -slewlimit;
    if (slew > slewlimit) slew = slewlimit;
    choice = (int)rint(x->avgfloat += slew / rate * samples);
    this_bits = blob_bytes[choice] * 8;
  }
  if (x->min_bitsper > 0 && this_bits < min_target) {
    while (x->minmax_reservoir - (min_target - this_bits) < 0) { choice++; if (choice >= VB200_PACKETBLOBS) break; this_bits = blob_bytes[choice] * 8; }
  }
  if (x->max_bitsper > 0 && this_bits > max_target) {
    while (x->minmax_reservoir + (this_bits - max_target) > reservoir_bits) { choice--; if (choice < 0) break; this_bits = blob_bytes[choice] * 8; }
  }
  long fb;
  if (choice < 0) {
    const long maxsize = (max_target + (reservoir_bits - x->minmax_reservoir)) / 8;
    choice = 0;
    fb = blob_bytes[0];
    if (fb > maxsize) fb = maxsize;             /* oggpack_writetrunc */
  } else {
    long minsize = (min_target - x->minmax_reservoir + 7) / 8;
    if (choice >= VB200_PACKETBLOBS) choice = VB200_PACKETBLOBS - 1;
    fb = blob_bytes[choice];
    if (minsize > fb) fb = minsize;             /* zero padding */
  }
  this_bits = fb * 8;
  if (x->min_bitsper > 0 || x->max_bitsper > 0) {
    if (max_target > 0 && this_bits > max_target) x->minmax_reservoir += (this_bits - max_target);
    else if (min_target > 0 && this_bits < min_target) x->minmax_reservoir += (this_bits - min_target);
    else if (x->minmax_reservoir > desired_fill) {
      if (max_target > 0) { x->minmax_reservoir += (this_bits - max_target); if (x->minmax_reservoir < desired_fill) x->minmax_reservoir = desired_fill; }
      else x->minmax_reservoir = desired_fill;
    } else {
      if (min_target > 0) { x->minmax_reservoir += (this_bits - min_target); if (x->minmax_reservoir > desired_fill) x->minmax_reservoir = desired_fill; }
      else x->minmax_reservoir = desired_fill;
    }
  }
  if (x->avg_bitsper > 0) {
    const long avg_target = W ? x->avg_bitsper * x->short_per_long : x->avg_bitsper;
    x->avg_reservoir += this_bits - avg_target;
  }
  *final_bytes = fb;
  return choice;
}

/* ------------------------------------------------------------------ bitrate manager (device) */
namespace {

/* vorbis_bitrate_addblock for every block of every stream: one thread per stream walks its blocks in order
   (the reservoirs carry from block to block), streams run side by side.  Only the sizes of the 15 packet
   blobs are read; the chosen blob's final size and a slot in the output pool come out. */
__global__ void k_bitrate(BitrateParams P, int nstreams, const int *__restrict__ scan_start, const ScanBlock *__restrict__ scan, int ch,
                          int max_blocks, const long long *__restrict__ pk_off, const int *__restrict__ pk_bytes,
                          BitrateState *__restrict__ state, int *__restrict__ choice, int *__restrict__ final_bytes,
                          long long *__restrict__ out_off, unsigned long long *__restrict__ out_top, unsigned long long out_cap) {
  const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= nstreams) return;
  BitrateState x = state[sidx];
  for (int pos = scan_start[sidx]; pos < scan_start[sidx + 1]; pos++) {
    const int W = scan[pos].W, idx = scan[pos].first_item / ch;
    long sz[VB200_PACKETBLOBS];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < VB200_PACKETBLOBS; k++) {
      const size_t o = ((size_t)k * 2 + W) * max_blocks + idx;
      sz[k] = pk_bytes[o];
      if (pk_off[o] < 0) ok = false;              /* a blob did not fit the packer's image */
    }
    long fb = 0;
    const int c = bitrate_choose(&x, P, W, sz, &fb);
    const size_t o = (size_t)W * max_blocks + idx;
    const unsigned long long need = ((unsigned long long)fb + 3ull) & ~3ull;
    const unsigned long long off = atomicAdd(out_top, need);
    if (off + need > out_cap) ok = false;
    choice[o] = c;
    final_bytes[o] = ok ? (int)fb : -1;
    out_off[o] = (long long)off;
  }
  state[sidx] = x;
}

/* copy the chosen blob of every block into the output pool: truncated (oggpack_writetrunc) or zero-padded to
   the size the manager decided (bitrate.c:178-215) */
__global__ void __launch_bounds__(128)
k_gather_choice(int W, int max_blocks, const int *__restrict__ choice, const int *__restrict__ final_bytes, const long long *__restrict__ out_off,
                const long long *__restrict__ pk_off, const int *__restrict__ pk_bytes, const unsigned char *__restrict__ pool,
                unsigned char *__restrict__ out_pool) {
  const size_t o = (size_t)W * max_blocks + blockIdx.x;
  const int fb = final_bytes[o];
  if (fb < 0) return;
  const size_t src_o = ((size_t)choice[o] * 2 + W) * max_blocks + blockIdx.x;
  const unsigned char *src = pool + pk_off[src_o];
  const int have = pk_bytes[src_o];
  unsigned char *dst = out_pool + out_off[o];
  for (int t = threadIdx.x; t < fb; t += 128) dst[t] = t < have ? src[t] : (unsigned char)0;
}

}  // namespace

/* ------------------------------------------------------------------ the run */
#define EV(name) do{ if (b->nev < 24) { b->evname[b->nev] = name; cudaEventRecord(b->ev[b->nev++], st); } }while(0)
#define RUN(x) do{ int r_ = (x); if (r_) return r_; }while(0)

static int batch_finish(vb200_batch *b);

extern "C" int vb200_batch_submit(vb200_batch *b, vb200_stream **streams, int nstreams,
                                  const vb200_block *blocks, int nblocks,
                                  const float *pcm, size_t pcm_elems, int pcm_on_device) {
  if (!b || !streams || !blocks || !pcm || nstreams < 1 || nblocks < 0) { vb200_set_error("vb200_batch_submit: bad arguments"); return VB200_EINVAL; }
  if (nstreams > PIPE_MAX_STREAM_SLOTS) { vb200_set_error("vb200_batch_run: too many streams"); return VB200_EINVAL; }
  if (nblocks > b->max_blocks) { vb200_set_error("vb200_batch_run: %d blocks > batch capacity %d", nblocks, b->max_blocks); return VB200_EINVAL; }
  vb200_setup *s = b->s;
  CUDA_TRY(cudaSetDevice(s->device));
  cudaStream_t st = b->st;
  if (b->pending) { vb200_set_error("vb200_batch_submit: the previous run has not been collected"); return VB200_EINVAL; }
  StatePool *pool = pool_of(s);
  const int ch = s->h.channels, managed = s->h.managed;
  const int nsel = managed ? 3 : 1;
  b->nev = 0;
  b->pk_off.clear(); b->pk_bytes.clear(); b->pk_gp.clear(); b->pk_eos.clear(); b->pk_stream.clear();
  b->bw.reset();
  b->run_streams.assign(streams, streams + nstreams); b->run_nblocks = nblocks;
  if (nblocks == 0) return VB200_OK;

  /* ---- host pre-pass over the block descriptors ---- */
  const double t_prep0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  b->hb.resize(nblocks);
  int nb[2] = {0, 0};
  b->pin.top = 0;
  struct HostClass { BlockDesc *desc; long long *last_addr, *last_addr2; int *last_rule, *last_rule2, *gate;
                     int *chain_start, *chain_len, *chain_items; int nchains, nchain_items; } hc[2];
  const bool hs = s->h.rate >= 26000;
  const bool gated = managed && hs;      /* managed mode at high sampling rates: the lastmdct links may need re-rounds (below) */
  for (int W = 0; W < 2; W++) {
    const size_t I = (size_t)nblocks * ch;
    memset(&hc[W], 0, sizeof(hc[W]));
    hc[W].desc = b->pin.take<BlockDesc>(nblocks);
    if (gated) {        /* per-ch-block link arrays: only a re-round fills and ships them */
      hc[W].last_addr = b->pin.take<long long>(I); hc[W].last_rule = b->pin.take<int>(I);
      hc[W].last_addr2 = b->pin.take<long long>(I); hc[W].last_rule2 = b->pin.take<int>(I); hc[W].gate = b->pin.take<int>(I);
    }
    if (W == 0) {       /* K4 chains: long blocks are their own chain (no list needed) */
      hc[W].chain_start = b->pin.take<int>(I); hc[W].chain_len = b->pin.take<int>(I); hc[W].chain_items = b->pin.take<int>(I);
    }
    hc[W].nchains = 0; hc[W].nchain_items = 0;
    if (!hc[W].desc || (gated && !hc[W].gate) || (W == 0 && !hc[W].chain_items)) { vb200_set_error("vb200_batch_run: pinned staging too small"); return VB200_EFAULT; }
  }
  ScanBlock *h_scan = b->pin.take<ScanBlock>(nblocks);
  int *h_scan_start = b->pin.take<int>(nstreams + 1), *h_slot = b->pin.take<int>(nstreams);
  const double t_prepA = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  /* (1) serial and light: validate, give every block its index inside its size class, bucket the blocks by stream
     (counting sort: blocks of one stream stay in submission order) */
  std::vector<int> sstart(nstreams + 1, 0), order(nblocks), idx_of(nblocks);      /* compact: this pass is memory-bound */
  for (int i = 0; i < nblocks; i++) {
    const vb200_block &q = blocks[i];
    if (q.stream < 0 || q.stream >= nstreams || (q.W != 0 && q.W != 1) || q.pcm_stride < 0 || q.pcm_stride > 0x7fffffff || q.blocktype < 0 || q.blocktype > 1) {
      vb200_set_error("vb200_batch_run: block %d malformed", i); return VB200_EINVAL; }
    sstart[q.stream + 1]++;
    idx_of[i] = nb[q.W]++;
  }
  const double t_prepB = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  for (int k = 0; k < nstreams; k++) sstart[k + 1] += sstart[k];
  {
    std::vector<int> fill(sstart.begin(), sstart.end() - 1);
    for (int i = 0; i < nblocks; i++) order[fill[blocks[i].stream]++] = i;
  }
  const double t_prep1 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  const int flr_of_end[2] = { s->h.flr[s->h.floor_of_ch[0][0]].info_n, s->h.flr[s->h.floor_of_ch[1][0]].info_n };
  /* (2) per stream, independent of the other streams (history, lastmdct links, K4 chains, scan list): host threads */
  struct StreamChains { std::vector<int> start[2], len[2], items[2]; bool impulse; };
  std::vector<StreamChains> sc(nstreams);
  /* Where every ch-block finds lastmdct, and what the stream's state slot must hold after the batch, given which
     ch-blocks WRITE lastmdct: all of them in VBR mode; in managed mode only those whose nominal floor fit exists
     (mapping0.c:1097 skips the sel = 2 mix otherwise, and only that call writes, psy.c:4165-4173).  A writer whose
     next block is short leaves 128 entries (copy of a short block / 8:1 min-pool of a transition block), every other
     writer 1024 (psy.c:4462-4501); a long reader that finds a 128-entry writer on top takes bins >= 128 from the newest
     1024-entry writer before it (second source of LastRef).  gate[W] == NULL: every block writes. */
  struct SaveOp { const float *src; float *dst; int rule, N; };
  std::vector<std::vector<SaveOp>> saveA(nstreams), saveB(nstreams);
  const int long_extent = s->h.bs[1] / 2;
  auto extent_of = [&](const HostBlock &h) { return (h.block_mode == 3 || h.nW) ? long_extent : LAST_SHORT_EXTENT; };
  auto rule_of = [&](const HostBlock &h) { return h.block_mode <= 1 ? (h.nW ? LAST_EXPAND8 : LAST_COPY) : ((h.block_mode == 2 && !h.nW) ? LAST_MINPOOL8 : LAST_COPY); };
  auto row_of = [&](const HostBlock &h, int k) -> const float * { const ClassBufs &pc = b->c[h.W]; return pc.logmdct + ((size_t)h.idx * ch + k) * pc.N2; };
  auto link_stream = [&](int sidx, int *const gate[2], bool fill_items) {
    vb200_stream *x = streams[sidx];
    const int *lst = order.data() + sstart[sidx];
    const int cnt = sstart[sidx + 1] - sstart[sidx];
    saveA[sidx].clear(); saveB[sidx].clear();
    for (int k = 0; k < ch; k++) {
      float *slot = pool->last + ((size_t)x->slot * ch + k) * 2048;
      int w1 = -1, w2 = -1;          /* newest writer so far; newest writer of 1024 entries (-1: the state slot) */
      for (int j = 0; j < cnt; j++) {
        const HostBlock &h = b->hb[lst[j]];
        HostClass &c = hc[h.W];
        const int it = h.idx * ch + k;
        if (!fill_items) { /* first round: every block writes, the links are the device's (k_expand_desc) */ }
        else if (!hs) { c.last_addr2[it] = 0; c.last_rule2[it] = LAST_NONE; c.last_addr[it] = (long long)(uintptr_t)slot; c.last_rule[it] = LAST_ZERO; }
        else if (w1 < 0) { c.last_addr2[it] = 0; c.last_rule2[it] = LAST_NONE; c.last_addr[it] = (long long)(uintptr_t)slot; c.last_rule[it] = LAST_COPY; }
        else {
          c.last_addr2[it] = 0; c.last_rule2[it] = LAST_NONE;
          const HostBlock &p = b->hb[w1];
          c.last_addr[it] = (long long)(uintptr_t)row_of(p, k); c.last_rule[it] = rule_of(p);
          if (h.W && extent_of(p) < long_extent) {
            if (w2 < 0) { c.last_addr2[it] = (long long)(uintptr_t)slot; c.last_rule2[it] = LAST_COPY; }
            else { const HostBlock &p2 = b->hb[w2]; c.last_addr2[it] = (long long)(uintptr_t)row_of(p2, k); c.last_rule2[it] = rule_of(p2); }
          }
        }
        if (!gated || !gate[h.W] || gate[h.W][it]) { w1 = lst[j]; if (extent_of(h) >= long_extent) w2 = lst[j]; }
      }
      if (hs && w1 >= 0) {
        const HostBlock &p = b->hb[w1];
        if (extent_of(p) >= long_extent) saveA[sidx].push_back(SaveOp{row_of(p, k), slot, rule_of(p), b->c[p.W].N2});
        else {
          if (w2 >= 0) { const HostBlock &p2 = b->hb[w2]; saveA[sidx].push_back(SaveOp{row_of(p2, k), slot, rule_of(p2), b->c[p2.W].N2}); }
          saveB[sidx].push_back(SaveOp{row_of(p, k), slot, rule_of(p), b->c[p.W].N2});
        }
      }
    }
  };
  int *const no_gate[2] = {nullptr, nullptr};
  auto do_stream = [&](int sidx) {
    vb200_stream *x = streams[sidx];
    const int *lst = order.data() + sstart[sidx];
    const int cnt = sstart[sidx + 1] - sstart[sidx];
    for (int j = 0; j < cnt; j++) {
      const int i = lst[j];
      const vb200_block &q = blocks[i];
      HostBlock &h = b->hb[i];
      const int W = q.W;
      h.stream = q.stream; h.W = W; h.lW = (int)q.lW; h.nW = (int)q.nW; h.blocktype = q.blocktype;
      h.block_mode = q.blocktype | (W << 1); h.psy = q.blocktype + (W ? 2 : 0);
      h.granulepos = q.granulepos; h.eof = q.eofflag; h.idx = idx_of[i];
      h.hist[0] = x->lW_block_mode; h.hist[1] = x->lW_no; h.hist[2] = x->impadnum;
      BlockDesc &d = hc[W].desc[h.idx];
      d.pcm_off = q.pcm_offset; d.pcm_stride = (int)q.pcm_stride; d.slot = x->slot; d.lW_no = x->lW_no;
      d.lW = (unsigned char)h.lW; d.nW = (unsigned char)h.nW; d.blocktype = (unsigned char)q.blocktype;
      d.lW_block_mode = (unsigned char)x->lW_block_mode; d.impadnum = (unsigned char)x->impadnum; d.pad0 = d.pad1 = 0;
      if (j == 0) { d.prev = -1; d.prev_rule = LAST_COPY; }
      else { const HostBlock &p = b->hb[lst[j - 1]]; d.prev = (p.W << 30) | p.idx; d.prev_rule = (unsigned char)rule_of(p); }
      /* aoTuV history, mapping0.c:1297-1305 -- once per packet blob */
      for (int rep = 0; rep < (managed ? VB200_PACKETBLOBS : 1); rep++) {
        if (h.block_mode >= 2) x->impadnum = 0;
        if (!x->lW_block_mode && h.block_mode == 1) x->impadnum = 1;
        else if (x->impadnum && x->impadnum < 8) x->impadnum++;
        if (x->lW_block_mode == h.block_mode) x->lW_no++; else x->lW_no = 1;
        x->lW_block_mode = h.block_mode;
      }
      ScanBlock &sb = h_scan[sstart[sidx] + j];
      sb.W = W; sb.first_item = h.idx * ch; sb.block_mode = h.block_mode; sb.lW_block_mode = h.hist[0]; sb.psy = h.psy;
    }
    /* chains for K4: runs of consecutive impulse blocks of one stream, per channel; everything else is its own
       chain.  Offsets are local to the stream here and rebased when the streams are concatenated. */
    link_stream(sidx, no_gate, false);
    StreamChains &C = sc[sidx];
    C.impulse = false;
    for (int j = 0; j < cnt; j++) if (b->hb[lst[j]].W == 0 && b->hb[lst[j]].block_mode == 0) { C.impulse = true; break; }
    /* (long blocks: every ch-block is its own chain, K4 needs no list for them) */
    for (int k = 0; k < ch; k++) {
      if (managed) {
        /* managed mode: the sel = 1 pass of a run's FIRST impulse block reads tempmdct before the sel = 2 pass resets
           it (mdctbuf_flag == 0, psy.c:4165-4210) -- it sees what the previous run left.  The recurrence therefore
           runs through every impulse block of the stream: one chain per (stream, channel), the only chain that
           touches the state slot; every other block is a chain of its own that neither loads nor stores it. */
        const int start0 = (int)C.items[0].size();
        for (int j = 0; j < cnt; j++) {
          const HostBlock &h = b->hb[lst[j]];
          if (h.W == 0 && h.block_mode == 0) C.items[0].push_back(h.idx * ch + k);
        }
        if ((int)C.items[0].size() > start0) { C.start[0].push_back(start0); C.len[0].push_back((int)C.items[0].size() - start0); }
        for (int j = 0; j < cnt; j++) {
          const HostBlock &h = b->hb[lst[j]];
          if (h.W || h.block_mode == 0) continue;
          C.start[0].push_back((int)C.items[0].size());
          C.items[0].push_back(h.idx * ch + k);
          C.len[0].push_back(1 | (1 << 30) | (1 << 29));
        }
        continue;
      }
      int j = 0;
      while (j < cnt) {
        const HostBlock &h = b->hb[lst[j]];
        const int W = h.W;
        if (W) { j++; continue; }
        const int start = (int)C.items[W].size();
        C.items[W].push_back(h.idx * ch + k);
        int e = j + 1;
        if (W == 0 && h.block_mode == 0)
          while (e < cnt && b->hb[lst[e]].W == 0 && b->hb[lst[e]].block_mode == 0) { C.items[W].push_back(b->hb[lst[e]].idx * ch + k); e++; }
        /* only the chain holding the stream's first block may read tempmdct from the state slot, and only
           the one holding its last block leaves it there */
        C.start[W].push_back(start);
        C.len[W].push_back(((int)C.items[W].size() - start) | (j == 0 ? 0 : (1 << 30)) | (e == cnt ? 0 : (1 << 29)));
        j = e;
      }
    }
  };
  {
    int nthr = nblocks >= 8192 ? (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (nthr > nstreams) nthr = nstreams;
    if (nthr <= 1) { for (int sidx = 0; sidx < nstreams; sidx++) do_stream(sidx); }
    else {
      std::vector<std::thread> th;
      std::atomic<int> next(0);
      for (int t = 0; t < nthr; t++) th.emplace_back([&]() { for (int k; (k = next.fetch_add(1)) < nstreams;) do_stream(k); });
      for (auto &t : th) t.join();
    }
  }
  const double t_prep2 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  /* (3) concatenate the chains */
  for (int sidx = 0; sidx < nstreams; sidx++)
    for (int W = 0; W < 1; W++) {
      HostClass &c = hc[W];
      const StreamChains &C = sc[sidx];
      const int base = c.nchain_items;
      memcpy(c.chain_items + base, C.items[W].data(), sizeof(int) * C.items[W].size());
      for (size_t q = 0; q < C.start[W].size(); q++) { c.chain_start[c.nchains] = base + C.start[W][q]; c.chain_len[c.nchains] = C.len[W][q]; c.nchains++; }
      c.nchain_items += (int)C.items[W].size();
    }
  for (int sidx = 0; sidx < nstreams; sidx++) { h_scan_start[sidx] = sstart[sidx]; h_slot[sidx] = streams[sidx]->slot; }
  h_scan_start[nstreams] = nblocks;
  b->n_blocks[0] = nb[0]; b->n_blocks[1] = nb[1]; b->n_items[0] = nb[0] * ch; b->n_items[1] = nb[1] * ch;

  if (getenv("VB200_TRACE")) fprintf(stderr, "[vb200 batch] host pre-pass %.2f ms for %d blocks (setup %.2f, count %.2f, order %.2f, per-stream %.2f)\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_prep0, nblocks, t_prepA - t_prep0, t_prepB - t_prepA, t_prep1 - t_prepB, t_prep2 - t_prep1);
  /* ---- uploads ---- */
  if (b->after && b->after->ev_done_set) CUDA_TRY(cudaStreamWaitEvent(st, b->after->ev_done, 0));   /* its kernels first */
  EV("h2d");
  const float *d_pcm = pcm;
  if (!pcm_on_device) {
    if (b->pcm_cap < pcm_elems) {
      if (b->d_pcm) cudaFree(b->d_pcm);
      b->d_pcm = nullptr; b->pcm_cap = 0;
      CUDA_TRY(cudaMalloc(&b->d_pcm, pcm_elems * sizeof(float)));
      b->pcm_cap = pcm_elems;
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_pcm, pcm, pcm_elems * sizeof(float), cudaMemcpyHostToDevice, st));
    d_pcm = b->d_pcm;
  }
  /* every source below is 256-byte aligned pinned staging (b->pin), every destination a 256-byte aligned arena slice,
     every size a multiple of 4 bytes */
  FetchTable ft; ft.n = 0;
  auto fetch_flush = [&]() -> int {
    if (!ft.n) return VB200_OK;
    k_fetch<<<dim3(96, ft.n), 256, 0, st>>>(ft);
    ft.n = 0;
    LAUNCH_CHECK();
    return VB200_OK;
  };
  auto fetch_add = [&](void *dst, const void *src, size_t bytes) -> int {
    if (!bytes) return VB200_OK;
    if (bytes >> 32) { vb200_set_error("vb200_batch_submit: descriptor array too large"); return VB200_EINVAL; }
    ft.dst[ft.n] = dst; ft.src[ft.n] = src; ft.bytes[ft.n] = (unsigned)bytes; ft.n++;
    return ft.n == FETCH_MAX ? fetch_flush() : VB200_OK;
  };
#define UP(dst, src, n) RUN(fetch_add(dst, src, sizeof(*(src)) * (size_t)(n)))
  for (int W = 0; W < 2; W++) {
    if (!nb[W]) continue;
    ClassBufs &c = b->c[W]; HostClass &h = hc[W];
    UP(c.desc, h.desc, nb[W]);
    if (W == 0) { UP(c.chain_start, h.chain_start, h.nchains); UP(c.chain_len, h.chain_len, h.nchains); UP(c.chain_items, h.chain_items, h.nchain_items); }
  }
  ScanBlock *d_scan = b->d_scan; int *d_scan_start = b->d_scan_start, *d_slot = b->d_slot; float **d_ptrs = b->d_ptrs;
  {
    UP(d_scan, h_scan, nblocks); UP(d_scan_start, h_scan_start, nstreams + 1); UP(d_slot, h_slot, nstreams);
    float **h_ptrs = b->pin.take<float *>(8);
    h_ptrs[0] = b->c[0].local_ampmax; h_ptrs[1] = b->c[1].local_ampmax; h_ptrs[2] = b->c[0].global_ampmax; h_ptrs[3] = b->c[1].global_ampmax;
    h_ptrs[4] = b->c[0].logmdct; h_ptrs[5] = b->c[1].logmdct; h_ptrs[6] = b->c[0].lowcomp; h_ptrs[7] = b->c[1].lowcomp;
    UP(d_ptrs, h_ptrs, 8);
  }

  RUN(fetch_flush());
  /* every per-ch-block descriptor array, from the 32-byte block records */
  for (int W = 0; W < 2; W++) {
    if (!nb[W]) continue;
    ClassBufs &c = b->c[W];
    ExpandArgs a;
    a.psy_item = c.psy_item; a.psy_blk = c.psy_blk; a.block_mode = c.block_mode; a.lW_block_mode = c.lW_block_mode; a.lW = c.lW; a.nW = c.nW; a.params = c.params;
    a.pcm_off = c.pcm_off; a.last_addr = c.last_addr; a.last_addr2 = gated ? c.last_addr2 : nullptr; a.temp_off = c.temp_off;
    a.last_rule = c.last_rule; a.last_rule2 = gated ? c.last_rule2 : nullptr;
    a.logmdct[0] = b->c[0].logmdct; a.logmdct[1] = b->c[1].logmdct; a.N2[0] = b->c[0].N2; a.N2[1] = b->c[1].N2;
    a.last_pool = pool->last; a.W = W; a.ch = ch; a.managed = managed; a.hs = hs ? 1 : 0; a.flr_end = flr_of_end[W]; a.nblocks = nb[W];
    k_expand_desc<<<(nb[W] * ch + 255) / 256, 256, 0, st>>>(c.desc, a);
    LAUNCH_CHECK();
  }
  /* ---- K1: window + MDCT + FFT/dB ---- */
  EV("K1 transform");
  auto run_k1 = [&]() -> int {
    for (int W = 0; W < 2; W++) {
      ClassBufs &c = b->c[W];
      RUN(vb200_launch_transform(s, st, W, nb[W] * ch, d_pcm, c.pcm_off, c.lW, c.nW, c.block_mode, c.lW_block_mode,
                                 c.gmdct, c.logfft, c.logmdct, c.local_ampmax, c.poste));
    }
    return VB200_OK;
  };
  RUN(run_k1());
  EV("scan ampmax");
  {
    bool loudnoise_off = true;          /* lb_loudnoise_fix returns -1 for every psy of the setup (psy.c:5162) */
    for (int k = 0; k < s->h.psys; k++) if (!(s->h.psy[k].m_val < 0.5f || s->h.psy[k].normal_thresh > .45)) loudnoise_off = false;
    static const bool general = getenv("VB200_SCAN_GENERAL") != nullptr;
    if (loudnoise_off && !general)
      k_stream_scan_ampmax<<<nstreams, SC_T, 0, st>>>(s->d, nstreams, d_slot, d_scan_start, d_scan, pool->scal,
                                                    (const float *const *)d_ptrs, (float *const *)(d_ptrs + 2), (float *const *)(d_ptrs + 6));
    else
      k_stream_scan<<<(nstreams + 3) / 4, 128, 0, st>>>(s->d, nstreams, d_slot, d_scan_start, d_scan, pool->scal,
                                                        (const float *const *)d_ptrs, (float *const *)(d_ptrs + 2),
                                                        (const float *const *)(d_ptrs + 4), (float *const *)(d_ptrs + 6));
  }
  LAUNCH_CHECK();
  /* ---- K2 / K3 / K4.  Managed mode at high sampling rates: which ch-blocks write lastmdct / tempmdct depends on
     their own masks (the nominal-fit gate), so the links assumed by the pre-pass -- every block writes -- are checked
     against the gates K4 finds and, if a gate is closed (digital silence, mostly), the links are rebuilt and K1
     (K4 scales gmdct in place), K2 and K4 run again until gates and links agree.  The prefix of every stream up to
     its first changed gate is final after each round, so this ends; streams without closed gates take one round. ---- */
  std::vector<int> gate_prev[2];
  int round = 0;
  for (;; round++) {
    if (round > 0) {
      int *const g[2] = {hc[0].gate, hc[1].gate};
      for (int sidx = 0; sidx < nstreams; sidx++) link_stream(sidx, g, true);
      for (int W = 0; W < 2; W++) {
        const int I = nb[W] * ch;
        if (!I) continue;
        ClassBufs &c = b->c[W]; HostClass &h = hc[W];
        UP(c.last_addr, h.last_addr, I); UP(c.last_rule, h.last_rule, I); UP(c.last_addr2, h.last_addr2, I); UP(c.last_rule2, h.last_rule2, I);
      }
      RUN(fetch_flush());
      RUN(run_k1());
    }
    if (round == 0) EV("K2 noisemask");
    for (int W = 0; W < 2; W++) {
      ClassBufs &c = b->c[W];
      LastRef last{nullptr, c.last_addr, c.last_rule, 0};
      RUN(vb200_launch_noisemask(s, st, W, nb[W] * ch, c.psy_item, c.block_mode, c.lowcomp, c.poste, c.logmdct, last,
                                 gated ? c.last_addr2 : nullptr, gated ? c.last_rule2 : nullptr, c.noise, c.epeak, c.npeak, c.nps));
    }
    if (round == 0) {
      EV("K3 tonemask");
      for (int W = 0; W < 2; W++) {
        ClassBufs &c = b->c[W];
        /* scratch rows: the LONG class's logmask / iwork rows, dead until K4 / K5b, for both size classes (short first) */
        RUN(vb200_launch_tonemask(s, st, W, nb[W] * ch, c.psy_item, c.logfft, c.global_ampmax, c.local_ampmax, c.tone, b->c[1].logmask, b->c[1].iwork, b->c[1].N2));
      }
      EV("K4 offset_and_mix");
    }
    for (int W = 0; W < 2; W++) {
      ClassBufs &c = b->c[W];
      if (!nb[W]) continue;
      LastRef last{nullptr, c.last_addr, c.last_rule, 0};      /* M3 reads lastmdct in short blocks only: one source */
      RUN(vb200_launch_mix(s, st, W, W ? nb[W] * ch : hc[W].nchains, W ? nullptr : c.chain_start, W ? nullptr : c.chain_len, W ? nullptr : c.chain_items, c.psy_item, c.params, c.lowcomp,
                           c.noise, c.tone, c.logmdct, c.gmdct, last, pool->temp, gated ? pool->temp2 : pool->temp, c.temp_off, c.npeak, c.nps,
                           c.logmask, nsel, c.N2 * nsel, nullptr, managed ? c.fit_gate : nullptr, -1));
    }
    if (!gated) break;
    for (int W = 0; W < 2; W++) if (nb[W]) CUDA_TRY(cudaMemcpyAsync(hc[W].gate, b->c[W].fit_gate, sizeof(int) * (size_t)nb[W] * ch, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    bool same = true;
    for (int W = 0; W < 2; W++) {
      const size_t I = (size_t)nb[W] * ch;
      if (round == 0) { for (size_t q = 0; q < I && same; q++) if (!hc[W].gate[q]) same = false; }
      else if (memcmp(gate_prev[W].data(), hc[W].gate, sizeof(int) * I)) same = false;
    }
    if (same) break;
    if (round == 15) { vb200_set_error("vb200_batch_submit: the floor-fit gates did not settle in 16 rounds"); return VB200_EFAULT; }
    for (int W = 0; W < 2; W++) gate_prev[W].assign(hc[W].gate, hc[W].gate + (size_t)nb[W] * ch);
  }
  b->link_rounds = round + 1;
  if (gated) {
    /* tempmdct as the final round left it, for the streams that had an impulse block */
    long long *h_off = b->pin.take<long long>((size_t)nstreams * ch);
    int n = 0;
    for (int sidx = 0; sidx < nstreams; sidx++)
      if (sc[sidx].impulse) for (int k = 0; k < ch; k++) h_off[n++] = ((long long)streams[sidx]->slot * ch + k) * 256;
    if (n) {
      long long *d_off = b->d_commit;
      UP(d_off, h_off, n);
      RUN(fetch_flush());
      k_commit_temp<<<n, 128, 0, st>>>(d_off, pool->temp2, pool->temp);
      LAUNCH_CHECK();
    }
  }
  /* ---- K5a ---- */
  EV("K5 floor fit");
  for (int W = 0; W < 2; W++) {
    ClassBufs &c = b->c[W];
    const int I = nb[W] * ch;
    if (!I) continue;
    /* nominal curve: blob PACKETBLOBS/2 */
    RUN(vb200_launch_floor_fit(s, st, W, I, -1, c.logmdct, c.logmask, c.N2, c.N2 * nsel, nullptr, c.posts[1], c.fitted[1]));
    if (managed) {
      RUN(vb200_launch_floor_fit(s, st, W, I, -1, c.logmdct, c.logmask + c.N2, c.N2, c.N2 * nsel, c.fitted[1], c.posts[2], c.fitted[2]));
      RUN(vb200_launch_floor_fit(s, st, W, I, -1, c.logmdct, c.logmask + 2 * c.N2, c.N2, c.N2 * nsel, c.fitted[1], c.posts[0], c.fitted[0]));
    }
  }
  /* stream state for the next batch: lastmdct as the writers of this batch leave it (the newest 1024-entry writer, then
     a newer 128-entry writer on top of it; link_stream), tempmdct is already in place */
  for (int pass = 0; pass < 2; pass++) {
    const std::vector<std::vector<SaveOp>> &ops = pass ? saveB : saveA;
    size_t total = 0;
    for (int sidx = 0; sidx < nstreams; sidx++) total += ops[sidx].size();
    if (!total) continue;
    const float **h_src = b->pin.take<const float *>(total); float **h_dst = b->pin.take<float *>(total);
    int *h_rule = b->pin.take<int>(total), *h_N = b->pin.take<int>(total);
    int n = 0;
    for (int sidx = 0; sidx < nstreams; sidx++)
      for (const SaveOp &o : ops[sidx]) { h_src[n] = o.src; h_dst[n] = o.dst; h_rule[n] = o.rule; h_N[n] = o.N; n++; }
    const size_t base = (size_t)pass * PIPE_MAX_STREAM_SLOTS * VB200_MAX_CHANNELS;
    const float **d_src = b->d_save_src + base; float **d_dst = b->d_save_dst + base; int *d_rule = b->d_save_rule + base, *d_N = b->d_save_N + base;
    UP(d_src, h_src, n); UP(d_dst, h_dst, n); UP(d_rule, h_rule, n); UP(d_N, h_N, n);
    RUN(fetch_flush());
    k_save_last<<<n, 128, 0, st>>>(n, d_src, d_rule, d_N, d_dst);
    LAUNCH_CHECK();
  }

  /* ---- per packet blob: K5b, K6, K7, then the bit packer (device) or the symbol D2H (host emission) ---- */
  const int k_lo = managed ? 0 : VB200_PACKETBLOBS / 2, k_hi = managed ? VB200_PACKETBLOBS - 1 : VB200_PACKETBLOBS / 2;
  const int nblobs = k_hi - k_lo + 1;
  const bool host_emit = b->host_emit != 0;
  struct HostRes { int *vals, *nonzero, *nsyms, *syms; } hr[2] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  std::vector<long> &blob_sizes = b->blob_sizes;
  if (managed) blob_sizes.assign((size_t)nblocks * VB200_PACKETBLOBS, 0);
  if (host_emit) {
    for (int W = 0; W < 2; W++) {
      const size_t I = (size_t)nb[W] * ch;
      const DevMap &M = s->h.map[s->h.mode_mapping[W]];
      hr[W].vals = b->pin.take<int>(I * VB200_MAX_POSTS + 1); hr[W].nonzero = b->pin.take<int>(I + 1);
      hr[W].nsyms = b->pin.take<int>((size_t)nb[W] * M.submaps + 1);
      hr[W].syms = b->pin.take<int>((size_t)nb[W] * M.submaps * b->c[W].max_syms * 2 + 1);
      if (!hr[W].syms) { vb200_set_error("vb200_batch_run: pinned staging too small for host emission (use a smaller batch)"); return VB200_EFAULT; }
    }
  } else {
    CUDA_TRY(cudaMemsetAsync(b->d_pool_top, 0, sizeof(unsigned long long), st));
  }
  /* chlist per (W, submap) on device */
  int *d_chlist[2][VB200_MAX_SUBMAPS]; int nch_sub[2][VB200_MAX_SUBMAPS];
  {
    int *h_cl = b->pin.take<int>(2 * VB200_MAX_SUBMAPS * VB200_MAX_CHANNELS);
    int *d_cl = b->d_chl;
    for (int W = 0; W < 2; W++) {
      const DevMap &M = s->h.map[s->h.mode_mapping[W]];
      for (int sm = 0; sm < M.submaps; sm++) {
        int n = 0;
        for (int k = 0; k < ch; k++) if (M.chmux[k] == sm) h_cl[(W * VB200_MAX_SUBMAPS + sm) * VB200_MAX_CHANNELS + n++] = k;
        nch_sub[W][sm] = n;
        d_chlist[W][sm] = d_cl + (W * VB200_MAX_SUBMAPS + sm) * VB200_MAX_CHANNELS;
      }
    }
    UP(d_cl, h_cl, 2 * VB200_MAX_SUBMAPS * VB200_MAX_CHANNELS);
    RUN(fetch_flush());
  }

  std::vector<BitWriter> &bws = b->bws; std::vector<std::vector<size_t>> &boff = b->boff;
  bws.clear(); boff.clear();
  if (managed && host_emit) { bws.resize(VB200_PACKETBLOBS); boff.resize(VB200_PACKETBLOBS); }
  for (int k = k_lo; k <= k_hi; k++) {
    if (k != k_lo) EV("blob");
    for (int W = 0; W < 2; W++) {
      ClassBufs &c = b->c[W];
      const int I = nb[W] * ch;
      if (!I) continue;
      if (k == k_lo) EV("K5b floor encode");
      const DevMap &M = s->h.map[s->h.mode_mapping[W]];
      /* posts of blob k (mapping0.c:1090-1181) into posts_k (floor1_encode mutates them) */
      const int H = VB200_PACKETBLOBS / 2;
      if (!managed || k == H) { k_copy_i32<<<copy_grid((size_t)I * VB200_MAX_POSTS), 256, 0, st>>>(c.posts[1], c.posts_k, (size_t)I * VB200_MAX_POSTS); k_copy_i32<<<copy_grid((size_t)I), 256, 0, st>>>(c.fitted[1], c.fitted_k, I); vb200_count_launch(2); }
      else if (k == 0) { k_copy_i32<<<copy_grid((size_t)I * VB200_MAX_POSTS), 256, 0, st>>>(c.posts[0], c.posts_k, (size_t)I * VB200_MAX_POSTS); k_copy_i32<<<copy_grid((size_t)I), 256, 0, st>>>(c.fitted[0], c.fitted_k, I); vb200_count_launch(2); }
      else if (k == VB200_PACKETBLOBS - 1) { k_copy_i32<<<copy_grid((size_t)I * VB200_MAX_POSTS), 256, 0, st>>>(c.posts[2], c.posts_k, (size_t)I * VB200_MAX_POSTS); k_copy_i32<<<copy_grid((size_t)I), 256, 0, st>>>(c.fitted[2], c.fitted_k, I); vb200_count_launch(2); }
      else if (k < H) RUN(vb200_launch_floor_interp(s, st, W, I, -1, c.posts[0], c.posts[1], c.fitted[0], c.fitted[1], k * 65536 / H, c.posts_k, c.fitted_k));
      else RUN(vb200_launch_floor_interp(s, st, W, I, -1, c.posts[1], c.posts[2], c.fitted[1], c.fitted[2], (k - H) * 65536 / H, c.posts_k, c.fitted_k));
      /* floor1_encode's return value (the pre-coupling nonzero flag) is what the packet carries */
      RUN(vb200_launch_floor_encode(s, st, W, I, -1, c.posts_k, c.fitted_k, c.vals, c.iwork, c.floor_nz));
      k_copy_i32<<<copy_grid((size_t)I), 256, 0, st>>>(c.floor_nz, c.nonzero, I); vb200_count_launch(1);
      if (host_emit) {
        CUDA_TRY(cudaMemcpyAsync(hr[W].vals, c.vals, sizeof(int) * (size_t)I * VB200_MAX_POSTS, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(hr[W].nonzero, c.floor_nz, sizeof(int) * (size_t)I, cudaMemcpyDeviceToHost, st));
      }
      if (k == k_lo) EV("K6 couple");
      RUN(vb200_launch_couple(s, st, W, nb[W], c.psy_blk, k, c.gmdct, c.epeak, c.npeak, c.nps, c.iwork, c.nonzero));
      if (k == k_lo) EV("K7 residue");
      for (int sm = 0; sm < M.submaps; sm++) {
        const int resnum = M.residuesubmap[sm];
        const DevResidue &R = s->h.res[resnum];
        const int pv = (R.end - R.begin) / R.grouping;
        int *pw = c.partword + (size_t)sm * c.cap_blocks * ch * pv;    /* [submap][block][ch][pv] with this residue's pv */
        int *cl = c.classified + (size_t)sm * c.cap_blocks;
        int *sy = c.syms + (size_t)sm * c.cap_blocks * c.max_syms * 2;
        int *ns = c.nsyms + (size_t)sm * c.cap_blocks;
        RUN(vb200_launch_residue_class(s, st, W, nb[W], resnum, nch_sub[W][sm], d_chlist[W][sm], ch, c.iwork, c.nonzero, pw, cl));
        RUN(vb200_launch_residue_forward(s, st, W, nb[W], resnum, nch_sub[W][sm], d_chlist[W][sm], ch, c.iwork, c.nonzero, pw, sy, c.max_syms, ns));
        if (host_emit) {
          CUDA_TRY(cudaMemcpyAsync(hr[W].nsyms + (size_t)sm * nb[W], ns, sizeof(int) * (size_t)nb[W], cudaMemcpyDeviceToHost, st));
          CUDA_TRY(cudaMemcpyAsync(hr[W].syms + (size_t)sm * nb[W] * c.max_syms * 2, sy, sizeof(int) * (size_t)nb[W] * c.max_syms * 2, cudaMemcpyDeviceToHost, st));
        }
      }
      if (!host_emit) {
        if (k == k_lo) EV("K8 bitpack");
        RUN(vb200_launch_pack(s, st, W, nb[W], c.lW, c.nW, c.vals, c.floor_nz, M.submaps, c.syms, c.nsyms, c.cap_blocks, c.max_syms,
                              b->d_pool, b->pool_cap, b->d_pool_top,
                              b->d_pk_off + ((size_t)(k - k_lo) * 2 + W) * b->max_blocks, b->d_pk_bytes + ((size_t)(k - k_lo) * 2 + W) * b->max_blocks));
      }
    }
    if (!host_emit) continue;
    if (k == k_lo) EV("d2h+emit");
    CUDA_TRY(cudaStreamSynchronize(st));
    /* ---- host emission of blob k for every block (mapping0.c:1204-1295 order) ---- */
    BitWriter &w = managed ? bws[k] : b->bw;
    for (int i = 0; i < nblocks; i++) {
      const HostBlock &h = b->hb[i];
      const int W = h.W;
      const DevMap &M = s->h.map[s->h.mode_mapping[W]];
      const size_t start = w.begin_packet();
      vb200_emit_header(s, w, W, h.lW, h.nW);
      for (int c2 = 0; c2 < ch; c2++) {
        const size_t it = (size_t)h.idx * ch + c2;
        vb200_emit_floor(s, w, s->h.floor_of_ch[W][c2], hr[W].nonzero[it], hr[W].vals + it * VB200_MAX_POSTS);
      }
      for (int sm = 0; sm < M.submaps; sm++) {
        const int n = hr[W].nsyms[(size_t)sm * nb[W] + h.idx];
        vb200_emit_syms(s, w, hr[W].syms + ((size_t)sm * nb[W] + h.idx) * b->c[W].max_syms * 2, n);
      }
      if (managed) { boff[k].push_back(start); blob_sizes[(size_t)i * VB200_PACKETBLOBS + k] = (long)w.bytes_since(start); }
      else {
        b->pk_off.push_back(start); b->pk_bytes.push_back(w.bytes_since(start));
        b->pk_gp.push_back(h.granulepos); b->pk_eos.push_back(h.eof); b->pk_stream.push_back(h.stream);
      }
    }
  }
  b->run_nb[0] = nb[0]; b->run_nb[1] = nb[1];
  bool done_marked = false;
  if (!host_emit && managed) {
    /* the bitrate manager runs where the blobs are: only the chosen packets leave the device */
    EV("bitrate manager");
    for (int i = 0; i < nstreams; i++) b->h_br[i] = streams[i]->br;
    CUDA_TRY(cudaMemcpyAsync(b->d_br, b->h_br, sizeof(BitrateState) * nstreams, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(b->d_out_top, 0, 8, st));
    k_bitrate<<<(nstreams + 31) / 32, 32, 0, st>>>(bitrate_params(s), nstreams, b->d_scan_start, b->d_scan, ch, b->max_blocks, b->d_pk_off, b->d_pk_bytes,
                                                   b->d_br, b->d_choice, b->d_final, b->d_out_off, b->d_out_top, (unsigned long long)b->out_cap);
    LAUNCH_CHECK();
    for (int W = 0; W < 2; W++) if (nb[W]) {
      k_gather_choice<<<nb[W], 128, 0, st>>>(W, b->max_blocks, b->d_choice, b->d_final, b->d_out_off, b->d_pk_off, b->d_pk_bytes, b->d_pool, b->d_out_pool);
      LAUNCH_CHECK();
    }
    EV("d2h packets");
    CUDA_TRY(cudaEventRecord(b->ev_done, st)); done_marked = true;      /* the kernels end here: a batch ordered behind this one may start */
    CUDA_TRY(cudaMemcpyAsync(b->h_out_top, b->d_out_top, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(b->h_pool_top, b->d_pool_top, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(b->h_br, b->d_br, sizeof(BitrateState) * nstreams, cudaMemcpyDeviceToHost, st));
    for (int W = 0; W < 2; W++) if (nb[W]) {
      const size_t o = (size_t)W * b->max_blocks;
      CUDA_TRY(cudaMemcpyAsync(b->h_choice + o, b->d_choice + o, sizeof(int) * nb[W], cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(b->h_final + o, b->d_final + o, sizeof(int) * nb[W], cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(b->h_out_off + o, b->d_out_off + o, sizeof(long long) * nb[W], cudaMemcpyDeviceToHost, st));
    }
  } else if (!host_emit) {
    /* first half of the only D2H of the path: packet offsets / sizes and the pool fill level */
    EV("d2h packets");
    CUDA_TRY(cudaEventRecord(b->ev_done, st)); done_marked = true;
    CUDA_TRY(cudaMemcpyAsync(b->h_pool_top, b->d_pool_top, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    for (int kk = 0; kk < nblobs; kk++) for (int W = 0; W < 2; W++) if (nb[W]) {
      const size_t o = ((size_t)kk * 2 + W) * b->max_blocks;
      CUDA_TRY(cudaMemcpyAsync(b->h_pk_off + o, b->d_pk_off + o, sizeof(long long) * nb[W], cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaMemcpyAsync(b->h_pk_bytes + o, b->d_pk_bytes + o, sizeof(int) * nb[W], cudaMemcpyDeviceToHost, st));
    }
  }
  if (!done_marked) CUDA_TRY(cudaEventRecord(b->ev_done, st));
  b->ev_done_set = true;
  b->pending = true;
  if (host_emit) return batch_finish(b);
  return VB200_OK;
}

/* second half: wait for the kernels, fetch the packets, run the bitrate manager */
static int batch_finish(vb200_batch *b) {
  vb200_setup *s = b->s;
  CUDA_TRY(cudaSetDevice(s->device));
  cudaStream_t st = b->st;
  const int managed = s->h.managed, nblocks = b->run_nblocks;
  const bool host_emit = b->host_emit != 0;
  std::vector<long> &blob_sizes = b->blob_sizes;
  std::vector<BitWriter> &bws = b->bws; std::vector<std::vector<size_t>> &boff = b->boff;
  vb200_stream **streams = b->run_streams.data();
  b->pending = false;
  b->pk_base = nullptr;
  if (!host_emit && managed) {
    CUDA_TRY(cudaStreamSynchronize(st));
    if (*b->h_pool_top > b->pool_cap) { vb200_set_error("vb200_batch_collect: packet pool overflow (%llu > %llu bytes)", *b->h_pool_top, (unsigned long long)b->pool_cap); return VB200_EFAULT; }
    const unsigned long long top = *b->h_out_top;
    if (top > b->out_cap) { vb200_set_error("vb200_batch_collect: chosen-packet pool overflow (%llu > %llu bytes)", top, (unsigned long long)b->out_cap); return VB200_EFAULT; }
    CUDA_TRY(cudaMemcpyAsync(b->h_pool, b->d_out_pool, top, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    b->pk_off.resize(nblocks); b->pk_bytes.resize(nblocks); b->pk_gp.resize(nblocks); b->pk_eos.resize(nblocks); b->pk_stream.resize(nblocks);
    for (int i = 0; i < nblocks; i++) {
      const HostBlock &h = b->hb[i];
      const size_t o = (size_t)h.W * b->max_blocks + h.idx;
      if (b->h_final[o] < 0) { vb200_set_error("vb200_batch_collect: packet of block %d does not fit the packer's image or the pool", i); return VB200_EFAULT; }
      b->pk_off[i] = (size_t)b->h_out_off[o]; b->pk_bytes[i] = (size_t)b->h_final[o];
      b->pk_gp[i] = h.granulepos; b->pk_eos[i] = h.eof; b->pk_stream[i] = h.stream;
    }
    for (size_t i = 0; i < b->run_streams.size(); i++) streams[i]->br = b->h_br[i];
    b->pk_base = b->h_pool;
  } else if (!host_emit) {
    CUDA_TRY(cudaStreamSynchronize(st));
    const unsigned long long top = *b->h_pool_top;
    if (top > b->pool_cap) { vb200_set_error("vb200_batch_collect: packet pool overflow (%llu > %llu bytes)", top, (unsigned long long)b->pool_cap); return VB200_EFAULT; }
    CUDA_TRY(cudaMemcpyAsync(b->h_pool, b->d_pool, top, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    b->pk_off.resize(nblocks); b->pk_bytes.resize(nblocks); b->pk_gp.resize(nblocks); b->pk_eos.resize(nblocks); b->pk_stream.resize(nblocks);
    for (int i = 0; i < nblocks; i++) {
      const HostBlock &h = b->hb[i];
      const size_t o = (size_t)h.W * b->max_blocks + h.idx;
      if (b->h_pk_off[o] < 0) { vb200_set_error("vb200_batch_collect: packet of block %d does not fit the packer's image", i); return VB200_EFAULT; }
      b->pk_off[i] = (size_t)b->h_pk_off[o]; b->pk_bytes[i] = (size_t)b->h_pk_bytes[o];
      b->pk_gp[i] = h.granulepos; b->pk_eos[i] = h.eof; b->pk_stream[i] = h.stream;
    }
    b->pk_base = b->h_pool;
  }
  if (managed && host_emit) {
    /* bitrate.c: choose one blob per block, in stream order (blocks arrive in stream order) */
    const BitrateParams bparams = bitrate_params(s);
    for (int i = 0; i < nblocks; i++) {
      const HostBlock &h = b->hb[i];
      long fb = 0;
      const int choice = bitrate_choose(&streams[h.stream]->br, bparams, h.W, &blob_sizes[(size_t)i * VB200_PACKETBLOBS], &fb);
      const size_t start = b->bw.begin_packet();
      const unsigned char *src = host_emit ? bws[choice].buf + boff[choice][i]
                                           : b->h_pool + b->h_pk_off[((size_t)choice * 2 + h.W) * b->max_blocks + h.idx];
      const long have = blob_sizes[(size_t)i * VB200_PACKETBLOBS + choice];
      for (long t = 0; t < fb; t++) b->bw.put(t < have ? src[t] : 0, 8);
      b->pk_off.push_back(start); b->pk_bytes.push_back((size_t)fb);
      b->pk_gp.push_back(h.granulepos); b->pk_eos.push_back(h.eof); b->pk_stream.push_back(h.stream);
    }
  }
  if (!b->pk_base) b->pk_base = b->bw.buf;
  EV("end");
  CUDA_TRY(cudaStreamSynchronize(st));
  b->total_ms = 0;
  for (int i = 0; i + 1 < b->nev; i++) { cudaEventElapsedTime(&b->ev_ms[i], b->ev[i], b->ev[i + 1]); }
  if (b->nev > 1) cudaEventElapsedTime(&b->total_ms, b->ev[0], b->ev[b->nev - 1]);
  if (getenv("VB200_TRACE") && b->nev > 1) {
    /* device-side timeline of this batch against a process-wide reference event (recorded at the first traced collect) */
    static cudaEvent_t ref = nullptr;
    if (!ref) { cudaEventCreate(&ref); cudaEventRecord(ref, st); cudaEventSynchronize(ref); }
    fprintf(stderr, "[vb200 batch] device timeline (ms, vs ref):");
    for (int i = 0; i < b->nev; i++) { float t = 0; cudaEventElapsedTime(&t, ref, b->ev[i]); fprintf(stderr, " %s@%.1f", b->evname[i], t); }
    fprintf(stderr, "\n");
  }
  return VB200_OK;
}

/* Order b's device work behind prev's: b can be submitted (host pre-pass, launches) while prev still runs, and its
   kernels start when prev's have finished -- successive jobs overlap host and device halves without sharing the SMs. */
extern "C" int vb200_batch_set_rank(vb200_batch *b, int rank) {
  if (!b || rank < 0) { vb200_set_error("vb200_batch_set_rank: bad arguments"); return VB200_EINVAL; }
  if (b->pending) { vb200_set_error("vb200_batch_set_rank: the batch is running"); return VB200_EINVAL; }
  CUDA_TRY(cudaSetDevice(b->s->device));
  int least = 0, greatest = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));      /* numerically lower = served first */
  int prio = greatest + 1 + rank;
  if (prio > least) prio = least;
  cudaStream_t st;
  CUDA_TRY(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio));
  CUDA_TRY(cudaStreamSynchronize(b->st));
  cudaStreamDestroy(b->st);
  b->st = st;
  return VB200_OK;
}

extern "C" int vb200_batch_after(vb200_batch *b, vb200_batch *prev) {
  if (!b || prev == b) { vb200_set_error("vb200_batch_after: bad arguments"); return VB200_EINVAL; }
  b->after = prev;
  return VB200_OK;
}

extern "C" int vb200_batch_collect(vb200_batch *b) {
  if (!b) return VB200_EINVAL;
  if (!b->pending) return VB200_OK;
  return batch_finish(b);
}

extern "C" int vb200_batch_run(vb200_batch *b, vb200_stream **streams, int nstreams,
                               const vb200_block *blocks, int nblocks,
                               const float *pcm, size_t pcm_elems, int pcm_on_device) {
  const int r = vb200_batch_submit(b, streams, nstreams, blocks, nblocks, pcm, pcm_elems, pcm_on_device);
  if (r) return r;
  return vb200_batch_collect(b);
}

extern "C" int vb200_batch_set_host_emit(vb200_batch *b, int on) {
  if (!b) return VB200_EINVAL;
  if (on && !b->host_emit_ready) {
    /* grow the pinned staging so that it can also hold vals + symbols of a whole batch */
    const vb200_setup *s = b->s;
    const size_t ch = s->h.channels;
    size_t extra = 0;
    for (int W = 0; W < 2; W++) {
      const size_t nsub = s->h.map[s->h.mode_mapping[W]].submaps;
      extra += (size_t)b->max_blocks * (ch * (VB200_MAX_POSTS + 2) * 4 + nsub * ((size_t)b->c[W].max_syms * 8 + 8)) + 4096;
    }
    unsigned char *nb = nullptr;
    if (cudaMallocHost(&nb, b->pin.cap + extra) != cudaSuccess) { cudaGetLastError(); vb200_set_error("vb200_batch_set_host_emit: cannot allocate %zu bytes of pinned staging", b->pin.cap + extra); return VB200_EFAULT; }
    cudaFreeHost(b->pin.base);
    b->pin.base = nb; b->pin.cap += extra; b->host_emit_ready = 1;
  }
  b->host_emit = on;
  return VB200_OK;
}

extern "C" int vb200_batch_packet(vb200_batch *b, int i, const unsigned char **data, long *bytes,
                                  int64_t *granulepos, int *eos, int *stream) {
  if (!b || i < 0 || i >= (int)b->pk_off.size()) return VB200_EINVAL;
  if (data) *data = b->pk_base + b->pk_off[i];
  if (bytes) *bytes = (long)b->pk_bytes[i];
  if (granulepos) *granulepos = b->pk_gp[i];
  if (eos) *eos = b->pk_eos[i];
  if (stream) *stream = b->pk_stream[i];
  return VB200_OK;
}

extern "C" int vb200_batch_packet_table(vb200_batch *b, const unsigned char **base, const size_t **off, const size_t **bytes,
                                        const int64_t **granulepos, const int **eos, const int **stream) {
  if (!b) return VB200_EINVAL;
  if (base) *base = b->pk_base;
  if (off) *off = b->pk_off.data();
  if (bytes) *bytes = b->pk_bytes.data();
  if (granulepos) *granulepos = b->pk_gp.data();
  if (eos) *eos = b->pk_eos.data();
  if (stream) *stream = b->pk_stream.data();
  return (int)b->pk_off.size();
}

extern "C" float vb200_batch_kernel_ms(vb200_batch *b) { return b ? b->total_ms : 0.f; }
extern "C" int vb200_batch_link_rounds(vb200_batch *b) { return b ? b->link_rounds : 0; }
/* device-side timing of a region that spans several runs (and several batches): a mark is a CUDA event recorded on the
   batch's own stream, behind everything the batch has been given so far */
extern "C" int vb200_batch_mark(vb200_batch *b) {
  if (!b) { vb200_set_error("vb200_batch_mark: bad arguments"); return VB200_EINVAL; }
  CUDA_TRY(cudaSetDevice(b->s->device));
  CUDA_TRY(cudaEventRecord(b->ev_mark, b->st));
  return VB200_OK;
}
extern "C" int vb200_batch_mark_elapsed(vb200_batch *from, vb200_batch *to, float *ms) {
  if (!from || !to || !ms) { vb200_set_error("vb200_batch_mark_elapsed: bad arguments"); return VB200_EINVAL; }
  CUDA_TRY(cudaSetDevice(to->s->device));
  CUDA_TRY(cudaEventSynchronize(from->ev_mark));
  CUDA_TRY(cudaEventSynchronize(to->ev_mark));
  CUDA_TRY(cudaEventElapsedTime(ms, from->ev_mark, to->ev_mark));
  return VB200_OK;
}

extern "C" int vb200_batch_kernel_times(vb200_batch *b, const char **names, float *ms, int cap) {
  if (!b) return 0;
  int n = 0;
  for (int i = 0; i + 1 < b->nev && n < cap; i++, n++) { names[n] = b->evname[i]; ms[n] = b->ev_ms[i]; }
  return n;
}

extern "C" int vb200_batch_get(vb200_batch *b, int what, int block, int chn, void *out, int count) {
  if (!b || !out || block < 0 || block >= (int)b->hb.size() || chn < 0 || chn >= b->s->h.channels) return VB200_EINVAL;
  const HostBlock &h = b->hb[block];
  const ClassBufs &c = b->c[h.W];
  const size_t it = (size_t)h.idx * b->s->h.channels + chn;
  const void *src = nullptr; size_t n = c.N2, esz = 4;
  switch (what) {
    case VB200_GET_GMDCT: src = c.gmdct + it * c.N2; break;
    case VB200_GET_LOGFFT: src = c.logfft + it * c.N2; break;
    case VB200_GET_LOGMDCT: src = c.logmdct + it * c.N2; break;
    case VB200_GET_NOISE: src = c.noise + it * c.N2; break;
    case VB200_GET_TONE: src = c.tone + it * c.N2; break;
    case VB200_GET_LOGMASK: src = c.logmask + it * c.N2 * (b->s->h.managed ? 3 : 1); n = (size_t)c.N2 * (b->s->h.managed ? 3 : 1); break;
    case VB200_GET_EPEAK: src = c.epeak + it * c.N2; break;
    case VB200_GET_IWORK: src = c.iwork + it * c.N2; break;
    case VB200_GET_ILOGMASK: src = c.iwork + it * c.N2; break;
    case VB200_GET_POSTS: src = c.posts[1] + it * VB200_MAX_POSTS; n = VB200_MAX_POSTS; break;
    case VB200_GET_POSTS_LO: src = c.posts[0] + it * VB200_MAX_POSTS; n = VB200_MAX_POSTS; break;
    case VB200_GET_POSTS_HI: src = c.posts[2] + it * VB200_MAX_POSTS; n = VB200_MAX_POSTS; break;
    case VB200_GET_NPEAK: src = c.npeak + it * c.nps; n = c.nps; break;
    case VB200_GET_BLOB_BYTES: {      /* sizes of the packet blobs of this block as the device packer left them (one per blob) */
      const int nblobs = b->s->h.managed ? VB200_PACKETBLOBS : 1;
      if (count < nblobs || b->host_emit) return VB200_EINVAL;
      for (int k = 0; k < nblobs; k++)
        if (cudaMemcpy((int *)out + k, b->d_pk_bytes + ((size_t)k * 2 + h.W) * b->max_blocks + h.idx, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
          vb200_set_error("vb200_batch_get: copy failed"); return VB200_EFAULT; }
      return nblobs;
    }
    default: return VB200_EINVAL;
  }
  if ((size_t)count < n) n = count;
  if (cudaMemcpy(out, src, n * esz, cudaMemcpyDeviceToHost) != cudaSuccess) { vb200_set_error("vb200_batch_get: copy failed"); return VB200_EFAULT; }
  return (int)n;
}
