/* Stream segmentation: PCM of whole streams -> the list of analysis blocks.
 *
 * Replaces, for offline (whole-stream) encoding, the block layer that north_star keeps on the
 * host: vorbis_analysis_buffer/_wrote (block.c:411-553: pre-amplitude, LPC extrapolation of the
 * stream head and tail, lpc.c:60-159), _ve_envelope_search/_ve_amp/_ve_envelope_mark
 * (envelope.c:101-362: the pre/post-echo detector that picks short vs long windows and marks
 * impulse blocks) and the window sequencing of vorbis_analysis_blockout (block.c:557-812).
 *
 * Not a port: the reference streams through a sliding buffer that it shifts after every block;
 * here a whole stream sits on one absolute timeline (1024 extrapolated samples of lead-in, the
 * PCM, 3 long blocks of extrapolated tail), and blocks are descriptors that point into that
 * timeline -- nothing is copied per block, and the timeline is exactly the buffer the GPU path
 * reads.  Two drivers share the code:
 *   vb200_segment_stream   -- host only (works without a GPU);
 *   vb200_segmenter_*      -- many streams at once; the detector's heavy half (a 128-point MDCT and
 *                             12 band energies every 64 samples) runs as CUDA
 *                             kernels on the timeline that is uploaded ONCE and then handed to
 *                             vb200_batch_run in place; the host keeps the LPC fits and the window
 *                             sequencing.  Both drivers produce identical blocks.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <atomic>
#include <chrono>
#include <stdio.h>
#include <vector>
#include "vb200_internal.h"
#include "dev_util.cuh"
#include "ve_mdct128.cuh"

namespace {

constexpr int VE_PRE = 16, VE_WIN = 4, VE_POST = 2, VE_AMP = VE_PRE + VE_POST - 1, VE_BANDS = 12, VE_NEARDC = 15;
constexpr int VE_MINSTRETCH = 2, VE_MAXSTRETCH = 12;
constexpr int VE_N = 128, VE_STEP = 64;
/* band layout of the detector (envelope.c:36-37) */
const int kBandBegin[VE_BANDS] = {2, 4, 6, 9, 13, 17, 22, 12, 8, 3, 2, 1};
const int kBandEnd[VE_BANDS] = {4, 5, 6, 8, 8, 8, 8, 4, 4, 3, 2, 4};

__host__ __device__ inline float todB_hd(float x) {
#ifdef __CUDA_ARCH__
  const uint32_t i = __float_as_uint(x) & 0x7fffffffu;
  return __fadd_rn(__fmul_rn(__uint2float_rn(i), 7.17711438e-7f), -764.6161886f);
#else
  union { uint32_t i; float f; } u; u.f = x; u.i &= 0x7fffffffu;
  return (float)(u.i * 7.17711438e-7f - 764.6161886f);
#endif
}

/* Levinson-Durbin from the autocorrelation (lpc.c:60-130) */
void lpc_from_data(const float *data, float *out, int n, int m) {
  std::vector<double> aut(m + 1), lpc(m, 0.0);
  for (int j = m; j >= 0; j--) { double d = 0; for (int i = j; i < n; i++) d += (double)data[i] * data[i - j]; aut[j] = d; }
  double error = aut[0] * (1. + 1e-10);
  const double eps = 1e-9 * aut[0] + 1e-10;
  for (int i = 0; i < m; i++) {
    double r = -aut[i + 1];
    if (error < eps) { for (int k = i; k < m; k++) lpc[k] = 0; break; }
    for (int j = 0; j < i; j++) r -= lpc[j] * aut[i - j];
    r /= error;
    lpc[i] = r;
    int j = 0;
    for (; j < i / 2; j++) { const double t = lpc[j]; lpc[j] += r * lpc[i - 1 - j]; lpc[i - 1 - j] += r * t; }
    if (i & 1) lpc[j] += lpc[j] * r;
    error *= 1. - r * r;
  }
  double damp = .99;
  for (int j = 0; j < m; j++) { lpc[j] *= damp; damp *= .99; }
  for (int j = 0; j < m; j++) out[j] = (float)lpc[j];
}

/* all-pole prediction (lpc.c:132-159): prime[0..m) history, data[0..n) out */
void lpc_predict(const float *coeff, const float *prime, int m, float *data, long n) {
  std::vector<float> work(m + n);
  for (int i = 0; i < m; i++) work[i] = prime ? prime[i] : 0.f;
  for (long i = 0; i < n; i++) {
    float y = 0; long o = i; int p = m;
    for (int j = 0; j < m; j++) y -= work[o++] * coeff[--p];
    data[i] = work[o] = y;
  }
}

/* tables of the detector, built once (envelope.c:42-82).  The 128-point MDCT is the reference's own (ve_mdct128.cuh:
   mdct_forward for n = 128, operation by operation), on the host and on the device alike. */
struct Detector {
  float win[VE_N];
  float bandwin[VE_BANDS][8], bandtotal[VE_BANDS];
  float T[VE_N + VE_N / 4];            /* mdct_init(128), mdct.c:54-92 */
  Detector() {
    for (int i = 0; i < VE_N; i++) { const float t = (float)sin(i / (VE_N - 1.) * M_PI); win[i] = t * t; }
    for (int j = 0; j < VE_BANDS; j++) {
      const int n = kBandEnd[j]; float tot = 0;
      for (int i = 0; i < 8; i++) bandwin[j][i] = 0.f;
      for (int i = 0; i < n; i++) { bandwin[j][i] = (float)sin((i + .5) / n * M_PI); tot += bandwin[j][i]; }
      bandtotal[j] = (float)(1. / tot);
    }
    int bitrev[VE_N / 4];
    vb200_exact_tables_mdct(VE_N, T, bitrev);
    for (int i = 0; i < VE_N / 8; i++)       /* ve_mdct128.cuh has the bit reversal in closed form */
      if (bitrev[2 * i] != ve128::rev0(i) || bitrev[2 * i + 1] != ve128::rev1(i)) abort();
  }
};
const Detector &detector() { static Detector d; return d; }

/* ---- the detector step, split into its four parts (envelope.c:101-217) ---- */
struct NearDC { float ring[VE_NEARDC]; float acc, partial; int ptr; };
struct BandFilter { float ampbuf[VE_AMP]; int ampptr; };

/* (1) windowed 128-point MDCT (envelope.c:139-186, scalar path: vec[i] = data[i]*mdct_win[i]; mdct_forward) */
inline void ve_mdct(const Detector &D, const float *data, float *vec) {
  ve128::forward([&](int t) { return VE_FM(data[t], D.win[t]); }, vec, D.T);
}
__host__ __device__ inline float ve_neartemp(float v0, float v1, float v2) {
  return (float)((double)(v0 * v0) + .7 * (double)v1 * (double)v1 + .2 * (double)v2 * (double)v2);
}
/* (2) near-DC running sum with its periodic refresh -> decay */
__host__ __device__ inline float ve_neardc(NearDC *s, float temp) {
  float decay;
  const int ptr = s->ptr;
  if (ptr == 0) { decay = s->acc = s->partial + temp; s->partial = temp; }
  else { decay = s->acc += temp; s->partial += temp; }
  s->acc -= s->ring[ptr];
  s->ring[ptr] = temp;
  decay = (float)((double)decay * (1. / (VE_NEARDC + 1)));
  s->ptr++;
  if (s->ptr >= VE_NEARDC) s->ptr = 0;
  return (float)((double)todB_hd(decay) * .5 - (double)15.f);
}
/* (3) spread/limit the spectrum and take the 12 band amplitudes */
inline void ve_bands(const Detector &D, const float *vec, float decay, float minV, float *acc_out) {
  float v2[VE_N / 4];
  for (int i = 0; i < VE_N / 2; i += 2) {
    float val = vec[i] * vec[i] + vec[i + 1] * vec[i + 1];
    val = todB_hd(val) * .5f;
    if (val < decay) val = decay;
    if (val < minV) val = minV;
    v2[i >> 1] = val;
    decay = (float)((double)decay - 8.);
  }
  for (int j = 0; j < VE_BANDS; j++) {
    float acc = 0.f;
    for (int i = 0; i < kBandEnd[j]; i++) acc += v2[i + kBandBegin[j]] * D.bandwin[j][i];
    acc_out[j] = acc * D.bandtotal[j];
  }
}
/* (4) pre/post-echo trigger of one band */
__host__ __device__ inline int ve_trigger(BandFilter *f, float acc, int stretch, float penalty, float pre_th, float post_th) {
  const int cur = f->ampptr;
  int p = cur - 1; if (p < 0) p += VE_AMP;
  const float postmax = acc > f->ampbuf[p] ? acc : f->ampbuf[p], postmin = acc < f->ampbuf[p] ? acc : f->ampbuf[p];
  float premax = -99999.f, premin = 99999.f;
  for (int i = 0; i < stretch; i++) {
    p--; if (p < 0) p += VE_AMP;
    if (f->ampbuf[p] > premax) premax = f->ampbuf[p];
    if (f->ampbuf[p] < premin) premin = f->ampbuf[p];
  }
  const float valmin = postmin - premin, valmax = postmax - premax;
  f->ampbuf[cur] = acc;
  f->ampptr++; if (f->ampptr >= VE_AMP) f->ampptr = 0;
  int ret = 0;
  if (valmax > pre_th + penalty) ret |= 1 | 4;
  if (valmin < post_th - penalty) ret |= 2;
  return ret;
}
__host__ __device__ inline void ve_stretch_params(int stretch_state, float stretch_penalty, int *stretch, float *penalty) {
  *stretch = stretch_state / 2 > VE_MINSTRETCH ? stretch_state / 2 : VE_MINSTRETCH;
  float p = stretch_penalty - (float)(stretch_state / 2 - VE_MINSTRETCH);
  if (p < 0.f) p = 0.f;
  if (p > stretch_penalty) p = stretch_penalty;
  *penalty = p;
}

struct EnvParams { float stretch_penalty, minenergy, pre_amp; float pre_th[VE_BANDS], post_th[VE_BANDS]; };

bool env_params(const vb200_setup *s, EnvParams *e) {
  uint32_t cnt = 0;
  const float *gf = (const float *)vb200_blob_find(s->blob, s->blob_bytes, "psyg.f32", nullptr, &cnt);
  if (!gf || cnt < 4 + 2 * VE_BANDS) return false;
  e->stretch_penalty = gf[1]; e->minenergy = gf[2]; e->pre_amp = gf[3];
  for (int i = 0; i < VE_BANDS; i++) { e->pre_th[i] = gf[4 + i]; e->post_th[i] = gf[4 + VE_BANDS + i]; }
  return true;
}

/* One stream on its timeline: LPC head/tail, detector state, window sequencing. */
struct StreamSeg {
  const vb200_setup *s; EnvParams env;
  int ch, bs0, bs1, stream_index;
  long lead, nsamples, T, eof_abs;
  float *pcm;                       /* [ch][T], owned by the caller */
  std::vector<unsigned char> mark;
  std::vector<NearDC> near; std::vector<BandFilter> filt;
  int stretch = 0;
  long next_step = 0, ve_current = 0;
  int lW = 0, W = 0;
  long centerW, cursor, curmark = 0;
  int64_t granulepos = 0;
  bool finished = false;
  std::vector<vb200_block> blocks;
  long pcm_base = 0;               /* offset of this stream's timeline inside the batch buffer */

  void init(const vb200_setup *s_, const EnvParams &e, long nsamples_, float *timeline, int index) {
    s = s_; env = e; ch = s->h.channels; bs0 = s->h.bs[0]; bs1 = s->h.bs[1]; stream_index = index;
    lead = bs1 / 2; nsamples = nsamples_; T = lead + nsamples + 3L * bs1; eof_abs = lead + nsamples; pcm = timeline;
    mark.assign((size_t)(T / VE_STEP) + VE_WIN + VE_POST + 4, 0);
    near.assign(ch, NearDC()); filt.assign((size_t)ch * VE_BANDS, BandFilter());
    memset(near.data(), 0, sizeof(NearDC) * near.size()); memset(filt.data(), 0, sizeof(BandFilter) * filt.size());
    centerW = lead; cursor = bs1 / 2;
  }
  static long timeline_len(const vb200_setup *s, long nsamples) { return s->h.bs[1] / 2 + nsamples + 3L * s->h.bs[1]; }

  /* timeline: lead-in | pre-amplified PCM | tail; head = reverse LPC extrapolation from the signal that
     had arrived when the reference first had more than one long block (block.c:438-471,545-548) */
  /* SRC = float (the reference's own input format) or short (16-bit PCM as a WAV file holds it: the value is
     sample / 32768.f, the conversion examples/encoder_example.c:183-190 does before vorbis_analysis_wrote) */
  static inline float sample_of(float v) { return v; }
  static inline float sample_of(short v) { return (float)v / 32768.f; }
  template <typename SRC>
  void fill_and_head(const SRC *src, long stride, int chunk, bool sparse = false) {
    /* sparse: the body of the signal goes to the device straight from the caller's (pinned) memory and
       only the two windows the host itself reads are staged: the head the LPC looks at and the tail the
       end-of-stream extrapolation and the last detector steps look at */
    long head_n = nsamples, tail_n = 0;
    if (sparse) {
      long sn = 0;
      while (sn < nsamples && !(sn > bs1)) sn += (nsamples - sn < chunk ? nsamples - sn : chunk);
      const long want_tail = bs1 + (VE_WIN + 4) * VE_STEP;
      if (sn + want_tail < nsamples) { head_n = sn; tail_n = want_tail; }
    }
    for (int c = 0; c < ch; c++) {
      float *d = pcm + (size_t)c * T; const SRC *q = src + (size_t)c * stride;
      for (long i = 0; i < lead; i++) d[i] = 0.f;
      for (long i = 0; i < head_n; i++) d[lead + i] = sample_of(q[i]) * env.pre_amp;
      for (long i = nsamples - tail_n; i < nsamples; i++) d[lead + i] = sample_of(q[i]) * env.pre_amp;
      for (long i = eof_abs; i < T; i++) d[i] = 0.f;
    }
    long seen = 0;
    while (seen < nsamples && !(seen > bs1)) seen += (nsamples - seen < chunk ? nsamples - seen : chunk);
    const long cur = lead + seen;
    if (cur - lead > 32) {
      std::vector<float> work(cur); float lpc[16];
      for (int c = 0; c < ch; c++) {
        float *d = pcm + (size_t)c * T;
        for (long j = 0; j < cur; j++) work[j] = d[cur - j - 1];
        lpc_from_data(work.data(), lpc, (int)(cur - lead), 16);
        lpc_predict(lpc, work.data() + cur - lead - 16, 16, work.data() + cur - lead, lead);
        for (long j = 0; j < cur; j++) d[cur - j - 1] = work[j];
      }
    }
  }
  /* tail: forward LPC extrapolation of 3 long blocks (block.c:497-537) from the signal still held in the
     reference's sliding buffer, i.e. back to centerW - blocksize[1]/2 */
  void tail() {
    float lpc[32];
    const long held = eof_abs - (centerW - lead);
    for (int c = 0; c < ch; c++) {
      float *d = pcm + (size_t)c * T;
      if (held > 64) {
        long n = held; if (n > bs1) n = bs1;
        lpc_from_data(d + eof_abs - n, lpc, (int)n, 32);
        lpc_predict(lpc, d + eof_abs - 32, 32, d + eof_abs, T - eof_abs);
      }
    }
  }
  long steps_complete(long avail) const { const long l = avail / VE_STEP - VE_WIN; return l > 0 ? l : 0; }

  void apply_ret(long j, int ret) {
    mark[j + VE_POST] = 0;
    if (ret & 1) { mark[j] = 1; mark[j + 1] = 1; }
    if (ret & 2) { mark[j] = 1; if (j > 0) mark[j - 1] = 1; }
    if (ret & 4) stretch = -1;
  }
  /* detector over every complete 64-sample step (envelope.c:569-616), host evaluation */
  void detect_host(long avail) {
    const Detector &D = detector();
    const long last = steps_complete(avail);
    for (long j = next_step; j < last; j++) {
      int ret = 0;
      stretch++;
      if (stretch > VE_MAXSTRETCH * 2) stretch = VE_MAXSTRETCH * 2;
      int st; float pen;
      ve_stretch_params(stretch, env.stretch_penalty, &st, &pen);
      for (int c = 0; c < ch; c++) {
        float vec[VE_N / 2], acc[VE_BANDS];
        ve_mdct(D, pcm + (size_t)c * T + VE_STEP * j, vec);
        const float decay = ve_neardc(&near[c], ve_neartemp(vec[0], vec[1], vec[2]));
        ve_bands(D, vec, decay, env.minenergy, acc);
        for (int b = 0; b < VE_BANDS; b++) ret |= ve_trigger(&filt[(size_t)c * VE_BANDS + b], acc[b], st, pen, env.pre_th[b], env.post_th[b]);
      }
      apply_ret(j, ret);
    }
    if (last > next_step) next_step = last;
    if (last > 0) ve_current = last * VE_STEP;
  }
  /* the detector ran elsewhere (device) for steps [0,last): adopt its marks and its final state */
  void adopt(long last, const unsigned char *marks, int stretch_final, const float *acc_tail /*[ch][17][12]*/,
             const float *temp_tail /*[ch][45]*/) {
    for (long j = 0; j < last + VE_POST && j < (long)mark.size(); j++) mark[j] = marks[j];
    stretch = stretch_final;
    for (int c = 0; c < ch; c++) {
      for (int b = 0; b < VE_BANDS; b++) {
        BandFilter &f = filt[(size_t)c * VE_BANDS + b];
        for (int k = 0; k < VE_AMP; k++) f.ampbuf[k] = 0.f;
        for (long sidx = last - VE_AMP; sidx < last; sidx++) if (sidx >= 0) f.ampbuf[sidx % VE_AMP] = acc_tail[((size_t)c * VE_AMP + (sidx - (last - VE_AMP))) * VE_BANDS + b];
        f.ampptr = (int)(last % VE_AMP);
      }
      /* near-DC state: start from the last refresh step r0 that has a full cycle behind it (the state at
         a refresh depends only on the 15 terms before it), then replay up to `last` */
      NearDC &n = near[c];
      memset(&n, 0, sizeof(n));
      auto T_ = [&](long step) { return step >= 0 ? temp_tail[(size_t)c * 45 + (step - (last - 45))] : 0.f; };
      const long r0 = last >= VE_NEARDC ? ((last - VE_NEARDC) / VE_NEARDC) * VE_NEARDC : 0;
      if (r0 >= VE_NEARDC) {
        for (int k = 0; k < VE_NEARDC; k++) n.ring[k] = T_(r0 - VE_NEARDC + k);
        n.partial = T_(r0 - VE_NEARDC);
        for (long u = r0 - VE_NEARDC + 1; u < r0; u++) n.partial += T_(u);
      }
      n.ptr = 0;
      for (long sidx = r0; sidx < last; sidx++) (void)ve_neardc(&n, T_(sidx));
    }
    next_step = last;
    if (last > 0) ve_current = last * VE_STEP;
  }

  /* window sequencing (block.c:557-812, envelope.c:617-706) on the absolute timeline */
  void sequence(long avail, bool eof) {
    const int bsz[2] = {bs0, bs1};
    while (!finished) {
      int nW;
      {
        long bp = -1;
        const long testW = centerW + bsz[W] / 4 + bs1 / 2 + bs0 / 4;
        long j = cursor;
        while (j < ve_current - VE_STEP) {
          if (j >= testW) { bp = 1; break; }
          cursor = j;
          if (mark[j / VE_STEP] && j > centerW) { curmark = j; bp = (j >= testW) ? 1 : 0; break; }
          j += VE_STEP;
        }
        if (bp == -1) { if (!eof) return; nW = 0; }
        else nW = (bs0 == bs1) ? 0 : (int)bp;
      }
      const long centerNext = centerW + bsz[W] / 4 + bsz[nW] / 4;
      if (avail < centerNext + bsz[nW] / 2) return;
      vb200_block blk;
      memset(&blk, 0, sizeof(blk));
      blk.stream = stream_index; blk.W = W; blk.lW = lW; blk.nW = nW;
      if (W) blk.blocktype = (!lW || !nW) ? 0 : 1;
      else {
        const long beginW = centerW - bs0 / 4 - bs0 / 4, endW = centerW + bs0 / 4 + bs0 / 4;
        int hit = (curmark >= beginW && curmark < endW);
        if (!hit) for (long i = beginW / VE_STEP; i < endW / VE_STEP; i++) if (i >= 0 && mark[i]) { hit = 1; break; }
        blk.blocktype = hit ? 0 : 1;
      }
      blk.granulepos = granulepos;
      blk.pcm_offset = pcm_base + centerW - bsz[W] / 2;
      blk.pcm_stride = T;
      if (eof && centerW >= eof_abs) { blk.eofflag = 1; blocks.push_back(blk); finished = true; return; }
      blocks.push_back(blk);
      const long movement = centerNext - centerW;
      if (eof && centerNext >= eof_abs) granulepos += movement - (centerNext - eof_abs);
      else granulepos += movement;
      lW = W; W = nW; centerW = centerNext;
    }
  }
};

}  // namespace

/* ================================================================ host-only driver */
struct vb200_segments {
  int channels; long total;
  std::vector<float> pcm;
  std::vector<vb200_block> blocks;
};

extern "C" int vb200_segment_stream(const vb200_setup *s, const float *pcm, long nsamples, long stride, int chunk,
                                    int stream_index, vb200_segments **out) {
  if (!s || !pcm || !out || nsamples < 0) { vb200_set_error("vb200_segment_stream: bad arguments"); return VB200_EINVAL; }
  EnvParams env;
  if (!env_params(s, &env)) { vb200_set_error("vb200_segment_stream: setup blob lacks the envelope parameters"); return VB200_EINVAL; }
  if (chunk <= 0) chunk = 1024;
  vb200_segments *g = new vb200_segments();
  g->channels = s->h.channels; g->total = StreamSeg::timeline_len(s, nsamples);
  g->pcm.assign((size_t)g->channels * g->total, 0.f);
  StreamSeg q;
  q.init(s, env, nsamples, g->pcm.data(), stream_index);
  q.fill_and_head(pcm, stride, chunk);
  q.detect_host(q.eof_abs); q.sequence(q.eof_abs, false);      /* all real signal arrived, no EOS yet */
  q.tail();
  q.detect_host(q.T); q.sequence(q.T, true);                    /* end of stream */
  g->blocks.swap(q.blocks);
  *out = g;
  return VB200_OK;
}

/* A/B hook: the detector's windowed 128-point transform of one step (envelope.c:139-186), host code */
extern "C" int vb200_stage_detector_mdct(const float *pcm128, float *coef64) {
  if (!pcm128 || !coef64) { vb200_set_error("vb200_stage_detector_mdct: bad arguments"); return VB200_EINVAL; }
  ve_mdct(detector(), pcm128, coef64);
  return VB200_OK;
}

extern "C" void vb200_segments_free(vb200_segments *g) { delete g; }
extern "C" int vb200_segments_count(const vb200_segments *g) { return g ? (int)g->blocks.size() : VB200_EINVAL; }
extern "C" const vb200_block *vb200_segments_blocks(const vb200_segments *g) { return g ? g->blocks.data() : nullptr; }
extern "C" const float *vb200_segments_pcm(const vb200_segments *g, long *elems) { if (!g) return nullptr; if (elems) *elems = (long)g->pcm.size(); return g->pcm.data(); }

/* ================================================================ device half of the detector */
namespace {

struct EnvDev {            /* parameters in HBM */
  float minenergy, stretch_penalty, pre_th[VE_BANDS], post_th[VE_BANDS];
};
/* tables every detector thread reads at compile-time indices: constant bank (they depend on nothing but the
   detector's fixed geometry, so one copy per device serves every setup) */
__constant__ float c_ve_T[VE_N + VE_N / 4];      /* mdct_init(128) */
__constant__ float c_ve_win[VE_N];
__constant__ float c_ve_bandwin[VE_BANDS][8];
__constant__ float c_ve_bandtotal[VE_BANDS];
__host__ __device__ constexpr int band_begin(int b) { constexpr int t[VE_BANDS] = {2, 4, 6, 9, 13, 17, 22, 12, 8, 3, 2, 1}; return t[b]; }
__host__ __device__ constexpr int band_end(int b) { constexpr int t[VE_BANDS] = {4, 5, 6, 8, 8, 8, 8, 4, 4, 3, 2, 4}; return t[b]; }

struct EnvStream { long long pcm_base; long long T; int nsteps; long long out_base; /* step-major output offset */ long long eof;
                   long long s16_base; /* 16-bit ingest: the stream's samples in the device staging buffer, [ch][eof - lead] */ };

constexpr int EV_TILE = 128;     /* steps per CTA */
constexpr int EV_HALO = 30;      /* the near-DC sum of a step depends on at most the 30 previous steps */
constexpr int EV_THREADS = 160;  /* one thread per transformed step: halo + tile (two spare) */
constexpr int EV_ROW = EV_THREADS + 1;   /* row stride of the transposed sample tile: odd, conflict-free both ways */

/* K0a: per (stream, channel, tile of 128 steps): one thread per step does the whole 128-point MDCT of the reference
   (ve_mdct128.cuh) on registers -- every sample of the step is read once, from a tile of the timeline that the CTA
   stages transposed (sample s at row s mod 64, column s div 64: lanes = consecutive steps read consecutive words) --
   then the near-DC decay (replayed from the 30 halo steps the CTA transforms as well), the spread/limit pass and
   the 12 band amplitudes, all in the reference's float operations (envelope.c:139-216, scalar path).  Everything
   here is a pure function of the PCM, so tiles are independent. */
__global__ void __launch_bounds__(EV_THREADS, 2)
k_env_bands(const EnvDev *__restrict__ E, const EnvStream *__restrict__ streams, int ch, const float *__restrict__ pcm,
            float *__restrict__ acc_out /*[stream][ch][step][12]*/, float *__restrict__ temp_out /*[stream][ch][step]*/) {
  __shared__ float tile[VE_STEP * EV_ROW];
  __shared__ float temps[EV_THREADS];
  const EnvStream S = streams[blockIdx.z];
  const int c = blockIdx.y, tid = threadIdx.x;
  const int j0 = blockIdx.x * EV_TILE;
  if (j0 >= S.nsteps) return;
  const int jn = min(EV_TILE, S.nsteps - j0);
  const int h0 = max(0, j0 - EV_HALO);                      /* first step that is transformed here */
  const int nh = j0 + jn - h0, toff = j0 - h0;
  const float *src = pcm + S.pcm_base + (long long)c * S.T + (long long)h0 * VE_STEP;
  for (int i = tid; i < (nh + 1) * VE_STEP; i += EV_THREADS) tile[(i & (VE_STEP - 1)) * EV_ROW + (i >> 6)] = __ldg(src + i);
  __syncthreads();
  float vec[VE_N / 2];
  if (tid < nh) {
    ve128::forward([&](int t) { return __fmul_rn(tile[(t & (VE_STEP - 1)) * EV_ROW + tid + (t >> 6)], c_ve_win[t]); }, vec, c_ve_T);
    temps[tid] = ve_neartemp(vec[0], vec[1], vec[2]);
  }
  __syncthreads();
  if (tid < toff || tid >= nh) return;
  const int j = j0 + tid - toff;
  float dec;
  {
    /* decay: replay the refresh cycle the step sits in (<= 44 float ops) */
    const int ptr = j % VE_NEARDC, r = j - ptr;      /* r = last refresh step */
    auto T_ = [&](int step) { return step >= 0 ? temps[step - h0] : 0.f; };
    float partial = 0.f;
    if (r >= VE_NEARDC) { partial = T_(r - VE_NEARDC); for (int w = r - VE_NEARDC + 1; w < r; w++) partial = __fadd_rn(partial, T_(w)); }
    float acc = __fadd_rn(partial, T_(r));
    dec = acc;
    acc = __fsub_rn(acc, r >= VE_NEARDC ? T_(r - VE_NEARDC) : 0.f);
    for (int w = r + 1; w <= j; w++) {
      acc = __fadd_rn(acc, T_(w)); dec = acc;
      acc = __fsub_rn(acc, w >= VE_NEARDC ? T_(w - VE_NEARDC) : 0.f);
    }
    dec = (float)((double)dec * (1. / (VE_NEARDC + 1)));
    dec = (float)((double)todB_hd(dec) * .5 - (double)15.f);
    temp_out[S.out_base * ch + (long long)c * S.nsteps + j] = temps[tid];
  }
  /* spread / limit (envelope.c:189-197): pair i of the spectrum -> v2[i], overwriting vec[i] as the reference does */
  const float minV = E->minenergy;
#pragma unroll
  for (int i = 0; i < VE_N / 4; i++) {
    float val = __fadd_rn(__fmul_rn(vec[2 * i], vec[2 * i]), __fmul_rn(vec[2 * i + 1], vec[2 * i + 1]));
    val = __fmul_rn(todB_hd(val), .5f);
    if (val < dec) val = dec;
    if (val < minV) val = minV;
    vec[i] = val;
    dec = (float)((double)dec - 8.);
  }
  float *ao = acc_out + (S.out_base * ch + (long long)c * S.nsteps + j) * VE_BANDS;
#pragma unroll
  for (int b = 0; b < VE_BANDS; b++) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < band_end(b); i++) acc = __fadd_rn(acc, __fmul_rn(vec[i + band_begin(b)], c_ve_bandwin[b][i]));
    ao[b] = __fmul_rn(acc, c_ve_bandtotal[b]);
  }
}

/* K0b: the trigger of one band at step j reads the band's own amplitude history (a pure function of the PCM)
   and the stretch counter, which is the only state carried from step to step.  The counter selects one of
   11 (window, penalty) pairs, so the answer for all 11 is tabulated here, one thread per step, and the
   serial walk below only looks its state up.  Word layout: bits 2v / 2v+1 = pre / post trigger of variant v
   (window v+2 steps), OR-ed over channels and bands. */
constexpr int EV_VARIANTS = VE_MAXSTRETCH - VE_MINSTRETCH + 1;
constexpr int PIPE_SEG_STREAM_SLACK = 64;      /* spare table rows of the parallel trigger */
/* The 14 steps x 12 bands a step looks at are ONE run of 168 consecutive floats of acc ([step][band]), and the next
   step's run starts 12 floats later: a CTA stages the 141 steps its 128 threads need (coalesced) with a row stride of
   13 words -- consecutive steps then sit in different banks -- and every thread reads its history from there.
   (Every one of the 336 history reads per thread used to be a global load whose 32 lanes touched 12 cache lines.) */
constexpr int EF_HIST = VE_MAXSTRETCH + 1;        /* steps before j that a trigger reads: j-1 .. j-13 */
constexpr int EF_ROW = VE_BANDS + 1;
__global__ void __launch_bounds__(128)
k_env_flags(const EnvDev *__restrict__ E, const EnvStream *__restrict__ streams, int ch, const float *__restrict__ acc_in,
            unsigned *__restrict__ flags /*[out_base + step]*/) {
  __shared__ float hist[(128 + EF_HIST) * EF_ROW];
  const EnvStream S = streams[blockIdx.y];
  const int j0 = blockIdx.x * 128, tid = threadIdx.x;
  if (j0 >= S.nsteps) return;
  const int j = j0 + tid;
  const int jn = min(128, S.nsteps - j0);
  float pen[EV_VARIANTS];
#pragma unroll
  for (int v = 0; v < EV_VARIANTS; v++) { int st; ve_stretch_params(2 * (v + VE_MINSTRETCH), E->stretch_penalty, &st, &pen[v]); }
  unsigned w = 0;
  for (int c = 0; c < ch; c++) {
    const float *a = acc_in + (S.out_base * ch + (long long)c * S.nsteps) * VE_BANDS;
    __syncthreads();                                     /* the previous channel's history is no longer read */
    /* rows r = 0 .. jn + 12 hold steps j0 - 13 + r; steps before the stream's first are zero (envelope.c: the
       filters start from a zeroed history) */
    for (int e = tid; e < (jn + EF_HIST) * VE_BANDS; e += 128) {
      const int r = e / VE_BANDS, bnd = e - r * VE_BANDS, step = j0 - EF_HIST + r;
      hist[r * EF_ROW + bnd] = step >= 0 ? __ldg(a + (long long)step * VE_BANDS + bnd) : 0.f;
    }
    __syncthreads();
    if (tid < jn) {
      const float *h = hist + (tid + EF_HIST) * EF_ROW;   /* row of step j; step j - k is k rows up */
#pragma unroll
      for (int b = 0; b < VE_BANDS; b++) {
        const float cur = h[b], prev = h[b - EF_ROW];
        const float postmax = cur > prev ? cur : prev, postmin = cur < prev ? cur : prev;
        const float pre_th = E->pre_th[b], post_th = E->post_th[b];
        float premax = -99999.f, premin = 99999.f;
#pragma unroll
        for (int i = 0; i < VE_MAXSTRETCH; i++) {
          const float x = h[b - (2 + i) * EF_ROW];
          if (x > premax) premax = x;
          if (x < premin) premin = x;
          if (i + 1 >= VE_MINSTRETCH) {
            const int v = i + 1 - VE_MINSTRETCH;
            const float valmin = __fsub_rn(postmin, premin), valmax = __fsub_rn(postmax, premax);
            if (valmax > __fadd_rn(pre_th, pen[v])) w |= 1u << (2 * v);
            if (valmin < __fsub_rn(post_th, pen[v])) w |= 2u << (2 * v);
          }
        }
      }
    }
  }
  if (tid < jn) flags[S.out_base + j] = w;
}

/* K0c: the order-dependent half: the stretch counter, one warp per stream, 32 steps fetched per pass.
   Writes the marks of steps [0,nsteps] (the buffer is zeroed beforehand) and the final stretch. */
__global__ void __launch_bounds__(32)
k_env_trigger(const EnvStream *__restrict__ streams, const unsigned *__restrict__ flags,
              unsigned char *__restrict__ marks, const long long *__restrict__ mark_base, int *__restrict__ stretch_out) {
  const EnvStream S = streams[blockIdx.x];
  const int lane = threadIdx.x;
  unsigned char *mk = marks + mark_base[blockIdx.x];
  int stretch = 0;
  unsigned carry = 0;                                   /* a pre-echo hit at step j also marks j+1 */
  for (int j0 = 0; j0 < S.nsteps; j0 += 32) {
    const int n = min(32, S.nsteps - j0);
    const unsigned w = lane < n ? flags[S.out_base + j0 + lane] : 0u;
    unsigned pre = 0, post = 0;
#pragma unroll 8
    for (int k = 0; k < n; k++) {
      const unsigned wk = __shfl_sync(FULLMASK, w, k);
      stretch = min(stretch + 1, VE_MAXSTRETCH * 2);
      const int v = max(stretch >> 1, VE_MINSTRETCH) - VE_MINSTRETCH;
      const unsigned bits = (wk >> (2 * v)) & 3u;
      pre |= (bits & 1u) << k; post |= (bits >> 1) << k;
      if (bits & 1u) stretch = -1;
    }
    const unsigned m = pre | post | (pre << 1) | (post >> 1) | carry;
    if ((m >> lane) & 1u) mk[j0 + lane] = 1;
    if (lane == 0 && (post & 1u) && j0 > 0) mk[j0 - 1] = 1;
    carry = (pre >> (n - 1)) & 1u;
    if (n < 32 && lane == 0 && carry) mk[j0 + n] = 1;
  }
  if (lane == 0) {
    if (carry && (S.nsteps & 31) == 0 && S.nsteps > 0) mk[S.nsteps] = 1;
    stretch_out[blockIdx.x] = stretch;
  }
}

/* K0c, parallel form.  The counter has 26 states (u = stretch + 1 = 0..25) and a step maps u to u' as a function of the
   step's flag word only, so a chunk of 32 steps is a 26-entry table and tables compose:
     k_env_maps   one thread per chunk: the table of its 32 steps (row = 32 bytes: entries 0..25);
     k_env_walk   one CTA per stream: warps compose 32 tables each into a super-table, thread 0 walks the (few dozen)
                  super-tables, warps walk their chunks from the states that gives (start state -> byte 31 of the row),
                  threads replay their chunk's 32 steps for the pre/post-echo hits and write the marks.
   ~40 us instead of the 1.4 ms a single warp needs to walk a 60 s stream; same marks, same final counter. */
__device__ __forceinline__ int env_step(int u, unsigned w, unsigned *bits) {
  const int s1 = u < VE_MAXSTRETCH * 2 ? u : VE_MAXSTRETCH * 2;            /* stretch after the increment */
  const int v = max(s1 >> 1, VE_MINSTRETCH) - VE_MINSTRETCH;
  *bits = (w >> (2 * v)) & 3u;
  return (*bits & 1u) ? 0 : s1 + 1;                                          /* a pre-echo hit sets stretch = -1 */
}
__host__ __device__ inline long long env_chunk_base(long long out_base, int stream) { return (out_base >> 5) + stream; }

__global__ void __launch_bounds__(128)
k_env_maps(const EnvStream *__restrict__ streams, const unsigned *__restrict__ flags, unsigned char *__restrict__ maps) {
  const EnvStream S = streams[blockIdx.y];
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int nchunks = (S.nsteps + 31) >> 5;
  if (c >= nchunks) return;
  const int n = min(32, S.nsteps - c * 32);
  unsigned w[32];
#pragma unroll
  for (int k = 0; k < 32; k++) w[k] = k < n ? flags[S.out_base + (long long)c * 32 + k] : 0u;
  unsigned out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int u0 = 0; u0 <= VE_MAXSTRETCH * 2 + 1; u0++) {
    int u = u0;
#pragma unroll
    for (int k = 0; k < 32; k++) { unsigned b; const int nu = env_step(u, w[k], &b); u = k < n ? nu : u; }
    out[u0 >> 2] |= (unsigned)u << (8 * (u0 & 3));
  }
  uint4 *row = reinterpret_cast<uint4 *>(maps + (env_chunk_base(S.out_base, blockIdx.y) + c) * 32);
  row[0] = make_uint4(out[0], out[1], out[2], out[3]);
  row[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

constexpr int EW_THREADS = 1024, EW_MAXSUPER = 1024;      /* up to 1024 * 32 * 32 steps (~23 min at 48 kHz) per stream; longer streams take the serial walk */
__global__ void __launch_bounds__(EW_THREADS)
k_env_walk(const EnvStream *__restrict__ streams, const unsigned *__restrict__ flags, unsigned char *__restrict__ maps,
           unsigned char *__restrict__ marks, const long long *__restrict__ mark_base, int *__restrict__ stretch_out) {
  __shared__ unsigned char super[EW_MAXSUPER][32];
  const EnvStream S = streams[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (S.nsteps + 31) >> 5, nsuper = (nchunks + 31) >> 5;
  unsigned char *rows = maps + env_chunk_base(S.out_base, blockIdx.x) * 32;
  unsigned char *mk = marks + mark_base[blockIdx.x];
  for (int sc = warp; sc < nsuper; sc += EW_THREADS / 32) {
    int x = lane;
    const int c1 = min(nchunks, sc * 32 + 32);
    if (lane <= VE_MAXSTRETCH * 2 + 1) for (int c = sc * 32; c < c1; c++) x = rows[(size_t)c * 32 + x];
    super[sc][lane] = (unsigned char)x;
  }
  __syncthreads();
  if (tid == 0) {
    int u = 1;                                        /* stretch = 0 */
    for (int sc = 0; sc < nsuper; sc++) { const int nu = super[sc][u]; super[sc][31] = (unsigned char)u; u = nu; }
    stretch_out[blockIdx.x] = u - 1;
  }
  __syncthreads();
  for (int sc = warp; sc < nsuper; sc += EW_THREADS / 32) {
    if (lane == 0) {
      int x = super[sc][31];
      const int c1 = min(nchunks, sc * 32 + 32);
      for (int c = sc * 32; c < c1; c++) { const int nx = rows[(size_t)c * 32 + x]; rows[(size_t)c * 32 + 31] = (unsigned char)x; x = nx; }
    }
  }
  __syncthreads();
  /* replay: pre / post hits of every chunk, left in the first two words of its row (the table is dead by now) */
  for (int c = tid; c < nchunks; c += EW_THREADS) {
    int u = rows[(size_t)c * 32 + 31];
    const int n = min(32, S.nsteps - c * 32);
    unsigned pre = 0, post = 0;
    for (int k = 0; k < n; k++) {
      unsigned b;
      u = env_step(u, flags[S.out_base + (long long)c * 32 + k], &b);
      pre |= (b & 1u) << k; post |= (b >> 1) << k;
    }
    unsigned *r32 = reinterpret_cast<unsigned *>(rows + (size_t)c * 32);
    r32[0] = pre; r32[1] = post;
  }
  __syncthreads();
  /* marks (envelope.c:601-614): a hit marks its own step; a pre-echo hit also the next, a post-echo hit the one before */
  for (int c = tid; c < nchunks; c += EW_THREADS) {
    const unsigned *r32 = reinterpret_cast<const unsigned *>(rows + (size_t)c * 32);
    const unsigned pre = r32[0], post = r32[1];
    const unsigned ppre = c > 0 ? r32[-8] : 0u, npost = c + 1 < nchunks ? r32[9] : 0u;
    const unsigned m = pre | post | (pre << 1) | (ppre >> 31) | (post >> 1) | (npost << 31);
    for (unsigned mm = m; mm; mm &= mm - 1) { const int j = c * 32 + __ffs(mm) - 1; if (j <= S.nsteps) mk[j] = 1; }
    if (c == nchunks - 1 && (S.nsteps & 31) == 0 && (pre >> 31)) mk[S.nsteps] = 1;
  }
}

/* direct upload: the caller's samples already sit at [lead, eof) of every channel row; zero the lead-in and
   the tail and apply the setup's pre-amplitude in place */
__global__ void k_timeline_finish(const EnvStream *__restrict__ streams, int ch, int lead, float pre_amp, float *__restrict__ pcm) {
  const EnvStream S = streams[blockIdx.y / ch];
  float *d = pcm + S.pcm_base + (long long)(blockIdx.y % ch) * S.T;
  const bool scale = pre_amp != 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S.T; i += (long long)gridDim.x * blockDim.x) {
    if (i < lead || i >= S.eof) d[i] = 0.f;
    else if (scale) d[i] = __fmul_rn(d[i], pre_amp);
  }
}

/* 16-bit ingest: the caller's samples went up as they are (half the bytes of the float path); this writes them into
   the timeline as sample / 32768.f -- an exact float operation, the one the reference's caller performs on the host */
__global__ void k_timeline_from_s16(const EnvStream *__restrict__ streams, int ch, int lead, const short *__restrict__ src, float *__restrict__ pcm) {
  const EnvStream S = streams[blockIdx.y / ch];
  const int c = blockIdx.y % ch;
  const long long n = S.eof - lead;
  float *d = pcm + S.pcm_base + (long long)c * S.T + lead;
  const short *q = src + S.s16_base + (long long)c * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = __fdiv_rn((float)q[i], 32768.f);
}

/* the extrapolated tails (a few thousand samples per channel) are fetched by a kernel straight from the pinned host
   timeline: a copy-engine transfer would queue behind the bulk uploads other segmenters have already issued */
__global__ void k_tail_upload(const EnvStream *__restrict__ streams, int ch, const float *__restrict__ host_pcm, float *__restrict__ pcm) {
  const EnvStream S = streams[blockIdx.x / ch];
  const long long o = S.pcm_base + (long long)(blockIdx.x % ch) * S.T + S.eof;
  const int n = (int)(S.T - S.eof);
  for (int i = threadIdx.x; i < n; i += blockDim.x) pcm[o + i] = host_pcm[o + i];
}

/* the last 17 band amplitudes and 45 near-DC terms of every (stream, channel): what the host needs to
   continue the detector over the extrapolated tail */
__global__ void k_env_tail(const EnvStream *__restrict__ streams, int ch, const float *__restrict__ acc_in, const float *__restrict__ temp_in,
                           float *__restrict__ acc_tail, float *__restrict__ temp_tail) {
  const EnvStream S = streams[blockIdx.x];
  const int c = blockIdx.y;
  const long long o = S.out_base * ch + (long long)c * S.nsteps;
  for (int i = threadIdx.x; i < VE_AMP * VE_BANDS; i += blockDim.x) {
    const int k = i / VE_BANDS, b = i % VE_BANDS;
    const int step = S.nsteps - VE_AMP + k;
    acc_tail[((size_t)blockIdx.x * ch + c) * VE_AMP * VE_BANDS + i] = step >= 0 ? acc_in[(o + step) * VE_BANDS + b] : 0.f;
  }
  for (int i = threadIdx.x; i < 45; i += blockDim.x) {
    const int step = S.nsteps - 45 + i;
    temp_tail[((size_t)blockIdx.x * ch + c) * 45 + i] = step >= 0 ? temp_in[o + step] : 0.f;
  }
}

}  // namespace

struct SegRun {            /* what a begun run keeps until it is finished */
  std::vector<StreamSeg> seg; std::vector<long long> base, mark_base;
  int nstreams, threads; size_t total, nmark; double t0, t1;
};

struct vb200_segmenter {
  vb200_setup *s; EnvParams env; SegRun *run;
  EnvDev *d_env;
  float *h_pcm, *d_pcm; size_t pcm_cap;          /* pinned host timeline + device timeline */
  short *d_pcm16; size_t pcm16_cap;              /* 16-bit ingest: the samples as uploaded */
  float *d_acc, *d_temp; unsigned *d_flags; unsigned char *d_maps; size_t step_cap;
  unsigned char *d_marks, *h_marks; size_t mark_cap;
  EnvStream *d_streams, *h_streams; long long *d_mark_base, *h_mark_base; int *d_stretch, *h_stretch;
  float *d_acc_tail, *h_acc_tail, *d_temp_tail, *h_temp_tail; int stream_cap;
  cudaStream_t st; int direct;
  cudaEvent_t ev_h2d; const vb200_segmenter *after;   /* uploads are ordered behind another segmenter's (see _after) */
  std::vector<vb200_block> blocks; size_t pcm_elems;
};

extern "C" int vb200_segmenter_create(vb200_setup *s, vb200_segmenter **out) {
  if (!s || !out) { vb200_set_error("vb200_segmenter_create: bad arguments"); return VB200_EINVAL; }
  if (s->device < 0 || !s->d) { vb200_set_error("vb200_segmenter_create: setup has no CUDA device (no CPU fallback; use vb200_segment_stream on the host)"); return VB200_EFAULT; }
  CUDA_TRY(cudaSetDevice(s->device));
  vb200_segmenter *g = new vb200_segmenter();
  memset((void *)g, 0, sizeof(vb200_segmenter) - sizeof(std::vector<vb200_block>) - sizeof(size_t));
  g->s = s;
  if (!env_params(s, &g->env)) { delete g; vb200_set_error("vb200_segmenter_create: setup blob lacks the envelope parameters"); return VB200_EINVAL; }
  const Detector &D = detector();
  EnvDev e;
  CUDA_TRY(cudaMalloc(&g->d_env, sizeof(EnvDev)));
  CUDA_TRY(cudaMemcpyToSymbol(c_ve_T, D.T, sizeof(D.T)));
  CUDA_TRY(cudaMemcpyToSymbol(c_ve_win, D.win, sizeof(D.win)));
  CUDA_TRY(cudaMemcpyToSymbol(c_ve_bandwin, D.bandwin, sizeof(D.bandwin)));
  CUDA_TRY(cudaMemcpyToSymbol(c_ve_bandtotal, D.bandtotal, sizeof(D.bandtotal)));
  for (int i = 0; i < VE_BANDS; i++) {
    if (kBandBegin[i] != band_begin(i) || kBandEnd[i] != band_end(i)) { vb200_set_error("vb200_segmenter_create: band tables out of step"); return VB200_EFAULT; }
    e.pre_th[i] = g->env.pre_th[i]; e.post_th[i] = g->env.post_th[i];
  }
  e.minenergy = g->env.minenergy; e.stretch_penalty = g->env.stretch_penalty;
  CUDA_TRY(cudaMemcpy(g->d_env, &e, sizeof(e), cudaMemcpyHostToDevice));
  {
    /* served before any batch stream: the detector kernels gate the host's block-size decision, the batch kernels of
       earlier groups would otherwise keep them waiting for SM slots */
    int least = 0, greatest = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CUDA_TRY(cudaStreamCreateWithPriority(&g->st, cudaStreamNonBlocking, greatest));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&g->ev_h2d, cudaEventDisableTiming));
  *out = g;
  return VB200_OK;
}

extern "C" void vb200_segmenter_destroy(vb200_segmenter *g) {
  if (!g) return;
  cudaSetDevice(g->s->device);
  cudaStreamDestroy(g->st);
  cudaEventDestroy(g->ev_h2d);
  delete g->run;
  cudaFree(g->d_env); cudaFree(g->d_pcm); cudaFree(g->d_pcm16); cudaFreeHost(g->h_pcm);
  cudaFree(g->d_acc); cudaFree(g->d_temp); cudaFree(g->d_flags); cudaFree(g->d_maps); cudaFree(g->d_marks); cudaFreeHost(g->h_marks);
  cudaFree(g->d_streams); cudaFreeHost(g->h_streams); cudaFree(g->d_mark_base); cudaFreeHost(g->h_mark_base);
  cudaFree(g->d_stretch); cudaFreeHost(g->h_stretch); cudaFree(g->d_acc_tail); cudaFreeHost(g->h_acc_tail);
  cudaFree(g->d_temp_tail); cudaFreeHost(g->h_temp_tail);
  delete g;
}

template <typename F> static void parallel_for(int n, int threads, F f) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads <= 1) { for (int i = 0; i < n; i++) f(i); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) th.emplace_back([=]() { for (int i = t; i < n; i += threads) f(i); });
  for (auto &x : th) x.join();
}

#define SEG_ALLOC(ptr, bytes, host) do{ if (host) { CUDA_TRY(cudaMallocHost((void **)&(ptr), (bytes))); } else { CUDA_TRY(cudaMalloc((void **)&(ptr), (bytes))); } }while(0)

/* Segment `nstreams` streams at once.  pcm[i]: planar [channels][nsamples[i]] with stride[i].
   Afterwards the timeline of all streams sits in HBM (vb200_segmenter_device_pcm) and the blocks index it. */
template <typename SRC>
static int segmenter_begin_impl(vb200_segmenter *g, int nstreams, const SRC *const *pcm, const long *nsamples,
                                const long *stride, int chunk, int threads) {
  constexpr bool S16 = sizeof(SRC) == 2;
  if (!g || nstreams < 1 || !pcm || !nsamples || !stride) { vb200_set_error("vb200_segmenter_begin: bad arguments"); return VB200_EINVAL; }
  vb200_setup *s = g->s;
  CUDA_TRY(cudaSetDevice(s->device));
  cudaStream_t st = g->st;
  const int ch = s->h.channels;
  if (chunk <= 0) chunk = 1024;
  if (g->run) { vb200_set_error("vb200_segmenter_begin: the previous run was not finished"); return VB200_EINVAL; }
  SegRun *R = new SegRun();
  g->run = R;
  R->nstreams = nstreams; R->threads = threads;
  R->seg.resize(nstreams); R->base.assign(nstreams + 1, 0); R->mark_base.assign(nstreams + 1, 0);
  std::vector<StreamSeg> &seg = R->seg;
  std::vector<long long> &base = R->base, &mark_base = R->mark_base;
  std::vector<long long> step_base(nstreams + 1, 0);
  for (int i = 0; i < nstreams; i++) {
    const long T = StreamSeg::timeline_len(s, nsamples[i]);
    base[i + 1] = base[i] + (long long)T * ch;
    const long steps = (s->h.bs[1] / 2 + nsamples[i]) / VE_STEP - VE_WIN;
    step_base[i + 1] = step_base[i] + (steps > 0 ? steps : 0);
    mark_base[i + 1] = mark_base[i] + ((T / VE_STEP + VE_WIN + VE_POST + 4 + 15) & ~15LL);
  }
  const size_t total = (size_t)base[nstreams];
  R->total = total;
  if (g->pcm_cap < total) {
    cudaFree(g->d_pcm); cudaFreeHost(g->h_pcm); g->d_pcm = nullptr; g->h_pcm = nullptr; g->pcm_cap = 0;
    SEG_ALLOC(g->h_pcm, total * sizeof(float), true); SEG_ALLOC(g->d_pcm, total * sizeof(float), false);
    g->pcm_cap = total;
  }
  const size_t nstep_tot = (size_t)step_base[nstreams] * ch;
  if (g->step_cap < nstep_tot) {
    cudaFree(g->d_acc); cudaFree(g->d_temp); cudaFree(g->d_flags); cudaFree(g->d_maps); g->d_acc = g->d_temp = nullptr; g->d_flags = nullptr; g->d_maps = nullptr; g->step_cap = 0;
    SEG_ALLOC(g->d_acc, (nstep_tot + 1) * VE_BANDS * sizeof(float), false); SEG_ALLOC(g->d_temp, (nstep_tot + 1) * sizeof(float), false);
    SEG_ALLOC(g->d_flags, ((size_t)step_base[nstreams] + 1) * sizeof(unsigned), false);
    SEG_ALLOC(g->d_maps, ((size_t)(step_base[nstreams] >> 5) + nstreams + PIPE_SEG_STREAM_SLACK) * 32, false);
    g->step_cap = nstep_tot;
  }
  const size_t nmark = (size_t)mark_base[nstreams];
  if (g->mark_cap < nmark) {
    cudaFree(g->d_marks); cudaFreeHost(g->h_marks); g->mark_cap = 0;
    SEG_ALLOC(g->d_marks, nmark, false); SEG_ALLOC(g->h_marks, nmark, true);
    g->mark_cap = nmark;
  }
  if (g->stream_cap < nstreams) {
    cudaFree(g->d_streams); cudaFreeHost(g->h_streams); cudaFree(g->d_mark_base); cudaFreeHost(g->h_mark_base);
    cudaFree(g->d_stretch); cudaFreeHost(g->h_stretch); cudaFree(g->d_acc_tail); cudaFreeHost(g->h_acc_tail); cudaFree(g->d_temp_tail); cudaFreeHost(g->h_temp_tail);
    g->stream_cap = 0;
    SEG_ALLOC(g->d_streams, sizeof(EnvStream) * nstreams, false); SEG_ALLOC(g->h_streams, sizeof(EnvStream) * nstreams, true);
    SEG_ALLOC(g->d_mark_base, sizeof(long long) * nstreams, false); SEG_ALLOC(g->h_mark_base, sizeof(long long) * nstreams, true);
    SEG_ALLOC(g->d_stretch, sizeof(int) * nstreams, false); SEG_ALLOC(g->h_stretch, sizeof(int) * nstreams, true);
    SEG_ALLOC(g->d_acc_tail, sizeof(float) * nstreams * ch * VE_AMP * VE_BANDS, false); SEG_ALLOC(g->h_acc_tail, sizeof(float) * nstreams * ch * VE_AMP * VE_BANDS, true);
    SEG_ALLOC(g->d_temp_tail, sizeof(float) * nstreams * ch * 45, false); SEG_ALLOC(g->h_temp_tail, sizeof(float) * nstreams * ch * 45, true);
    g->stream_cap = nstreams;
  }
  const bool trace = getenv("VB200_TRACE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now(), t1;
  /* inputs already in page-locked memory go up from where they are; otherwise they are staged through
     the segmenter's own pinned timeline */
  bool direct = getenv("VB200_SEG_STAGED") == nullptr;
  for (int i = 0; direct && i < nstreams; i++) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, pcm[i]) != cudaSuccess) { (void)cudaGetLastError(); direct = false; }
    else if (at.type != cudaMemoryTypeHost) direct = false;
  }
  g->direct = direct ? 1 : 0;
  std::vector<long long> s16_base(nstreams + 1, 0);
  if (S16 && direct) {
    for (int i = 0; i < nstreams; i++) s16_base[i + 1] = s16_base[i] + (long long)nsamples[i] * ch;
    const size_t need = (size_t)s16_base[nstreams] + 8;
    if (g->pcm16_cap < need) { cudaFree(g->d_pcm16); g->d_pcm16 = nullptr; g->pcm16_cap = 0; SEG_ALLOC(g->d_pcm16, need * sizeof(short), false); g->pcm16_cap = need; }
  }
  const int lead = s->h.bs[1] / 2;
  int max_steps = 0;
  for (int i = 0; i < nstreams; i++) {
    EnvStream &e = g->h_streams[i];
    const long T = StreamSeg::timeline_len(s, nsamples[i]), eof = lead + nsamples[i];
    const long l = eof / VE_STEP - VE_WIN;
    e.pcm_base = base[i]; e.T = T; e.nsteps = (int)(l > 0 ? l : 0); e.out_base = step_base[i]; e.eof = eof;
    e.s16_base = s16_base[i];
    g->h_mark_base[i] = mark_base[i];
    if (e.nsteps > max_steps) max_steps = e.nsteps;
  }
  if (g->after) CUDA_TRY(cudaStreamWaitEvent(st, g->after->ev_h2d, 0));   /* the copy engine would otherwise share the link between both */
  CUDA_TRY(cudaMemcpyAsync(g->d_streams, g->h_streams, sizeof(EnvStream) * nstreams, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(g->d_mark_base, g->h_mark_base, sizeof(long long) * nstreams, cudaMemcpyHostToDevice, st));
  /* host: timelines (pre-amplified PCM + LPC head) into the pinned buffer; staged mode sends every finished
     stream to the device at once, so the H2D of stream i overlaps the fill of the streams after it */
  {
    std::vector<std::atomic<int>> done(nstreams);
    for (int i = 0; i < nstreams; i++) done[i].store(0);
    int nthr = threads < 1 ? 1 : (threads > nstreams ? nstreams : threads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; t++)
      th.emplace_back([&, t]() {
        for (int i = t; i < nstreams; i += nthr) {
          seg[i].init(s, g->env, nsamples[i], g->h_pcm + base[i], i);
          seg[i].pcm_base = (long)base[i];
          seg[i].fill_and_head(pcm[i], stride[i], chunk, direct);
          done[i].store(1, std::memory_order_release);
        }
      });
    cudaError_t cerr = cudaSuccess;
    if (direct) {
      for (int i = 0; i < nstreams && cerr == cudaSuccess; i++)
        if (nsamples[i] > 0) {
          if (S16) cerr = cudaMemcpy2DAsync(g->d_pcm16 + s16_base[i], sizeof(short) * nsamples[i], pcm[i], sizeof(short) * stride[i],
                                            sizeof(short) * nsamples[i], ch, cudaMemcpyHostToDevice, st);
          else cerr = cudaMemcpy2DAsync(g->d_pcm + base[i] + lead, sizeof(float) * g->h_streams[i].T, pcm[i], sizeof(float) * stride[i],
                                        sizeof(float) * nsamples[i], ch, cudaMemcpyHostToDevice, st);
        }
      if (S16 && cerr == cudaSuccess) {
        k_timeline_from_s16<<<dim3(64, nstreams * ch), 256, 0, st>>>(g->d_streams, ch, lead, g->d_pcm16, g->d_pcm);
        cerr = cudaGetLastError();
      }
      if (cerr == cudaSuccess) {
        k_timeline_finish<<<dim3(64, nstreams * ch), 256, 0, st>>>(g->d_streams, ch, lead, g->env.pre_amp, g->d_pcm);
        cerr = cudaGetLastError();
      }
    }
    for (int i = 0; i < nstreams; i++) {
      while (!done[i].load(std::memory_order_acquire)) std::this_thread::yield();
      if (cerr != cudaSuccess) continue;
      if (!direct)
        cerr = cudaMemcpyAsync(g->d_pcm + base[i], g->h_pcm + base[i], (size_t)(base[i + 1] - base[i]) * sizeof(float), cudaMemcpyHostToDevice, st);
      else
        cerr = cudaMemcpy2DAsync(g->d_pcm + base[i], sizeof(float) * seg[i].T, g->h_pcm + base[i], sizeof(float) * seg[i].T,
                                 sizeof(float) * lead, ch, cudaMemcpyHostToDevice, st);      /* the LPC head */
    }
    for (auto &x : th) x.join();
    CUDA_TRY(cerr);
  }
  CUDA_TRY(cudaEventRecord(g->ev_h2d, st));
  t1 = now();
  /* the detector's heavy half on the device */
  CUDA_TRY(cudaMemsetAsync(g->d_marks, 0, nmark, st));
  if (max_steps > 0) {
    dim3 grid((max_steps + EV_TILE - 1) / EV_TILE, ch, nstreams);
    k_env_bands<<<grid, EV_THREADS, 0, st>>>(g->d_env, g->d_streams, ch, g->d_pcm, g->d_acc, g->d_temp);
    LAUNCH_CHECK();
    k_env_flags<<<dim3((max_steps + 127) / 128, nstreams), 128, 0, st>>>(g->d_env, g->d_streams, ch, g->d_acc, g->d_flags);
    LAUNCH_CHECK();
    if (((max_steps + 31) >> 5) <= EW_MAXSUPER * 32 && getenv("VB200_SEG_SERIAL_TRIGGER") == nullptr) {
      k_env_maps<<<dim3((((max_steps + 31) >> 5) + 127) / 128, nstreams), 128, 0, st>>>(g->d_streams, g->d_flags, g->d_maps);
      LAUNCH_CHECK();
      k_env_walk<<<nstreams, EW_THREADS, 0, st>>>(g->d_streams, g->d_flags, g->d_maps, g->d_marks, g->d_mark_base, g->d_stretch);
      LAUNCH_CHECK();
    } else {
      k_env_trigger<<<nstreams, 32, 0, st>>>(g->d_streams, g->d_flags, g->d_marks, g->d_mark_base, g->d_stretch);
      LAUNCH_CHECK();
    }
    k_env_tail<<<dim3(nstreams, ch), 64, 0, st>>>(g->d_streams, ch, g->d_acc, g->d_temp, g->d_acc_tail, g->d_temp_tail);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaMemcpyAsync(g->h_marks, g->d_marks, nmark, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g->h_stretch, g->d_stretch, sizeof(int) * nstreams, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g->h_acc_tail, g->d_acc_tail, sizeof(float) * nstreams * ch * VE_AMP * VE_BANDS, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(g->h_temp_tail, g->d_temp_tail, sizeof(float) * nstreams * ch * 45, cudaMemcpyDeviceToHost, st));
  R->t0 = t0; R->t1 = t1; R->nmark = nmark;
  return VB200_OK;
}

extern "C" int vb200_segmenter_begin(vb200_segmenter *g, int nstreams, const float *const *pcm, const long *nsamples,
                                     const long *stride, int chunk, int threads) {
  return segmenter_begin_impl<float>(g, nstreams, pcm, nsamples, stride, chunk, threads);
}
/* 16-bit PCM (planar, [channels][nsamples] with stride, as de-interleaved from a WAV file): half the host-to-device
   bytes of the float entry point; the device converts (sample / 32768.f).  Same timeline, same blocks and the same
   packets as vb200_segmenter_begin on the converted samples. */
extern "C" int vb200_segmenter_begin_s16(vb200_segmenter *g, int nstreams, const short *const *pcm, const long *nsamples,
                                         const long *stride, int chunk, int threads) {
  return segmenter_begin_impl<short>(g, nstreams, pcm, nsamples, stride, chunk, threads);
}

/* second half: wait for the marks, sequence the windows, extrapolate and upload the tails */
extern "C" int vb200_segmenter_finish(vb200_segmenter *g) {
  if (!g || !g->run) { vb200_set_error("vb200_segmenter_finish: nothing was begun"); return VB200_EINVAL; }
  vb200_setup *s = g->s;
  CUDA_TRY(cudaSetDevice(s->device));
  cudaStream_t st = g->st;
  SegRun *R = g->run;
  std::vector<StreamSeg> &seg = R->seg;
  std::vector<long long> &mark_base = R->mark_base;
  const int nstreams = R->nstreams, threads = R->threads, ch = s->h.channels;
  const size_t total = R->total;
  const bool trace = getenv("VB200_TRACE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = R->t0, t1 = R->t1;
  double t2, t3, t4;
  const double tw = now();
  if (trace) { CUDA_TRY(cudaStreamSynchronize(st)); }
  t2 = now();
  CUDA_TRY(cudaStreamSynchronize(st));
  /* host: sequencing up to the end of the signal, LPC tail, detector over the tail, final sequencing */
  parallel_for(nstreams, threads, [&](int i) {
    StreamSeg &q = seg[i];
    const long last = g->h_streams[i].nsteps;
    if (last > 0) q.adopt(last, g->h_marks + mark_base[i], g->h_stretch[i], g->h_acc_tail + (size_t)i * ch * VE_AMP * VE_BANDS, g->h_temp_tail + (size_t)i * ch * 45);
    q.sequence(q.eof_abs, false);
    q.tail();
    q.detect_host(q.T); q.sequence(q.T, true);
  });
  t3 = now();
  /* the extrapolated tails go up as well (small), then the blocks */
  k_tail_upload<<<nstreams * ch, 256, 0, st>>>(g->d_streams, ch, g->h_pcm, g->d_pcm);
  LAUNCH_CHECK();
  g->blocks.clear();
  for (int i = 0; i < nstreams; i++) g->blocks.insert(g->blocks.end(), seg[i].blocks.begin(), seg[i].blocks.end());
  g->pcm_elems = total;
  CUDA_TRY(cudaStreamSynchronize(st));
  t4 = now();
  if (trace) fprintf(stderr, "[vb200 segmenter] begin (fill+head, enqueue) %.1f ms | waited %.1f ms for h2d+detector+d2h (%.1f ms after begin) | sequencing+tail %.1f ms | tail h2d+collect %.1f ms\n", t1 - t0, t2 - tw, t2 - t1, t3 - t2, t4 - t3);
  delete R;
  g->run = nullptr;
  return VB200_OK;
}

extern "C" int vb200_segmenter_run(vb200_segmenter *g, int nstreams, const float *const *pcm, const long *nsamples,
                                   const long *stride, int chunk, int threads) {
  const int r = vb200_segmenter_begin(g, nstreams, pcm, nsamples, stride, chunk, threads);
  if (r) { if (g && g->run) { delete g->run; g->run = nullptr; } return r; }
  return vb200_segmenter_finish(g);
}

/* order this segmenter's uploads behind `prev`'s (NULL: no ordering).  Several segmenters begun back to back
   would otherwise share the host link and all finish late; chained, the first one's data arrives first. */
extern "C" int vb200_segmenter_after(vb200_segmenter *g, const vb200_segmenter *prev) {
  if (!g || prev == g) { vb200_set_error("vb200_segmenter_after: bad arguments"); return VB200_EINVAL; }
  g->after = prev;
  return VB200_OK;
}

extern "C" int vb200_segmenter_count(const vb200_segmenter *g) { return g ? (int)g->blocks.size() : VB200_EINVAL; }
extern "C" const vb200_block *vb200_segmenter_blocks(const vb200_segmenter *g) { return g ? g->blocks.data() : nullptr; }
extern "C" const float *vb200_segmenter_device_pcm(const vb200_segmenter *g, size_t *elems) { if (!g) return nullptr; if (elems) *elems = g->pcm_elems; return g->d_pcm; }
extern "C" const float *vb200_segmenter_host_pcm(const vb200_segmenter *g) { return g ? g->h_pcm : nullptr; }
extern "C" int vb200_segmenter_read_pcm(const vb200_segmenter *g, float *out, size_t elems) {
  if (!g || !out || elems > g->pcm_elems) { vb200_set_error("vb200_segmenter_read_pcm: bad arguments"); return VB200_EINVAL; }
  CUDA_TRY(cudaSetDevice(g->s->device));
  CUDA_TRY(cudaMemcpy(out, g->d_pcm, elems * sizeof(float), cudaMemcpyDeviceToHost));
  return VB200_OK;
}
extern "C" int vb200_pinned_alloc(size_t bytes, void **out) {
  if (!out) { vb200_set_error("vb200_pinned_alloc: bad arguments"); return VB200_EINVAL; }
  CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1));
  return VB200_OK;
}
extern "C" void vb200_pinned_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int vb200_segmenter_direct(const vb200_segmenter *g) { return g ? g->direct : VB200_EINVAL; }
