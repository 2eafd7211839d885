/* Host-side stream segmentation: PCM of a whole stream -> the list of analysis blocks.
 *
 * Replaces, for offline (whole-stream) encoding, the block layer that north_star keeps on the
 * host: vorbis_analysis_buffer/_wrote (block.c:411-553: pre-amplitude, LPC extrapolation of the
 * stream head and tail, lpc.c:60-159), _ve_envelope_search/_ve_amp/_ve_envelope_mark
 * (envelope.c:101-362: the pre/post-echo detector that picks short vs long windows and marks
 * impulse blocks) and the window sequencing of vorbis_analysis_blockout (block.c:557-812).
 *
 * Not a port: the reference streams through a sliding buffer that it shifts after every block;
 * here the whole stream sits on one absolute timeline (1024 extrapolated samples of lead-in, the
 * PCM, 3 long blocks of extrapolated tail), the detector runs once over every 64-sample step,
 * and blocks are descriptors that point into that timeline -- nothing is copied per block, and
 * the timeline is exactly the buffer the GPU path reads.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "vb200_internal.h"

namespace {

constexpr int VE_PRE = 16, VE_WIN = 4, VE_POST = 2, VE_AMP = VE_PRE + VE_POST - 1, VE_BANDS = 12, VE_NEARDC = 15;
constexpr int VE_MINSTRETCH = 2, VE_MAXSTRETCH = 12;
constexpr int VE_N = 128, VE_STEP = 64;
/* band layout of the detector (envelope.c:36-37) */
const int kBandBegin[VE_BANDS] = {2, 4, 6, 9, 13, 17, 22, 12, 8, 3, 2, 1};
const int kBandEnd[VE_BANDS] = {4, 5, 6, 8, 8, 8, 8, 4, 4, 3, 2, 4};

inline float todB_h(float x) {
  union { uint32_t i; float f; } u; u.f = x; u.i &= 0x7fffffffu;
  return (float)(u.i * 7.17711438e-7f - 764.6161886f);
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

struct Detector {
  float win[VE_N];
  float bandwin[VE_BANDS][8], bandtotal[VE_BANDS];
  double cs[VE_N / 2][VE_N];       /* the 128-point MDCT as a matrix (64 x 128), built once */
  Detector() {
    for (int i = 0; i < VE_N; i++) { const float t = (float)sin(i / (VE_N - 1.) * M_PI); win[i] = t * t; }
    for (int j = 0; j < VE_BANDS; j++) {
      const int n = kBandEnd[j]; float tot = 0;
      for (int i = 0; i < n; i++) { bandwin[j][i] = (float)sin((i + .5) / n * M_PI); tot += bandwin[j][i]; }
      bandtotal[j] = (float)(1. / tot);
    }
    for (int k = 0; k < VE_N / 2; k++)
      for (int t = 0; t < VE_N; t++) cs[k][t] = (4.0 / VE_N) * cos(M_PI / (2.0 * VE_N) * (2 * t + 1 + VE_N / 2) * (2 * k + 1));
  }
};
const Detector &detector() { static Detector d; return d; }

struct FilterState { float ampbuf[VE_AMP]; int ampptr; float nearDC[VE_NEARDC]; float nearDC_acc, nearDC_partialacc; int nearptr; };

/* one detector step on 128 samples of one channel (envelope.c:101-217) */
int ve_amp(const Detector &D, const float *data, FilterState *nearst, FilterState *filters, int stretch_state,
           float minV, float stretch_penalty, const float *pre_th, const float *post_th) {
  float vec[VE_N];
  int ret = 0;
  const int stretch = stretch_state / 2 > VE_MINSTRETCH ? stretch_state / 2 : VE_MINSTRETCH;
  float penalty = stretch_penalty - (stretch_state / 2 - VE_MINSTRETCH);
  if (penalty < 0.f) penalty = 0.f;
  if (penalty > stretch_penalty) penalty = stretch_penalty;
  {
    float w[VE_N];
    for (int i = 0; i < VE_N; i++) w[i] = data[i] * D.win[i];
    for (int k = 0; k < VE_N / 2; k++) { double a = 0; for (int t = 0; t < VE_N; t++) a += D.cs[k][t] * w[t]; vec[k] = (float)a; }
  }
  float decay;
  {
    const float temp = (float)(vec[0] * vec[0] + .7 * vec[1] * vec[1] + .2 * vec[2] * vec[2]);
    const int ptr = nearst->nearptr;
    if (ptr == 0) { decay = nearst->nearDC_acc = nearst->nearDC_partialacc + temp; nearst->nearDC_partialacc = temp; }
    else { decay = nearst->nearDC_acc += temp; nearst->nearDC_partialacc += temp; }
    nearst->nearDC_acc -= nearst->nearDC[ptr];
    nearst->nearDC[ptr] = temp;
    decay = (float)(decay * (1. / (VE_NEARDC + 1)));
    nearst->nearptr++;
    if (nearst->nearptr >= VE_NEARDC) nearst->nearptr = 0;
    decay = (float)(todB_h(decay) * .5 - 15.f);
  }
  for (int i = 0; i < VE_N / 2; i += 2) {
    float val = vec[i] * vec[i] + vec[i + 1] * vec[i + 1];
    val = todB_h(val) * .5f;
    if (val < decay) val = decay;
    if (val < minV) val = minV;
    vec[i >> 1] = val;
    decay = (float)(decay - 8.);
  }
  for (int j = 0; j < VE_BANDS; j++) {
    float acc = 0.f;
    for (int i = 0; i < kBandEnd[j]; i++) acc += vec[i + kBandBegin[j]] * D.bandwin[j][i];
    acc *= D.bandtotal[j];
    FilterState &f = filters[j];
    const int cur = f.ampptr;
    int p = cur - 1; if (p < 0) p += VE_AMP;
    const float postmax = acc > f.ampbuf[p] ? acc : f.ampbuf[p], postmin = acc < f.ampbuf[p] ? acc : f.ampbuf[p];
    float premax = -99999.f, premin = 99999.f;
    for (int i = 0; i < stretch; i++) {
      p--; if (p < 0) p += VE_AMP;
      if (f.ampbuf[p] > premax) premax = f.ampbuf[p];
      if (f.ampbuf[p] < premin) premin = f.ampbuf[p];
    }
    const float valmin = postmin - premin, valmax = postmax - premax;
    f.ampbuf[cur] = acc;
    f.ampptr++; if (f.ampptr >= VE_AMP) f.ampptr = 0;
    if (valmax > pre_th[j] + penalty) ret |= 1 | 4;
    if (valmin < post_th[j] - penalty) ret |= 2;
  }
  return ret;
}

}  // namespace

struct vb200_segments {
  int channels; long total;                  /* timeline length per channel */
  std::vector<float> pcm;                    /* [channels][total] */
  std::vector<vb200_block> blocks;
};

/* Segment one whole stream.  pcm: planar [channels][nsamples] (stride between channels),
   chunk: the granularity the caller would have fed vorbis_analysis_wrote with (it decides how much
   signal the head extrapolation sees, block.c:545-548; the reference's examples use 1024). */
extern "C" int vb200_segment_stream(const vb200_setup *s, const float *pcm, long nsamples, long stride, int chunk,
                                    int stream_index, vb200_segments **out) {
  if (!s || !pcm || !out || nsamples < 0) { vb200_set_error("vb200_segment_stream: bad arguments"); return VB200_EINVAL; }
  uint32_t cnt = 0;
  const float *gf = (const float *)vb200_blob_find(s->blob, s->blob_bytes, "psyg.f32", nullptr, &cnt);
  if (!gf || cnt < 4 + 2 * VE_BANDS) { vb200_set_error("vb200_segment_stream: setup blob lacks the envelope parameters"); return VB200_EINVAL; }
  const float stretch_penalty = gf[1], minenergy = gf[2], pre_amp = gf[3];
  const float *pre_th = gf + 4, *post_th = gf + 4 + VE_BANDS;
  const int ch = s->h.channels, bs0 = s->h.bs[0], bs1 = s->h.bs[1];
  const long lead = bs1 / 2, tail = 3L * bs1;
  if (chunk <= 0) chunk = 1024;
  vb200_segments *g = new vb200_segments();
  g->channels = ch; g->total = lead + nsamples + tail;
  g->pcm.assign((size_t)ch * g->total, 0.f);
  const long T = g->total;
  /* timeline: lead-in | pre-amplified PCM | tail */
  for (int c = 0; c < ch; c++) { float *d = &g->pcm[(size_t)c * T + lead]; const float *q = pcm + (size_t)c * stride; for (long i = 0; i < nsamples; i++) d[i] = q[i] * pre_amp; }
  /* head: reverse LPC extrapolation over the lead-in, from the signal that had arrived when the
     reference first had more than one long block (block.c:438-471,545-548) */
  {
    long seen = 0;
    while (seen < nsamples && !(seen > bs1)) seen += (nsamples - seen < chunk ? nsamples - seen : chunk);
    const long cur = lead + seen;                 /* pcm_current at that moment; centerW == lead */
    if (cur - lead > 32) {
      std::vector<float> work(cur); float lpc[16];
      for (int c = 0; c < ch; c++) {
        float *d = &g->pcm[(size_t)c * T];
        for (long j = 0; j < cur; j++) work[j] = d[cur - j - 1];
        lpc_from_data(work.data(), lpc, (int)(cur - lead), 16);
        lpc_predict(lpc, work.data() + cur - lead - 16, 16, work.data() + cur - lead, lead);
        for (long j = 0; j < cur; j++) d[cur - j - 1] = work[j];
      }
    }
  }
  const long eof_abs = lead + nsamples;           /* first sample past the real signal */
  const Detector &D = detector();
  std::vector<unsigned char> mark((size_t)(T / VE_STEP) + VE_WIN + VE_POST + 4, 0);
  std::vector<FilterState> filt((size_t)VE_BANDS * ch);
  memset(filt.data(), 0, sizeof(FilterState) * filt.size());
  int stretch = 0;
  long next_step = 0, ve_current = 0;
  /* block sequencing state (block.c:557-812) on the absolute timeline */
  int lW = 0, W = 0;
  long centerW = lead, cursor = bs1 / 2, curmark = 0;
  int64_t granulepos = 0;
  const int bsz[2] = {bs0, bs1};
  bool finished = false;

  /* everything the reference can do once `avail` samples of the timeline exist */
  auto advance = [&](long avail, bool eof) {
    /* detector over every 64-sample step that is complete (envelope.c:569-616) */
    const long last = avail / VE_STEP - VE_WIN;
    for (long j = next_step; j < last; j++) {
      int ret = 0;
      stretch++;
      if (stretch > VE_MAXSTRETCH * 2) stretch = VE_MAXSTRETCH * 2;
      for (int c = 0; c < ch; c++)
        ret |= ve_amp(D, &g->pcm[(size_t)c * T + VE_STEP * j], &filt[(size_t)c * VE_BANDS], &filt[(size_t)c * VE_BANDS], stretch,
                      minenergy, stretch_penalty, pre_th, post_th);
      mark[j + VE_POST] = 0;
      if (ret & 1) { mark[j] = 1; mark[j + 1] = 1; }
      if (ret & 2) { mark[j] = 1; if (j > 0) mark[j - 1] = 1; }
      if (ret & 4) stretch = -1;
    }
    if (last > next_step) next_step = last;
    if (last > 0) ve_current = last * VE_STEP;
    while (!finished) {
      int nW;
      {
        /* _ve_envelope_search (envelope.c:617-681): first marked step between the cursor and the
           long-block horizon */
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
      if (avail < centerNext + bsz[nW] / 2) return;      /* not enough data yet */
      vb200_block blk;
      memset(&blk, 0, sizeof(blk));
      blk.stream = stream_index; blk.W = W; blk.lW = lW; blk.nW = nW;
      if (W) blk.blocktype = (!lW || !nW) ? 0 : 1;
      else {
        /* _ve_envelope_mark (envelope.c:683-706) */
        const long beginW = centerW - bs0 / 4 - bs0 / 4, endW = centerW + bs0 / 4 + bs0 / 4;
        int hit = (curmark >= beginW && curmark < endW);
        if (!hit) for (long i = beginW / VE_STEP; i < endW / VE_STEP; i++) if (i >= 0 && mark[i]) { hit = 1; break; }
        blk.blocktype = hit ? 0 : 1;
      }
      blk.granulepos = granulepos;
      blk.pcm_offset = centerW - bsz[W] / 2;
      blk.pcm_stride = T;
      if (eof && centerW >= eof_abs) { blk.eofflag = 1; g->blocks.push_back(blk); finished = true; return; }
      g->blocks.push_back(blk);
      const long movement = centerNext - centerW;
      /* granule position: the padding past the end of the stream is not counted (block.c:793-803) */
      if (eof && centerNext >= eof_abs) granulepos += movement - (centerNext - eof_abs);
      else granulepos += movement;
      lW = W; W = nW; centerW = centerNext;
    }
  };

  /* phase 1: all real signal has arrived, end of stream not yet signalled */
  advance(eof_abs, false);
  /* phase 2: end of stream.  Forward LPC extrapolation of 3 long blocks (block.c:497-537) from the
     signal still held in the reference's sliding buffer, i.e. back to centerW - blocksize[1]/2 */
  {
    float lpc[32];
    const long held = eof_abs - (centerW - lead);        /* v->eofflag at vorbis_analysis_wrote(0) */
    for (int c = 0; c < ch; c++) {
      float *d = &g->pcm[(size_t)c * T];
      if (held > 64) {
        long n = held; if (n > bs1) n = bs1;
        lpc_from_data(d + eof_abs - n, lpc, (int)n, 32);
        lpc_predict(lpc, d + eof_abs - 32, 32, d + eof_abs, T - eof_abs);
      }
    }
  }
  advance(T, true);
  *out = g;
  return VB200_OK;
}

extern "C" void vb200_segments_free(vb200_segments *g) { delete g; }
extern "C" int vb200_segments_count(const vb200_segments *g) { return g ? (int)g->blocks.size() : VB200_EINVAL; }
extern "C" const vb200_block *vb200_segments_blocks(const vb200_segments *g) { return g ? g->blocks.data() : nullptr; }
extern "C" const float *vb200_segments_pcm(const vb200_segments *g, long *elems) { if (!g) return nullptr; if (elems) *elems = (long)g->pcm.size(); return g->pcm.data(); }
