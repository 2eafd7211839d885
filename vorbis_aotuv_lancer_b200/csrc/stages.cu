/* Stage calls of the C ABI: host pointers in/out around the device launchers.
 * These are the A/B injection points -- tests feed them the reference's own
 * stage inputs (captured with the --wrap taps) and compare the outputs. */
#include "vb200_internal.h"
#include "host_util.cuh"

extern "C" int vb200_stage_transform(vb200_setup *s, int W, int items,
                                     const float *pcm, const int32_t *lW, const int32_t *nW,
                                     const int32_t *block_mode, const int32_t *lW_block_mode,
                                     float *gmdct, float *logfft, float *logmdct,
                                     float *local_ampmax, float *poste) {
  STAGE_PROLOGUE(s);
  if ((W != 0 && W != 1) || items < 0 || !pcm || !lW || !nW || !block_mode || !lW_block_mode) { vb200_set_error("vb200_stage_transform: bad arguments"); return VB200_EINVAL; }
  if (items == 0) return VB200_OK;
  const size_t n = s->h.bs[W], N2 = n / 2;
  DevArr<float> d_pcm(pcm, items * n, st); CHECK_ALLOC(d_pcm);
  DevArr<int> d_lW(lW, items, st), d_nW(nW, items, st), d_bm(block_mode, items, st), d_lbm(lW_block_mode, items, st);
  CHECK_ALLOC(d_lW); CHECK_ALLOC(d_nW); CHECK_ALLOC(d_bm); CHECK_ALLOC(d_lbm);
  DevArr<float> d_g(items * N2), d_lf(items * N2), d_lm(items * N2), d_amp(items), d_po(items);
  CHECK_ALLOC(d_g); CHECK_ALLOC(d_lf); CHECK_ALLOC(d_lm); CHECK_ALLOC(d_amp); CHECK_ALLOC(d_po);
  int r = vb200_launch_transform(s, st, W, items, d_pcm, nullptr, d_lW, d_nW, d_bm, d_lbm, d_g, d_lf, d_lm, d_amp, d_po);
  if (r) return r;
  CUDA_TRY(d_g.down(gmdct, st)); CUDA_TRY(d_lf.down(logfft, st)); CUDA_TRY(d_lm.down(logmdct, st));
  CUDA_TRY(d_amp.down(local_ampmax, st)); CUDA_TRY(d_po.down(poste, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}
