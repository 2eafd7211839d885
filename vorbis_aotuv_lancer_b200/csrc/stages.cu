/* Stage calls of the C ABI: host pointers in/out around the device launchers.
 * These are the A/B injection points -- tests feed them the reference's own
 * stage inputs (captured with the --wrap taps) and compare the outputs. */
#include "vb200_internal.h"
#include "host_util.cuh"

#define BAD(cond, msg) do{ if (cond) { vb200_set_error("%s: %s", __func__, msg); return VB200_EINVAL; } }while(0)

static int psy_partition(const vb200_setup *s, int W) {
  /* all psys of one block size share normal_partition in every setup the reference builds */
  for (int k = 0; k < s->h.psys; k++) if (s->h.psy[k].blockflag == W) return s->h.psy[k].partition;
  return 16;
}
static int check_psy(const vb200_setup *s, int W, const int32_t *psy, int items) {
  for (int i = 0; i < items; i++)
    if (psy[i] < 0 || psy[i] >= s->h.psys || s->h.psy[psy[i]].blockflag != W) { vb200_set_error("psy index %d invalid for W=%d", psy[i], W); return VB200_EINVAL; }
  return 0;
}

extern "C" int vb200_stage_transform(vb200_setup *s, int W, int items,
                                     const float *pcm, const int32_t *lW, const int32_t *nW,
                                     const int32_t *block_mode, const int32_t *lW_block_mode,
                                     float *gmdct, float *logfft, float *logmdct,
                                     float *local_ampmax, float *poste) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || items < 0 || !pcm || !lW || !nW || !block_mode || !lW_block_mode, "bad arguments");
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

extern "C" int vb200_stage_noisemask(vb200_setup *s, int W, int items, const int32_t *psy,
                                     const int32_t *block_mode, const float *lowcomp, const float *poste,
                                     const float *logmdct, const float *lastmdct,
                                     float *noise, float *epeak, float *npeak) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || items < 0 || !psy || !block_mode || !lowcomp || !poste || !logmdct || !lastmdct, "bad arguments");
  if (items == 0) return VB200_OK;
  if (check_psy(s, W, psy, items)) return VB200_EINVAL;
  const size_t N2 = s->h.bs[W] / 2;
  const int nps = (int)N2 / psy_partition(s, W);
  DevArr<int> d_psy(psy, items, st), d_bm(block_mode, items, st);
  DevArr<float> d_lc(lowcomp, items, st), d_po(poste, items, st), d_lm(logmdct, items * N2, st), d_last(lastmdct, items * N2, st);
  DevArr<float> d_no(items * N2), d_ep(items * N2), d_np((size_t)items * nps);
  CHECK_ALLOC(d_psy); CHECK_ALLOC(d_bm); CHECK_ALLOC(d_lc); CHECK_ALLOC(d_po); CHECK_ALLOC(d_lm); CHECK_ALLOC(d_last);
  CHECK_ALLOC(d_no); CHECK_ALLOC(d_ep); CHECK_ALLOC(d_np);
  LastRef last{d_last, nullptr, nullptr, (long long)N2};
  int r = vb200_launch_noisemask(s, st, W, items, d_psy, d_bm, d_lc, d_po, d_lm, last, nullptr, nullptr, d_no, d_ep, d_np, nps);
  if (r) return r;
  CUDA_TRY(d_no.down(noise, st)); CUDA_TRY(d_ep.down(epeak, st)); CUDA_TRY(d_np.down(npeak, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_tonemask(vb200_setup *s, int W, int items, const int32_t *psy,
                                    const float *logfft, const float *global_ampmax, const float *local_ampmax,
                                    float *tone) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || items < 0 || !psy || !logfft || !global_ampmax || !local_ampmax, "bad arguments");
  if (items == 0) return VB200_OK;
  if (check_psy(s, W, psy, items)) return VB200_EINVAL;
  const size_t N2 = s->h.bs[W] / 2;
  DevArr<int> d_psy(psy, items, st);
  DevArr<float> d_lf(logfft, items * N2, st), d_g(global_ampmax, items, st), d_l(local_ampmax, items, st), d_t(items * N2);
  const size_t RN = s->h.bs[1] / 2;                                   /* scratch rows of the three-kernel path: long-block rows for both sizes */
  DevArr<float> d_ra(items * RN); DevArr<int> d_rb(items * RN);
  CHECK_ALLOC(d_psy); CHECK_ALLOC(d_lf); CHECK_ALLOC(d_g); CHECK_ALLOC(d_l); CHECK_ALLOC(d_t); CHECK_ALLOC(d_ra); CHECK_ALLOC(d_rb);
  int r = vb200_launch_tonemask(s, st, W, items, d_psy, d_lf, d_g, d_l, d_t, d_ra, d_rb, (int)RN);
  if (r) return r;
  CUDA_TRY(d_t.down(tone, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_offset_and_mix(vb200_setup *s, int W, int items, const int32_t *psy,
                                          const int32_t *params, const float *low_compand,
                                          const float *noise, const float *tone, const float *logmdct,
                                          float *mdct, float *lastmdct, float *tempmdct, float *npeak,
                                          float *logmask) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || items < 0 || !psy || !params || !low_compand || !noise || !tone || !logmdct || !mdct || !lastmdct || !tempmdct || !npeak, "bad arguments");
  if (items == 0) return VB200_OK;
  if (check_psy(s, W, psy, items)) return VB200_EINVAL;
  const size_t N2 = s->h.bs[W] / 2;
  const int nps = (int)N2 / psy_partition(s, W);
  const int floor_index = s->h.map[s->h.mode_mapping[W]].floorsubmap[0];
  DevArr<int> d_psy(psy, items, st), d_par(params, (size_t)items * 8, st);
  DevArr<float> d_lc(low_compand, items, st), d_no(noise, items * N2, st), d_to(tone, items * N2, st), d_lm(logmdct, items * N2, st);
  DevArr<float> d_md(mdct, items * N2, st), d_last(lastmdct, (size_t)items * 2048, st), d_lout(lastmdct, (size_t)items * 2048, st);
  DevArr<float> d_tmp(tempmdct, (size_t)items * 256, st), d_np(npeak, (size_t)items * nps, st), d_mask(items * N2);
  CHECK_ALLOC(d_psy); CHECK_ALLOC(d_par); CHECK_ALLOC(d_lc); CHECK_ALLOC(d_no); CHECK_ALLOC(d_to); CHECK_ALLOC(d_lm);
  CHECK_ALLOC(d_md); CHECK_ALLOC(d_last); CHECK_ALLOC(d_lout); CHECK_ALLOC(d_tmp); CHECK_ALLOC(d_np); CHECK_ALLOC(d_mask);
  LastRef last{d_last, nullptr, nullptr, 2048};
  int r = vb200_launch_mix(s, st, W, items, nullptr, nullptr, nullptr, d_psy, d_par, d_lc, d_no, d_to, d_lm, d_md, last,
                           d_tmp, d_tmp, nullptr, d_np, nps, d_mask, 1, (int)N2, d_lout, nullptr, floor_index);
  if (r) return r;
  CUDA_TRY(d_md.down(mdct, st)); CUDA_TRY(d_lout.down(lastmdct, st)); CUDA_TRY(d_tmp.down(tempmdct, st));
  CUDA_TRY(d_np.down(npeak, st)); CUDA_TRY(d_mask.down(logmask, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_floor1_fit(vb200_setup *s, int W, int items, int floor_index,
                                      const float *logmdct, const float *logmask,
                                      int32_t *posts, int32_t *fitted) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || items < 0 || floor_index < 0 || floor_index >= s->h.floors || !logmdct || !logmask, "bad arguments");
  if (items == 0) return VB200_OK;
  const size_t N2 = s->h.bs[W] / 2;
  DevArr<float> d_lm(logmdct, items * N2, st), d_mk(logmask, items * N2, st);
  DevArr<int> d_posts((size_t)items * VB200_MAX_POSTS), d_fit(items);
  CHECK_ALLOC(d_lm); CHECK_ALLOC(d_mk); CHECK_ALLOC(d_posts); CHECK_ALLOC(d_fit);
  CUDA_TRY(cudaMemsetAsync(d_posts.p, 0, sizeof(int) * (size_t)items * VB200_MAX_POSTS, st));
  int r = vb200_launch_floor_fit(s, st, W, items, floor_index, d_lm, d_mk, (int)N2, (int)N2, nullptr, d_posts, d_fit);
  if (r) return r;
  CUDA_TRY(d_posts.down(posts, st)); CUDA_TRY(d_fit.down(fitted, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_floor1_encode(vb200_setup *s, int W, int items, int floor_index,
                                         int32_t *posts, const int32_t *fitted,
                                         int32_t *vals, int32_t *ilogmask, int32_t *nonzero) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || items < 0 || floor_index < 0 || floor_index >= s->h.floors || !posts || !fitted, "bad arguments");
  if (items == 0) return VB200_OK;
  const size_t N2 = s->h.bs[W] / 2;
  DevArr<int> d_posts(posts, (size_t)items * VB200_MAX_POSTS, st), d_fit(fitted, items, st);
  DevArr<int> d_vals((size_t)items * VB200_MAX_POSTS), d_il(items * N2), d_nz(items);
  CHECK_ALLOC(d_posts); CHECK_ALLOC(d_fit); CHECK_ALLOC(d_vals); CHECK_ALLOC(d_il); CHECK_ALLOC(d_nz);
  CUDA_TRY(cudaMemsetAsync(d_vals.p, 0, sizeof(int) * (size_t)items * VB200_MAX_POSTS, st));
  int r = vb200_launch_floor_encode(s, st, W, items, floor_index, d_posts, d_fit, d_vals, d_il, d_nz);
  if (r) return r;
  CUDA_TRY(d_posts.down(posts, st)); CUDA_TRY(d_vals.down(vals, st)); CUDA_TRY(d_il.down(ilogmask, st)); CUDA_TRY(d_nz.down(nonzero, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_couple(vb200_setup *s, int W, int blocks, const int32_t *psy, int blobno,
                                  const float *mdct, const float *epeak, float *npeak,
                                  int32_t *iwork, int32_t *nonzero) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || blocks < 0 || !psy || !mdct || !epeak || !npeak || !iwork || !nonzero, "bad arguments");
  if (blocks == 0) return VB200_OK;
  if (check_psy(s, W, psy, blocks)) return VB200_EINVAL;
  const size_t N2 = s->h.bs[W] / 2, ch = s->h.channels;
  const int nps = (int)N2 / psy_partition(s, W);
  DevArr<int> d_psy(psy, blocks, st), d_iw(iwork, blocks * ch * N2, st), d_nz(nonzero, blocks * ch, st);
  DevArr<float> d_md(mdct, blocks * ch * N2, st), d_ep(epeak, blocks * ch * N2, st), d_np(npeak, blocks * ch * nps, st);
  CHECK_ALLOC(d_psy); CHECK_ALLOC(d_iw); CHECK_ALLOC(d_nz); CHECK_ALLOC(d_md); CHECK_ALLOC(d_ep); CHECK_ALLOC(d_np);
  int r = vb200_launch_couple(s, st, W, blocks, d_psy, blobno, d_md, d_ep, d_np, nps, d_iw, d_nz);
  if (r) return r;
  CUDA_TRY(d_iw.down(iwork, st)); CUDA_TRY(d_nz.down(nonzero, st)); CUDA_TRY(d_np.down(npeak, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_residue_class(vb200_setup *s, int W, int blocks, int resnum, int ch,
                                         const int32_t *in, const int32_t *nonzero,
                                         int32_t *partword, int32_t *classified) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || blocks < 0 || resnum < 0 || resnum >= s->h.residues || ch < 1 || ch > VB200_MAX_CHANNELS || !in || !nonzero, "bad arguments");
  if (blocks == 0) return VB200_OK;
  const size_t N2 = s->h.bs[W] / 2;
  const DevResidue &R = s->h.res[resnum];
  BAD(R.type != 1 && R.type != 2, "residue type not built (types 1 and 2 only)");
  const int partvals = (R.end - R.begin) / R.grouping;
  int chl[VB200_MAX_CHANNELS]; for (int k = 0; k < ch; k++) chl[k] = k;
  DevArr<int> d_chl(chl, ch, st), d_in(in, blocks * ch * N2, st), d_nz(nonzero, (size_t)blocks * ch, st);
  DevArr<int> d_pw((size_t)blocks * ch * partvals), d_cl(blocks);
  CHECK_ALLOC(d_chl); CHECK_ALLOC(d_in); CHECK_ALLOC(d_nz); CHECK_ALLOC(d_pw); CHECK_ALLOC(d_cl);
  CUDA_TRY(cudaMemsetAsync(d_pw.p, 0, sizeof(int) * (size_t)blocks * ch * partvals, st));
  int r = vb200_launch_residue_class(s, st, W, blocks, resnum, ch, d_chl, ch, d_in, d_nz, d_pw, d_cl);
  if (r) return r;
  CUDA_TRY(d_pw.down(partword, st)); CUDA_TRY(d_cl.down(classified, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}

extern "C" int vb200_stage_residue_forward(vb200_setup *s, int W, int blocks, int resnum, int ch,
                                           int32_t *in, const int32_t *nonzero, const int32_t *partword,
                                           int32_t *syms, int32_t max_syms, int32_t *nsyms) {
  STAGE_PROLOGUE(s);
  BAD((W != 0 && W != 1) || blocks < 0 || resnum < 0 || resnum >= s->h.residues || ch < 1 || ch > VB200_MAX_CHANNELS || !in || !nonzero || !partword || max_syms < 1, "bad arguments");
  if (blocks == 0) return VB200_OK;
  const size_t N2 = s->h.bs[W] / 2;
  const DevResidue &R = s->h.res[resnum];
  BAD(R.type != 1 && R.type != 2, "residue type not built (types 1 and 2 only)");
  const int partvals = (R.end - R.begin) / R.grouping;
  int chl[VB200_MAX_CHANNELS]; for (int k = 0; k < ch; k++) chl[k] = k;
  DevArr<int> d_chl(chl, ch, st), d_in(in, blocks * ch * N2, st), d_nz(nonzero, (size_t)blocks * ch, st);
  DevArr<int> d_pw(partword, (size_t)blocks * ch * partvals, st), d_sy((size_t)blocks * max_syms * 2), d_ns(blocks);
  CHECK_ALLOC(d_chl); CHECK_ALLOC(d_in); CHECK_ALLOC(d_nz); CHECK_ALLOC(d_pw); CHECK_ALLOC(d_sy); CHECK_ALLOC(d_ns);
  int r = vb200_launch_residue_forward(s, st, W, blocks, resnum, ch, d_chl, ch, d_in, d_nz, d_pw, d_sy, max_syms, d_ns);
  if (r) return r;
  CUDA_TRY(d_in.down(in, st)); CUDA_TRY(d_sy.down(syms, st)); CUDA_TRY(d_ns.down(nsyms, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return VB200_OK;
}
