/* K4 -- offset & mix: noise + tone -> masking curve, aoTuV M1/M3/M4.
 *
 * Replaces _vp_offset_and_mix (psy.c:4274-4502) and set_m3p (psy.c:4148-4272).
 *
 * B200 mapping.  For long-size blocks the stage is purely elementwise: float4-friendly
 * coalesced loads of noise/tone/logmdct/mdct, one pass, coalesced stores -- HBM bound.
 * Impulse blocks carry aoTuV's M3 state: `tempmdct` is a true recurrence along a run of
 * consecutive impulse blocks of one channel, so a CTA owns a whole *chain*
 * of items and keeps tempmdct in shared memory across them; chains are independent.  In
 * managed-bitrate mode the three offset selections (1, 2, 0 in the reference's call order,
 * mapping0.c:1044-1149) run back to back in the same CTA, including the "nominal fit exists"
 * gate of mapping0.c:1097, which is a simple any-bin predicate (floor1.c:441-448).
 */
#include "vb200_internal.h"
#include "dev_util.cuh"

namespace {

struct M3 { int on, keep; float rate, rate_low, center, tone_rate; };

__device__ __forceinline__ int dbq(float x) {      /* vorbis_dBquant, floor1.c:285-298 */
  int i = (int)__fadd_rn(__fmul_rn(x, 7.3142857f), 1023.5f);
  return i > 1023 ? 1023 : (i < 0 ? 0 : i);
}

#ifndef KM_MINB
#define KM_MINB 6          /* 40 registers: 4 / 5 / 6 CTAs per SM: 2.56 / 2.26 / 2.11 ms */
#endif
template <int N, int T>
__global__ void __launch_bounds__(T, (N > 256 ? KM_MINB : 20))      /* short blocks: 48 registers as before (a bound of 1 let the compiler take 114) */
k_mix(const DevSetup *__restrict__ S, int W, const int *__restrict__ chain_start, const int *__restrict__ chain_len, const int *__restrict__ chain_items,
      const int *__restrict__ psyv, const int *__restrict__ params, const float *__restrict__ low_compandv,
      const float *__restrict__ noise, const float *__restrict__ tone, const float *__restrict__ logmdct,
      float *__restrict__ mdct, LastRef last, const float *__restrict__ temp_state, float *__restrict__ temp_state_out,
      const long long *__restrict__ temp_off, float *__restrict__ npeak, int npeak_stride, float *__restrict__ logmask, int nsel, int mask_stride,
      float *__restrict__ last_out, int *__restrict__ fit_gate, int floor_index, int ch) {
  __shared__ float temp[N > 256 ? 1 : 256];
  __shared__ float sp_slope[N > 256 ? 1 : 256], sp_lm[N > 256 ? 1 : 256];      /* cross-bin spread of M3 (short blocks) */
  __shared__ int sp_b[N > 256 ? 1 : 256];
  __shared__ int pflag[64];
  __shared__ int gate;
  const int tid = threadIdx.x;
  const int c0 = chain_start ? chain_start[blockIdx.x] : blockIdx.x;
  /* chain_len: low bits = links; bit 30 = do not load tempmdct from the state slot, bit 29 = do not store it
     back (only a chain that starts / ends the stream's part of this batch touches the per-stream state) */
  const int clen_raw = chain_len ? chain_len[blockIdx.x] : 1;
  const int clen = clen_raw & 0xffffff;
  const bool t_load = !(clen_raw & (1 << 30)), t_save = !(clen_raw & (1 << 29));
  constexpr bool SHORT = (N <= 256);

  for (int ci = 0; ci < clen; ci++) {
    const int item = chain_items ? chain_items[c0 + ci] : c0 + ci;
    const DevPsy &P = S->psy[psyv[item]];
    const int *pr = params + item * 8;
    const int managed = pr[1], end_block = pr[2], block_mode = pr[3], nW = pr[4], lW_block_mode = pr[5], lW_no = pr[6], impadnum = pr[7];
    const int hs = P.rate < 26000 ? 0 : 1;
    const int part = P.partition;
    const float *nz = noise + (size_t)item * N, *tn = tone + (size_t)item * N, *lm = logmdct + (size_t)item * N;
    float *md = mdct + (size_t)item * N;
    float *np = npeak + (size_t)item * npeak_stride;
    const float *tstate = (SHORT && temp_state) ? temp_state + (temp_off ? temp_off[item] : (long long)item * 256) : nullptr;

    if (SHORT && ci == 0 && tstate && t_load) { for (int i = tid; i < N; i += T) temp[i] = tstate[i]; }
    if (tid == 0) gate = 1;
    __syncthreads();

    for (int si = 0; si < nsel; si++) {
      /* VBR: the single selection pr[0]; managed: 1, 2, 0 */
      const int sel = (nsel == 1) ? pr[0] : (si == 0 ? 1 : (si == 1 ? 2 : 0));
      if (si > 0 && !gate) break;       /* mapping0.c:1097: no nominal fit -> no hi/lo curves */
      const float toneatt = P.tone_masteratt[sel];
      float low_compand = low_compandv[item];
      if (low_compand < 0.f || toneatt < 25.f) low_compand = 0.f;
      else low_compand = (float)((double)low_compand * ((double)toneatt - 25.));

      /* ---- set_m3p ---- */
      M3 m3; m3.on = 0; m3.keep = 0; m3.rate = m3.rate_low = m3.center = m3.tone_rate = 0.f;
      bool m3_go = false;
      if (hs) {
        if (!managed || sel == 2) { m3.keep = 1; m3_go = true; }
        else if (sel != 0) m3_go = true;
        if (block_mode) m3_go = false;
      }
      if (SHORT && m3_go && (N == 128 || N == 256)) {
        const int *bfn = (N == 128) ? S->freq_bfn128 : S->freq_bfn256;
        const double drop = (N == 128) ? 5. : 10.;
        if (N == 128) {
          const int count = (toneatt < 3.f) ? 2 : 3;
          if (!lW_block_mode) {
            if (lW_no < 8) { m3.rate = (float)(0.7 - (double)((float)(lW_no - 1) / 17.f)); m3.center = (float)(lW_no * count); m3.tone_rate = (float)(8 - lW_no); }
            else { m3.rate = (float)0.3; m3.center = 25.f; m3.tone_rate = 0.f; if ((lW_no * count) < 24) m3.center = (float)(lW_no * count); }
          } else { m3.rate = (float)0.7; m3.center = 0.f; m3.tone_rate = 8.f; }
        } else {
          const int count = 6;
          if (!lW_block_mode) {
            if (lW_no < 4) { m3.rate = (float)(0.4 - (double)((float)(lW_no - 1) / 11.f)); m3.center = (float)(lW_no * count + 12); m3.tone_rate = (float)(8 - lW_no * 2); }
            else { m3.rate = (float)0.2; m3.center = 30.f; m3.tone_rate = 0.f; }
          } else { m3.rate = (float)0.6; m3.center = 12.f; m3.tone_rate = 8.f; }
        }
        if (m3.keep) {
          if (!lW_block_mode) { for (int i = tid; i < N; i += T) temp[i] = (float)((double)temp[i] - drop); }
          else { for (int i = tid; i < N; i += T) temp[i] = (float)((double)last.at(item, i) - drop); }
        }
        m3.rate_low = 0.f;
        m3.on = 1;
        if (impadnum) m3.rate = (float)((double)m3.rate * ((double)impadnum * (N == 128 ? 0.125 : 0.0625)));
        __syncthreads();
        if (m3.keep) {
          /* cross-bin spread (psy.c:4215-4222): bin i raises the bins i+1 .. i+bfn[i]-1.  The reference walks i
             upwards; every target bin t is only ever touched through temp[t], so the updates of one target
             form an independent chain in ascending i -- one thread per target, same order, no race */
          /* the slope 75 / bfn[i] belongs to the source bin: one division per bin (staged with logmdct and the
             widths), not one per (source, target) pair -- the same float division, so the same bits */
          for (int i = tid; i < N; i += T) { const int b = bfn[i]; sp_b[i] = b; sp_slope[i] = 75.f / (float)b; sp_lm[i] = __ldg(lm + i); }
          __syncthreads();
          for (int t = tid; t < N; t += T) {
            float v = temp[t];
            const double lift = drop / (double)(float)sp_b[t];
            for (int i = (t > 64 ? t - 64 : 0); i < t; i++) {
              const int j = t - i;
              if (j < sp_b[i]) {
                const float fb = __fsub_rn(sp_lm[i], __fmul_rn(sp_slope[i], (float)j));
                if (v < fb) v = (float)((double)v + lift);
              }
            }
            temp[t] = v;
          }
        }
        if (managed && sel == 0 && m3.on) m3.rate = (float)((double)m3.rate * 0.2);
      }
      if (tid < 64) pflag[tid] = 0;
      __syncthreads();

      /* ---- M4 set-up ---- */
      int m4_start = P.normal_start, m4_end = P.tonecomp_endp, m4_endblock = end_block + P.normal_partition;
      if (m4_endblock > N) m4_endblock = N;
      if (!hs) m4_end = m4_endblock;
      else if (P.normal_thresh > 1.) m4_start = 9999;
      const float m4_thres = P.tonecomp_thres;
      const float *noff = P.noiseoffset + sel * N;
      float *mk = logmask + (size_t)item * mask_stride + (size_t)si * N;
      int local_gate = 0;

      /* one bin: (noise, noise offset, tone, logmdct, mdct) in, mask and M1-scaled mdct out */
      auto elem = [&](const int i, const float md_in, const float nz_i, const float noff_i, const float tn_i, const float l, float &mk_o, float &md_o) {
        float val = __fadd_rn(nz_i, noff_i);
        float tval = __fadd_rn(tn_i, toneatt);
        if (i <= m4_start) tval = __fsub_rn(tval, low_compand);
        if (val > P.noisemaxsupp) val = P.noisemaxsupp;
        if (SHORT && m3.on && val > tval) {                                       /* aoTuV M3 */
          const float la = last.at(item, i);
          if (val > la && l > __fadd_rn(temp[i], m3.center)) {
            int toneac = 0;
            if (m3.keep) temp[i] = l;
            float rate = (l > la) ? m3.rate : m3.rate_low;
            if (!impadnum && i < P.tonecomp_endp && __fsub_rn(val, la) > 20.f) {
              const float dBsub = __fsub_rn(l, la);
              if (dBsub > 25.f) {
                toneac = 1;
                if (tval > -100.f && __fsub_rn(l, tval) < 48.f) {
                  float tr = m3.tone_rate;
                  if (dBsub < 35.f) tr = __fmul_rn(tr, __fmul_rn(__fsub_rn(35.f, dBsub), .1f));
                  tval = __fsub_rn(tval, tr);
                  if (tval < -100.f) tval = -100.f;
                  if (__fsub_rn(l, tval) > 48.f) tval = __fsub_rn(l, 48.f);
                }
              }
            }
            float mainth;
            if (i > P.m3n[0]) mainth = 30.f;
            else if (i > P.m3n[1]) mainth = 20.f;
            else if (i > P.m3n[2]) { mainth = 10.f; rate = __fmul_rn(rate, .5f); }
            else { mainth = 10.f; rate = __fmul_rn(rate, .3f); }
            float valmask;
            const float dv = __fsub_rn(val, tval);
            if (dv > mainth) valmask = __fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(dv, mainth), .1f), mainth), rate);
            else valmask = __fmul_rn(dv, rate);
            if (__fsub_rn(val, valmask) > la) val = __fsub_rn(val, valmask); else val = la;
            if (toneac) { const float t = __fsub_rn(val, fmaxf(la, -140.f)); if (t > 20.f) val = __fsub_rn(val, __fmul_rn(__fsub_rn(t, 20.f), .2f)); }
            atomicOr(&pflag[i / part], toneac ? 3 : 1);
          }
        }
        float out;                                                                /* aoTuV M4 */
        if (val > tval) out = val;
        else {
          if (i > m4_start && i < m4_end && l < tval) {
            if (l < val) tval = __fsub_rn(tval, __fmul_rn(__fsub_rn(tval, val), m4_thres)); else tval = l;
          }
          out = tval;
        }
        mk_o = out;
        md_o = md_in;
        if (sel == 1) {                                                           /* aoTuV M1 */
          const float coeffi = (float)-17.2;
          const float v = __fsub_rn(val, l);
          float de;
          if (v > coeffi) { de = (float)(1.0 - ((double)__fsub_rn(v, coeffi) * 0.005 * (double)P.m_val)); if (de < 0.f) de = (float)0.0001; }
          else de = (float)(1.0 - ((double)__fsub_rn(v, coeffi) * 0.0003 * (double)P.m_val));
          md_o = __fmul_rn(md_in, de);
          if (nsel > 1 && dbq(out) && __fadd_rn(l, S->flr[floor_index >= 0 ? floor_index : S->floor_of_ch[W][item % ch]].twofitatten) >= out) local_gate = 1;
        }
      };
      if constexpr (!SHORT && N == 4 * T) {
        /* long blocks: four consecutive bins per thread, every vector of the block as ONE 128-bit request per thread,
           all five in flight before the first is used (the stage is HBM-bound: 24 KB per ch-block) */
        const int i0 = tid * 4;
        const float4 m4 = (sel == 1) ? *reinterpret_cast<const float4 *>(md + i0) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 n4 = __ldg(reinterpret_cast<const float4 *>(nz + i0)), o4 = __ldg(reinterpret_cast<const float4 *>(noff + i0));
        const float4 t4 = __ldg(reinterpret_cast<const float4 *>(tn + i0)), l4 = __ldg(reinterpret_cast<const float4 *>(lm + i0));
        float4 mo, do_;
        elem(i0, m4.x, n4.x, o4.x, t4.x, l4.x, mo.x, do_.x);
        elem(i0 + 1, m4.y, n4.y, o4.y, t4.y, l4.y, mo.y, do_.y);
        elem(i0 + 2, m4.z, n4.z, o4.z, t4.z, l4.z, mo.z, do_.z);
        elem(i0 + 3, m4.w, n4.w, o4.w, t4.w, l4.w, mo.w, do_.w);
        *reinterpret_cast<float4 *>(mk + i0) = mo;
        if (sel == 1) *reinterpret_cast<float4 *>(md + i0) = do_;
      } else {
        for (int i = tid; i < N; i += T) {
          const float md_in = (sel == 1) ? md[i] : 0.f;      /* M1 scales it at the end of the iteration: request it now */
          float mo, do_;
          elem(i, md_in, __ldg(nz + i), __ldg(noff + i), __ldg(tn + i), __ldg(lm + i), mo, do_);
          mk[i] = mo;
          if (sel == 1) md[i] = do_;
        }
      }
      if (nsel > 1 && si == 0) {
        if (tid == 0) gate = 0;
        __syncthreads();
        if (local_gate) gate = 1;
      }
      __syncthreads();
      /* npeak exceptions for M8 (psy.c:4396-4398): any tone hit -> -1, else a hit clears a positive value */
      if (SHORT && m3.on) {
        for (int k = tid; k < N / part; k += T) {
          const int fl = pflag[k];
          if (fl & 2) np[k] = -1.f;
          else if (fl & 1) { if (np[k] > 0.f) np[k] = 0.f; }
        }
      }
      /* lastmdct as the reference leaves it (psy.c:4462-4501) -- only materialised on request */
      if (last_out && m3.keep) {
        float *lo_ = last_out + (size_t)item * 2048;
        if (block_mode <= 1) {
          if (nW) { for (int i = tid; i < N * 8; i += T) lo_[i] = __ldg(lm + (i >> 3)); }
          else { for (int i = tid; i < N; i += T) lo_[i] = __ldg(lm + i); }
        } else if (block_mode == 2 && !nW) {
          for (int i = tid; i < (N >> 3); i += T) {
            float mn = __ldg(lm + i * 8);
            for (int j = 1; j < 8; j++) { const float v = __ldg(lm + i * 8 + j); if (mn > v) mn = v; }
            lo_[i] = mn;
          }
        } else { for (int i = tid; i < N; i += T) lo_[i] = __ldg(lm + i); }
      }
      __syncthreads();
    }
    if (fit_gate && tid == 0) fit_gate[item] = gate;
    if (SHORT && tstate && t_save && ci == clen - 1) {
      /* leave the chain's tempmdct in the stream's state slot for the next batch (managed mode: in a staging copy of
         the slot, committed when the floor-fit gates have settled -- pipeline.cu) */
      float *dst = temp_state_out + (temp_off ? temp_off[item] : (long long)item * 256);
      for (int i = tid; i < N; i += T) dst[i] = temp[i];
    }
    __syncthreads();
  }
}

}  // namespace

int vb200_launch_mix(const vb200_setup *s, cudaStream_t st, int W, int nchains, const int *chain_start, const int *chain_len, const int *chain_items,
                     const int *psy, const int *params, const float *low_compand,
                     const float *noise, const float *tone, const float *logmdct, float *mdct, LastRef last,
                     const float *temp_state, float *temp_state_out, const long long *temp_off, float *npeak, int npeak_stride,
                     float *logmask, int nsel, int mask_stride, float *last_out, int *fit_gate, int floor_index) {
  if (nchains <= 0) return VB200_OK;
  if (W) k_mix<1024, 256><<<nchains, 256, 0, st>>>(s->d, W, chain_start, chain_len, chain_items, psy, params, low_compand, noise, tone, logmdct, mdct, last,
                                                   temp_state, temp_state_out, temp_off, npeak, npeak_stride, logmask, nsel, mask_stride, last_out, fit_gate, floor_index, s->h.channels);
  else   k_mix<128, 64><<<nchains, 64, 0, st>>>(s->d, W, chain_start, chain_len, chain_items, psy, params, low_compand, noise, tone, logmdct, mdct, last,
                                                temp_state, temp_state_out, temp_off, npeak, npeak_stride, logmask, nsel, mask_stride, last_out, fit_gate, floor_index, s->h.channels);
  LAUNCH_CHECK();
  return VB200_OK;
}
