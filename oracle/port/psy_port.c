/* TEST INFRASTRUCTURE -- CPU oracle, psychoacoustic stages (see orc.h).
 * Restated from the algorithm of /root/reference/lib/psy.c (scalar `#else` paths);
 * each function cites the lines it follows.  Compile with
 * -O2 -fno-fast-math -ffp-contract=off so every float op rounds once. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

#define NEGINF_ (-9999.f)
static inline float fmin_(float a, float b) { return a < b ? a : b; }
static inline float fmax_(float a, float b) { return a > b ? a : b; }

/* scales.h:43-51 */
float orc_todB(float x) {
  union { uint32_t i; float f; } u;
  u.f = x;
  u.i &= 0x7fffffffu;
  return (float)(u.i * 7.17711438e-7f - 764.6161886f);
}

/* psy.c:619-648 */
float orc_postnoise(const float *pcm, int n, int mode, int lw_mode) {
  int q = n >> 2, i;
  double a = 0, b = 0;
  float r = -1.0f;
  if (mode != 2 || lw_mode != 0 || n < 2048) return r;
  for (i = q; i < 2 * q; i++) a += fabs(pcm[i]);
  for (i = 2 * q; i < 3 * q; i++) b += fabs(pcm[i]);
  if (b / q > 0.01) return r;
  a *= a; b *= b; b *= 15;
  if (a > b) { r = a - b; if (r < 0.1) r = -1.0f; }
  return r;
}

/* psy.c:5152-5180 */
float orc_loudnoise_fix(const orc_psy *p, float level, const float *logmdct, int block_mode, int lW_block_mode) {
  int i; double acc = 0;
  if (p->m_val < 0.5) return -1;
  if (p->normal_thresh > .45) return -1;
  if (!((block_mode == 2 && lW_block_mode == 3) || (block_mode == 3 && lW_block_mode == 2))) return level;
  for (i = p->n25p; i < p->n75p; i++) acc += (logmdct[i] > -130) ? logmdct[i] : -130;
  acc /= p->n;
  if (acc > -40.) return -1;
  if (acc < -50.) return 1.;
  return 1. - ((acc + 50) / 10);
}

/* ------------------------------------------------------------------ noise mask */
/* bark_noise_hybridmp, psy.c:3480-3638.  Stated as: (1) per-bin weighted terms, (2) five serial
   float32 prefix sums (the summation order is part of the result), (3) an
   independent per-bin window regression whose window class (mirrored / interior / held) is a static
   property of the window table. */
typedef struct { float *N, *X, *XX, *Y, *XY; } sums_t;

static void prefix_terms(int n, const float *f, float offset, sums_t s) {
  float tN = 0, tX = 0, tXX = 0, tY = 0, tXY = 0, x;
  int i;
  for (i = 0, x = 0.f; i < n; i++, x += 1.f) {
    float y = f[i] + offset, w;
    if (y < 1.f) y = 1.f;
    if (i == 0) {            /* the first bin is half-weighted and enters X as if x were 1 */
      w = y * y * .5;
      tN += w; tX += w; tY += w * y;
    } else {
      w = y * y;
      tN += w; tX += w * x; tXX += w * x * x; tY += w * y; tXY += w * x * y;
    }
    s.N[i] = tN; s.X[i] = tX; s.XX[i] = tXX; s.Y[i] = tY; s.XY[i] = tXY;
  }
}

static void window_abd(sums_t s, int lo, int hi, float *A, float *B, float *D) {
  float tN, tX, tXX, tY, tXY;
  if (lo < 0) {  /* window reaches below bin 0: mirrored */
    tN = s.N[hi] + s.N[-lo]; tX = s.X[hi] - s.X[-lo]; tXX = s.XX[hi] + s.XX[-lo];
    tY = s.Y[hi] + s.Y[-lo]; tXY = s.XY[hi] - s.XY[-lo];
  } else {
    tN = s.N[hi] - s.N[lo]; tX = s.X[hi] - s.X[lo]; tXX = s.XX[hi] - s.XX[lo];
    tY = s.Y[hi] - s.Y[lo]; tXY = s.XY[hi] - s.XY[lo];
  }
  *A = tY * tXX - tX * tXY;
  *B = tN * tXY - tX * tY;
  *D = tN * tXX - tX * tX;
}

/* window bounds of bin i: pass 0 = bark table, pass 1 = fixed width */
static void bounds(const int *bark, int fixed, int pass, int i, int *lo, int *hi) {
  if (pass == 0) { *lo = bark[i] >> 16; *hi = bark[i] & 0xffff; }
  else { *hi = i + fixed / 2; *lo = *hi - fixed; }
}

static void noise_fit(int n, const int *bark, const float *f, float *noise, float offset, int fixed) {
  float *buf = malloc(sizeof(float) * 5 * n);
  sums_t s = { buf, buf + n, buf + 2 * n, buf + 3 * n, buf + 4 * n };
  float A = 0.f, B = 0.f, D = 1.f;   /* held values start at (0,0,1) and survive from pass to pass */
  int pass;
  prefix_terms(n, f, offset, s);
  for (pass = 0; pass < 2; pass++) {
    int i = 0, phase = 0;            /* phase 0: mirrored windows, 1: interior, 2: hold the last fit */
    if (pass == 1 && fixed <= 0) break;
    for (; i < n; i++) {
      int lo, hi; float R, x = (float)i;
      bounds(bark, fixed, pass, i, &lo, &hi);
      if (phase == 0 && !(hi < n && lo < 0 && (pass == 1 || -lo < n))) phase = 1;
      if (phase == 1 && !(hi < n && lo >= 0 && (pass == 1 || lo < n))) phase = 2;
      if (phase < 2) window_abd(s, lo, hi, &A, &B, &D);
      R = (A + x * B) / D;
      if (pass == 0) { if (R < 0.f) R = 0.f; noise[i] = R - offset; }
      else if (R - offset < noise[i]) noise[i] = R - offset;
    }
  }
  free(buf);
}

/* aoTuV M7, psy.c:3645-3768 */
static void tone_fix(const orc_psy *p, const float *spec, float *noise, int block_mode) {
  int nx = p->tonefix_end, n = p->n, i, j;
  float tmp[256], inm[256];
  const float *off1 = p->noiseoffset + n;       /* noiseoffset[1] */
  float limit = fabs(off1[0]);
  if (!nx) return;
  memset(tmp, 0, sizeof(tmp)); memset(inm, 0, sizeof(inm));
  if (block_mode <= 1) {
    const int up = 3, un = 4;
    int nxp = nx + un;
    float tol = (n == 256) ? 15.f : 9.f, strength = .6f;
    if (nxp > n) { nx = n; nxp = n - un; }
    for (i = 0; i < nxp; i++) inm[i] = (spec[i] < -70) ? -70 + (spec[i] + 70) * .1 : spec[i];
    for (i = un; i < nx; i++) {
      int ps, pe; float ss;
      if (!(spec[i] > spec[i - 1] && spec[i] > spec[i + 1])) continue;
      ps = i - 1; pe = i + 1;
      for (j = ps; j > i - up; j--) { if (spec[j + 1] < spec[j]) break; ps = j; }
      for (j = pe; j < i + un; j++) { if (spec[j - 1] < spec[j]) break; pe = j; }
      ss = fmax_(inm[i] - inm[ps], inm[i] - inm[pe]);
      if (ss > tol) {
        if (spec[i] > noise[i]) { ss -= tol; ss *= strength; }
        for (j = ps; j <= pe; j++) { tmp[j] = fmax_(ss, tmp[j]); if (tmp[j] < 0) tmp[j] = 0; }
      }
      i = pe;
    }
    for (i = un - 1; i < nx; i++) {
      float cap = fmin_(p->ntfix[i], off1[i] + limit);
      if (tmp[i] > cap) tmp[i] = cap;
      noise[i] -= tmp[i];
    }
  } else if (block_mode == 2) {
    int k, groups;
    for (i = 0, k = 0; i < nx; i += 8, k++) {
      double a = 0;
      for (j = 0; j < 8; j++) a += noise[i + j];
      tmp[k] = a / 8;
    }
    groups = nx / 8;
    for (i = 3; i < groups; i++) {
      float base, th; int a, b;
      if (!(tmp[i] > tmp[i - 1] && tmp[i] > tmp[i + 1])) continue;
      if (tmp[i - 1] > tmp[i - 2]) { base = tmp[i - 2]; a = i - 3; } else { base = tmp[i - 1]; a = i - 2; }
      b = i + 3;
      th = tmp[i] - base;
      if (th > 2.) {
        float cap = fmin_(p->ntfix[i * 8], off1[i * 8] + limit);
        th = fmin_(th - 2, cap);
        for (j = a * 8; j <= b * 8; j++) noise[j] -= th;
      }
    }
  }
}

/* psy.c:3770-4074 */
void orc_noisemask(const orc_psy *p, float level, const float *logmdct, const float *lastmdct,
                   float *epeak, float *npeak, float *noise, float poste, int block_mode) {
  int n = p->n, part = p->partition, i, j, k;
  float *work = malloc(sizeof(float) * n);
  const float *off1 = p->noiseoffset + n;
  noise_fit(n, p->bark, logmdct, noise, 140.f, -1);
  for (i = 0; i < n; i++) work[i] = logmdct[i] - noise[i];
  noise_fit(n, p->bark, work, noise, 0.f, p->noisewindowfixed);
  for (i = 0; i < n; i++) work[i] = logmdct[i] - work[i];
  tone_fix(p, logmdct, work, block_mode);
  for (i = 0; i < n; i++) {                       /* compander, psy.c:4005-4023 */
    int dB = noise[i] + .5;
    if (dB >= 40) dB = 39;
    if (dB < 0) dB = 0;
    epeak[i] = work[i] + p->stn_compand[dB];
    if (level > 0 && i < p->n33p)
      noise[i] = work[i] + p->compand[dB] - ((p->compand[dB] - p->compand_hi[dB]) * level);
    else
      noise[i] = work[i] + p->compand[dB];
  }
  for (i = 0, k = 0; i < n; i += part, k++) npeak[k] = 0.f;
  if (poste > 0) {                                /* aoTuV M2 post-processing, psy.c:4028-4036 */
    for (i = 0, k = 0; i < p->min_nn_lp; i += part, k++) {
      float t = fmin_(fmin_(poste, 30.f), off1[i] + 30.f);
      if (t <= 0) continue;
      npeak[k] = -1.f;
      for (j = 0; j < part; j++) noise[i + j] -= t;
    }
  }
  for (k = 0, i = 0; i < p->min_nn_lp; i += part, k++) {   /* aoTuV M8, psy.c:4038-4063 */
    const float nt = 4;
    float o = off1[i + part - 1] + 6, me = 0, av = 0;
    if (o <= 0) continue;
    if (npeak[k] < -0.5) continue;
    for (j = 0; j < part; j++) {
      float t = logmdct[i + j] - noise[i + j];
      if (me < t) me = t;
      av += logmdct[i + j];
    }
    if (av < (-95 * part)) continue;
    if (me < nt) npeak[k] = fmin_(o, nt - me) / nt;
  }
  i = 0;                                          /* aoTuV M9, psy.c:4065-4073 */
  if (block_mode > 1)
    for (; i < p->tonecomp_endp; i++) {
      float t = logmdct[i] - epeak[i];
      epeak[i] = 0.f;
      if (t >= 12.f) { float mi = logmdct[i] - lastmdct[i]; if (mi >= 1) epeak[i] = mi; }
    }
  for (; i < n; i++) epeak[i] = 0.f;
  free(work);
}

/* ------------------------------------------------------------------ tone mask */
/* seed_chase, psy.c:773-935 (scalar form): a stack of (position, amplitude) candidates over the seed
   vector, then a left-to-right fill. */
static void chase(float *seeds, int width, int n) {
  int *pos = malloc(sizeof(int) * n); float *amp = malloc(sizeof(float) * n);
  int top = 0, i, w = 0;
  for (i = 0; i < n; i++) {
    if (top >= 2) {
      while (!(seeds[i] < amp[top - 1])) {
        if (i < pos[top - 1] + width && top > 1 && amp[top - 1] <= amp[top - 2] && i < pos[top - 2] + width) { top--; continue; }
        break;
      }
    }
    pos[top] = i; amp[top] = seeds[i]; top++;
  }
  for (i = 0; i < top; i++) {
    int end = (i < top - 1 && amp[i + 1] > amp[i]) ? pos[i + 1] : pos[i] + width + 1;
    if (end > n) end = n;
    for (; w < end; w++) seeds[w] = amp[i];
  }
  free(pos); free(amp);
}

/* psy.c:4076-4142 with seed_loop :719, seed_curve :652, max_seeds :936 */
void orc_tonemask(const orc_psy *p, const float *f, float *tone, float global_specmax, float local_specmax) {
  int n = p->n, L = p->total_octave_lines, per = p->eighth, i;
  float *seed = malloc(sizeof(float) * L);
  float att = local_specmax + p->ath_adjatt;
  float dBoffset = p->max_curve_dB - global_specmax;
  for (i = 0; i < L; i++) seed[i] = NEGINF_;
  if (att < p->ath_maxatt) att = p->ath_maxatt;
  for (i = 0; i < n; i++) tone[i] = p->ath[i] + att;

  for (i = 0; i < n; i++) {                       /* one pass per run of bins sharing an octave line */
    float mx = f[i];
    int oc = p->octave[i];
    while (i + 1 < n && p->octave[i + 1] == oc) { i++; if (f[i] > mx) mx = f[i]; }
    if (mx + 6.f > tone[i]) {
      int band = oc >> p->shiftoc, choice, k, k1, sp;
      const float *posts;
      if (band >= 17) band = 16;
      if (band < 0) band = 0;
      choice = (int)((mx + dBoffset - 30.) * .1f);
      if (choice < 0) choice = 0;
      if (choice > 7) choice = 7;
      posts = p->tonecurves + (band * 8 + choice) * 58;
      k1 = (int)posts[1];
      sp = (p->octave[i] - p->firstoc) + (posts[0] - 16) * per - (per >> 1);
      for (k = posts[0]; k < k1; k++) {
        if (sp > 0) { float lin = mx + posts[2 + k]; if (seed[sp] < lin) seed[sp] = lin; }
        sp += per;
        if (sp >= L) break;
      }
    }
  }

  chase(seed, per, L);
  {
    int lin = 0, pos = p->octave[0] - p->firstoc - (per >> 1);
    while (lin + 1 < n) {
      float mv = seed[pos];
      int end = ((p->octave[lin] + p->octave[lin + 1]) >> 1) - p->firstoc;
      if (mv > p->tone_abs_limit) mv = p->tone_abs_limit;
      while (pos + 1 <= end) { pos++; if ((seed[pos] > NEGINF_ && seed[pos] < mv) || mv == NEGINF_) mv = seed[pos]; }
      end = pos + p->firstoc;
      for (; lin < n && p->octave[lin] <= end; lin++) if (tone[lin] < mv) tone[lin] = mv;
    }
    { float mv = seed[L - 1]; for (; lin < n; lin++) if (tone[lin] < mv) tone[lin] = mv; }
  }
  free(seed);
}

/* ------------------------------------------------------------------ offset & mix */
typedef struct { int on, keep; float rate, rate_low, center, tone_rate; } m3_t;

/* set_m3p, psy.c:4148-4272 */
static void m3_prepare(m3_t *m, const orc_psy *p, int lW_no, int impadnum, int hs, float toneatt,
                       const float *logmdct, const float *lastmdct, float *tempmdct,
                       int block_mode, int lW_block_mode, int managed, int sel) {
  int n = p->n, i, j, count;
  const int *bfn; double drop, lift;
  memset(m, 0, sizeof(*m));
  if (!hs) return;
  if (!managed || sel == 2) m->keep = 1;
  else if (sel == 0) return;
  if (block_mode) return;
  if (n == 128) {
    bfn = p->freq_bfn128; drop = 5; lift = 5.;
    count = (toneatt < 3) ? 2 : 3;
    if (!lW_block_mode) {
      if (lW_no < 8) { m->rate = 0.7 - (float)(lW_no - 1) / 17; m->center = (float)(lW_no * count); m->tone_rate = 8 - lW_no; }
      else { m->rate = 0.3; m->center = 25; m->tone_rate = 0; if ((lW_no * count) < 24) m->center = lW_no * count; }
    } else { m->rate = 0.7; m->center = 0; m->tone_rate = 8.; }
  } else if (n == 256) {
    bfn = p->freq_bfn256; drop = 10; lift = 10.;
    count = 6;
    if (!lW_block_mode) {
      if (lW_no < 4) { m->rate = 0.4 - (float)(lW_no - 1) / 11; m->center = (float)(lW_no * count + 12); m->tone_rate = 8 - lW_no * 2; }
      else { m->rate = 0.2; m->center = 30; m->tone_rate = 0; }
    } else { m->rate = 0.6; m->center = 12; m->tone_rate = 8.; }
  } else return;
  if (m->keep) {
    if (!lW_block_mode) for (i = 0; i < n; i++) tempmdct[i] -= drop;
    else for (i = 0; i < n; i++) tempmdct[i] = lastmdct[i] - drop;
  }
  m->rate_low = 0;
  m->on = 1;
  if (impadnum) m->rate *= (impadnum * (n == 128 ? 0.125 : 0.0625));
  for (i = 0; i < n; i++) {
    float cell = 75 / (float)bfn[i];
    for (j = 1; j < bfn[i]; j++) {
      float fb = logmdct[i] - (cell * j);
      if (tempmdct[i + j] < fb && m->keep) tempmdct[i + j] += (lift / (float)bfn[i + j]);
    }
  }
  if (managed && sel == 0 && m->on) m->rate *= 0.2;
}

/* psy.c:4274-4502 */
void orc_offset_and_mix(const orc_psy *p, const float *noise, const float *tone, int sel, int managed,
                        float *logmask, float *mdct, const float *logmdct, float *lastmdct, float *tempmdct,
                        float low_compand, float *npeak, int end_block, int block_mode, int nW,
                        int lW_block_mode, int lW_no, int impadnum) {
  int n = p->n, hs = (p->rate < 26000) ? 0 : 1, part = p->partition, i, j, k;
  float toneatt = p->tone_masteratt[sel];
  int m4_start = p->normal_start, m4_end = p->tonecomp_endp, m4_endblock = end_block;
  float m4_thres = p->tonecomp_thres;
  const float *noff = p->noiseoffset + sel * n;
  m3_t m3;
  if (low_compand < 0 || toneatt < 25.) low_compand = 0;
  else low_compand *= (toneatt - 25.);
  m3_prepare(&m3, p, lW_no, impadnum, hs, toneatt, logmdct, lastmdct, tempmdct, block_mode, lW_block_mode, managed, sel);
  m4_endblock += p->normal_partition;
  if (m4_endblock > n) m4_endblock = n;
  if (!hs) m4_end = m4_endblock;
  else if (p->normal_thresh > 1.) m4_start = 9999;

  for (i = 0; i < n; i++) {
    float val = noise[i] + noff[i], tval = tone[i] + toneatt;
    if (i <= m4_start) tval -= low_compand;
    if (val > p->noisemaxsupp) val = p->noisemaxsupp;
    if (m3.on && val > tval && val > lastmdct[i] && logmdct[i] > (tempmdct[i] + m3.center)) {   /* aoTuV M3 */
      int toneac = 0; float valmask, rate, mainth;
      if (m3.keep) tempmdct[i] = logmdct[i];
      rate = (logmdct[i] > lastmdct[i]) ? m3.rate : m3.rate_low;
      if (!impadnum && i < p->tonecomp_endp && (val - lastmdct[i]) > 20.f) {
        float dBsub = logmdct[i] - lastmdct[i];
        if (dBsub > 25.f) {
          toneac = 1;
          if (tval > -100.f && (logmdct[i] - tval) < 48.f) {
            float tr = m3.tone_rate;
            if (dBsub < 35.f) tr *= ((35.f - dBsub) * .1f);
            tval -= tr;
            if (tval < -100.f) tval = -100.f;
            if ((logmdct[i] - tval) > 48.f) tval = logmdct[i] - 48.f;
          }
        }
      }
      if (i > p->m3n[0]) mainth = 30.f;
      else if (i > p->m3n[1]) mainth = 20.f;
      else if (i > p->m3n[2]) { mainth = 10.f; rate *= .5f; }
      else { mainth = 10.f; rate *= .3f; }
      if ((val - tval) > mainth) valmask = ((val - tval - mainth) * .1f + mainth) * rate;
      else valmask = (val - tval) * rate;
      if ((val - valmask) > lastmdct[i]) val -= valmask; else val = lastmdct[i];
      if (toneac) { float t = val - fmax_(lastmdct[i], -140); if (t > 20.f) val -= (t - 20.f) * .2f; }
      if (toneac == 1) npeak[i / part] = -1.f;
      else if (npeak[i / part] > 0) npeak[i / part] = 0;
    }
    if (val > tval) logmask[i] = val;                                         /* aoTuV M4 */
    else {
      if (i > m4_start && i < m4_end && logmdct[i] < tval) {
        if (logmdct[i] < val) tval -= (tval - val) * m4_thres; else tval = logmdct[i];
      }
      logmask[i] = tval;
    }
    if (sel == 1) {                                                           /* aoTuV M1 */
      float coeffi = -17.2, de;
      val = val - logmdct[i];
      if (val > coeffi) { de = 1.0 - ((val - coeffi) * 0.005 * p->m_val); if (de < 0) de = 0.0001; }
      else de = 1.0 - ((val - coeffi) * 0.0003 * p->m_val);
      mdct[i] *= de;
    }
  }
  if (m3.keep) {                                                              /* psy.c:4462-4501 */
    if (block_mode <= 1) {
      if (nW) { for (i = 0, k = 0; i < n; i++, k += 8) for (j = 0; j < 8; j++) lastmdct[k + j] = logmdct[i]; }
      else memcpy(lastmdct, logmdct, n * sizeof(float));
    } else if (block_mode == 2 && !nW) {
      for (i = 0; i < (n >> 3); i++) {
        float mn = logmdct[i * 8];
        for (j = 1; j < 8; j++) if (mn > logmdct[i * 8 + j]) mn = logmdct[i * 8 + j];
        lastmdct[i] = mn;
      }
    } else memcpy(lastmdct, logmdct, n * sizeof(float));
  }
}
