/* TEST INFRASTRUCTURE -- CPU oracle, aoTuV channel coupling / quantisation / noise
 * normalisation (see orc.h).  Restated from the algorithm of
 * /root/reference/lib/psy.c:4584-5142 (_vp_couple_quantize_normalize and helpers). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

#define PMAX 32     /* largest normal_partition */
#define CHMAX 8

typedef struct {
  float raw[PMAX], quant[PMAX], floor[PMAX], res[PMAX];
  int flag[PMAX];
} lane_t;

static int opposite(float a, float b) { return (a > 0. && b < 0.) || (b > 0. && a < 0.); }
static int big(float x) { return x < -0.5 || x >= 0.5; }        /* existe(x,0.5), psy.c:33 */

/* flag_lossless, psy.c:4584-4624: residue = mdct/floor, then classify against the point-stereo
   thresholds, which ramp linearly across the partition that contains the point limit */
static void classify(int limit, float pre, float post, float pre_r, float post_r, lane_t *L,
                     const float *mdct, const float *enpeak, int i, int jn) {
  int left = limit - i, ramp = 0, j;
  float p1, p2, s1 = 0, s2 = 0;
  if (left > 0) {
    p1 = pre; p2 = pre_r;
    if (left - jn <= 0) { s1 = (post - pre) / jn; s2 = (post_r - pre_r) / jn; ramp = 1; }
  } else { p1 = post; p2 = post_r; }
  for (j = 0; j < jn; j++) {
    float t, r;
    if (ramp) { p1 += s1; p2 += s2; }
    L->res[j] = mdct[j] / L->floor[j];
    r = fabs(L->res[j]);
    t = p1 - enpeak[j];
    if (t < pre) t = pre;
    L->flag[j] = (r < t) ? ((r < p2) ? 0 : -1) : 1;
  }
}

static float unit(float x) {     /* unitnorm, scales.h:33-41 */
  union { uint32_t i; float f; } u; u.f = x; u.i = (u.i & 0x80000000u) | 0x3f800000u; return u.f;
}

/* noise_normalize, psy.c:4732-4854 */
static void normalize(const orc_psy *p, int limit, lane_t *L, int use_flags, float nepeak, int i, int n, int *out) {
  float *cand[PMAX]; int count = 0, j = 0, k;
  int start = p->normal_p ? p->normal_start - i : n;
  float acc = 0.f;
  if (start > n || nepeak < -0.5) start = n;
  for (; j < start; j++) {
    if (!use_flags) out[j] = rint(L->res[j]);
    else if (L->flag[j] != 1) {
      float ve = sqrt(L->quant[j] / L->floor[j]);
      if (L->raw[j] < 0) { out[j] = -rint(ve); L->res[j] = -ve; } else { out[j] = rint(ve); L->res[j] = ve; }
    }
  }
  for (; j < n; j++) {
    if (use_flags) {
      float ve;
      if (L->flag[j] == 1) continue;
      ve = L->quant[j] / L->floor[j];
      if (ve < .25f && j >= limit - i) {
        acc += ve; cand[count++] = L->quant + j;
        L->res[j] = (L->raw[j] < 0) ? -sqrt(ve) : sqrt(ve);
      } else {
        ve = sqrt(ve);
        if (L->raw[j] < 0) { out[j] = -rint(ve); L->res[j] = -ve; } else { out[j] = rint(ve); L->res[j] = ve; }
        L->quant[j] = out[j] * out[j] * L->floor[j];
      }
    } else {
      float ve = L->res[j] * L->res[j];
      if (ve < .25f) { acc += ve; cand[count++] = L->quant + j; }
      else { out[j] = rint(L->res[j]); L->quant[j] = out[j] * out[j] * L->floor[j]; }
    }
  }
  acc += acc * nepeak * nepeak;
  if (count) {
    int top = ((int)acc) + 1, a, b;
    if (top > n) top = n;
    if (count < top) top = count;
    for (a = 0; a < top; a++) {              /* partial selection sort, largest first (ssort, psy.c:4709) */
      int best = a; float *t;
      for (b = a + 1; b < count; b++) if (*cand[best] < *cand[b]) best = b;
      t = cand[a]; cand[a] = cand[best]; cand[best] = t;
    }
    for (k = 0; k < count; k++) {
      int e = (int)(cand[k] - L->quant);
      if (acc >= p->normal_thresh) { out[e] = unit(L->raw[e]); acc -= 1.f; L->quant[e] = L->floor[e]; }
      else { out[e] = 0; L->quant[e] = 0.f; }
    }
  }
}

/* min_indemnity_dipole_hypot, psy.c:4660-4674 */
static float dipole(float a, float b, float threv) {
  const float th = 0.94;
  float a2 = fabs(a * th), b2 = fabs(b * th);
  if (a > 0.) { if (b > 0.) return a2 + b2; if (a > -b) return a2 - b2 * threv; return -(b2 - a2 * threv); }
  if (b < 0.) return -(a2 + b2);
  if (-a > b) return -(a2 - b2 * threv);
  return b2 - a2 * threv;
}

void orc_couple(const orc_psy *p, const orc_couple_cfg *c, int blobno, float **mdct, float **epeak, float **npeak,
                int **iwork, int *nonzero, int sliding_lowpass, int ch, int lowpassr) {
  int n = p->n, part = p->partition, steps = c->coupling_steps, i, pi, k, j, step;
  int limit = c->pointlimit[p->blockflag * 15 + blobno];
  float pre = c->stereo_th[c->prepointamp[blobno]], post = c->stereo_th[c->postpointamp[blobno]];
  float pre_x = c->stereo_thx[c->prepointamp[blobno]], post_x = c->stereo_thx[c->postpointamp[blobno]];
  float prae = (steps == 1) ? 0.34 : 0.825;
  float resdef[CHMAX];
  lane_t L[CHMAX];
  int nz[CHMAX];
  if (pre_x < pre) pre_x = pre;
  if (post_x < pre) post_x = pre;
  for (k = 0; k < steps && k < CHMAX; k++) resdef[k] = -1.f;

  for (i = 0, pi = 0; i < lowpassr; i += part, pi++) {
    int jn = part > n - i ? n - i : part;
    memcpy(nz, nonzero, sizeof(int) * ch);
    for (k = 0; k < ch; k++) {
      int *io = iwork[k] + i;
      memset(L[k].flag, 0, sizeof(L[k].flag));
      if (nz[k]) {
        for (j = 0; j < jn; j++) L[k].floor[j] = c->fromdB[io[j]];
        classify(limit, pre, post, pre_x, post_x, &L[k], mdct[k] + i, epeak[k] + i, i, jn);
        for (j = 0; j < jn; j++) {
          float m = mdct[k][i + j];
          L[k].quant[j] = L[k].raw[j] = m * m;
          if (m < 0.f) L[k].raw[j] *= -1.f;
          L[k].floor[j] *= L[k].floor[j];
        }
        normalize(p, limit, &L[k], 0, npeak[k][pi], i, jn, io);
      } else {
        for (j = 0; j < jn; j++) { L[k].floor[j] = 1e-10f; L[k].raw[j] = L[k].quant[j] = L[k].res[j] = 0.f; io[j] = 0; }
      }
    }
    for (step = 0; step < steps; step++) {
      int Mi = c->mag[step], Ai = c->ang[step], point = 0;
      lane_t *M = &L[Mi], *A = &L[Ai];
      int *iM = iwork[Mi] + i, *iA = iwork[Ai] + i;
      if (!(nz[Mi] || nz[Ai])) continue;
      nz[Mi] = nz[Ai] = 1;
      if (p->tonefix_end > i) {                       /* aoTuV M6, psy.c:5006-5048 */
        int rp = 0, pp = 0, ap; float def = 0;
        for (j = 0; j < jn; j++)
          if (big(M->res[j]) || big(A->res[j])) {
            if (opposite(M->raw[j], A->raw[j])) rp++; else pp++;
            def += fabs(fabs(M->res[j]) - fabs(A->res[j]));
          }
        ap = rp + pp;
        if (ap != 0) {
          float cur = def = def / ap;
          if (resdef[step] > 0) def = cur * 0.5 + resdef[step] * 0.5;
          resdef[step] = cur;
          if (def > 1.f) for (j = 0; j < jn; j++) if (M->flag[j] == -1 || A->flag[j] == -1) M->flag[j] = 1;
          if ((float)rp / ap >= prae)
            for (j = 0; j < jn; j++)
              if ((M->flag[j] == -1 || A->flag[j] == -1) && opposite(M->raw[j], A->raw[j])) M->flag[j] = 1;
        } else resdef[step] = -1.f;
      }
      for (j = 0; j < jn; j++) {
        if (j < sliding_lowpass - i) {
          if (M->flag[j] == 1 || A->flag[j] == 1) {   /* lossless (psy.c:4626-4658) */
            float fa, fb; int a, b;
            M->raw[j] = fabs(M->raw[j]) + fabs(A->raw[j]);
            M->quant[j] = M->quant[j] + A->quant[j];
            M->flag[j] = A->flag[j] = 1;
            fa = M->res[j]; fb = A->res[j];
            if (fabs(fa) > fabs(fb)) A->res[j] = (fa > 0 ? fa - fb : fb - fa);
            else { A->res[j] = (fb > 0 ? fa - fb : fb - fa); M->res[j] = fb; }
            if (A->res[j] >= fabs(M->res[j]) * 2) { A->res[j] = -A->res[j]; M->res[j] = -M->res[j]; }
            a = iM[j]; b = iA[j];
            if (abs(a) > abs(b)) iA[j] = (a > 0 ? a - b : b - a);
            else { iA[j] = (b > 0 ? a - b : b - a); iM[j] = b; }
            if (iA[j] >= abs(iM[j]) * 2) { iA[j] = -iA[j]; iM[j] = -iM[j]; }
          } else {                                    /* lossy point coupling */
            float hL = .18f, hH = (steps == 1 || step == 3) ? .12f : .04f;
            M->raw[j] = dipole(M->raw[j], A->raw[j], (j < limit - i) ? hL : hH);
            M->quant[j] = fabs(M->raw[j]);
            A->raw[j] = A->quant[j] = 0.f;
            A->flag[j] = 1;
            iA[j] = 0;
            A->res[j] = 0;
            if (npeak[Mi][pi] < -0.5 || npeak[Ai][pi] < -0.5) npeak[Mi][pi] = -1;
            else npeak[Mi][pi] = npeak[Mi][pi] < npeak[Ai][pi] ? npeak[Mi][pi] : npeak[Ai][pi];
            point = 1;
          }
        }
        M->floor[j] = A->floor[j] = M->floor[j] + A->floor[j];
      }
      if (point) normalize(p, limit, M, 1, npeak[Mi][pi], i, jn, iM);
    }
  }
  if (lowpassr < n) for (k = 0; k < ch; k++) memset(iwork[k] + lowpassr, 0, sizeof(int) * (n - lowpassr));
  for (step = 0; step < steps; step++)
    if (nonzero[c->mag[step]] || nonzero[c->ang[step]]) nonzero[c->mag[step]] = nonzero[c->ang[step]] = 1;
}
