/* TEST INFRASTRUCTURE -- CPU oracle: the reference's real forward FFT (FFTPACK-style, smallft.c) and the dB spectrum
 * built on it, restated operation by operation so that the float32 result is reproduced BIT-EXACTLY
 * (transform_port.c states the spectrum from a double-precision FFT and is only within a dB-image tolerance).
 * Follows the scalar paths of /root/reference/lib/smallft.c:
 *   tables          drfti1   smallft.c:5576-5644   (factors 4 first, a factor 2 moved to the front; twiddles from float angles)
 *   radix-2 pass    dradf2   smallft.c:5652-5705
 *   radix-4 pass    dradf4   smallft.c:5707-5807
 *   pass driver     drftf1   smallft.c:6111-6170   (factors taken from the last to the first, ping-pong c <-> ch)
 * and of the spectrum loop /root/reference/lib/mapping0.c:847-888.
 * Every float operation is the reference's (same operands and order; build with -ffp-contract=off).  The form is
 * index-parallel: a pass is three loops over independent work items (k), (k, i) and (k) with every operand index in
 * closed form, out of place from `cc` to `ch` -- the shape of a CUDA kernel with one barrier per pass (DESIGN.md
 * section 9, item 3).  Only the block sizes of the built path are needed (n = 256, 2048: factors 4 and 2).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* drfti1: ifac[0] = n, ifac[1] = number of factors, ifac[2..] = factors; wa = twiddles of every pass but the last factor */
void orc_drft_tables(int n, float *wa, int *ifac) {
  static const int ntryh[4] = {4, 2, 3, 5};
  const float tpi = 6.28318530717958648f;
  int nl = n, nf = 0, j = 0, ntry = 4, i;
  while (nl != 1) {
    while (nl % ntry) { j++; ntry = j < 4 ? ntryh[j] : ntry + 2; }
    nf++;
    ifac[nf + 1] = ntry;
    nl /= ntry;
    if (ntry == 2 && nf != 1) {            /* a factor of two goes to the front */
      for (i = nf; i > 1; i--) ifac[i + 1] = ifac[i];
      ifac[2] = 2;
    }
  }
  ifac[0] = n; ifac[1] = nf;
  {
    const float argh = tpi / n;
    int is = 0, l1 = 1, k1;
    for (k1 = 0; k1 < nf - 1; k1++) {
      const int ip = ifac[k1 + 2], l2 = l1 * ip, ido = n / l2;
      int ld = 0, jj;
      for (jj = 0; jj < ip - 1; jj++) {
        int ii, w = is;
        float argld, fi = 0.f;
        ld += l1;
        argld = (float)ld * argh;
        for (ii = 2; ii < ido; ii += 2) {
          float arg;
          fi += 1.f;
          arg = fi * argld;
          wa[w++] = cos(arg);
          wa[w++] = sin(arg);
        }
        is += ido;
      }
      l1 = l2;
    }
  }
}

static void pass2(int ido, int l1, const float *cc, float *ch, const float *wa1) {
  const int t0 = l1 * ido;
  int k, i;
  for (k = 0; k < l1; k++) {                                   /* items (k): the i = 0 column */
    const int t1 = k * ido, t2 = t0 + k * ido;
    ch[2 * t1] = cc[t1] + cc[t2];
    ch[2 * t1 + 2 * ido - 1] = cc[t1] - cc[t2];
  }
  if (ido < 2) return;
  for (k = 0; k < l1; k++)                                     /* items (k, i) */
    for (i = 2; i < ido; i += 2) {
      const int t1 = k * ido, t3 = t0 + t1 + i, t4 = 2 * t1 + 2 * ido - i, t5 = t1 + i, t6 = 2 * t1 + i;
      const float tr2 = wa1[i - 2] * cc[t3 - 1] + wa1[i - 1] * cc[t3];
      const float ti2 = wa1[i - 2] * cc[t3] - wa1[i - 1] * cc[t3 - 1];
      ch[t6] = cc[t5] + ti2;
      ch[t4] = ti2 - cc[t5];
      ch[t6 - 1] = cc[t5 - 1] + tr2;
      ch[t4 - 1] = cc[t5 - 1] - tr2;
    }
  if (ido & 1) return;
  for (k = 0; k < l1; k++) {                                   /* items (k): the middle column */
    const int t1 = ido + 2 * k * ido, t3 = ido - 1 + k * ido, t2 = t3 + t0;
    ch[t1] = -cc[t2];
    ch[t1 - 1] = cc[t3];
  }
}

static void pass4(int ido, int l1, const float *cc, float *ch, const float *wa1, const float *wa2, const float *wa3) {
  const float hsqt2 = .70710678118654752f;
  const int t0 = l1 * ido, t6 = ido << 1;
  int k, i;
  for (k = 0; k < l1; k++) {                                   /* items (k): the i = 0 column */
    const int t3 = k * ido, t1 = t0 + t3, t4 = 2 * t0 + t3, t2 = 3 * t0 + t3, t5 = t3 << 2;
    const float tr1 = cc[t1] + cc[t2], tr2 = cc[t3] + cc[t4];
    ch[t5] = tr1 + tr2;
    ch[(ido << 2) + t5 - 1] = tr2 - tr1;
    ch[t5 + t6 - 1] = cc[t3] - cc[t4];
    ch[t5 + t6] = cc[t2] - cc[t1];
  }
  if (ido < 2) return;
  for (k = 0; k < l1; k++)                                     /* items (k, i) */
    for (i = 2; i < ido; i += 2) {
      const int t1 = k * ido, t2 = t1 + i, t4 = (t1 << 2) + i, t5 = t6 + (t1 << 2) - i;
      int t3 = t2 + t0;
      float cr2, ci2, cr3, ci3, cr4, ci4, tr1, tr2, tr3, tr4, ti1, ti2, ti3, ti4;
      cr2 = wa1[i - 2] * cc[t3 - 1] + wa1[i - 1] * cc[t3];
      ci2 = wa1[i - 2] * cc[t3] - wa1[i - 1] * cc[t3 - 1];
      t3 += t0;
      cr3 = wa2[i - 2] * cc[t3 - 1] + wa2[i - 1] * cc[t3];
      ci3 = wa2[i - 2] * cc[t3] - wa2[i - 1] * cc[t3 - 1];
      t3 += t0;
      cr4 = wa3[i - 2] * cc[t3 - 1] + wa3[i - 1] * cc[t3];
      ci4 = wa3[i - 2] * cc[t3] - wa3[i - 1] * cc[t3 - 1];
      tr1 = cr2 + cr4; tr4 = cr4 - cr2; ti1 = ci2 + ci4; ti4 = ci2 - ci4;
      ti2 = cc[t2] + ci3; ti3 = cc[t2] - ci3; tr2 = cc[t2 - 1] + cr3; tr3 = cc[t2 - 1] - cr3;
      ch[t4 - 1] = tr1 + tr2;      ch[t4] = ti1 + ti2;
      ch[t5 - 1] = tr3 - ti4;      ch[t5] = tr4 - ti3;
      ch[t4 + t6 - 1] = ti4 + tr3; ch[t4 + t6] = tr4 + ti3;
      ch[t5 + t6 - 1] = tr2 - tr1; ch[t5 + t6] = ti1 - ti2;
    }
  if (ido & 1) return;
  for (k = 0; k < l1; k++) {                                   /* items (k): the middle column */
    const int t1 = t0 + ido - 1 + k * ido, t2 = t1 + (t0 << 1), t4 = ido + k * (ido << 2), t5 = ido << 1, t6b = ido + k * ido;
    const float ti1 = -hsqt2 * (cc[t1] + cc[t2]), tr1 = hsqt2 * (cc[t1] - cc[t2]);
    ch[t4 - 1] = tr1 + cc[t6b - 1];
    ch[t4 + t5 - 1] = cc[t6b - 1] - tr1;
    ch[t4] = ti1 - cc[t1 + t0];
    ch[t4 + t5] = ti1 + cc[t1 + t0];
  }
}

/* drft_forward: data[0] = X0, data[2k-1] / data[2k] = Re / Im X_k, data[n-1] = X_{n/2}.  Returns -1 for a size whose
   factors are not 2 and 4 (the reference's general-radix pass is not restated) */
int orc_drft_forward_exact(float *data, int n) {
  float *wa = calloc(2 * n, sizeof(float)), *ch = malloc(sizeof(float) * n), *c = data;
  int ifac[32], nf, k1, l2 = n, iw = n, na = 1, rc = 0;
  orc_drft_tables(n, wa, ifac);
  nf = ifac[1];
  for (k1 = 0; k1 < nf; k1++) {
    const int ip = ifac[nf - k1 + 1], l1 = l2 / ip, ido = n / l2;
    iw -= (ip - 1) * ido;
    na = 1 - na;
    /* the reference hands the passes wa + iw - 1 with wa = trigcache + n, and indexes wa1[i - 2], i >= 2 */
    if (ip == 4) {
      const float *w1 = wa + iw - 1, *w2 = w1 + ido, *w3 = w2 + ido;
      if (na) pass4(ido, l1, ch, c, w1, w2, w3); else pass4(ido, l1, c, ch, w1, w2, w3);
    } else if (ip == 2) {
      if (na) pass2(ido, l1, ch, c, wa + iw - 1); else pass2(ido, l1, c, ch, wa + iw - 1);
    } else { rc = -1; break; }
    l2 = l1;
  }
  if (!rc && na != 1) memcpy(c, ch, sizeof(float) * n);
  free(wa); free(ch);
  return rc;
}

/* mapping0.c:784-795, 847-888 on the exact transform: logfft[0] from X0, logfft[k] from |X_k|^2, block maximum */
int orc_logfft_exact(const float *w, int n, float *logfft, float *local_ampmax) {
  float *d = malloc(sizeof(float) * n), scale = 4.f / n, scale_dB, mx;
  int j;
  memcpy(d, w, sizeof(float) * n);
  if (orc_drft_forward_exact(d, n)) { free(d); return -1; }
  scale_dB = orc_todB(scale) + .345;
  logfft[0] = scale_dB + orc_todB(d[0]) + .345;
  mx = logfft[0];
  for (j = 1; j < n - 1; j += 2) {
    float temp = d[j] * d[j] + d[j + 1] * d[j + 1];
    temp = logfft[(j + 1) >> 1] = scale_dB + .5f * orc_todB(temp) + .345;
    if (temp > mx) mx = temp;
  }
  if (mx > 0.f) mx = 0.f;
  *local_ampmax = mx;
  free(d);
  return 0;
}
