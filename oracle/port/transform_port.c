/* TEST INFRASTRUCTURE -- CPU oracle, front-end stages (see orc.h): window, forward MDCT,
 * FFT/dB spectra.  The transforms are stated from their definitions and evaluated in double
 * (the reference's float butterflies, mdct.c:1799-1869 / smallft.c, agree with the
 * definition to ~1e-6 of the block RMS -- SURVEY.md App. C-7 -- which is the tolerance the
 * parity tests state). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

/* window.c:2137-2261: zero head, rising half-window, flat top, falling half-window, zero tail */
void orc_window(float *d, int n, int ln, int rn, const float *wl, const float *wr) {
  int lb = n / 4 - ln / 4, le = lb + ln / 2, rb = n / 2 + n / 4 - rn / 4, re = rb + rn / 2, i, p;
  for (i = 0; i < lb; i++) d[i] = 0.f;
  for (p = 0; i < le; i++, p++) d[i] *= wl[p];
  for (i = rb, p = rn / 2 - 1; i < re; i++, p--) d[i] *= wr[p];
  for (; i < n; i++) d[i] = 0.f;
}

/* in-place iterative radix-2 complex FFT in double, forward sign */
static void fft_d(double *re, double *im, int m) {
  int i, j, len;
  for (i = 1, j = 0; i < m; i++) {
    int bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (len = 2; len <= m; len <<= 1) {
    double ang = -2.0 * M_PI / len;
    for (i = 0; i < m; i += len)
      for (j = 0; j < len / 2; j++) {
        double wr = cos(ang * j), wi = sin(ang * j);
        double xr = re[i + j + len / 2] * wr - im[i + j + len / 2] * wi;
        double xi = re[i + j + len / 2] * wi + im[i + j + len / 2] * wr;
        re[i + j + len / 2] = re[i + j] - xr; im[i + j + len / 2] = im[i + j] - xi;
        re[i + j] += xr; im[i + j] += xi;
      }
  }
}

/* out[k] = (4/n) sum_t in[t] cos(pi/(2n) (2t+1+n/2)(2k+1)),  k < n/2   (mdct.c forward + scale) */
void orc_mdct_forward(const float *in, float *out, int n) {
  int N = n / 2, q = N / 2, i, m;
  double *u = malloc(sizeof(double) * N), *re = malloc(sizeof(double) * q), *im = malloc(sizeof(double) * q);
  for (i = 0; i < q; i++) u[i] = -(double)in[3 * q - 1 - i] - (double)in[3 * q + i];
  for (i = q; i < N; i++) u[i] = (double)in[i - q] - (double)in[3 * q - 1 - i];
  for (m = 0; m < q; m++) {
    double a = -M_PI * (8.0 * m + 1.0) / (8.0 * N), c = cos(a), s = sin(a);
    double x = u[2 * m], y = u[N - 1 - 2 * m];
    re[m] = x * c - y * s; im[m] = x * s + y * c;
  }
  fft_d(re, im, q);
  for (m = 0; m < q; m++) {
    double a = -M_PI * (8.0 * m + 1.0) / (8.0 * N), c = cos(a), s = sin(a);
    double yr = re[m] * c - im[m] * s, yi = re[m] * s + im[m] * c;
    out[2 * m] = (float)(yr * (4.0 / n));
    out[N - 1 - 2 * m] = (float)(-yi * (4.0 / n));
  }
  free(u); free(re); free(im);
}

/* mapping0.c:784-795, 847-888: spectrum of the windowed block in the reference's pseudo-dB */
void orc_logfft(const float *w, int n, float *logfft, float *local_ampmax) {
  int k, N2 = n / 2;
  double *re = malloc(sizeof(double) * n), *im = malloc(sizeof(double) * n);
  float scale = 4.f / n, scale_dB, mx;
  for (k = 0; k < n; k++) { re[k] = w[k]; im[k] = 0; }
  fft_d(re, im, n);
  scale_dB = orc_todB(scale) + .345;
  { float x0 = (float)re[0]; logfft[0] = scale_dB + orc_todB(x0) + .345; }
  mx = logfft[0];
  for (k = 1; k < N2; k++) {
    float a = (float)re[k], b = (float)im[k];
    float t = a * a + b * b;
    logfft[k] = scale_dB + .5f * orc_todB(t) + .345;
    if (logfft[k] > mx) mx = logfft[k];
  }
  if (mx > 0.f) mx = 0.f;
  *local_ampmax = mx;
  free(re); free(im);
}

/* mapping0.c:933-951 */
void orc_logmdct(const float *mdct, int n2, float *logmdct) {
  int j;
  for (j = 0; j < n2; j++) logmdct[j] = orc_todB(mdct[j]) + .345;
}
