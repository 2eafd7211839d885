/* TEST INFRASTRUCTURE -- CPU oracle, floor-1 stages (see orc.h).
 * Restated from the algorithm of /root/reference/lib/floor1.c: quantise the mask to the
 * 1024-step dB scale, integer least-squares sums per post interval, greedy split with
 * double-precision line fits, then post quantisation/prediction and the integer floor
 * render.  Lines are evaluated in closed form (y at x) rather than by stepping. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

#define MAXP 65

/* floor1.c:285-298 */
static int dbq(float x) {
  int i = x * 7.3142857f + 1023.5f;
  if (i > 1023) return 1023;
  if (i < 0) return 0;
  return i;
}

/* value at x of the integer line (x0,y0)-(x1,y1) as render_line0 / inspect_error step it
   (floor1.c:397-425, 537-586): base slope plus one extra unit per accumulated carry */
static int line_at(int x0, int x1, int y0, int y1, int x) {
  int dy = y1 - y0, adx = x1 - x0, base = dy / adx;
  int rem = abs(dy) - abs(base * adx), k = x - x0;
  int carry = (rem * k) / adx;
  return y0 + base * k + (dy < 0 ? -carry : carry);
}

/* floor1.c:260-283 */
static int point_at(int x0, int x1, int y0, int y1, int x) {
  int dy, adx, off;
  y0 &= 0x7fff; y1 &= 0x7fff;
  dy = y1 - y0; adx = x1 - x0;
  off = (abs(dy) * (x - x0)) / adx;
  return dy < 0 ? y0 - off : y0 + off;
}

typedef struct { int x0, x1; int s[2][6]; } seg_t;   /* s[cls] = {sum x, sum y, sum xx, sum yy, sum xy, count} */

/* floor1.c:477-535: weighted LS line through a run of segments (+ optional pinned ends) */
static int fit(const seg_t *a, int cnt, int *y0, int *y1, float weight) {
  double xb = 0, yb = 0, x2b = 0, y2b = 0, xyb = 0, bn = 0;
  int x0 = a[0].x0, x1 = a[cnt - 1].x1, i;
  for (i = 0; i < cnt; i++) {
    double w = (a[i].s[1][5] + a[i].s[0][5]) * weight / (a[i].s[0][5] + 1) + 1.;
    xb += a[i].s[1][0] + a[i].s[0][0] * w;
    yb += a[i].s[1][1] + a[i].s[0][1] * w;
    x2b += a[i].s[1][2] + a[i].s[0][2] * w;
    y2b += a[i].s[1][3] + a[i].s[0][3] * w;
    xyb += a[i].s[1][4] + a[i].s[0][4] * w;
    bn += a[i].s[1][5] + a[i].s[0][5] * w;
  }
  if (*y0 >= 0) { xb += x0; yb += *y0; x2b += x0 * x0; y2b += *y0 * *y0; xyb += *y0 * x0; bn++; }
  if (*y1 >= 0) { xb += x1; yb += *y1; x2b += x1 * x1; y2b += *y1 * *y1; xyb += *y1 * x1; bn++; }
  {
    double den = bn * x2b - xb * xb;
    if (den > 0.) {
      double ic = (yb * x2b - xyb * xb) / den, sl = (bn * xyb - xb * yb) / den;
      *y0 = rint(ic + sl * x0);
      *y1 = rint(ic + sl * x1);
      if (*y0 > 1023) *y0 = 1023;
      if (*y1 > 1023) *y1 = 1023;
      if (*y0 < 0) *y0 = 0;
      if (*y1 < 0) *y1 = 0;
      return 0;
    }
    *y0 = 0; *y1 = 0;
    return 1;
  }
}

/* floor1.c:537-586 */
static int bad_fit(const orc_floor *f, int x0, int x1, int y0, int y1, const int *q, const unsigned char *loud) {
  int x, mse = 0, cnt = 0;
  for (x = x0; x < x1; x++) {
    int y = line_at(x0, x1, y0, y1, x), v = q[x];
    mse += (y - v) * (y - v);
    cnt++;
    if (loud[x] && (x == x0 || v)) {
      if (y + f->maxover < v) return 1;
      if (y - f->maxunder > v) return 1;
    }
  }
  if (f->maxover * f->maxover / cnt > f->maxerr) return 0;
  if (f->maxunder * f->maxunder / cnt > f->maxerr) return 0;
  if (mse / cnt > f->maxerr) return 1;
  return 0;
}

static int pick(const int *A, const int *B, int i) {
  if (A[i] < 0) return B[i];
  if (B[i] < 0) return A[i];
  return (A[i] + B[i]) >> 1;
}

/* floor1.c:597-750.  Returns 1 and fills posts[] when a fit exists, 0 when the reference returns NULL. */
int orc_floor1_fit(const orc_floor *f, const float *logmdct, const float *logmask, int *out) {
  int n = f->n, posts = f->posts, i, j, any = 0;
  int *q = malloc(sizeof(int) * n);
  unsigned char *loud = malloc(n);
  seg_t seg[MAXP];
  int A[MAXP], B[MAXP], lo[MAXP], hi[MAXP], memo[MAXP];
  for (i = 0; i < n; i++) { q[i] = dbq(logmask[i]); loud[i] = (logmdct[i] + f->twofitatten >= logmask[i]); }
  for (j = 0; j < posts - 1; j++) {
    int x0 = f->sorted[j], x1 = f->sorted[j + 1], xe = x1 >= n ? n - 1 : x1;
    memset(&seg[j], 0, sizeof(seg[j]));
    seg[j].x0 = x0; seg[j].x1 = x1;
    for (i = x0; i <= xe; i++) if (q[i]) {
      int *s = seg[j].s[loud[i] ? 0 : 1];
      s[0] += i; s[1] += q[i]; s[2] += i * i; s[3] += q[i] * q[i]; s[4] += i * q[i]; s[5]++;
    }
    any += seg[j].s[0][5];
  }
  if (!any) { free(q); free(loud); return 0; }
  for (i = 0; i < posts; i++) { A[i] = B[i] = -200; lo[i] = 0; hi[i] = 1; memo[i] = -1; }
  { int y0 = -200, y1 = -200; fit(seg, posts - 1, &y0, &y1, f->twofitweight); A[0] = B[0] = y0; A[1] = B[1] = y1; }
  for (i = 2; i < posts; i++) {
    int sp = f->reverse[i], ln = lo[sp], hn = hi[sp];
    if (memo[ln] == hn) continue;
    memo[ln] = hn;
    {
      int lsp = f->reverse[ln], hsp = f->reverse[hn];
      int lx = f->postlist[ln], hx = f->postlist[hn];
      int ly = pick(A, B, ln), hy = pick(A, B, hn);
      if (bad_fit(f, lx, hx, ly, hy, q, loud)) {
        int ly0 = -200, ly1 = -200, hy0 = -200, hy1 = -200;
        int r0 = fit(seg + lsp, sp - lsp, &ly0, &ly1, f->twofitweight);
        int r1 = fit(seg + sp, hsp - sp, &hy0, &hy1, f->twofitweight);
        if (r0) { ly0 = ly; ly1 = hy0; }
        if (r1) { hy0 = ly1; hy1 = hy; }
        if (r0 && r1) { A[i] = B[i] = -200; }
        else {
          B[ln] = ly0; if (ln == 0) A[ln] = ly0;
          A[i] = ly1; B[i] = hy0;
          A[hn] = hy1; if (hn == 1) B[hn] = hy1;
          if (ly1 >= 0 || hy0 >= 0) {
            for (j = sp - 1; j >= 0 && hi[j] == hn; j--) hi[j] = i;
            for (j = sp + 1; j < posts && lo[j] == ln; j++) lo[j] = i;
          }
        }
      } else { A[i] = B[i] = -200; }
    }
  }
  out[0] = pick(A, B, 0);
  out[1] = pick(A, B, 1);
  for (i = 2; i < posts; i++) {
    int ln = f->lonb[i - 2], hn = f->hinb[i - 2];
    int pred = point_at(f->postlist[ln], f->postlist[hn], out[ln], out[hn], f->postlist[i]);
    int vx = pick(A, B, i);
    out[i] = (vx >= 0 && pred != vx) ? vx : (pred | 0x8000);
  }
  free(q); free(loud);
  return 1;
}

/* floor1.c:752-771 */
void orc_floor1_interpolate(const orc_floor *f, const int *A, const int *B, int del, int *out) {
  int i;
  for (i = 0; i < f->posts; i++) {
    out[i] = ((65536 - del) * (A[i] & 0x7fff) + del * (B[i] & 0x7fff) + 32768) >> 16;
    if ((A[i] & 0x8000) && (B[i] & 0x8000)) out[i] |= 0x8000;
  }
}

/* floor1.c:774-974 minus the bit emission: quantise posts to the multiplier, code each post as a
   wrapped deviation from its neighbours' interpolation (-> vals[]), and render the integer floor. */
int orc_floor1_encode(const orc_floor *f, int *post, int fitted, int *vals, int *ilogmask, int n2) {
  int posts = f->posts, i, j;
  if (!fitted) { memset(ilogmask, 0, sizeof(int) * n2); return 0; }
  for (i = 0; i < posts; i++) {
    int v = post[i] & 0x7fff;
    switch (f->mult) { case 1: v >>= 2; break; case 2: v >>= 3; break; case 3: v /= 12; break; case 4: v >>= 4; break; }
    post[i] = v | (post[i] & 0x8000);
  }
  vals[0] = post[0]; vals[1] = post[1];
  for (i = 2; i < posts; i++) {
    int ln = f->lonb[i - 2], hn = f->hinb[i - 2];
    int pred = point_at(f->postlist[ln], f->postlist[hn], post[ln], post[hn], f->postlist[i]);
    if ((post[i] & 0x8000) || pred == post[i]) { post[i] = pred | 0x8000; vals[i] = 0; }
    else {
      int room = (f->quant_q - pred < pred) ? f->quant_q - pred : pred;
      int d = post[i] - pred;
      if (d < 0) d = (d < -room) ? room - d - 1 : -1 - (d << 1);
      else d = (d >= room) ? d + room : d << 1;
      vals[i] = d;
      post[ln] &= 0x7fff; post[hn] &= 0x7fff;
    }
  }
  {
    int lx = 0, hx = 0, ly = post[0] * f->mult;
    for (j = 1; j < posts; j++) {
      int cur = f->forward[j], hy = post[cur] & 0x7fff;
      if (hy == post[cur]) {
        int x, lim;
        hy *= f->mult; hx = f->postlist[cur];
        lim = n2 < hx ? n2 : hx;
        for (x = lx; x < lim; x++) ilogmask[x] = line_at(lx, hx, ly, hy, x);
        lx = hx; ly = hy;
      }
    }
    for (j = hx; j < n2; j++) ilogmask[j] = ly;
  }
  return 1;
}
