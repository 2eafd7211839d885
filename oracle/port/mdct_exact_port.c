/* TEST INFRASTRUCTURE -- CPU oracle: the reference's forward MDCT restated operation by operation, so that its
 * float32 result is reproduced BIT-EXACTLY (transform_port.c states the transform from its definition in double and
 * is only within ~1e-6 of the block RMS).  Follows the scalar (#else) paths of /root/reference/lib/mdct.c:
 *   tables            mdct_init              mdct.c:54-92
 *   fold + step 1     mdct_forward           mdct.c:1799-1846
 *   butterflies       mdct_butterflies       mdct.c:1105-1135  (first :854-895, generic :1032-1079,
 *                                                               32/16/8-point :602-659, :495-524, :432-453)
 *   bit reversal      mdct_bitreverse        mdct.c:1228-1273
 *   final rotation    mdct_forward           mdct.c:1855-1868
 * Every float operation is the reference's (same operands, same order, every product and sum rounded to float: build
 * with -ffp-contract=off).  The FORM is different on purpose: each stage is a loop over independent work items
 * (a "pair" = one complex value), with the item's operands given in closed form instead of walking pointers -- this
 * is the shape a CUDA kernel needs (one thread per pair, a barrier per stage), and what a bit-exact K1 would be built
 * from.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "orc.h"

#define C1 .92387953251128675613F   /* cos(pi/8)  */
#define C2 .70710678118654752441F   /* cos(2pi/8) */
#define C3 .38268343236508977175F   /* cos(3pi/8) */

/* mdct_init: T[0..n/2) = (cos, -sin)(4i pi/n), T[n/2..n) = (cos, sin)((2i+1) pi/2n), T[n..n+n/4) = .5 (cos, -sin)((4i+2) pi/n);
   bitrev[2i] = (~rev(i) & mask) - 1, bitrev[2i+1] = rev(i), rev over log2n-1 bits */
void orc_mdct_tables(int n, float *T, int *bitrev) {
  int i, j, n2 = n >> 1, log2n = (int)rint(log((float)n) / log(2.f));
  for (i = 0; i < n / 4; i++) {
    T[i * 2] = (float)cos((M_PI / n) * (4 * i));
    T[i * 2 + 1] = (float)-sin((M_PI / n) * (4 * i));
    T[n2 + i * 2] = (float)cos((M_PI / (2 * n)) * (2 * i + 1));
    T[n2 + i * 2 + 1] = (float)sin((M_PI / (2 * n)) * (2 * i + 1));
  }
  for (i = 0; i < n / 8; i++) {
    T[n + i * 2] = (float)(cos((M_PI / n) * (4 * i + 2)) * .5);
    T[n + i * 2 + 1] = (float)(-sin((M_PI / n) * (4 * i + 2)) * .5);
  }
  {
    int mask = (1 << (log2n - 1)) - 1, msb = 1 << (log2n - 2);
    for (i = 0; i < n / 8; i++) {
      int acc = 0;
      for (j = 0; msb >> j; j++) if ((msb >> j) & i) acc |= 1 << j;
      bitrev[i * 2] = ((~acc) & mask) - 1;
      bitrev[i * 2 + 1] = acc;
    }
  }
}

/* one radix-2 stage over sub-blocks of `points` floats: pair c (counted from the TOP of each half) of sub-block j
   combines upper u = base + points/2 + 2c' and lower l = base + 2c' (c' = points/4 - 1 - c) with the twiddle at
   T[c * trigint].  butterfly_first is the stage with one sub-block and trigint = 4. */
static void stage(float *x, int total, int points, int trigint, const float *T) {
  int pairs = points >> 2, g;
  for (g = 0; g < total >> 2; g++) {            /* independent work items */
    int j = g / pairs, c = g - j * pairs, cp = pairs - 1 - c;
    float *lo = x + j * points + 2 * cp, *hi = lo + (points >> 1);
    const float *t = T + c * trigint;
    float r0 = hi[0] - lo[0], r1 = hi[1] - lo[1];
    hi[0] += lo[0];
    hi[1] += lo[1];
    lo[0] = r1 * t[1] + r0 * t[0];
    lo[1] = r1 * t[0] - r0 * t[1];
  }
}

/* the hard-coded 32-point end of the recursion (mdct_butterfly_32 -> 2 x _16 -> 4 x _8), per 32-float group */
static void bfly8(float *x) {
  float r0 = x[6] + x[2], r1 = x[6] - x[2], r2 = x[4] + x[0], r3 = x[4] - x[0];
  x[6] = r0 + r2; x[4] = r0 - r2;
  r0 = x[5] - x[1]; r2 = x[7] - x[3];
  x[0] = r1 + r0; x[2] = r1 - r0;
  r0 = x[5] + x[1]; r1 = x[7] + x[3];
  x[3] = r2 + r3; x[1] = r2 - r3;
  x[7] = r1 + r0; x[5] = r1 - r0;
}
static void bfly16(float *x) {
  float r0 = x[1] - x[9], r1 = x[0] - x[8];
  x[8] += x[0]; x[9] += x[1];
  x[0] = (r0 + r1) * C2; x[1] = (r0 - r1) * C2;
  r0 = x[3] - x[11]; r1 = x[10] - x[2];
  x[10] += x[2]; x[11] += x[3];
  x[2] = r0; x[3] = r1;
  r0 = x[12] - x[4]; r1 = x[13] - x[5];
  x[12] += x[4]; x[13] += x[5];
  x[4] = (r0 - r1) * C2; x[5] = (r0 + r1) * C2;
  r0 = x[14] - x[6]; r1 = x[15] - x[7];
  x[14] += x[6]; x[15] += x[7];
  x[6] = r0; x[7] = r1;
  bfly8(x); bfly8(x + 8);
}
static void bfly32(float *x) {
  float r0 = x[30] - x[14], r1 = x[31] - x[15];
  x[30] += x[14]; x[31] += x[15]; x[14] = r0; x[15] = r1;
  r0 = x[28] - x[12]; r1 = x[29] - x[13];
  x[28] += x[12]; x[29] += x[13];
  x[12] = r0 * C1 - r1 * C3; x[13] = r0 * C3 + r1 * C1;
  r0 = x[26] - x[10]; r1 = x[27] - x[11];
  x[26] += x[10]; x[27] += x[11];
  x[10] = (r0 - r1) * C2; x[11] = (r0 + r1) * C2;
  r0 = x[24] - x[8]; r1 = x[25] - x[9];
  x[24] += x[8]; x[25] += x[9];
  x[8] = r0 * C3 - r1 * C1; x[9] = r1 * C3 + r0 * C1;
  r0 = x[22] - x[6]; r1 = x[7] - x[23];
  x[22] += x[6]; x[23] += x[7];
  x[6] = r1; x[7] = r0;
  r0 = x[4] - x[20]; r1 = x[5] - x[21];
  x[20] += x[4]; x[21] += x[5];
  x[4] = r1 * C1 + r0 * C3; x[5] = r1 * C3 - r0 * C1;
  r0 = x[2] - x[18]; r1 = x[3] - x[19];
  x[18] += x[2]; x[19] += x[3];
  x[2] = (r1 + r0) * C2; x[3] = (r1 - r0) * C2;
  r0 = x[0] - x[16]; r1 = x[1] - x[17];
  x[16] += x[0]; x[17] += x[1];
  x[0] = r1 * C3 + r0 * C1; x[1] = r1 * C1 - r0 * C3;
  bfly16(x); bfly16(x + 16);
}

void orc_mdct_forward_exact(const float *in, float *out, int n) {
  int n2 = n >> 1, n4 = n >> 2, n8 = n >> 3, log2n = (int)rint(log((float)n) / log(2.f));
  float *T = malloc(sizeof(float) * (n + n / 4)), *w = malloc(sizeof(float) * n), *w2 = w + n2;
  int *bitrev = malloc(sizeof(int) * (n / 4));
  int p, s;
  orc_mdct_tables(n, T, bitrev);

  /* fold + first rotation: output pair p (w2[2p], w2[2p+1]), p < n/4, takes in[a], in[b] around 3n/4 going down and
     in[c], in[d] going up from 3n/4+1 / 1, with the sign pattern of the three segments, twiddle T[n/2 - 2 - 2p] */
  for (p = 0; p < n4; p++) {
    const int i = 2 * p;
    const float *t = T + n2 - 2 - i;
    float r0, r1;
    if (i < n8) {                       /* x0 = in + 3n/4 - 4(p+1), x1 = in + 3n/4 + 1 + 4p */
      const float *x0 = in + n2 + n4 - 4 * (p + 1), *x1 = in + n2 + n4 + 1 + 4 * p;
      r0 = x0[2] + x1[0]; r1 = x0[0] + x1[2];
    } else if (i < n2 - n8) {           /* x1 restarts at in + 1 */
      const float *x0 = in + n2 + n4 - 4 * (p + 1), *x1 = in + 1 + 4 * (p - n8 / 2);
      r0 = x0[2] - x1[0]; r1 = x0[0] - x1[2];
    } else {                            /* x0 restarts at in + n */
      const float *x0 = in + n - 4 * (p - (n2 - n8) / 2 + 1), *x1 = in + 1 + 4 * (p - n8 / 2);
      r0 = -x0[2] - x1[0]; r1 = -x0[0] - x1[2];
    }
    w2[i] = r1 * t[1] + r0 * t[0];
    w2[i + 1] = r1 * t[0] - r0 * t[1];
  }

  /* butterflies on the upper half: log2n - 6 radix-2 stages (sub-block sizes n/2, n/4, ... 64), then 32-point groups */
  for (s = 0; s < log2n - 6; s++) stage(w2, n2, n2 >> s, 4 << s, T);
  for (p = 0; p < n2; p += 32) bfly32(w2 + p);

  /* bit reversal + rotation into the lower half: item q < n/8 reads the two complex values at bitrev[2q], bitrev[2q+1]
     of the upper half and writes w[2q], w[2q+1] and their mirror w[n/2-2-2q], w[n/2-1-2q] */
  {
    const float *Tb = T + n;
    int q;
    for (q = 0; q < n8; q++) {
      const float *x0 = w2 + bitrev[2 * q], *x1 = w2 + bitrev[2 * q + 1];
      float r0 = x0[1] - x1[1], r1 = x0[0] + x1[0];
      float r2 = r1 * Tb[2 * q] + r0 * Tb[2 * q + 1], r3 = r1 * Tb[2 * q + 1] - r0 * Tb[2 * q];
      r0 = (x0[1] + x1[1]) * .5f;
      r1 = (x0[0] - x1[0]) * .5f;
      w[2 * q] = r0 + r2;
      w[n2 - 2 - 2 * q] = r0 - r2;
      w[2 * q + 1] = r1 + r3;
      w[n2 - 1 - 2 * q] = r3 - r1;
    }
  }

  /* final rotation and scale */
  {
    const float scale = 4.f / n, *t = T + n2;
    for (p = 0; p < n4; p++) {
      out[p] = (w[2 * p] * t[2 * p] + w[2 * p + 1] * t[2 * p + 1]) * scale;
      out[n2 - 1 - p] = (w[2 * p] * t[2 * p + 1] - w[2 * p + 1] * t[2 * p]) * scale;
    }
  }
  free(T); free(w); free(bitrev);
}
