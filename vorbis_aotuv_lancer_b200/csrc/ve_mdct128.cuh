/* The block-size detector's 128-point forward MDCT with the reference's OWN float operations (mdct_forward for n = 128:
 * mdct.c:1799-1869, butterflies :1105-1135 -> first stage :854-895, 32-point groups :602-659 / :495-524 / :432-453, bit
 * reversal :1228-1273, tables mdct_init :54-92) -- same operands, same order, every product and sum rounded to float on
 * its own, so that the spectrum _ve_amp (envelope.c:101-217) looks at is bit-identical to the reference's and the
 * block-size decisions are the reference's by construction, on the host (vb200_segment_stream) and on the device (K0)
 * alike.  Same dataflow as K1x (k_transform_exact.cu) in its n = 128 form: one "thread" does the whole transform on
 * arrays with compile-time indices (registers on the device).
 *
 * T: the table of vb200_exact_tables_mdct(128): T[0..64) (cos, -sin)(4i pi/n), T[64..128) (cos, sin)((2i+1) pi/2n),
 * T[128..160) .5 (cos, -sin)((4i+2) pi/n).
 */
#ifndef VB200_VE_MDCT128_CUH
#define VB200_VE_MDCT128_CUH

#ifdef __CUDA_ARCH__
#define VE_FM(a, b) __fmul_rn((a), (b))
#define VE_FA(a, b) __fadd_rn((a), (b))
#define VE_FS(a, b) __fsub_rn((a), (b))
#else
/* host: the library is built without FMA contraction (Makefile: -ffp-contract=off) */
#define VE_FM(a, b) ((float)((a) * (b)))
#define VE_FA(a, b) ((float)((a) + (b)))
#define VE_FS(a, b) ((float)((a) - (b)))
#endif
#define VE_MAD2(a, b, c, d) VE_FA(VE_FM(a, b), VE_FM(c, d))      /* a*b + c*d, three roundings */
#define VE_MSU2(a, b, c, d) VE_FS(VE_FM(a, b), VE_FM(c, d))      /* a*b - c*d */

namespace ve128 {

constexpr float C1 = .92387953251128675613F, C2 = .70710678118654752441F, C3 = .38268343236508977175F;

/* bit reversal targets of mdct_init for n = 128: item i < 16 reads the complex values at float offsets rev0(i), rev1(i) of
   the upper work half: rev1 = the 4 bits of i reversed into bits 5..2, rev0 = 62 - rev1 */
__host__ __device__ constexpr int rev1(int i) { return ((i & 1) << 5) | ((i & 2) << 3) | ((i & 4) << 1) | ((i & 8) >> 1); }
__host__ __device__ constexpr int rev0(int i) { return 62 - rev1(i); }

__host__ __device__ inline void bfly8(float *x) {
  float r0 = VE_FA(x[6], x[2]), r1 = VE_FS(x[6], x[2]), r2 = VE_FA(x[4], x[0]), r3 = VE_FS(x[4], x[0]);
  x[6] = VE_FA(r0, r2); x[4] = VE_FS(r0, r2);
  r0 = VE_FS(x[5], x[1]); r2 = VE_FS(x[7], x[3]);
  x[0] = VE_FA(r1, r0); x[2] = VE_FS(r1, r0);
  r0 = VE_FA(x[5], x[1]); r1 = VE_FA(x[7], x[3]);
  x[3] = VE_FA(r2, r3); x[1] = VE_FS(r2, r3);
  x[7] = VE_FA(r1, r0); x[5] = VE_FS(r1, r0);
}
__host__ __device__ inline void bfly16(float *x) {
  float r0 = VE_FS(x[1], x[9]), r1 = VE_FS(x[0], x[8]);
  x[8] = VE_FA(x[8], x[0]); x[9] = VE_FA(x[9], x[1]);
  x[0] = VE_FM(VE_FA(r0, r1), C2); x[1] = VE_FM(VE_FS(r0, r1), C2);
  r0 = VE_FS(x[3], x[11]); r1 = VE_FS(x[10], x[2]);
  x[10] = VE_FA(x[10], x[2]); x[11] = VE_FA(x[11], x[3]);
  x[2] = r0; x[3] = r1;
  r0 = VE_FS(x[12], x[4]); r1 = VE_FS(x[13], x[5]);
  x[12] = VE_FA(x[12], x[4]); x[13] = VE_FA(x[13], x[5]);
  x[4] = VE_FM(VE_FS(r0, r1), C2); x[5] = VE_FM(VE_FA(r0, r1), C2);
  r0 = VE_FS(x[14], x[6]); r1 = VE_FS(x[15], x[7]);
  x[14] = VE_FA(x[14], x[6]); x[15] = VE_FA(x[15], x[7]);
  x[6] = r0; x[7] = r1;
  bfly8(x); bfly8(x + 8);
}
__host__ __device__ inline void bfly32(float *x) {
  float r0 = VE_FS(x[30], x[14]), r1 = VE_FS(x[31], x[15]);
  x[30] = VE_FA(x[30], x[14]); x[31] = VE_FA(x[31], x[15]); x[14] = r0; x[15] = r1;
  r0 = VE_FS(x[28], x[12]); r1 = VE_FS(x[29], x[13]);
  x[28] = VE_FA(x[28], x[12]); x[29] = VE_FA(x[29], x[13]);
  x[12] = VE_MSU2(r0, C1, r1, C3); x[13] = VE_MAD2(r0, C3, r1, C1);
  r0 = VE_FS(x[26], x[10]); r1 = VE_FS(x[27], x[11]);
  x[26] = VE_FA(x[26], x[10]); x[27] = VE_FA(x[27], x[11]);
  x[10] = VE_FM(VE_FS(r0, r1), C2); x[11] = VE_FM(VE_FA(r0, r1), C2);
  r0 = VE_FS(x[24], x[8]); r1 = VE_FS(x[25], x[9]);
  x[24] = VE_FA(x[24], x[8]); x[25] = VE_FA(x[25], x[9]);
  x[8] = VE_MSU2(r0, C3, r1, C1); x[9] = VE_MAD2(r1, C3, r0, C1);
  r0 = VE_FS(x[22], x[6]); r1 = VE_FS(x[7], x[23]);
  x[22] = VE_FA(x[22], x[6]); x[23] = VE_FA(x[23], x[7]);
  x[6] = r1; x[7] = r0;
  r0 = VE_FS(x[4], x[20]); r1 = VE_FS(x[5], x[21]);
  x[20] = VE_FA(x[20], x[4]); x[21] = VE_FA(x[21], x[5]);
  x[4] = VE_MAD2(r1, C1, r0, C3); x[5] = VE_MSU2(r1, C3, r0, C1);
  r0 = VE_FS(x[2], x[18]); r1 = VE_FS(x[3], x[19]);
  x[18] = VE_FA(x[18], x[2]); x[19] = VE_FA(x[19], x[3]);
  x[2] = VE_FM(VE_FA(r1, r0), C2); x[3] = VE_FM(VE_FS(r1, r0), C2);
  r0 = VE_FS(x[0], x[16]); r1 = VE_FS(x[1], x[17]);
  x[16] = VE_FA(x[16], x[0]); x[17] = VE_FA(x[17], x[1]);
  x[0] = VE_MAD2(r1, C3, r0, C1); x[1] = VE_MSU2(r1, C1, r0, C3);
  bfly16(x); bfly16(x + 16);
}

/* in(t): windowed sample t of the step (t = 0..127), each read exactly once; out[64]: the coefficients; T: the table.
   IN is a callable so that the device can read its samples from a transposed shared-memory tile and window them on
   the way. */
template <typename IN>
__host__ __device__ inline void forward(IN in, float *out, const float *__restrict__ T) {
  float w2[64];
  /* fold + first rotation (mdct.c:1817-1846): pair p of the upper work half */
#pragma unroll
  for (int p = 0; p < 32; p++) {
    const int i = 2 * p;
    const float t0 = T[62 - i], t1 = T[63 - i];
    float r0, r1;
    if (p < 8) {
      const int x0 = 96 - 4 * (p + 1), x1 = 97 + 4 * p;
      r0 = VE_FA(in(x0 + 2), in(x1)); r1 = VE_FA(in(x0), in(x1 + 2));
    } else if (p < 24) {
      const int x0 = 96 - 4 * (p + 1), x1 = 1 + 4 * (p - 8);
      r0 = VE_FS(in(x0 + 2), in(x1)); r1 = VE_FS(in(x0), in(x1 + 2));
    } else {
      const int x0 = 128 - 4 * (p - 24 + 1), x1 = 1 + 4 * (p - 8);
      r0 = VE_FS(-in(x0 + 2), in(x1)); r1 = VE_FS(-in(x0), in(x1 + 2));
    }
    w2[i] = VE_MAD2(r1, t1, r0, t0);
    w2[i + 1] = VE_MSU2(r1, t0, r0, t1);
  }
  /* the one radix-2 stage of n = 128 (mdct_butterfly_first, 64 points, trigint 4) */
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const int lo = 2 * (15 - c), hi = lo + 32;
    const float t0 = T[4 * c], t1 = T[4 * c + 1];
    const float r0 = VE_FS(w2[hi], w2[lo]), r1 = VE_FS(w2[hi + 1], w2[lo + 1]);
    w2[hi] = VE_FA(w2[hi], w2[lo]);
    w2[hi + 1] = VE_FA(w2[hi + 1], w2[lo + 1]);
    w2[lo] = VE_MAD2(r1, t1, r0, t0);
    w2[lo + 1] = VE_MSU2(r1, t0, r0, t1);
  }
  bfly32(w2); bfly32(w2 + 32);
  /* bit reversal + rotation (mdct.c:1228-1273) and final rotation + scale (mdct.c:1855-1868): item q yields the work
     pairs q and 31-q, i.e. the coefficients q, 63-q, 31-q, 32+q */
  const float scale = 4.f / 128.f;
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const float x00 = w2[rev0(q)], x01 = w2[rev0(q) + 1], x10 = w2[rev1(q)], x11 = w2[rev1(q) + 1];
    const float tb0 = T[128 + 2 * q], tb1 = T[128 + 2 * q + 1];
    float r0 = VE_FS(x01, x11), r1 = VE_FA(x00, x10);
    const float r2 = VE_MAD2(r1, tb0, r0, tb1), r3 = VE_MSU2(r1, tb1, r0, tb0);
    r0 = VE_FM(VE_FA(x01, x11), .5f);
    r1 = VE_FM(VE_FS(x00, x10), .5f);
    const float a0 = VE_FA(r0, r2), b0 = VE_FA(r1, r3);      /* w[2q], w[2q+1] */
    const float a1 = VE_FS(r0, r2), b1 = VE_FS(r3, r1);      /* w[62-2q], w[63-2q] */
    const float ta0 = T[64 + 2 * q], ta1 = T[64 + 2 * q + 1], tc0 = T[64 + 2 * (31 - q)], tc1 = T[64 + 2 * (31 - q) + 1];
    out[q] = VE_FM(VE_MAD2(a0, ta0, b0, ta1), scale);
    out[63 - q] = VE_FM(VE_MSU2(a0, ta1, b0, ta0), scale);
    out[31 - q] = VE_FM(VE_MAD2(a1, tc0, b1, tc1), scale);
    out[32 + q] = VE_FM(VE_MSU2(a1, tc1, b1, tc0), scale);
  }
}

}  // namespace ve128
#endif
