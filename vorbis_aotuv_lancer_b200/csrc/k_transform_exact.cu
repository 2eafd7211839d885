/* K1x -- the front end with OPERATION-EXACT transforms (the default; vb200_setup_set_exact_transform(s, 0) /
 * VB200_EXACT_TRANSFORM=0 select K1).
 *
 * K1 (k_transform.cu) computes the MDCT and the real FFT with its own radix-4 Stockham formulation; the results agree
 * with the reference's to ~1e-6 of the block RMS (north_star's tolerance), which is enough to flip a quantiser decision
 * in a few percent of the packets.  This kernel instead performs the reference's OWN float operations -- same operands,
 * same order, every product and sum rounded separately -- so that gmdct / logmdct / logfft / local_ampmax, and with them
 * every later stage, are bit-identical to the reference's C path:
 *   mdct_forward   mdct.c:1799-1869  (butterflies :1105-1135, :854-895, :1032-1079, :602-659, :495-524, :432-453;
 *                                     bit reversal :1228-1273; tables mdct_init :54-92)
 *   drft_forward   smallft.c:6770    (drftf1 :6111-6170, dradf4 :5707-5807, dradf2 :5652-5705; tables drfti1 :5576-5644)
 *   spectrum loop  mapping0.c:847-888, logmdct :933-951, _postnoise_detection psy.c:619-648, window window.c:2137-2261
 * The dataflow is the reference's, the form is not.  Every stage is a set of independent work items with their operand
 * indices in closed form (validated on the CPU as oracle/port/mdct_exact_port.c and drft_exact_port.c, bit-exact on
 * every block of every test configuration).  B200 mapping, one CTA of n/8 threads per ch-block:
 *   - the block goes from HBM to registers as float4 quads, windowed on the way; fold item p and its mirror n/4-1-p
 *     need exactly the same two quads, so one thread folds both straight from registers (and drops the windowed
 *     quads into shared memory for the real FFT) -- no scalar staging pass, no strided reads of the block;
 *   - the MDCT's work vector holds complex values as aligned float2 slots, two pad slots per 16: the radix-2
 *     butterfly stages move 8 bytes per access conflict-free, the 32-point groups (one thread per group, registers)
 *     move 16; the twiddles of stage s are regrouped at unit stride (vb200_exact_tables_k1x);
 *   - bit reversal and final rotation are one item: the reversed pair q feeds output pairs q and n/4-1-q directly;
 *   - the real FFT's passes are compiled per (ido, l1): no index division, the three loops (k), (k, i), (k) of a pass
 *     are one item space of exactly n/8 items; its buffers store element e at e + 1 + 2*(e / block) -- the (odd, even)
 *     pairs every pass reads and writes become aligned float2 slots and the block stride becomes odd in slots, so both
 *     the k-major and the i-major lane orders are conflict-free; the first two passes (ido = 1, 4) are a 16-point
 *     transform on registers;
 *   - an MDCT stage and a real-FFT pass share each barrier interval: 8 barriers per long ch-block (was 16).
 * HBM traffic is K1's: n*4 bytes in, 3*N2*4 bytes out.
 */
#include "vb200_internal.h"
#include "dev_util.cuh"
#include <math.h>

/* ---- tables (host): the reference's own values, float conversions of double cos/sin (mdct_init) and double cos/sin of
   float angles (drfti1) ---- */
void vb200_exact_tables_mdct(int n, float *T, int *bitrev) {
  const int n2 = n >> 1;
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  for (int i = 0; i < n / 4; i++) {
    T[i * 2] = (float)cos((M_PI / n) * (4 * i));
    T[i * 2 + 1] = (float)-sin((M_PI / n) * (4 * i));
    T[n2 + i * 2] = (float)cos((M_PI / (2 * n)) * (2 * i + 1));
    T[n2 + i * 2 + 1] = (float)sin((M_PI / (2 * n)) * (2 * i + 1));
  }
  for (int i = 0; i < n / 8; i++) {
    T[n + i * 2] = (float)(cos((M_PI / n) * (4 * i + 2)) * .5);
    T[n + i * 2 + 1] = (float)(-sin((M_PI / n) * (4 * i + 2)) * .5);
  }
  const int mask = (1 << (log2n - 1)) - 1, msb = 1 << (log2n - 2);
  for (int i = 0; i < n / 8; i++) {
    int acc = 0;
    for (int j = 0; msb >> j; j++) if ((msb >> j) & i) acc |= 1 << j;
    bitrev[i * 2] = ((~acc) & mask) - 1;
    bitrev[i * 2 + 1] = acc;
  }
}

/* twiddles of the real FFT for n = 2^k (factors 4...4 and, for an odd k, one 2 that the reference moves to the front);
   wa must hold 2n floats */
void vb200_exact_tables_drft(int n, float *wa) {
  int fac[32], nf = 0, nl = n;
  while (nl % 4 == 0) { fac[nf++] = 4; nl /= 4; }
  if (nl == 2) { for (int i = nf; i > 0; i--) fac[i] = fac[i - 1]; fac[0] = 2; nf++; }
  const float tpi = 6.28318530717958648f;
  const float argh = tpi / n;
  int is = 0, l1 = 1;
  for (int k1 = 0; k1 < nf - 1; k1++) {
    const int ip = fac[k1], l2 = l1 * ip, ido = n / l2;
    int ld = 0;
    for (int j = 0; j < ip - 1; j++) {
      int w = is;
      float fi = 0.f;
      ld += l1;
      const float argld = (float)ld * argh;
      for (int ii = 2; ii < ido; ii += 2) {
        fi += 1.f;
        const float arg = fi * argld;
        wa[w++] = (float)cos((double)arg);
        wa[w++] = (float)sin((double)arg);
      }
      is += ido;
    }
    l1 = l2;
  }
}


/* per-stage twiddles of the MDCT butterflies at unit stride -- stage s, pair c: (T[c*(4<<s)], T[c*(4<<s)+1]), stages
   concatenated (n/8 >> s entries each) -- and the bit-reversal targets as padded slot indices (x0 slot, x1 slot).
   stage_tw must hold n/4 float2, rev n/8 int2. */
void vb200_exact_tables_k1x(int n, const float *T, const int *bitrev, float2 *stage_tw, int2 *rev) {
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  int o = 0;
  for (int s = 0; s < log2n - 6; s++) {
    const int pairs = (n / 8) >> s, trigint = 4 << s;
    for (int c = 0; c < pairs; c++) { stage_tw[o].x = T[c * trigint]; stage_tw[o].y = T[c * trigint + 1]; o++; }
  }
  for (int q = 0; q < n / 8; q++) {
    const int u0 = bitrev[2 * q] >> 1, u1 = bitrev[2 * q + 1] >> 1;      /* complex slots of the upper work half */
    rev[q].x = u0 + 2 * (u0 >> 4);
    rev[q].y = u1 + 2 * (u1 >> 4);
  }
}

namespace {

#define C1_8 .92387953251128675613F
#define C2_8 .70710678118654752441F
#define C3_8 .38268343236508977175F

/* a*b + c*d and a*b - c*d with the three roundings of the C expression */
__device__ __forceinline__ float mad2(float a, float b, float c, float d) { return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }
__device__ __forceinline__ float msu2(float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }

template <int THREADS> __device__ __forceinline__ void cta_sync() {
  if (THREADS <= 32) __syncwarp(); else __syncthreads();
}

/* the hard-coded 32-point end of the butterfly recursion on registers (mdct.c:602-659 -> :495-524 -> :432-453) */
__device__ __forceinline__ void bfly8(float *x) {
  float r0 = add(x[6], x[2]), r1 = sub(x[6], x[2]), r2 = add(x[4], x[0]), r3 = sub(x[4], x[0]);
  x[6] = add(r0, r2); x[4] = sub(r0, r2);
  r0 = sub(x[5], x[1]); r2 = sub(x[7], x[3]);
  x[0] = add(r1, r0); x[2] = sub(r1, r0);
  r0 = add(x[5], x[1]); r1 = add(x[7], x[3]);
  x[3] = add(r2, r3); x[1] = sub(r2, r3);
  x[7] = add(r1, r0); x[5] = sub(r1, r0);
}
__device__ __forceinline__ void bfly16(float *x) {
  float r0 = sub(x[1], x[9]), r1 = sub(x[0], x[8]);
  x[8] = add(x[8], x[0]); x[9] = add(x[9], x[1]);
  x[0] = mul(add(r0, r1), C2_8); x[1] = mul(sub(r0, r1), C2_8);
  r0 = sub(x[3], x[11]); r1 = sub(x[10], x[2]);
  x[10] = add(x[10], x[2]); x[11] = add(x[11], x[3]);
  x[2] = r0; x[3] = r1;
  r0 = sub(x[12], x[4]); r1 = sub(x[13], x[5]);
  x[12] = add(x[12], x[4]); x[13] = add(x[13], x[5]);
  x[4] = mul(sub(r0, r1), C2_8); x[5] = mul(add(r0, r1), C2_8);
  r0 = sub(x[14], x[6]); r1 = sub(x[15], x[7]);
  x[14] = add(x[14], x[6]); x[15] = add(x[15], x[7]);
  x[6] = r0; x[7] = r1;
  bfly8(x); bfly8(x + 8);
}
__device__ __forceinline__ void bfly32(float *x) {
  float r0 = sub(x[30], x[14]), r1 = sub(x[31], x[15]);
  x[30] = add(x[30], x[14]); x[31] = add(x[31], x[15]); x[14] = r0; x[15] = r1;
  r0 = sub(x[28], x[12]); r1 = sub(x[29], x[13]);
  x[28] = add(x[28], x[12]); x[29] = add(x[29], x[13]);
  x[12] = msu2(r0, C1_8, r1, C3_8); x[13] = mad2(r0, C3_8, r1, C1_8);
  r0 = sub(x[26], x[10]); r1 = sub(x[27], x[11]);
  x[26] = add(x[26], x[10]); x[27] = add(x[27], x[11]);
  x[10] = mul(sub(r0, r1), C2_8); x[11] = mul(add(r0, r1), C2_8);
  r0 = sub(x[24], x[8]); r1 = sub(x[25], x[9]);
  x[24] = add(x[24], x[8]); x[25] = add(x[25], x[9]);
  x[8] = msu2(r0, C3_8, r1, C1_8); x[9] = mad2(r1, C3_8, r0, C1_8);
  r0 = sub(x[22], x[6]); r1 = sub(x[7], x[23]);
  x[22] = add(x[22], x[6]); x[23] = add(x[23], x[7]);
  x[6] = r1; x[7] = r0;
  r0 = sub(x[4], x[20]); r1 = sub(x[5], x[21]);
  x[20] = add(x[20], x[4]); x[21] = add(x[21], x[5]);
  x[4] = mad2(r1, C1_8, r0, C3_8); x[5] = msu2(r1, C3_8, r0, C1_8);
  r0 = sub(x[2], x[18]); r1 = sub(x[3], x[19]);
  x[18] = add(x[18], x[2]); x[19] = add(x[19], x[3]);
  x[2] = mul(add(r1, r0), C2_8); x[3] = mul(sub(r1, r0), C2_8);
  r0 = sub(x[0], x[16]); r1 = sub(x[1], x[17]);
  x[16] = add(x[16], x[0]); x[17] = add(x[17], x[1]);
  x[0] = mad2(r1, C3_8, r0, C1_8); x[1] = msu2(r1, C1_8, r0, C3_8);
  bfly16(x); bfly16(x + 16);
}

/* slot of complex value u of the MDCT's work vector: two pad slots per 16, so that the threads that own one 32-point
   group each (stride 16 slots) spread over the banks while every group stays 16-byte aligned and contiguous */
__device__ __forceinline__ int PS(int u) { return u + 2 * (u >> 4); }

/* float offset of the real FFT's twiddles for the pass with this ido (the reference walks the table backwards,
   smallft.c:6111-6170): n = 2048 has factors 2,4,4,4,4,4 (the 2 in front: ido 1024, then 256, 64, 16, 4), n = 256 has
   4,4,4,4 (ido 64, 16, 4); the last factor (ido 1) has no twiddles */
template <int N> __host__ __device__ constexpr int wa_off(int ido) {
  int off = 0, l1 = 1, n = N, first = 1;
  int lg = 0;
  while ((1 << lg) < N) lg++;
  while (n > 1) {
    const int ip = (first && (lg & 1)) ? 2 : 4;
    first = 0;
    const int l2 = l1 * ip, id = N / l2;
    if (id == ido) return off;
    off += (ip - 1) * id;
    l1 = l2;
    n /= ip;
  }
  return -1;
}

/* ---- MDCT: one radix-2 butterfly stage over sub-blocks of (N/2 >> S) floats, one thread per pair (mdct.c:854-895,
   :1032-1079).  Pair c, counted from the top of each half, of sub-block j combines the complex values at slots
   L = j*points/2 + (pairs-1-c) and L + points/4 with the twiddle T[c*(4<<S)]. ---- */
template <int N, int S>
__device__ __forceinline__ void mdct_stage(float2 *wz, const float2 *__restrict__ stw) {
  constexpr int points = (N / 2) >> S, pairs = points >> 2;
  int off = 0;
#pragma unroll
  for (int q = 0; q < S; q++) off += (N / 8) >> q;
  const int g = threadIdx.x;
  const int j = g / pairs, c = g % pairs, cp = pairs - 1 - c;
  const int L = PS(j * (points / 2) + cp), H = PS(j * (points / 2) + cp + points / 4);
  const float2 t = __ldg(stw + off + c);
  const float2 lo = wz[L], hi = wz[H];
  const float r0 = sub(hi.x, lo.x), r1 = sub(hi.y, lo.y);
  wz[H] = make_float2(add(hi.x, lo.x), add(hi.y, lo.y));
  wz[L] = make_float2(mad2(r1, t.y, r0, t.x), msu2(r1, t.x, r0, t.y));
}

/* ---- real FFT, passes ido = 1 and ido = 4 on registers (smallft.c dradf4 twice): thread k < N/16 reads the 16 block
   values x[k + m*N/16] and writes elements 16k .. 16k+15 of the next buffer (blocks of 16, see the layout note) ---- */
template <int N>
__device__ __forceinline__ void rfft_first16(const float *x, float *ch, const float *__restrict__ wa, int k) {
  constexpr int Q = N / 16;
  float y[4][4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const float a0 = x[k + r * Q], a1 = x[k + r * Q + N / 4], a2 = x[k + r * Q + N / 2], a3 = x[k + r * Q + 3 * (N / 4)];
    const float tr1 = add(a1, a3), tr2 = add(a0, a2);
    y[r][0] = add(tr1, tr2); y[r][3] = sub(tr2, tr1); y[r][1] = sub(a0, a2); y[r][2] = sub(a3, a1);
  }
  const float *wa1 = wa + wa_off<N>(4);
  const float2 w1 = __ldg(reinterpret_cast<const float2 *>(wa1)), w2 = __ldg(reinterpret_cast<const float2 *>(wa1 + 4)),
               w3 = __ldg(reinterpret_cast<const float2 *>(wa1 + 8));
  float o[16];
  {  /* i = 0 column */
    const float tr1 = add(y[1][0], y[3][0]), tr2 = add(y[0][0], y[2][0]);
    o[0] = add(tr1, tr2); o[15] = sub(tr2, tr1); o[7] = sub(y[0][0], y[2][0]); o[8] = sub(y[3][0], y[1][0]);
  }
  {  /* i = 2 */
    const float cr2 = mad2(w1.x, y[1][1], w1.y, y[1][2]), ci2 = msu2(w1.x, y[1][2], w1.y, y[1][1]);
    const float cr3 = mad2(w2.x, y[2][1], w2.y, y[2][2]), ci3 = msu2(w2.x, y[2][2], w2.y, y[2][1]);
    const float cr4 = mad2(w3.x, y[3][1], w3.y, y[3][2]), ci4 = msu2(w3.x, y[3][2], w3.y, y[3][1]);
    const float tr1 = add(cr2, cr4), tr4 = sub(cr4, cr2), ti1 = add(ci2, ci4), ti4 = sub(ci2, ci4);
    const float ti2 = add(y[0][2], ci3), ti3 = sub(y[0][2], ci3), tr2 = add(y[0][1], cr3), tr3 = sub(y[0][1], cr3);
    o[1] = add(tr1, tr2);  o[2] = add(ti1, ti2);
    o[5] = sub(tr3, ti4);  o[6] = sub(tr4, ti3);
    o[9] = add(ti4, tr3);  o[10] = add(tr4, ti3);
    o[13] = sub(tr2, tr1); o[14] = sub(ti1, ti2);
  }
  {  /* middle column */
    const float hsqt2 = .70710678118654752f;
    const float ti1 = mul(-hsqt2, add(y[1][3], y[3][3])), tr1 = mul(hsqt2, sub(y[1][3], y[3][3]));
    o[3] = add(tr1, y[0][3]); o[11] = sub(y[0][3], tr1); o[4] = sub(ti1, y[2][3]); o[12] = add(ti1, y[2][3]);
  }
  float *b = ch + k * 18;               /* element e of the block at b[e + 1] */
  b[1] = o[0];
#pragma unroll
  for (int m = 1; m < 8; m++) *reinterpret_cast<float2 *>(b + 2 * m) = make_float2(o[2 * m - 1], o[2 * m]);
  b[16] = o[15];
}

/* ---- real FFT, one radix-4 pass cc -> ch (dradf4, smallft.c:5707-5807) with IDO >= 16.  Buffer layout: element e of a
   buffer whose blocks have B elements lives at float e + 1 + 2*(e / B); cc has blocks of IDO, ch of 4*IDO.  Item
   (k, j): j = 0 is the i = 0 column plus the middle column of block k, j >= 1 the pair i = 2j.  KFAST: k is the fast
   lane index (l1 >= 16), else j. ---- */
template <int N, int IDO, bool KFAST>
__device__ __forceinline__ void rfft_pass4(const float *cc, float *ch, const float *__restrict__ wa) {
  constexpr int L1 = N / (4 * IDO), H = IDO / 2;
  constexpr int SBI = H + 1, SBO = 2 * IDO + 1;          /* block strides in float2 slots */
  const float2 *c2 = reinterpret_cast<const float2 *>(cc);
  float2 *o2 = reinterpret_cast<float2 *>(ch);
  const int t = threadIdx.x;
  const int k = KFAST ? t % L1 : t / H, j = KFAST ? t / L1 : t % H;
  const int ob = k * SBO;
  if (j == 0) {
    const float *f0 = cc + k * (IDO + 2), *f1 = f0 + L1 * (IDO + 2), *f2 = f1 + L1 * (IDO + 2), *f3 = f2 + L1 * (IDO + 2);
    {
      const float c3 = f0[1], c1 = f1[1], c4 = f2[1], c2v = f3[1];
      const float tr1 = add(c1, c2v), tr2 = add(c3, c4);
      ch[2 * ob + 1] = add(tr1, tr2);
      ch[2 * ob + 4 * IDO] = sub(tr2, tr1);
      o2[ob + IDO] = make_float2(sub(c3, c4), sub(c2v, c1));
    }
    {
      const float hsqt2 = .70710678118654752f;
      const float d0 = f0[IDO], d1 = f1[IDO], d2 = f2[IDO], d3 = f3[IDO];
      const float ti1 = mul(-hsqt2, add(d1, d3)), tr1 = mul(hsqt2, sub(d1, d3));
      o2[ob + H] = make_float2(add(tr1, d0), sub(ti1, d2));
      o2[ob + 3 * H] = make_float2(sub(d0, tr1), add(ti1, d2));
    }
  } else {
    const int sb = k * SBI + j;
    const float2 p0 = c2[sb], p1 = c2[sb + L1 * SBI], p2 = c2[sb + 2 * L1 * SBI], p3 = c2[sb + 3 * L1 * SBI];
    const float *wa1 = wa + wa_off<N>(IDO) + 2 * (j - 1);
    const float2 w1 = __ldg(reinterpret_cast<const float2 *>(wa1)), w2 = __ldg(reinterpret_cast<const float2 *>(wa1 + IDO)),
                 w3 = __ldg(reinterpret_cast<const float2 *>(wa1 + 2 * IDO));
    const float cr2 = mad2(w1.x, p1.x, w1.y, p1.y), ci2 = msu2(w1.x, p1.y, w1.y, p1.x);
    const float cr3 = mad2(w2.x, p2.x, w2.y, p2.y), ci3 = msu2(w2.x, p2.y, w2.y, p2.x);
    const float cr4 = mad2(w3.x, p3.x, w3.y, p3.y), ci4 = msu2(w3.x, p3.y, w3.y, p3.x);
    const float tr1 = add(cr2, cr4), tr4 = sub(cr4, cr2), ti1 = add(ci2, ci4), ti4 = sub(ci2, ci4);
    const float ti2 = add(p0.y, ci3), ti3 = sub(p0.y, ci3), tr2 = add(p0.x, cr3), tr3 = sub(p0.x, cr3);
    o2[ob + j] = make_float2(add(tr1, tr2), add(ti1, ti2));
    o2[ob + IDO - j] = make_float2(sub(tr3, ti4), sub(tr4, ti3));
    o2[ob + IDO + j] = make_float2(add(ti4, tr3), add(tr4, ti3));
    o2[ob + 2 * IDO - j] = make_float2(sub(tr2, tr1), sub(ti1, ti2));
  }
}

/* ---- real FFT, the radix-2 pass that ends a transform with an odd log2 (dradf2, smallft.c:5652-5705; l1 = 1,
   ido = N/2): items j < N/4, two per thread ---- */
template <int N, int THREADS>
__device__ __forceinline__ void rfft_pass2_last(const float *cc, float *ch, const float *__restrict__ wa) {
  constexpr int IDO = N / 2, H = IDO / 2;
  const float2 *c2 = reinterpret_cast<const float2 *>(cc);
  float2 *o2 = reinterpret_cast<float2 *>(ch);
#pragma unroll
  for (int rep = 0; rep < H / THREADS; rep++) {
    const int j = threadIdx.x + rep * THREADS;
    if (j == 0) {
      const float c0 = cc[1], c1 = cc[IDO + 2 + 1];
      ch[1] = add(c0, c1);
      ch[2 * IDO] = sub(c0, c1);
      o2[H] = make_float2(cc[IDO], -cc[IDO + 2 + IDO]);
    } else {
      const float2 q0 = c2[j], q1 = c2[H + 1 + j];
      const float2 w = __ldg(reinterpret_cast<const float2 *>(wa + wa_off<N>(IDO) + 2 * (j - 1)));
      const float tr2 = mad2(w.x, q1.x, w.y, q1.y), ti2 = msu2(w.x, q1.y, w.y, q1.x);
      o2[j] = make_float2(add(q0.x, tr2), add(q0.y, ti2));
      o2[IDO - j] = make_float2(sub(q0.x, tr2), sub(ti2, q0.y));
    }
  }
}

template <int N, int THREADS>
__global__ void __launch_bounds__(THREADS, (N == 2048 ? 6 : 16))
k_transform_exact(const DevSetup *__restrict__ S, int W, const float *__restrict__ pcm, const long long *__restrict__ pcm_off,
                  const int *__restrict__ lWv, const int *__restrict__ nWv,
                  const int *__restrict__ block_mode, const int *__restrict__ lW_block_mode,
                  float *__restrict__ gmdct, float *__restrict__ logfft, float *__restrict__ logmdct,
                  float *__restrict__ local_ampmax, float *__restrict__ poste) {
  constexpr int N2 = N / 2, N4 = N / 4, N8 = N / 8, N16 = N / 16;
  static_assert(N == 2048 || N == 256, "block sizes of the built path");
  static_assert(THREADS == N8, "one thread per fold pair / butterfly / bit-reversal item");
  /* real FFT ping-pong: A holds the windowed block (natural order), then the buffers with blocks of 64 and 1024;
     B the buffers with blocks of 16 and 256 and the spectrum */
  __shared__ __align__(16) float bufA[N + 2 * (N / 64) + 4];
  __shared__ __align__(16) float bufB[N + 2 * (N / 16) + 4];
  __shared__ __align__(16) float2 wz[N4 + 2 * (N4 / 16)];       /* the MDCT's upper work half as complex slots, PS() */
  __shared__ double red[2][(THREADS + 31) / 32];
  __shared__ float redf[(THREADS + 31) / 32];

  const int item = blockIdx.x, tid = threadIdx.x;
  const float *src = pcm + (pcm_off ? pcm_off[item] : (long long)item * N);
  const int lW = W ? lWv[item] : 0, nW = W ? nWv[item] : 0;
  const int ln = S->bs[lW], rn = S->bs[nW];
  const float *wl = S->win[lW], *wr = S->win[nW];
  const int leftbegin = N / 4 - ln / 4, leftend = leftbegin + ln / 2;
  const int rightbegin = N / 2 + N / 4 - rn / 4, rightend = rightbegin + rn / 2;

  /* ---- aoTuV M2 post-echo detector on the un-windowed PCM (psy.c:619-648), as in K1 ---- */
  if (poste) {
    float ret = -1.f;
    const int mode = block_mode[item], lwm = lW_block_mode[item];
    if (mode == 2 && lwm == 0 && N >= 2048) {
      const int sn = N >> 2, mn = sn + sn, en = sn + (N >> 1);
      double up = 0., un = 0.;
      for (int i = sn + tid; i < mn; i += THREADS) up += (double)fabsf(__ldg(src + i));
      for (int i = mn + tid; i < en; i += THREADS) un += (double)fabsf(__ldg(src + i));
      up = warp_sum_d(up); un = warp_sum_d(un);
      if ((tid & 31) == 0) { red[0][tid >> 5] = up; red[1][tid >> 5] = un; }
      cta_sync<THREADS>();
      if (tid == 0) {
        double upt = 0., unt = 0.;
        for (int q = 0; q < (THREADS + 31) / 32; q++) { upt += red[0][q]; unt += red[1][q]; }
        if (!(unt / sn > 0.01)) {
          upt *= upt; unt *= unt; unt *= 15;
          if (upt > unt) { ret = (float)(upt - unt); if (ret < 0.1) ret = -1.0f; }
        }
        poste[item] = ret;
      }
    } else if (tid == 0) poste[item] = ret;
  }

  const float *T = S->xm_trig[W];
  const float2 *T2 = reinterpret_cast<const float2 *>(T);
  const float2 *stw = S->xm_stage[W];
  const float *wa = S->xf_wa[W];

  /* ---- window (window.c:2137-2261) + fold + first rotation (mdct.c:1817-1846).  Pair p and its mirror N/4-1-p read the
     even / the odd samples of the same two quads qa, qb of the block; every window boundary is a multiple of 4 ---- */
  {
    const bool aligned = (((uintptr_t)src) & 15) == 0;
    auto quad = [&](int q) -> float4 {
      const int i = 4 * q;
      float4 v;
      if (aligned) v = __ldg(reinterpret_cast<const float4 *>(src) + q);
      else v = make_float4(__ldg(src + i), __ldg(src + i + 1), __ldg(src + i + 2), __ldg(src + i + 3));
      if (i < leftbegin || i >= rightend) v = make_float4(0.f, 0.f, 0.f, 0.f);
      else if (i < leftend) {
        const float4 w = __ldg(reinterpret_cast<const float4 *>(wl + (i - leftbegin)));
        v.x = mul(v.x, w.x); v.y = mul(v.y, w.y); v.z = mul(v.z, w.z); v.w = mul(v.w, w.w);
      } else if (i >= rightbegin) {       /* the falling half reads the table backwards */
        const float4 w = __ldg(reinterpret_cast<const float4 *>(wr + (rn / 2 - 4 - (i - rightbegin))));
        v.x = mul(v.x, w.w); v.y = mul(v.y, w.z); v.z = mul(v.z, w.y); v.w = mul(v.w, w.x);
      }
      return v;
    };
    const int p = tid;
    const bool outer = p < N16;                     /* p in the first segment, its mirror in the third */
    const int qa = 3 * N16 - 1 - p, qb = outer ? 3 * N16 + p : p - N16;
    const float4 A = quad(qa), B = quad(qb);
    reinterpret_cast<float4 *>(bufA)[qa] = A;
    reinterpret_cast<float4 *>(bufA)[qb] = B;
    float r0, r1, s0, s1;
    if (outer) { r0 = add(A.z, B.y); r1 = add(A.x, B.w); s0 = sub(-B.z, A.y); s1 = sub(-B.x, A.w); }
    else       { r0 = sub(A.z, B.y); r1 = sub(A.x, B.w); s0 = sub(B.z, A.y);  s1 = sub(B.x, A.w); }
    const float2 t = __ldg(T2 + N4 - 1 - p), u = __ldg(T2 + p);
    wz[PS(p)] = make_float2(mad2(r1, t.y, r0, t.x), msu2(r1, t.x, r0, t.y));
    wz[PS(N4 - 1 - p)] = make_float2(mad2(s1, u.y, s0, u.x), msu2(s1, u.x, s0, u.y));
  }
  cta_sync<THREADS>();

  /* ---- MDCT butterfly stages and real-FFT passes, one of each per barrier interval ---- */
  if (N == 2048) {
    mdct_stage<N, 0>(wz, stw); if (tid < N16) rfft_first16<N>(bufA, bufB, wa, tid);   cta_sync<THREADS>();
    mdct_stage<N, 1>(wz, stw); rfft_pass4<N, 16, true>(bufB, bufA, wa);               cta_sync<THREADS>();
    mdct_stage<N, 2>(wz, stw); rfft_pass4<N, 64, false>(bufA, bufB, wa);              cta_sync<THREADS>();
    mdct_stage<N, 3>(wz, stw); rfft_pass4<N, 256, false>(bufB, bufA, wa);             cta_sync<THREADS>();
    mdct_stage<N, 4>(wz, stw); rfft_pass2_last<N, THREADS>(bufA, bufB, wa);           cta_sync<THREADS>();
  } else {
    mdct_stage<N, 0>(wz, stw); if (tid < N16) rfft_first16<N>(bufA, bufB, wa, tid);   cta_sync<THREADS>();
    mdct_stage<N, 1>(wz, stw); rfft_pass4<N, 16, false>(bufB, bufA, wa);              cta_sync<THREADS>();
    rfft_pass4<N, 64, false>(bufA, bufB, wa);                                         /* joins the next interval */
  }

  /* ---- 32-point groups on registers, one thread per group (mdct.c:602-659) ---- */
  if (tid < N2 / 32) {
    float4 *g4 = reinterpret_cast<float4 *>(wz + 18 * tid);          /* = PS(16 * tid) */
    float r[32];
#pragma unroll
    for (int i = 0; i < 8; i++) { const float4 v = g4[i]; r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w; }
    bfly32(r);
#pragma unroll
    for (int i = 0; i < 8; i++) g4[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  }
  cta_sync<THREADS>();

  /* ---- dB spectrum and block maximum (mapping0.c:784-795, 847-888): the spectrum is in B, element e at B[e + 1] ---- */
  {
    const float scale_dB = todB_p345(4.f / N);
    float *lf = logfft + (size_t)item * N2;
    const float2 *X = reinterpret_cast<const float2 *>(bufB);
    float mx = -3.0e38f;
#pragma unroll
    for (int rep = 0; rep < N2 / THREADS; rep++) {
      const int k = tid + rep * THREADS;
      const float2 z = X[k];
      float v;
      if (k == 0) v = (float)((double)add(scale_dB, todB_raw(z.y)) + .345);
      else {
        const float pw = add(mul(z.x, z.x), mul(z.y, z.y));
        v = (float)((double)add(scale_dB, mul(.5f, todB_raw(pw))) + .345);
      }
      lf[k] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max_f(mx);
    if ((tid & 31) == 0) redf[tid >> 5] = mx;
  }

  /* ---- bit reversal + rotation (mdct.c:1228-1273) and final rotation + scale (mdct.c:1855-1868), logmdct
     (mapping0.c:933-951): item q produces the work pairs q and N/4-1-q, i.e. outputs q, N/2-1-q, N/4-1-q, N/4+q ---- */
  {
    const int q = tid;
    const int2 rv = __ldg(S->xm_rev[W] + q);
    const float2 X0 = wz[rv.x], X1 = wz[rv.y];
    const float2 tb = __ldg(T2 + N2 + q);                              /* T[N + 2q], T[N + 2q + 1] */
    float r0 = sub(X0.y, X1.y), r1 = add(X0.x, X1.x);
    const float r2 = mad2(r1, tb.x, r0, tb.y), r3 = msu2(r1, tb.y, r0, tb.x);
    r0 = mul(add(X0.y, X1.y), .5f);
    r1 = mul(sub(X0.x, X1.x), .5f);
    const float a0 = add(r0, r2), b0 = add(r1, r3);                     /* w[2q], w[2q+1] */
    const float a1 = sub(r0, r2), b1 = sub(r3, r1);                     /* w[N/2-2-2q], w[N/2-1-2q] */
    const float2 ta = __ldg(T2 + N4 + q), tc = __ldg(T2 + N4 + N4 - 1 - q);  /* T[N/2 + 2p], p = q and N/4-1-q */
    const float scale = 4.f / N;
    float *gm = gmdct + (size_t)item * N2, *lm = logmdct + (size_t)item * N2;
    const float v0 = mul(mad2(a0, ta.x, b0, ta.y), scale), v1 = mul(msu2(a0, ta.y, b0, ta.x), scale);
    const float v2 = mul(mad2(a1, tc.x, b1, tc.y), scale), v3 = mul(msu2(a1, tc.y, b1, tc.x), scale);
    gm[q] = v0;          lm[q] = todB_p345(v0);
    gm[N2 - 1 - q] = v1; lm[N2 - 1 - q] = todB_p345(v1);
    gm[N4 - 1 - q] = v2; lm[N4 - 1 - q] = todB_p345(v2);
    gm[N4 + q] = v3;     lm[N4 + q] = todB_p345(v3);
  }
  cta_sync<THREADS>();
  if (tid == 0) {
    float m = redf[0];
    for (int q = 1; q < (THREADS + 31) / 32; q++) m = fmaxf(m, redf[q]);
    if (m > 0.f) m = 0.f;
    local_ampmax[item] = m;
  }
}

}  // namespace

#ifdef __CUDACC__      /* (tests/emu compiles the kernel above as a host program to check its index arithmetic) */
int vb200_launch_transform_exact(const vb200_setup *s, cudaStream_t st, int W, int items,
                                 const float *pcm, const long long *pcm_off, const int *lW, const int *nW,
                                 const int *block_mode, const int *lW_block_mode,
                                 float *gmdct, float *logfft, float *logmdct, float *local_ampmax, float *poste) {
  if (items <= 0) return VB200_OK;
  if (W) k_transform_exact<2048, 256><<<items, 256, 0, st>>>(s->d, 1, pcm, pcm_off, lW, nW, block_mode, lW_block_mode, gmdct, logfft, logmdct, local_ampmax, poste);
  else   k_transform_exact<256, 32><<<items, 32, 0, st>>>(s->d, 0, pcm, pcm_off, lW, nW, block_mode, lW_block_mode, gmdct, logfft, logmdct, local_ampmax, poste);
  LAUNCH_CHECK();
  return VB200_OK;
}
#endif
