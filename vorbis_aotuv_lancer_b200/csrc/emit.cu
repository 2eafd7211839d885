/* Host-side codeword emission: turns the per-block results of the CUDA path into the Vorbis
 * audio packet, in the exact bit order of the reference.
 *
 * Replaces the emitting halves of mapping0_forward (packet type / mode / window bits,
 * mapping0.c:1209-1218), floor1_encode (floor1.c:850-942) and vorbis_book_encode
 * (codebook.c:402-410) with our own LSb-first bit writer (the bit layout is the one libogg's
 * oggpack_write defines; libogg itself is not used).  north_star keeps this on the host.
 */
#include <stdlib.h>
#include <string.h>
#include "vb200_internal.h"
#include "emit.h"

void BitWriter::grow(size_t need) {
  size_t nc = cap ? cap * 2 : 4096;
  while (nc < need) nc *= 2;
  buf = (unsigned char *)realloc(buf, nc);
  memset(buf + cap, 0, nc - cap);
  cap = nc;
}

/* append `bits` (<=32) low bits of value, LSb first */
void BitWriter::put(uint32_t value, int bits) {
  if (bits <= 0) return;
  const size_t byte = bitpos >> 3;
  if (byte + 9 > cap) grow(byte + 9);
  if (bits < 32) value &= (1u << bits) - 1u;
  uint64_t v = (uint64_t)value << (bitpos & 7);
  unsigned char *p = buf + byte;
  for (int k = 0; v; k++, v >>= 8) p[k] |= (unsigned char)(v & 0xff);
  bitpos += bits;
}

static inline int ilog_(unsigned v) { int r = 0; while (v) { r++; v >>= 1; } return r; }

static inline void book_put(const vb200_setup *s, BitWriter &w, int book, int entry) {
  const HostBook &b = s->book[book];
  if (entry < 0 || entry >= b.entries) return;          /* vorbis_book_encode, codebook.c:403 */
  w.put(b.code[entry], b.len[entry]);
}

void vb200_emit_header(const vb200_setup *s, BitWriter &w, int W, int lW, int nW) {
  w.put(0, 1);                        /* packet type: audio */
  w.put((uint32_t)W, s->h.modebits);  /* mode number == block size flag in every setup the reference builds */
  if (W) { w.put((uint32_t)lW, 1); w.put((uint32_t)nW, 1); }
}

/* floor1_encode, floor1.c:850-942: `vals` are the wrapped post deviations the device computed */
void vb200_emit_floor(const vb200_setup *s, BitWriter &w, int floor_index, int nonzero, const int *vals) {
  if (!nonzero) { w.put(0, 1); return; }
  const DevFloor &F = s->h.flr[floor_index];
  const HostFloor &H = s->hflr[floor_index];
  w.put(1, 1);
  const int qb = ilog_((unsigned)(F.quant_q - 1));
  w.put((uint32_t)vals[0], qb);
  w.put((uint32_t)vals[1], qb);
  for (int i = 0, j = 2; i < H.partitions; i++) {
    const int cls = H.partclass[i], cdim = H.class_dim[cls], csubbits = H.class_subs[cls], csub = 1 << csubbits;
    int bookas[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cval = 0, cshift = 0;
    if (csubbits) {
      int maxval[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < csub; k++) {
        const int bn = H.class_subbook[cls * 8 + k];
        maxval[k] = bn < 0 ? 1 : s->book[bn].entries;
      }
      for (int k = 0; k < cdim; k++) {
        for (int l = 0; l < csub; l++) if (vals[j + k] < maxval[l]) { bookas[k] = l; break; }
        cval |= bookas[k] << cshift;
        cshift += csubbits;
      }
      book_put(s, w, H.class_book[cls], cval);
    }
    for (int k = 0; k < cdim; k++) {
      const int book = H.class_subbook[cls * 8 + bookas[k]];
      if (book >= 0 && vals[j + k] < s->book[book].entries) book_put(s, w, book, vals[j + k]);
    }
    j += cdim;
  }
}

void vb200_emit_syms(const vb200_setup *s, BitWriter &w, const int *syms, int nsyms) {
  for (int i = 0; i < nsyms; i++) book_put(s, w, syms[i * 2], syms[i * 2 + 1]);
}
