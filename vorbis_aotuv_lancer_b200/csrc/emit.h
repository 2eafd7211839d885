/* Host bit writer + packet emission helpers (see emit.cu). */
#ifndef VB200_EMIT_H
#define VB200_EMIT_H
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "vb200_internal.h"

struct BitWriter {
  unsigned char *buf = nullptr;
  size_t cap = 0;
  size_t bitpos = 0;
  void grow(size_t need);
  void put(uint32_t value, int bits);
  /* start a new packet at the next byte boundary; returns its byte offset */
  size_t begin_packet() { bitpos = (bitpos + 7) & ~(size_t)7; return bitpos >> 3; }
  size_t bytes_since(size_t start) const { return ((bitpos + 7) >> 3) - start; }
  void reset() { if (buf) memset(buf, 0, (bitpos + 7) / 8 + 8 < cap ? (bitpos + 7) / 8 + 8 : cap); bitpos = 0; }
  ~BitWriter() { free(buf); }
};

void vb200_emit_header(const vb200_setup *s, BitWriter &w, int W, int lW, int nW);
void vb200_emit_floor(const vb200_setup *s, BitWriter &w, int floor_index, int nonzero, const int *vals);
void vb200_emit_syms(const vb200_setup *s, BitWriter &w, const int *syms, int nsyms);
#endif
