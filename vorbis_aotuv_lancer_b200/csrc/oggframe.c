/* Ogg page framing + CRC for the packets the batch returns -- SURVEY.md section 8(f4).
 *
 * The reference leaves this to libogg (external, unpinned: configure.ac:247), called from the encoder's outer
 * loop (examples/encoder_example.c:141-176 headers + ogg_stream_flush, :224-235 packetin/pageout).  libogg is not
 * in this image, so this is a restatement of the published format (RFC 3533) with libogg 1.3's paging policy:
 *   - a packet is laid out as bytes/255 lacing values of 255 followed by one value < 255;
 *   - the first page carries only the first packet (the identification header) and the BOS flag;
 *   - afterwards a page is closed when it holds more than 4096 body bytes and at least four packets have been
 *     completed on it, when its 255 lacing values are used up, on an explicit flush, or at end of stream;
 *   - granule position of a page = that of the last packet completed on it (-1 if none), flag 0x01 when the page
 *     starts inside a packet, 0x04 on the last page;
 *   - CRC-32, polynomial 0x04c11db7, initial value 0, no reflection, over the page with the CRC field zero.
 * Host-only, one pass per stream; parity is unpinned (no libogg to compare with): tests/test_oggframe.py checks
 * the stream with an independent demultiplexer and bitwise CRC.
 */
#include <stdlib.h>
#include <string.h>
#include "vorbis_b200.h"

extern void vb200_set_error(const char *fmt, ...);

static uint32_t g_crc[256];
static int g_crc_ready = 0;
static void crc_init(void) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t r = i << 24;
    for (int k = 0; k < 8; k++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04c11db7u : (r << 1);
    g_crc[i] = r;
  }
  g_crc_ready = 1;
}
static uint32_t crc_update(uint32_t crc, const unsigned char *p, size_t n) {
  if (!g_crc_ready) crc_init();
  for (size_t i = 0; i < n; i++) crc = (crc << 8) ^ g_crc[((crc >> 24) & 0xff) ^ p[i]];
  return crc;
}
uint32_t vb200_ogg_crc(const unsigned char *data, size_t bytes) { return crc_update(0, data, bytes); }

struct vb200_oggmux {
  int32_t serial; uint32_t pageno; int bos_done, eos;
  /* pending segments */
  unsigned char *body; size_t body_fill, body_cap, body_returned;
  int *lacing; int64_t *granule; size_t lacing_fill, lacing_cap, granule_cap;      /* lacing | 0x100 marks a packet start */
  /* finished pages */
  unsigned char *out; size_t out_fill, out_cap;
};

int vb200_oggmux_create(int32_t serialno, vb200_oggmux **out) {
  if (!out) { vb200_set_error("vb200_oggmux_create: bad arguments"); return VB200_EINVAL; }
  vb200_oggmux *m = (vb200_oggmux *)calloc(1, sizeof(*m));
  if (!m) return VB200_EFAULT;
  m->serial = serialno;
  *out = m;
  return VB200_OK;
}
void vb200_oggmux_destroy(vb200_oggmux *m) {
  if (!m) return;
  free(m->body); free(m->lacing); free(m->granule); free(m->out); free(m);
}

static int grow(void **p, size_t *cap, size_t need, size_t elem) {
  if (need <= *cap) return 1;
  size_t nc = need * 2 + 1024;
  void *q = realloc(*p, nc * elem);
  if (!q) return 0;
  *p = q; *cap = nc;
  return 1;
}

/* close one page if the policy (or `force`) says so; returns 1 when a page was written */
static int page_out(vb200_oggmux *m, int force) {
  const size_t maxvals = m->lacing_fill > 255 ? 255 : m->lacing_fill;
  size_t vals = 0, bytes = 0, acc = 0;
  int64_t gp = -1;
  if (!maxvals) return 0;
  if (!m->bos_done) {                       /* the first page holds the first packet and nothing else */
    gp = 0;
    for (vals = 0; vals < maxvals; vals++) if ((m->lacing[vals] & 0xff) < 255) { vals++; break; }
    force = 1;
  } else {
    int done = 0, just_done = 0;
    for (vals = 0; vals < maxvals; vals++) {
      if (acc > 4096 && just_done >= 4) { force = 1; break; }
      acc += (size_t)(m->lacing[vals] & 0xff);
      if ((m->lacing[vals] & 0xff) < 255) { gp = m->granule[vals]; just_done = ++done; }
      else just_done = 0;
    }
    if (vals == 255) force = 1;
  }
  if (!force) return 0;
  for (size_t i = 0; i < vals; i++) bytes += (size_t)(m->lacing[i] & 0xff);
  const size_t hlen = 27 + vals;
  {
    void *p = m->out;
    if (!grow(&p, &m->out_cap, m->out_fill + hlen + bytes, 1)) return -1;
    m->out = (unsigned char *)p;
  }
  unsigned char *h = m->out + m->out_fill;
  memcpy(h, "OggS", 4);
  h[4] = 0;
  h[5] = 0;
  if (!(m->lacing[0] & 0x100)) h[5] |= 0x01;                                   /* continued packet */
  if (!m->bos_done) h[5] |= 0x02;                                              /* first page */
  if (m->eos && m->lacing_fill == vals) h[5] |= 0x04;                          /* last page */
  for (int i = 0; i < 8; i++) h[6 + i] = (unsigned char)(((uint64_t)gp >> (8 * i)) & 0xff);
  for (int i = 0; i < 4; i++) h[14 + i] = (unsigned char)(((uint32_t)m->serial >> (8 * i)) & 0xff);
  for (int i = 0; i < 4; i++) h[18 + i] = (unsigned char)((m->pageno >> (8 * i)) & 0xff);
  h[22] = h[23] = h[24] = h[25] = 0;
  h[26] = (unsigned char)vals;
  for (size_t i = 0; i < vals; i++) h[27 + i] = (unsigned char)(m->lacing[i] & 0xff);
  memcpy(h + hlen, m->body + m->body_returned, bytes);
  const uint32_t crc = crc_update(0, h, hlen + bytes);
  for (int i = 0; i < 4; i++) h[22 + i] = (unsigned char)((crc >> (8 * i)) & 0xff);
  m->out_fill += hlen + bytes;
  m->pageno++;
  m->bos_done = 1;
  /* drop what went out */
  m->lacing_fill -= vals;
  memmove(m->lacing, m->lacing + vals, m->lacing_fill * sizeof(int));
  memmove(m->granule, m->granule + vals, m->lacing_fill * sizeof(int64_t));
  m->body_returned += bytes;
  if (m->body_returned == m->body_fill) { m->body_returned = 0; m->body_fill = 0; }
  return 1;
}

/* ogg_stream_packetin followed by the encoder loop's `while (ogg_stream_pageout(...))` */
int vb200_oggmux_packet(vb200_oggmux *m, const unsigned char *data, long bytes, int64_t granulepos, int eos) {
  if (!m || bytes < 0 || (bytes && !data)) { vb200_set_error("vb200_oggmux_packet: bad arguments"); return VB200_EINVAL; }
  if (m->eos) { vb200_set_error("vb200_oggmux_packet: the stream has ended"); return VB200_EINVAL; }
  const size_t nl = (size_t)bytes / 255 + 1;
  {
    void *p = m->body;
    if (!grow(&p, &m->body_cap, m->body_fill + (size_t)bytes, 1)) return VB200_EFAULT;
    m->body = (unsigned char *)p;
    p = m->lacing;
    if (!grow(&p, &m->lacing_cap, m->lacing_fill + nl, sizeof(int))) return VB200_EFAULT;
    m->lacing = (int *)p;
    p = m->granule;
    if (!grow(&p, &m->granule_cap, m->lacing_fill + nl, sizeof(int64_t))) return VB200_EFAULT;
    m->granule = (int64_t *)p;
  }
  if (bytes) memcpy(m->body + m->body_fill, data, (size_t)bytes);
  m->body_fill += (size_t)bytes;
  size_t i;
  for (i = 0; i + 1 < nl; i++) { m->lacing[m->lacing_fill + i] = 255; m->granule[m->lacing_fill + i] = -1; }
  m->lacing[m->lacing_fill + i] = (int)(bytes % 255);
  m->granule[m->lacing_fill + i] = granulepos;
  m->lacing[m->lacing_fill] |= 0x100;
  m->lacing_fill += nl;
  if (eos) m->eos = 1;
  for (;;) {
    const int force = (m->eos && m->lacing_fill) || (m->lacing_fill && !m->bos_done);
    const int r = page_out(m, force);
    if (r < 0) { vb200_set_error("vb200_oggmux_packet: out of memory"); return VB200_EFAULT; }
    if (!r) break;
  }
  return VB200_OK;
}

/* ogg_stream_flush until empty: the encoder does this after the three headers so that audio starts a page */
int vb200_oggmux_flush(vb200_oggmux *m) {
  if (!m) return VB200_EINVAL;
  for (;;) {
    const int r = page_out(m, 1);
    if (r < 0) { vb200_set_error("vb200_oggmux_flush: out of memory"); return VB200_EFAULT; }
    if (!r) break;
  }
  return VB200_OK;
}

const unsigned char *vb200_oggmux_data(const vb200_oggmux *m, size_t *bytes) {
  if (!m) return NULL;
  if (bytes) *bytes = m->out_fill;
  return m->out;
}
long vb200_oggmux_pages(const vb200_oggmux *m) { return m ? (long)m->pageno : VB200_EINVAL; }
