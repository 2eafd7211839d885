/* TEST INFRASTRUCTURE (oracle build only) -- minimal stand-in for libogg's
 * <ogg/ogg.h>: the bit-packer struct + the packet struct + the nine
 * oggpack_* entry points the reference's encode/decode core links against
 * (SURVEY.md section 8c).  Ogg page framing is not needed: the parity harness
 * exchanges raw packets. */
#ifndef VB200_SHIM_OGG_H
#define VB200_SHIM_OGG_H
#include <stddef.h>
#include <ogg/os_types.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  long endbyte;
  int  endbit;
  unsigned char *buffer;
  unsigned char *ptr;
  long storage;
} oggpack_buffer;

typedef struct {
  unsigned char *packet;
  long  bytes;
  long  b_o_s;
  long  e_o_s;
  ogg_int64_t granulepos;
  ogg_int64_t packetno;
} ogg_packet;

extern void  oggpack_writeinit(oggpack_buffer *b);
extern void  oggpack_writetrunc(oggpack_buffer *b,long bits);
extern void  oggpack_writeclear(oggpack_buffer *b);
extern void  oggpack_reset(oggpack_buffer *b);
extern void  oggpack_write(oggpack_buffer *b,unsigned long value,int bits);
extern void  oggpack_readinit(oggpack_buffer *b,unsigned char *buf,int bytes);
extern long  oggpack_read(oggpack_buffer *b,int bits);
extern long  oggpack_look(oggpack_buffer *b,int bits);
extern void  oggpack_adv(oggpack_buffer *b,int bits);
extern long  oggpack_bytes(oggpack_buffer *b);
extern long  oggpack_bits(oggpack_buffer *b);
extern unsigned char *oggpack_get_buffer(oggpack_buffer *b);
#ifdef __cplusplus
}
#endif
#endif
