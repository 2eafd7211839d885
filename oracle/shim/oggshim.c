/* TEST INFRASTRUCTURE (oracle build only).
 * LSb-first bit packer with the documented semantics of libogg's oggpack_*
 * API, so that the reference (which needs libogg only for bit packing,
 * SURVEY.md section 8c) links in this container.  Keeps >= 8 spare bytes past the
 * write pointer because the reference's private writer (lib/codebook.c:80-110)
 * stores 32 bits at ptr[0] and one byte at ptr[4]. */
#include <string.h>
#include <stdlib.h>
#include <ogg/ogg.h>

#define GROW 256
#define SPARE 8

static unsigned long lowmask(int bits){
  return bits>=32 ? 0xffffffffUL : ((1UL<<bits)-1UL);
}

void oggpack_writeinit(oggpack_buffer *b){
  memset(b,0,sizeof(*b));
  b->buffer=calloc(GROW+SPARE,1);
  b->ptr=b->buffer;
  b->storage=GROW;
}

void oggpack_reset(oggpack_buffer *b){
  if(!b->buffer)return;
  b->ptr=b->buffer;
  b->buffer[0]=0;
  b->endbit=0;
  b->endbyte=0;
}

void oggpack_writeclear(oggpack_buffer *b){
  if(b->buffer)free(b->buffer);
  memset(b,0,sizeof(*b));
}

void oggpack_writetrunc(oggpack_buffer *b,long bits){
  long bytes=bits>>3;
  if(!b->buffer)return;
  bits-=bytes*8;
  b->ptr=b->buffer+bytes;
  b->endbit=(int)bits;
  b->endbyte=bytes;
  *b->ptr&=(unsigned char)lowmask((int)bits);
}

void oggpack_write(oggpack_buffer *b,unsigned long value,int bits){
  int k,total;
  if(bits<0||bits>32||!b->buffer)return;
  if(b->endbyte>=b->storage-4){
    b->buffer=realloc(b->buffer,b->storage+GROW+SPARE);
    b->storage+=GROW;
    b->ptr=b->buffer+b->endbyte;
  }
  value&=lowmask(bits);
  total=bits+b->endbit;
  /* byte 0 keeps its low endbit bits; following bytes are assigned */
  b->ptr[0]=(unsigned char)((b->ptr[0]&lowmask(b->endbit))|((value<<b->endbit)&0xff));
  for(k=1;k*8<total+8&&k<=4;k++){
    int sh=k*8-b->endbit;
    b->ptr[k]=(unsigned char)(sh<64?((unsigned long long)value>>sh)&0xff:0);
  }
  b->endbyte+=total/8;
  b->ptr+=total/8;
  b->endbit=total&7;
}

void oggpack_readinit(oggpack_buffer *b,unsigned char *buf,int bytes){
  memset(b,0,sizeof(*b));
  b->buffer=b->ptr=buf;
  b->storage=bytes;
}

long oggpack_look(oggpack_buffer *b,int bits){
  unsigned long long acc=0;
  int total,k;
  if(bits<0||bits>32)return -1;
  total=bits+b->endbit;
  if(b->endbyte>=b->storage-4){
    if(b->endbyte>b->storage-((total+7)>>3))return -1;
    else if(!total)return 0L;
  }
  for(k=0;k*8<total;k++)
    acc|=(unsigned long long)b->ptr[k]<<(8*k);
  return (long)((acc>>b->endbit)&lowmask(bits));
}

void oggpack_adv(oggpack_buffer *b,int bits){
  bits+=b->endbit;
  if(b->endbyte>b->storage-((bits+7)>>3)){
    b->ptr=NULL; b->endbyte=b->storage; b->endbit=1; return;
  }
  b->ptr+=bits/8;
  b->endbyte+=bits/8;
  b->endbit=bits&7;
}

long oggpack_read(oggpack_buffer *b,int bits){
  long ret;
  int total;
  if(!b->ptr)return -1;
  if(bits<0||bits>32)goto fail;
  total=bits+b->endbit;
  if(b->endbyte>b->storage-((total+7)>>3))goto fail;
  ret=total?oggpack_look(b,bits):0L;
  oggpack_adv(b,bits);
  return ret;
fail:
  b->ptr=NULL; b->endbyte=b->storage; b->endbit=1;
  return -1;
}

long oggpack_bytes(oggpack_buffer *b){ return b->endbyte+(b->endbit+7)/8; }
long oggpack_bits(oggpack_buffer *b){ return b->endbyte*8+b->endbit; }
unsigned char *oggpack_get_buffer(oggpack_buffer *b){ return b->buffer; }
