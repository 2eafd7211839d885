/* TEST INFRASTRUCTURE -- never linked into the product.
 *
 * Thin driver around the UNMODIFIED reference (compiled from
 * /root/reference/lib/*.c into oracle/_ref/): runs the canonical caller loop
 * of examples/encoder_example.c:180-238 on planar float PCM, collects the raw
 * packets, and decodes raw packets with the reference's own decoder following
 * the pattern of test/write_read.c:132-296.  Used by tests/ and by
 * bench.py's cpu_baseline / --impl reference legs only.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ogg/ogg.h>
#include "vorbis/codec.h"
#include "vorbis/vorbisenc.h"
#include "codec_internal.h"
#include "ref_taps.h"

typedef struct {
  vorbis_info vi;
  vorbis_comment vc;
  vorbis_dsp_state vd;
  vorbis_block vb;
  int open;
  /* collected packets */
  unsigned char *pk; long pklen, pkcap;
  long *pkoff; long *pkbytes; long long *pkgp; int *pkeos; long npk, cappk;
  /* headers */
  unsigned char *hdr[3]; long hdrlen[3];
  /* block log */
  int *blk; long nblk, capblk; /* W lW nW blocktype per block */
} ref_enc;

static void keep_packet(ref_enc *e, ogg_packet *op){
  if(e->npk>=e->cappk){
    e->cappk=e->cappk*2+1024;
    e->pkoff=realloc(e->pkoff,e->cappk*sizeof(long));
    e->pkbytes=realloc(e->pkbytes,e->cappk*sizeof(long));
    e->pkgp=realloc(e->pkgp,e->cappk*sizeof(long long));
    e->pkeos=realloc(e->pkeos,e->cappk*sizeof(int));
  }
  if(e->pklen+op->bytes>e->pkcap){
    e->pkcap=(e->pklen+op->bytes)*2+65536;
    e->pk=realloc(e->pk,e->pkcap);
  }
  memcpy(e->pk+e->pklen,op->packet,op->bytes);
  e->pkoff[e->npk]=e->pklen;
  e->pkbytes[e->npk]=op->bytes;
  e->pkgp[e->npk]=op->granulepos;
  e->pkeos[e->npk]=(int)op->e_o_s;
  e->pklen+=op->bytes;
  e->npk++;
}

/* managed_bitrate<=0: vorbis_encode_init_vbr(quality); else vorbis_encode_init(-1,nominal,-1) */
void *ref_enc_open(int channels,long rate,float quality,long managed_bitrate){
  ref_enc *e=calloc(1,sizeof(*e));
  int ret,i;
  ogg_packet h[3];
  vorbis_info_init(&e->vi);
  if(managed_bitrate>0) ret=vorbis_encode_init(&e->vi,channels,rate,-1,managed_bitrate,-1);
  else ret=vorbis_encode_init_vbr(&e->vi,channels,rate,quality);
  if(ret){ vorbis_info_clear(&e->vi); free(e); return NULL; }
  vorbis_comment_init(&e->vc);
  vorbis_comment_add_tag(&e->vc,"ENCODER","vb200-oracle");
  vorbis_analysis_init(&e->vd,&e->vi);
  vorbis_block_init(&e->vd,&e->vb);
  vorbis_analysis_headerout(&e->vd,&e->vc,&h[0],&h[1],&h[2]);
  for(i=0;i<3;i++){
    e->hdr[i]=malloc(h[i].bytes);
    memcpy(e->hdr[i],h[i].packet,h[i].bytes);
    e->hdrlen[i]=h[i].bytes;
  }
  e->open=1;
  return e;
}

/* The block-size detector's transform as the reference's own state runs it (envelope.c:139-186, scalar path): the
   handle's envelope_lookup supplies the window and the 128-point mdct_lookup.  data: 128 samples, out: 64 coefficients.
   (Any size n the reference can initialise: ref_mdct_forward below.) */
int ref_envelope_mdct(void *h,const float *data,float *out){
  ref_enc *e=h;
  private_state *b=e->vd.backend_state;
  envelope_lookup *ve=b->ve;
  float vec[1024];
  int i,n=ve->winlength;
  if(n>1024)return -1;
  for(i=0;i<n;i++)vec[i]=data[i]*ve->mdct_win[i];
  mdct_forward(&ve->mdct,vec,vec);
  memcpy(out,vec,sizeof(float)*(n/2));
  return n;
}
void ref_mdct_forward(int n,const float *in,float *out){
  mdct_lookup m;
  float *buf=malloc(sizeof(float)*n);
  memcpy(buf,in,sizeof(float)*n);
  mdct_init(&m,n);
  mdct_forward(&m,buf,buf);
  memcpy(out,buf,sizeof(float)*(n/2));
  mdct_clear(&m);
  free(buf);
}

void ref_enc_close(void *h){
  ref_enc *e=h; int i;
  if(!e)return;
  vorbis_block_clear(&e->vb);
  vorbis_dsp_clear(&e->vd);
  vorbis_comment_clear(&e->vc);
  vorbis_info_clear(&e->vi);
  for(i=0;i<3;i++)free(e->hdr[i]);
  free(e->pk); free(e->pkoff); free(e->pkbytes); free(e->pkgp); free(e->pkeos); free(e->blk);
  free(e);
}

const unsigned char *ref_enc_header(void *h,int i,long *bytes){
  ref_enc *e=h; *bytes=e->hdrlen[i]; return e->hdr[i];
}

vorbis_info *ref_enc_info(void *h){ return &((ref_enc*)h)->vi; }
vorbis_dsp_state *ref_enc_dsp(void *h){ return &((ref_enc*)h)->vd; }

static void drain(ref_enc *e){
  ogg_packet op;
  while(vorbis_analysis_blockout(&e->vd,&e->vb)==1){
    vorbis_block_internal *vbi=e->vb.internal;
    if(e->nblk>=e->capblk){ e->capblk=e->capblk*2+1024; e->blk=realloc(e->blk,e->capblk*4*sizeof(int)); }
    e->blk[e->nblk*4+0]=(int)e->vb.W; e->blk[e->nblk*4+1]=(int)e->vb.lW;
    e->blk[e->nblk*4+2]=(int)e->vb.nW; e->blk[e->nblk*4+3]=vbi->blocktype;
    e->nblk++;
    vorbis_analysis(&e->vb,NULL);
    vorbis_bitrate_addblock(&e->vb);
    while(vorbis_bitrate_flushpacket(&e->vd,&op)){
      ref_tap_arm(&e->vb,1);
      if(ref_tap_on(TAP_PACKET))
        ref_tap_put(TAP_PACKET,(int)op.bytes,(int)op.e_o_s,(int)(op.granulepos&0xffffffff),(int)(op.granulepos>>32),2,op.packet,op.bytes);
      ref_tap_arm(&e->vb,0);
      keep_packet(e,&op);
    }
  }
}

/* pcm: planar [channels][nsamples]; chunk = samples per vorbis_analysis_wrote
   (encoder_example.c uses 1024); eos!=0 also signals end of stream. Returns packets so far. */
long ref_enc_write(void *h,const float *pcm,long nsamples,long stride,int chunk,int eos){
  ref_enc *e=h;
  long pos=0; int c;
  if(chunk<=0)chunk=1024;
  while(pos<nsamples){
    long m=nsamples-pos<chunk?nsamples-pos:chunk;
    float **buf=vorbis_analysis_buffer(&e->vd,(int)m);
    for(c=0;c<e->vi.channels;c++)
      memcpy(buf[c],pcm+c*stride+pos,m*sizeof(float));
    vorbis_analysis_wrote(&e->vd,(int)m);
    pos+=m;
    drain(e);
  }
  if(eos){
    vorbis_analysis_wrote(&e->vd,0);
    drain(e);
  }
  return e->npk;
}

long ref_enc_npackets(void *h){ return ((ref_enc*)h)->npk; }
long ref_enc_nblocks(void *h){ return ((ref_enc*)h)->nblk; }
const int *ref_enc_blocks(void *h){ return ((ref_enc*)h)->blk; }
const unsigned char *ref_enc_packets(void *h,const long **off,const long **bytes,const long long **gp,const int **eos,long *total){
  ref_enc *e=h;
  *off=e->pkoff; *bytes=e->pkbytes; *gp=e->pkgp; *eos=e->pkeos; *total=e->pklen;
  return e->pk;
}

/* ------------------------------------------------------------------ decode */
/* Decode raw packets (3 headers + audio packets) with the reference decoder.
   out: planar [channels][cap]; returns samples decoded per channel or <0. */
long ref_decode(const unsigned char *h0,long n0,const unsigned char *h1,long n1,const unsigned char *h2,long n2,
                const unsigned char *pk,const long *off,const long *bytes,const long long *gp,long npk,
                float *out,long cap,int *channels_out,long *rate_out){
  vorbis_info vi; vorbis_comment vc; vorbis_dsp_state vd; vorbis_block vb;
  ogg_packet op; long i,total=0; int c;
  const unsigned char *hp[3]; long hn[3];
  hp[0]=h0;hp[1]=h1;hp[2]=h2; hn[0]=n0;hn[1]=n1;hn[2]=n2;
  vorbis_info_init(&vi); vorbis_comment_init(&vc);
  for(i=0;i<3;i++){
    memset(&op,0,sizeof(op));
    op.packet=(unsigned char*)hp[i]; op.bytes=hn[i]; op.b_o_s=(i==0); op.packetno=i;
    if(vorbis_synthesis_headerin(&vi,&vc,&op)<0){ vorbis_comment_clear(&vc); vorbis_info_clear(&vi); return -1; }
  }
  if(vorbis_synthesis_init(&vd,&vi)){ vorbis_comment_clear(&vc); vorbis_info_clear(&vi); return -2; }
  vorbis_block_init(&vd,&vb);
  *channels_out=vi.channels; *rate_out=vi.rate;
  for(i=0;i<npk;i++){
    float **pcm; int samples;
    memset(&op,0,sizeof(op));
    op.packet=(unsigned char*)pk+off[i]; op.bytes=bytes[i]; op.packetno=3+i;
    op.granulepos=gp?gp[i]:-1; op.e_o_s=(i==npk-1);
    if(vorbis_synthesis(&vb,&op)==0) vorbis_synthesis_blockin(&vd,&vb);
    else { total=-3-i; break; }
    while((samples=vorbis_synthesis_pcmout(&vd,&pcm))>0){
      int take=samples;
      if(total+take>cap)take=(int)(cap-total);
      for(c=0;c<vi.channels;c++) memcpy(out+c*cap+total,pcm[c],take*sizeof(float));
      total+=take;
      vorbis_synthesis_read(&vd,samples);
      if(take<samples)break;
    }
  }
  vorbis_block_clear(&vb); vorbis_dsp_clear(&vd); vorbis_comment_clear(&vc); vorbis_info_clear(&vi);
  return total;
}
