/* TEST INFRASTRUCTURE -- the drop-in boundary, exercised for real.
 *
 * This file is the reference-side binding of INTEGRATION.md, compiled against the reference's own headers where
 * they lie and linked with the UNMODIFIED reference objects into oracle/_ref/libvorbis_ref_drop.so.  In that
 * flavour the reference's own vorbis_analysis() (lib/analysis.c:29-63) reaches, through the mapping table
 * (_mapping_P[0]->forward, lib/registry.c:41-44, lib/backends.h:121-128), not mapping0_forward
 * (lib/mapping0.c:738) but b200_forward() below, which queues the block for libvorbis_b200.so; the packets
 * come back through vorbis_bitrate_flushpacket (lib/bitrate.c:229), exactly as SURVEY.md section 8(b) describes.
 * Everything else -- vorbis_analysis_init, _buffer/_wrote, the envelope search and _blockout, the headers --
 * is the reference's own code.  tests/test_gpu_dropin.py drives it.
 *
 * Nothing here is product code; the product never links this. */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ogg/ogg.h>
#include "vorbis/codec.h"
#include "codec_internal.h"
#include "registry.h"
#include "vb200_blob.h"
#include "vorbis_b200.h"

extern const vorbis_func_mapping mapping0_exportbundle;
extern unsigned char *vb200_export_setup(vorbis_dsp_state *vd,const unsigned char *hdr[3],const long hdrlen[3],long *bytes_out);
extern int  __real_vorbis_bitrate_addblock(vorbis_block *vb);
extern int  __real_vorbis_bitrate_flushpacket(vorbis_dsp_state *vd,ogg_packet *op);
extern void __real_vorbis_dsp_clear(vorbis_dsp_state *v);

/* file-scope constant tables of psy.c / floor1.c that no look object reaches (static in the reference): taken
   from a shipped setup blob, where tools/make_setups.py put them (they do not depend on the configuration) */
static unsigned char *g_const=NULL; static long g_const_bytes=0;
static int g_batch_blocks=256;
static char g_err[256];

void ref_dropin_constants(const unsigned char *blob,long bytes){
  free(g_const); g_const=malloc(bytes); memcpy(g_const,blob,bytes); g_const_bytes=bytes;
}
void ref_dropin_batch_blocks(int n){ if(n>0) g_batch_blocks=n; }
const char *ref_dropin_error(void){ return g_err; }

static const vb200_blob_section *blob_find(const unsigned char *blob,long bytes,const char *name,const unsigned char **payload){
  const vb200_blob_header *H=(const vb200_blob_header*)blob;
  long pos=sizeof(*H); uint32_t i;
  for(i=0;i<H->nsections&&pos+(long)sizeof(vb200_blob_section)<=bytes;i++){
    const vb200_blob_section *s=(const vb200_blob_section*)(blob+pos);
    long nb=(long)s->count*vb200_dtype_size(s->dtype);
    if(!strncmp(s->name,name,sizeof(s->name))){ *payload=blob+pos+sizeof(*s); return s; }
    pos+=sizeof(*s)+((nb+7)&~7L);
  }
  return NULL;
}
static unsigned char *blob_append(unsigned char *blob,long *bytes,const vb200_blob_section *s,const unsigned char *payload){
  long nb=(long)s->count*vb200_dtype_size(s->dtype), padded=(nb+7)&~7L;
  vb200_blob_header H;
  blob=realloc(blob,*bytes+sizeof(*s)+padded);
  memcpy(blob+*bytes,s,sizeof(*s));
  memset(blob+*bytes+sizeof(*s),0,padded);
  memcpy(blob+*bytes+sizeof(*s),payload,nb);
  *bytes+=sizeof(*s)+padded;
  memcpy(&H,blob,sizeof(H)); H.nsections++; H.total_bytes=(uint32_t)*bytes; memcpy(blob,&H,sizeof(H));
  return blob;
}

typedef struct {
  vorbis_dsp_state *vd;
  vb200_setup *gs; vb200_stream *st; vb200_batch *bt;
  vb200_block *q; int nq;
  float *pcm; size_t npcm,cappcm;
  int eof_seen, next_packet, npackets;
  ogg_int64_t packetno;
} dropin;
#define MAXD 16
static dropin g_d[MAXD];

static dropin *find(vorbis_dsp_state *vd){ int i; for(i=0;i<MAXD;i++) if(g_d[i].vd==vd) return &g_d[i]; return NULL; }

static dropin *attach(vorbis_dsp_state *vd){
  static const char *names[]={"fromdB","aotuv.stereo_th","aotuv.stereo_thx","aotuv.freq_bfn128","aotuv.freq_bfn256","aotuv.stn_compand"};
  static const unsigned char nothing[1]={0};
  const unsigned char *hp[3]={nothing,nothing,nothing}; const long hn[3]={1,1,1};   /* the reference emits its own headers: placeholders */
  dropin *d=NULL; unsigned char *blob; long bytes=0; int i;
  for(i=0;i<MAXD;i++) if(!g_d[i].vd){ d=&g_d[i]; break; }
  if(!d){ snprintf(g_err,sizeof(g_err),"too many encoder states"); return NULL; }
  if(!g_const){ snprintf(g_err,sizeof(g_err),"ref_dropin_constants was not called"); return NULL; }
  memset(d,0,sizeof(*d));
  /* the tables vorbis_analysis_init left behind -> HBM, once */
  blob=vb200_export_setup(vd,hp,hn,&bytes);
  for(i=0;i<6;i++){
    const unsigned char *pl; const vb200_blob_section *s=blob_find(g_const,g_const_bytes,names[i],&pl);
    if(!s){ snprintf(g_err,sizeof(g_err),"constant table %s missing",names[i]); free(blob); return NULL; }
    blob=blob_append(blob,&bytes,s,pl);
  }
  if(vb200_setup_create(blob,bytes,0,&d->gs)||vb200_stream_create(d->gs,&d->st)||vb200_batch_create(d->gs,g_batch_blocks,&d->bt)){
    snprintf(g_err,sizeof(g_err),"libvorbis_b200: %s",vb200_last_error()); free(blob); return NULL;
  }
  free(blob);
  d->q=malloc(sizeof(vb200_block)*g_batch_blocks);
  d->vd=vd; d->packetno=3;                       /* the three header packets come first (block.c:400, info.c:638) */
  return d;
}

/* what stands in mapping0_exportbundle.forward's place.  The arena of a vorbis_block is
   recycled at the next blockout (block.c:615), so everything mapping0_forward reads is copied out here. */
static int b200_forward(vorbis_block *vb){
  vorbis_dsp_state *vd=vb->vd;
  vorbis_block_internal *vbi=vb->internal;
  dropin *d=find(vd);
  int ch=vd->vi->channels, n=vb->pcmend, c;
  vb200_block *b;
  if(!d&&!(d=attach(vd))) return OV_EFAULT;
  if(d->nq>=g_batch_blocks){ snprintf(g_err,sizeof(g_err),"queue full: flushpacket was not called"); return OV_EFAULT; }
  if(d->npcm+(size_t)n*ch>d->cappcm){ d->cappcm=(d->npcm+(size_t)n*ch)*2+65536; d->pcm=realloc(d->pcm,d->cappcm*sizeof(float)); }
  b=&d->q[d->nq++];
  memset(b,0,sizeof(*b));
  b->stream=0; b->W=(int)vb->W; b->lW=(int)vb->lW; b->nW=(int)vb->nW;
  b->blocktype=vbi->blocktype; b->eofflag=vb->eofflag; b->granulepos=vb->granulepos;
  b->pcm_offset=(int64_t)d->npcm; b->pcm_stride=n;
  for(c=0;c<ch;c++) memcpy(d->pcm+d->npcm+(size_t)c*n,vb->pcm[c],sizeof(float)*n);
  d->npcm+=(size_t)n*ch;
  if(vb->eofflag) d->eof_seen=1;
  return 0;
}

/* the bitrate manager runs inside the batch: nothing to add per block */
int __wrap_vorbis_bitrate_addblock(vorbis_block *vb){
  if(find(vb->vd)) return 0;
  return __real_vorbis_bitrate_addblock(vb);
}

int __wrap_vorbis_bitrate_flushpacket(vorbis_dsp_state *vd,ogg_packet *op){
  dropin *d=find(vd);
  const unsigned char *data; long bytes; int64_t gp; int eos,stream;
  if(!d) return __real_vorbis_bitrate_flushpacket(vd,op);
  if(d->nq&&(d->nq>=g_batch_blocks||d->eof_seen)&&d->next_packet>=d->npackets){
    int r=vb200_batch_run(d->bt,&d->st,1,d->q,d->nq,d->pcm,d->npcm,0);
    if(r){ snprintf(g_err,sizeof(g_err),"libvorbis_b200: %s",vb200_last_error()); return 0; }
    d->npackets=d->nq; d->next_packet=0; d->nq=0; d->npcm=0;
  }
  if(d->next_packet>=d->npackets) return 0;
  if(vb200_batch_packet(d->bt,d->next_packet++,&data,&bytes,&gp,&eos,&stream)) return 0;
  if(op){
    op->packet=(unsigned char*)data; op->bytes=bytes; op->b_o_s=0; op->e_o_s=eos;
    op->granulepos=gp; op->packetno=d->packetno++;          /* memory owned by the batch until its next run */
  }
  return 1;
}

void __wrap_vorbis_dsp_clear(vorbis_dsp_state *v){
  dropin *d=find(v);
  if(d){
    vb200_batch_destroy(d->bt); vb200_stream_destroy(d->st); vb200_setup_destroy(d->gs);
    free(d->q); free(d->pcm); memset(d,0,sizeof(*d));
  }
  __real_vorbis_dsp_clear(v);
}

/* the mapping table entry ref_taps.c's replacement registry points at in this flavour */
vorbis_func_mapping dropin_mapping_bundle;
__attribute__((constructor)) static void dropin_init(void){
  dropin_mapping_bundle=mapping0_exportbundle;
  dropin_mapping_bundle.forward=&b200_forward;
}
