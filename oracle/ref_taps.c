/* TEST INFRASTRUCTURE -- never linked into the product.
 * --wrap taps around the reference's stage functions; see ref_taps.h.
 * Compiled against the reference's own headers where they lie
 * (/root/reference/lib, /root/reference/include); outputs go to oracle/_ref/.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ogg/ogg.h>
#include "vorbis/codec.h"
#include "codec_internal.h"
#include "registry.h"
#include "psy.h"
#include "mdct.h"
#include "smallft.h"
#include "window.h"
#include "codebook.h"
#include "ref_taps.h"

/* ------------------------------------------------------------------ log */
static unsigned char *g_log=NULL;
static long g_len=0, g_cap=0;
static unsigned char g_enable[TAP_COUNT];
static long g_lo=0, g_hi=-1;
static vorbis_block *g_vb=NULL;
static int g_armed=0;

/* per-block symbol list (book,entry) */
static int *g_syms=NULL; static long g_nsyms=0, g_capsyms=0;
/* per-block call counters */
static int g_cnt_pcm, g_cnt_win, g_cnt_mdct, g_cnt_loud, g_cnt_nm, g_cnt_tm, g_cnt_mix, g_cnt_fit, g_cnt_fenc, g_cnt_res;
static int g_cur_blob;

void ref_tap_config(const unsigned char *enable,long lo,long hi){
  memcpy(g_enable,enable,TAP_COUNT);
  g_lo=lo; g_hi=hi;
}
void ref_tap_reset(void){ g_len=0; }
const unsigned char *ref_tap_data(long *bytes){ *bytes=g_len; return g_log; }

static long cur_seq(void){ return g_vb?(long)g_vb->sequence:-1; }

int ref_tap_on(int tag){
  long s;
  if(!g_armed||!g_enable[tag])return 0;
  s=cur_seq();
  return s>=g_lo && s<g_hi;
}

void ref_tap_put(int tag,int a0,int a1,int a2,int a3,int dtype,const void *p,long count){
  long payload=(dtype==2)?((count+3)&~3L):count*4;
  long need=g_len+32+payload;
  int hdr[8];
  if(need>g_cap){
    g_cap=need*2+(1<<20);
    g_log=realloc(g_log,g_cap);
  }
  hdr[0]=tag; hdr[1]=(int)cur_seq(); hdr[2]=a0; hdr[3]=a1; hdr[4]=a2; hdr[5]=dtype; hdr[6]=(int)count; hdr[7]=a3;
  memcpy(g_log+g_len,hdr,32);
  if(payload){
    memset(g_log+g_len+32,0,payload);
    if(p)memcpy(g_log+g_len+32,p,dtype==2?count:count*4);
  }
  g_len+=32+payload;
}
#define PUTF(tag,a0,a1,a2,p,n) do{ if(ref_tap_on(tag)) ref_tap_put(tag,a0,a1,a2,0,0,p,n); }while(0)
#define PUTI(tag,a0,a1,a2,p,n) do{ if(ref_tap_on(tag)) ref_tap_put(tag,a0,a1,a2,0,1,p,n); }while(0)

void ref_tap_arm(void *vb,int armed){
  g_vb=(vorbis_block*)vb; g_armed=armed;
}

static private_state *PS(void){ return (private_state*)g_vb->vd->backend_state; }
static codec_setup_info *CI(void){ return (codec_setup_info*)g_vb->vd->vi->codec_setup; }
static int psy_index(const vorbis_look_psy *p){ return (int)(p-PS()->psy); }
static int floor_index(const void *look){
  int i; for(i=0;i<CI()->floors;i++) if(PS()->flr[i]==look) return i; return -1;
}
static int residue_index(const void *look){
  int i; for(i=0;i<CI()->residues;i++) if(PS()->residue[i]==look) return i; return -1;
}
static int psy_partition(const vorbis_look_psy *p){ return p->vi->normal_p ? p->vi->normal_partition : 16; }

/* ---------------------------------------------------------- vorbis_analysis */
int __real_vorbis_analysis(vorbis_block *vb, ogg_packet *op);
int __wrap_vorbis_analysis(vorbis_block *vb, ogg_packet *op){
  int ret,i;
  vorbis_block_internal *vbi=vb->internal;
  private_state *b=vb->vd->backend_state;
  float amp[2];
  ref_tap_arm(vb,1);
  g_nsyms=0;
  g_cnt_pcm=g_cnt_win=g_cnt_mdct=g_cnt_loud=g_cnt_nm=g_cnt_tm=g_cnt_mix=g_cnt_fit=g_cnt_fenc=g_cnt_res=0;
  g_cur_blob=vorbis_bitrate_managed(vb)?0:PACKETBLOBS/2;
  if(ref_tap_on(TAP_BLOCK)){
    int v[12];
    v[0]=vb->W; v[1]=vb->lW; v[2]=vb->nW; v[3]=vbi->blocktype; v[4]=vb->pcmend; v[5]=vb->eofflag;
    v[6]=(int)(vb->granulepos&0xffffffff); v[7]=(int)(vb->granulepos>>32);
    v[8]=b->lW_block_mode; v[9]=b->lW_no; v[10]=b->impadnum; v[11]=vorbis_bitrate_managed(vb);
    ref_tap_put(TAP_BLOCK,vb->vd->vi->channels,0,0,0,1,v,12);
  }
  amp[0]=vbi->ampmax;
  ret=__real_vorbis_analysis(vb,op);
  amp[1]=vbi->ampmax;
  PUTF(TAP_AMPMAX,0,0,0,amp,2);
  if(ref_tap_on(TAP_BLOCK_SYMS)) ref_tap_put(TAP_BLOCK_SYMS,0,0,0,0,1,g_syms,g_nsyms*2);
  if(ref_tap_on(TAP_BLOBBYTES)){
    int v[PACKETBLOBS];
    for(i=0;i<PACKETBLOBS;i++)v[i]=(int)oggpack_bytes(vbi->packetblob[i]);
    ref_tap_put(TAP_BLOBBYTES,0,0,0,0,1,v,PACKETBLOBS);
  }
  g_armed=0;
  return ret;
}

/* ---------------------------------------------------------- front end */
float __real__postnoise_detection(float *pcm,int nn,int mode,int lw_mode);
float __wrap__postnoise_detection(float *pcm,int nn,int mode,int lw_mode){
  float r;
  int ch=g_cnt_pcm++;
  PUTF(TAP_PCM,ch,mode,lw_mode,pcm,nn);
  r=__real__postnoise_detection(pcm,nn,mode,lw_mode);
  PUTF(TAP_POSTE,ch,0,0,&r,1);
  return r;
}

void __real__vorbis_apply_window(float *d,int *winno,long *blocksizes,int lW,int W,int nW);
void __wrap__vorbis_apply_window(float *d,int *winno,long *blocksizes,int lW,int W,int nW){
  __real__vorbis_apply_window(d,winno,blocksizes,lW,W,nW);
  if(g_armed){ int ch=g_cnt_win++; PUTF(TAP_WINDOWED,ch,0,0,d,blocksizes[W]); }
}

void __real_mdct_forward(mdct_lookup *init, DATA_TYPE *in, DATA_TYPE *out);
void __wrap_mdct_forward(mdct_lookup *init, DATA_TYPE *in, DATA_TYPE *out){
  __real_mdct_forward(init,in,out);
  if(g_armed){ int ch=g_cnt_mdct++; PUTF(TAP_MDCT,ch,0,0,out,init->n/2); }
}

float __real_lb_loudnoise_fix(const vorbis_look_psy *p,float level,const float *logmdct,const int block_mode,const int lW_block_mode);
float __wrap_lb_loudnoise_fix(const vorbis_look_psy *p,float level,const float *logmdct,const int block_mode,const int lW_block_mode){
  float v[2];
  int ch=g_cnt_loud++;
  v[0]=level;
  v[1]=__real_lb_loudnoise_fix(p,level,logmdct,block_mode,lW_block_mode);
  PUTF(TAP_LOUDFIX,ch,block_mode,lW_block_mode,v,2);
  return v[1];
}

/* ---------------------------------------------------------- psy */
void __real__vp_noisemask(const vorbis_look_psy *p,const float level,const float *logmdct,const float *lastmdct,
                          float *epeak,float *npeak,float *logmask,float poste,int block_mode);
void __wrap__vp_noisemask(const vorbis_look_psy *p,const float level,const float *logmdct,const float *lastmdct,
                          float *epeak,float *npeak,float *logmask,float poste,int block_mode){
  int ch=g_cnt_nm++, n=p->n, pi=psy_index(p);
  float sc[2];
  sc[0]=level; sc[1]=poste;
  PUTF(TAP_NM_LOGMDCT,ch,pi,block_mode,logmdct,n);
  PUTF(TAP_NM_LASTMDCT,ch,pi,block_mode,lastmdct,n);
  PUTF(TAP_NM_SCAL,ch,pi,block_mode,sc,2);
  __real__vp_noisemask(p,level,logmdct,lastmdct,epeak,npeak,logmask,poste,block_mode);
  PUTF(TAP_NM_EPEAK,ch,pi,block_mode,epeak,n);
  PUTF(TAP_NM_NPEAK,ch,pi,block_mode,npeak,n/psy_partition(p));
  PUTF(TAP_NM_NOISE,ch,pi,block_mode,logmask,n);
}

void __real__vp_tonemask(const vorbis_look_psy *p,const float *logfft,float *logmask,const float global_specmax,const float local_specmax);
void __wrap__vp_tonemask(const vorbis_look_psy *p,const float *logfft,float *logmask,const float global_specmax,const float local_specmax){
  int ch=g_cnt_tm++, n=p->n, pi=psy_index(p);
  float sc[2];
  sc[0]=global_specmax; sc[1]=local_specmax;
  PUTF(TAP_TM_LOGFFT,ch,pi,0,logfft,n);
  PUTF(TAP_TM_SCAL,ch,pi,0,sc,2);
  __real__vp_tonemask(p,logfft,logmask,global_specmax,local_specmax);
  PUTF(TAP_TM_TONE,ch,pi,0,logmask,n);
}

void __real__vp_offset_and_mix(const vorbis_look_psy *p,const float *noise,const float *tone,const int offset_select,
                               const int bit_managed,float *logmask,float *mdct,float *logmdct,float *lastmdct,float *tempmdct,
                               float low_compand,float *npeak,const int end_block,const int block_mode,const int nW_modenumber,
                               const int lW_block_mode,const int lW_no,const int impadnum);
void __wrap__vp_offset_and_mix(const vorbis_look_psy *p,const float *noise,const float *tone,const int offset_select,
                               const int bit_managed,float *logmask,float *mdct,float *logmdct,float *lastmdct,float *tempmdct,
                               float low_compand,float *npeak,const int end_block,const int block_mode,const int nW_modenumber,
                               const int lW_block_mode,const int lW_no,const int impadnum){
  int call=g_cnt_mix++, n=p->n, pi=psy_index(p), np=n/psy_partition(p);
  int ch=-1;
  int iv[8];
  /* channel = which mblock slot lastmdct points at */
  ch=(int)((lastmdct-PS()->mblock)/2048);
  (void)call;
  iv[0]=offset_select; iv[1]=bit_managed; iv[2]=end_block; iv[3]=block_mode; iv[4]=nW_modenumber;
  iv[5]=lW_block_mode; iv[6]=lW_no; iv[7]=impadnum;
  PUTI(TAP_MIX_I32,ch,pi,offset_select,iv,8);
  PUTF(TAP_MIX_F32,ch,pi,offset_select,&low_compand,1);
  PUTF(TAP_MIX_NOISE,ch,pi,offset_select,noise,n);
  PUTF(TAP_MIX_TONE,ch,pi,offset_select,tone,n);
  PUTF(TAP_MIX_MDCT_IN,ch,pi,offset_select,mdct,n);
  PUTF(TAP_MIX_LOGMDCT,ch,pi,offset_select,logmdct,n);
  PUTF(TAP_MIX_LAST_IN,ch,pi,offset_select,lastmdct,2048);
  PUTF(TAP_MIX_TEMP_IN,ch,pi,offset_select,tempmdct,256);
  PUTF(TAP_MIX_NPEAK_IN,ch,pi,offset_select,npeak,np);
  __real__vp_offset_and_mix(p,noise,tone,offset_select,bit_managed,logmask,mdct,logmdct,lastmdct,tempmdct,
                            low_compand,npeak,end_block,block_mode,nW_modenumber,lW_block_mode,lW_no,impadnum);
  PUTF(TAP_MIX_LOGMASK,ch,pi,offset_select,logmask,n);
  PUTF(TAP_MIX_MDCT_OUT,ch,pi,offset_select,mdct,n);
  PUTF(TAP_MIX_LAST_OUT,ch,pi,offset_select,lastmdct,2048);
  PUTF(TAP_MIX_TEMP_OUT,ch,pi,offset_select,tempmdct,256);
  PUTF(TAP_MIX_NPEAK_OUT,ch,pi,offset_select,npeak,np);
}

/* ---------------------------------------------------------- floor1 */
int *__real_floor1_fit(vorbis_block *vb,vorbis_look_floor1 *look,const float *logmdct,const float *logmask);
int *__wrap_floor1_fit(vorbis_block *vb,vorbis_look_floor1 *look,const float *logmdct,const float *logmask){
  int call=g_cnt_fit++, fi=floor_index(look);
  int n=vb->pcmend/2;
  int *r;
  PUTF(TAP_FIT_LOGMDCT,call,fi,0,logmdct,n);
  PUTF(TAP_FIT_LOGMASK,call,fi,0,logmask,n);
  r=__real_floor1_fit(vb,look,logmdct,logmask);
  PUTI(TAP_FIT_POSTS,call,fi,0,r,r?look->posts:0);
  return r;
}

int *__real_floor1_interpolate_fit(vorbis_block *vb,vorbis_look_floor1 *look,int *A,int *B,int del);
int *__wrap_floor1_interpolate_fit(vorbis_block *vb,vorbis_look_floor1 *look,int *A,int *B,int del){
  int *r=__real_floor1_interpolate_fit(vb,look,A,B,del);
  if(ref_tap_on(TAP_INTERP) && r){
    int tmp[3*(VIF_POSIT+2)], posts=look->posts;
    memcpy(tmp,A,posts*sizeof(int));
    memcpy(tmp+posts,B,posts*sizeof(int));
    memcpy(tmp+2*posts,r,posts*sizeof(int));
    ref_tap_put(TAP_INTERP,floor_index(look),del,posts,0,1,tmp,3*posts);
  }
  return r;
}

int __real_floor1_encode(oggpack_buffer *opb,vorbis_block *vb,vorbis_look_floor1 *look,int *post,int *ilogmask);
int __wrap_floor1_encode(oggpack_buffer *opb,vorbis_block *vb,vorbis_look_floor1 *look,int *post,int *ilogmask){
  int call=g_cnt_fenc++, fi=floor_index(look), ret;
  int chs=vb->vd->vi->channels;
  int ch=call%chs, blob;
  vorbis_block_internal *vbi=vb->internal;
  for(blob=0;blob<PACKETBLOBS;blob++) if(vbi->packetblob[blob]==opb)break;
  g_cur_blob=blob;
  PUTI(TAP_FENC_IN,ch,fi,blob,post,post?look->posts:0);
  ret=__real_floor1_encode(opb,vb,look,post,ilogmask);
  PUTI(TAP_FENC_POST,ch,fi,blob,post,post?look->posts:0);
  PUTI(TAP_FENC_ILOG,ch,fi,blob,ilogmask,vb->pcmend/2);
  PUTI(TAP_FENC_RET,ch,fi,blob,&ret,1);
  return ret;
}

/* ---------------------------------------------------------- coupling */
void __real__vp_couple_quantize_normalize(int blobno,vorbis_info_psy_global *g,vorbis_look_psy *p,vorbis_info_mapping0 *vi,
                                          float **mdct,float **enpeak,float **nepeak,int **iwork,int *nonzero,
                                          int sliding_lowpass,int ch,int lowpassr);
static void put_rows_f(int tag,int a0,float **rows,int ch,int n){
  if(ref_tap_on(tag)){
    float *t=malloc(sizeof(float)*ch*n); int i;
    for(i=0;i<ch;i++)memcpy(t+i*n,rows[i],n*sizeof(float));
    ref_tap_put(tag,a0,ch,n,0,0,t,(long)ch*n);
    free(t);
  }
}
static void put_rows_i(int tag,int a0,int a1,int **rows,int ch,int n){
  if(ref_tap_on(tag)){
    int *t=malloc(sizeof(int)*ch*n); int i;
    for(i=0;i<ch;i++)memcpy(t+i*n,rows[i],n*sizeof(int));
    ref_tap_put(tag,a0,a1,n,0,1,t,(long)ch*n);
    free(t);
  }
}
void __wrap__vp_couple_quantize_normalize(int blobno,vorbis_info_psy_global *g,vorbis_look_psy *p,vorbis_info_mapping0 *vi,
                                          float **mdct,float **enpeak,float **nepeak,int **iwork,int *nonzero,
                                          int sliding_lowpass,int ch,int lowpassr){
  int n=p->n, np=n/psy_partition(p);
  int iv[5];
  iv[0]=blobno; iv[1]=sliding_lowpass; iv[2]=ch; iv[3]=lowpassr; iv[4]=psy_index(p);
  PUTI(TAP_CPL_I32,blobno,0,0,iv,5);
  put_rows_f(TAP_CPL_MDCT,blobno,mdct,ch,n);
  put_rows_f(TAP_CPL_EPEAK,blobno,enpeak,ch,n);
  put_rows_f(TAP_CPL_NPEAK_IN,blobno,nepeak,ch,np);
  put_rows_i(TAP_CPL_IWORK_IN,blobno,ch,iwork,ch,n);
  PUTI(TAP_CPL_NZ_IN,blobno,0,0,nonzero,ch);
  __real__vp_couple_quantize_normalize(blobno,g,p,vi,mdct,enpeak,nepeak,iwork,nonzero,sliding_lowpass,ch,lowpassr);
  put_rows_i(TAP_CPL_IWORK_OUT,blobno,ch,iwork,ch,n);
  PUTI(TAP_CPL_NZ_OUT,blobno,0,0,nonzero,ch);
  put_rows_f(TAP_CPL_NPEAK_OUT,blobno,nepeak,ch,np);
}

/* ---------------------------------------------------------- book encode */
int __real_vorbis_book_encode(codebook *book,int a,oggpack_buffer *b);
int __wrap_vorbis_book_encode(codebook *book,int a,oggpack_buffer *b){
  if(g_armed && g_vb){
    if(g_nsyms>=g_capsyms){ g_capsyms=g_capsyms*2+4096; g_syms=realloc(g_syms,g_capsyms*2*sizeof(int)); }
    g_syms[g_nsyms*2]=(int)(book-CI()->fullbooks);
    g_syms[g_nsyms*2+1]=a;
    g_nsyms++;
  }
  return __real_vorbis_book_encode(book,a,b);
}

/* ---------------------------------------------------------- residue (via the replacement registry) */
typedef struct { vorbis_info_residue0 *info; } look_residue_head; /* first member of res0.c's private look struct (res0.c:37-39) */

extern const vorbis_func_floor     floor0_exportbundle;
extern const vorbis_func_floor     floor1_exportbundle;
extern const vorbis_func_residue   residue0_exportbundle;
extern const vorbis_func_residue   residue1_exportbundle;
extern const vorbis_func_residue   residue2_exportbundle;
extern const vorbis_func_mapping   mapping0_exportbundle;

static int g_res_type_of_call=0;
static long **tap_class(int type,struct vorbis_block *vb,vorbis_look_residue *vl,int **in,int *nonzero,int ch){
  const vorbis_func_residue *real=(type==1?&residue1_exportbundle:&residue2_exportbundle);
  vorbis_info_residue0 *info=((look_residue_head*)vl)->info;
  int n=vb->pcmend/2, i, used=0, call=g_cnt_res;
  int partvals=(info->end-info->begin)/info->grouping;
  long **pw;
  int iv[5];
  for(i=0;i<ch;i++)if(nonzero[i])used++;
  iv[0]=residue_index(vl); iv[1]=type; iv[2]=ch; iv[3]=used; iv[4]=call;
  PUTI(TAP_RES_I32,g_cur_blob,call,0,iv,5);
  put_rows_i(TAP_RES_IN,g_cur_blob,call,in,ch,n);
  PUTI(TAP_RES_NZ,g_cur_blob,call,0,nonzero,ch);
  pw=real->class(vb,vl,in,nonzero,ch);
  if(ref_tap_on(TAP_RES_PARTWORD)){
    int rows=pw?(type==2?1:used):0;
    int *t=malloc(sizeof(int)*(rows*partvals+1)); int r,k;
    for(r=0;r<rows;r++)for(k=0;k<partvals;k++)t[r*partvals+k]=(int)pw[r][k];
    ref_tap_put(TAP_RES_PARTWORD,g_cur_blob,call,rows,0,1,t,(long)rows*partvals);
    free(t);
  }
  g_res_type_of_call=type;
  return pw;
}
static int tap_forward(int type,oggpack_buffer *opb,struct vorbis_block *vb,vorbis_look_residue *vl,
                       int **in,int *nonzero,int ch,long **partword,int submap){
  const vorbis_func_residue *real=(type==1?&residue1_exportbundle:&residue2_exportbundle);
  long s0=g_nsyms;
  int call=g_cnt_res++, ret;
  int n=vb->pcmend/2;
  int *keep[256]; int i;
  for(i=0;i<ch&&i<256;i++)keep[i]=in[i];  /* res1_forward compacts in[] */
  ret=real->forward(opb,vb,vl,in,nonzero,ch,partword,submap);
  if(ref_tap_on(TAP_RES_SYMS)) ref_tap_put(TAP_RES_SYMS,g_cur_blob,call,0,0,1,g_syms+s0*2,(g_nsyms-s0)*2);
  put_rows_i(TAP_RES_OUT,g_cur_blob,call,keep,ch,n);
  return ret;
}
static long **c1(struct vorbis_block *vb,vorbis_look_residue *vl,int **in,int *nz,int ch){ return tap_class(1,vb,vl,in,nz,ch); }
static long **c2(struct vorbis_block *vb,vorbis_look_residue *vl,int **in,int *nz,int ch){ return tap_class(2,vb,vl,in,nz,ch); }
static int f1(oggpack_buffer *o,struct vorbis_block *vb,vorbis_look_residue *vl,int **in,int *nz,int ch,long **pw,int sm){ return tap_forward(1,o,vb,vl,in,nz,ch,pw,sm); }
static int f2(oggpack_buffer *o,struct vorbis_block *vb,vorbis_look_residue *vl,int **in,int *nz,int ch,long **pw,int sm){ return tap_forward(2,o,vb,vl,in,nz,ch,pw,sm); }

/* The replacement registry (stands in for lib/registry.c:31-44): same three
   tables, but the residue 1/2 bundles' class/forward go through the taps. */
static vorbis_func_residue tapped1, tapped2;
const vorbis_func_floor *const _floor_P[]={ &floor0_exportbundle, &floor1_exportbundle };
const vorbis_func_residue *const _residue_P[]={ &residue0_exportbundle, &tapped1, &tapped2 };
#ifdef REF_DROPIN
/* drop-in flavour (ref_dropin.c): the reference's vorbis_analysis() reaches the B200 library through this entry */
extern vorbis_func_mapping dropin_mapping_bundle;
const vorbis_func_mapping *const _mapping_P[]={ &dropin_mapping_bundle };
#else
const vorbis_func_mapping *const _mapping_P[]={ &mapping0_exportbundle };
#endif

__attribute__((constructor)) static void tap_registry_init(void){
  tapped1=residue1_exportbundle; tapped1.class=c1; tapped1.forward=f1;
  tapped2=residue2_exportbundle; tapped2.class=c2; tapped2.forward=f2;
  memset(g_enable,0,sizeof(g_enable));
}
