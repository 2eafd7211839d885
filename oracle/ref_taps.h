/* TEST INFRASTRUCTURE -- never linked into the product.
 *
 * A/B injection taps for the reference build under oracle/_ref/.  The
 * reference's stage functions are `extern` and called across object files, so
 * `-Wl,--wrap=<sym>` lets these taps see every stage's inputs and outputs
 * without editing a reference source file.  Records are
 * appended to one in-memory log that Python parses (oracle/refbuild.py).
 *
 * Record = 8 x int32 header {tag, seq, a0, a1, a2, dtype, count, a3} followed
 * by `count` elements (f32 / i32, or u8 padded to 4 bytes).
 */
#ifndef VB200_REF_TAPS_H
#define VB200_REF_TAPS_H

enum {
  TAP_BLOCK=1,      /* i32: W lW nW blocktype pcmend eofflag gp_lo gp_hi lW_block_mode lW_no impadnum managed; a0=channels */
  TAP_AMPMAX,       /* f32: ampmax_in ampmax_out */
  TAP_PCM,          /* f32 pcm[n] un-windowed; a0=ch a1=mode a2=lw_mode */
  TAP_POSTE,        /* f32 ret; a0=ch */
  TAP_WINDOWED,     /* f32 [n]; a0=ch (call order) */
  TAP_MDCT,         /* f32 [n/2]; a0=ch (call order) */
  TAP_LOUDFIX,      /* f32: level_in ret; a0=ch a1=block_mode a2=lW_block_mode */
  TAP_NM_LOGMDCT,   /* f32 [n]; a0=ch a1=psy a2=block_mode */
  TAP_NM_LASTMDCT,  /* f32 [n] */
  TAP_NM_SCAL,      /* f32: level poste */
  TAP_NM_EPEAK,     /* f32 [n] out */
  TAP_NM_NPEAK,     /* f32 [n/partition] out */
  TAP_NM_NOISE,     /* f32 [n] out */
  TAP_TM_LOGFFT,    /* f32 [n]; a0=ch a1=psy */
  TAP_TM_SCAL,      /* f32: global local */
  TAP_TM_TONE,      /* f32 [n] out */
  TAP_MIX_I32,      /* i32: sel managed end_block block_mode nW lW_block_mode lW_no impadnum; a0=ch a1=psy a2=sel */
  TAP_MIX_F32,      /* f32: low_compand */
  TAP_MIX_NOISE, TAP_MIX_TONE, TAP_MIX_MDCT_IN, TAP_MIX_LOGMDCT,
  TAP_MIX_LAST_IN,  /* f32 [2048] */
  TAP_MIX_TEMP_IN,  /* f32 [256] */
  TAP_MIX_NPEAK_IN,
  TAP_MIX_LOGMASK, TAP_MIX_MDCT_OUT,
  TAP_MIX_LAST_OUT, TAP_MIX_TEMP_OUT, TAP_MIX_NPEAK_OUT,
  TAP_FIT_LOGMDCT,  /* f32 [n]; a0=ch a1=floor index a2=call# within block */
  TAP_FIT_LOGMASK,
  TAP_FIT_POSTS,    /* i32 [posts], count 0 == NULL */
  TAP_INTERP,       /* i32: A[posts] B[posts] out[posts] ; a0=floor a1=del */
  TAP_FENC_IN,      /* i32 posts in (count 0 == NULL); a0=ch-in-order a1=floor a2=blob */
  TAP_FENC_POST,    /* i32 posts out (mutated) */
  TAP_FENC_ILOG,    /* i32 [n] */
  TAP_FENC_RET,     /* i32 [1] */
  TAP_CPL_I32,      /* i32: blobno sliding_lowpass ch lowpassr psy; a0=blob */
  TAP_CPL_MDCT, TAP_CPL_EPEAK, TAP_CPL_NPEAK_IN, TAP_CPL_IWORK_IN, TAP_CPL_NZ_IN,  /* [ch][..] flattened */
  TAP_CPL_IWORK_OUT, TAP_CPL_NZ_OUT, TAP_CPL_NPEAK_OUT,
  TAP_RES_I32,      /* i32: resnum type ch used submap ; a0=blob a1=submap-call */
  TAP_RES_IN,       /* i32 [ch][n] */
  TAP_RES_NZ,       /* i32 [ch] */
  TAP_RES_PARTWORD, /* i32 [rows][partvals] (count 0 == NULL); a2=rows */
  TAP_RES_SYMS,     /* i32 pairs (book,entry) emitted by ->forward */
  TAP_RES_OUT,      /* i32 [ch][n] after forward (cascade residual) */
  TAP_BLOCK_SYMS,   /* i32 pairs (book,entry) for the whole block, blob order */
  TAP_PACKET,       /* u8 packet bytes; a0=bytes a1=eos a2=gp_lo a3=gp_hi */
  TAP_BLOBBYTES,    /* i32 [15] bytes of every packet blob after analysis */
  TAP_COUNT
};

#ifdef __cplusplus
extern "C" {
#endif
void   ref_tap_config(const unsigned char *enable /*TAP_COUNT flags*/, long seq_lo, long seq_hi);
void   ref_tap_reset(void);
const unsigned char *ref_tap_data(long *bytes);
/* used by the driver */
void   ref_tap_arm(void *vb, int armed);
int    ref_tap_on(int tag);
void   ref_tap_put(int tag,int a0,int a1,int a2,int a3,int dtype,const void *p,long count);
#ifdef __cplusplus
}
#endif
#endif
