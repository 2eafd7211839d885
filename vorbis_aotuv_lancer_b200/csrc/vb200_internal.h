/* Internal device/host structures of libvorbis_b200 (not part of the C ABI). */
#ifndef VB200_INTERNAL_H
#define VB200_INTERNAL_H
#include <stdint.h>
#include <stddef.h>
#include "vorbis_b200.h"
#include "vb200_blob.h"

#define VB200_P_BANDS 17
#define VB200_P_LEVELS 8
#define VB200_EHMER 58            /* EHMER_MAX+2: two fenceposts + 56 curve points (psy.c:336-344) */
#define VB200_COMPAND_LEVELS 40
#define VB200_MAX_PSY 4
#define VB200_MAX_FLOORS 4
#define VB200_MAX_RESIDUES 6
#define VB200_MAX_MAPS 2
#define VB200_MAX_COUPLING 8
#define VB200_MAX_BOOKS 128
#define VB200_MAX_SUBMAPS 4

/* ---- device-resident tables (App. F of SURVEY.md) ---- */
struct DevPsy {
  int n, blockflag, firstoc, shiftoc, eighth, total_octave_lines, rate;
  int m3n[3];
  int tonecomp_endp, min_nn_lp, tonefix_end, n25p, n33p, n75p;
  int noisewindowfixed, normal_p, normal_start, normal_partition, partition;
  float m_val, tonecomp_thres, ath_adjatt, ath_maxatt, tone_masteratt[3];
  float tone_abs_limit, max_curve_dB, noisemaxsupp;
  double normal_thresh;
  const float *ath;          /* [n] */
  const int   *octave;       /* [n] */
  const int   *bark;         /* [n] packed ((lo-1)<<16)+(hi-1) */
  const float *tonecurves;   /* [17][8][58] */
  const float *noiseoffset;  /* [3][n] */
  const float *ntfix;        /* [n] */
  float compand[VB200_COMPAND_LEVELS], compand_hi[VB200_COMPAND_LEVELS];
  /* derived at setup time (static properties of the tables) */
  int nf_b1[2], nf_b2[2];    /* noise regression: first bin of the interior / held window class, per pass */
  const int4 *tone_groups;   /* max_seeds: (seed pos0, pos1, bin0, bin1) per group */
  int n_tone_groups, tone_tail;
  const int2 *tone_runs;     /* seed_loop: (first bin, last bin) of every run of bins sharing an octave line */
  int n_tone_runs;
};

struct DevFloor {
  int type, posts, n, quant_q, mult, info_n;
  float maxover, maxunder, maxerr, twofitweight, twofitatten;
  int postlist[VB200_MAX_POSTS], sorted[VB200_MAX_POSTS], forward[VB200_MAX_POSTS], reverse[VB200_MAX_POSTS];
  int lonb[VB200_MAX_POSTS], hinb[VB200_MAX_POSTS];
  /* posts 2.. grouped by their depth in the neighbour tree (a post's prediction reads only posts of lower levels):
     lv_order[lv_start[l] .. lv_start[l+1]) are the posts of level l = 1..nlevels (setup.cu) */
  int lv_order[VB200_MAX_POSTS], lv_start[VB200_MAX_POSTS + 2], nlevels;
  /* emission tables (vorbis_info_floor1, backends.h:60-70) for the device bit packer */
  int partitions, partclass[31], class_dim[16], class_subs[16], class_book[16], class_subbook[16 * 8];
};

struct DevResidue {
  int type, begin, end, grouping, partitions, partvals, groupbook, stages, phrase_dim;
  int secondstages[64], classmetric1[64], classmetric2[64];
  int partbooks[64*8];
};

struct DevBook {
  int dim, entries, quantvals, minval, delta;
  unsigned delta_rcp;        /* ceil(2^32 / delta): |x| / delta == umulhi(|x|, delta_rcp) for |x| < 2^21, delta < 1024 */
  int used;                     /* number of entries with a codeword */
  const uint8_t  *len;          /* [entries] codeword lengths (0 = unused) */
  const uint32_t *code;         /* [entries] */
  const int      *used_list;    /* [used] entry indices in ascending order */
  const int      *used_vals;    /* [used][8] lattice value of each used entry (res0.c:343-370 enumeration) */
  const short    *nearest;      /* [entries] for books with unused entries: index into used_list of the used entry
                                   nearest (first minimum) to lattice point i -- the exhaustive fallback of
                                   res0.c:343-370 precomputed for inputs that sit exactly on the lattice */
};

struct DevMap {
  int submaps, coupling_steps;
  int chmux[VB200_MAX_CHANNELS];
  int floorsubmap[16], residuesubmap[16];
  int mag[VB200_MAX_COUPLING], ang[VB200_MAX_COUPLING];
};

struct DevSetup {
  int channels, rate, bs[2], modebits, managed, lowpassr[2];
  int psys, floors, residues, books, maps, modes;
  int mode_blockflag[8], mode_mapping[8];
  float ampmax_att_per_sec;
  int pointlimit[2][VB200_PACKETBLOBS], prepointamp[VB200_PACKETBLOBS], postpointamp[VB200_PACKETBLOBS];
  int sliding_lowpass[2][VB200_PACKETBLOBS];
  int stn_compand[VB200_COMPAND_LEVELS];
  int freq_bfn128[128], freq_bfn256[256];
  double stereo_th[9], stereo_thx[9];
  float fromdB[256];
  int floor_of_ch[2][VB200_MAX_CHANNELS];   /* floorsubmap[chmuxlist[c]] of the mapping of mode W */
  DevPsy psy[VB200_MAX_PSY];
  DevFloor flr[VB200_MAX_FLOORS];
  DevResidue res[VB200_MAX_RESIDUES];
  DevMap map[VB200_MAX_MAPS];
  const DevBook *book;          /* [books] (device) */
  const float *win[2];          /* half windows, bs/2 each */
  /* transform tables per W: computed in double on the host at setup time */
  const float2 *mdct_tw[2];     /* [n/4]  exp(-i*pi*(8m+1)/(8*N2)) */
  const float2 *fft_root[2];    /* [n/4]  exp(-2*pi*i*k/(n/2)), k<n/4 : roots for the n/2-point FFT */
  const float2 *rfft_tw[2];     /* [n/2]  exp(-2*pi*i*k/n) */
  const float2 *fft_pass_tw[2]; /* fft_root regrouped per radix-4 pass (setup.cu), unit-stride reads */
  int fft_r2_off[2];            /* float2 offset of the radix-2 pass's twiddles in fft_pass_tw */
  /* tables of the operation-exact transforms (k_transform_exact.cu): the reference's own values */
  const float *xm_trig[2];      /* [n + n/4]  mdct_init, mdct.c:54-92 */
  const int *xm_bitrev[2];      /* [n/4] */
  const float *xf_wa[2];        /* [2n]  drfti1, smallft.c:5576-5644 */
  const float2 *xm_stage[2];    /* [n/4]  xm_trig regrouped per butterfly stage at unit stride (vb200_exact_tables_k1x) */
  const int2 *xm_rev[2];        /* [n/8]  xm_bitrev as padded complex-slot indices of K1x's work vector */
};

/* ---- host-side object behind vb200_setup ---- */
struct HostBook {
  int dim, entries, quantvals, minval, delta, maptype;
  const uint8_t *len; const uint32_t *code;   /* point into the owned blob copy */
};
struct HostFloor {
  int partitions;
  int partclass[31], class_dim[16], class_subs[16], class_book[16], class_subbook[16*8];
};

struct vb200_setup {
  int device;                 /* -1: host-only (no tables uploaded) */
  unsigned char *blob; size_t blob_bytes;
  DevSetup h;                 /* host copy (pointers are device pointers) */
  DevSetup *d;                /* device copy */
  HostBook book[VB200_MAX_BOOKS];
  HostFloor hflr[VB200_MAX_FLOORS];
  const unsigned char *hdr[3]; long hdrlen[3];
  int bi_i32[4]; double bi_f64[2];
  void *stream;               /* cudaStream_t */
  void **allocs; int nallocs, capallocs;
  int exact_transform;        /* K1x (default) or K1 (vb200_setup_set_exact_transform / VB200_EXACT_TRANSFORM=0) */
  void *state_pool;           /* per-setup stream state in HBM (pipeline.cu: StatePool), created with the first stream */
};

#ifdef __cplusplus
extern "C" {
#endif
void vb200_set_error(const char *fmt, ...);
void vb200_count_launch(int n);
void vb200_state_pool_free(void *pool);                          /* pipeline.cu */
void vb200_exact_tables_mdct(int n, float *T, int *bitrev);      /* k_transform_exact.cu */
void vb200_exact_tables_drft(int n, float *wa);
void vb200_exact_tables_k1x(int n, const float *T, const int *bitrev, float2 *stage_tw, int2 *rev);
const void *vb200_blob_find(const unsigned char *blob, size_t bytes, const char *name, uint32_t *dtype, uint32_t *count);
#ifdef __cplusplus
}
#endif

#ifdef __CUDACC__
#include <cuda_runtime.h>
#include "k_common.cuh"
#define CUDA_TRY(expr) do{ cudaError_t e_=(expr); if(e_!=cudaSuccess){ \
  vb200_set_error("%s:%d %s: %s",__FILE__,__LINE__,#expr,cudaGetErrorString(e_)); return VB200_EFAULT; } }while(0)
#define LAUNCH_CHECK() do{ vb200_count_launch(1); CUDA_TRY(cudaGetLastError()); }while(0)

/* stage launchers on device pointers (used by both the stage calls and the batched path) */
int vb200_launch_transform(const vb200_setup *s, cudaStream_t st, int W, int items,
                           const float *pcm, const long long *pcm_off, const int *lW, const int *nW,
                           const int *block_mode, const int *lW_block_mode,
                           float *gmdct, float *logfft, float *logmdct, float *local_ampmax, float *poste);
int vb200_launch_transform_exact(const vb200_setup *s, cudaStream_t st, int W, int items,
                           const float *pcm, const long long *pcm_off, const int *lW, const int *nW,
                           const int *block_mode, const int *lW_block_mode,
                           float *gmdct, float *logfft, float *logmdct, float *local_ampmax, float *poste);   /* K1x, k_transform_exact.cu */
int vb200_launch_noisemask(const vb200_setup *s, cudaStream_t st, int W, int items, const int *psy,
                           const int *block_mode, const float *lowcomp, const float *poste,
                           const float *logmdct, LastRef last, const long long *off2, const int *rule2,
                           float *noise, float *epeak, float *npeak, int npeak_stride);
int vb200_launch_tonemask(const vb200_setup *s, cudaStream_t st, int W, int items, const int *psy,
                          const float *logfft, const float *global_ampmax, const float *local_ampmax, float *tone,
                          float *rowA, int *rowB, int rowN);
int vb200_launch_mix(const vb200_setup *s, cudaStream_t st, int W, int nchains, const int *chain_start, const int *chain_len, const int *chain_items,
                     const int *psy, const int *params, const float *low_compand,
                     const float *noise, const float *tone, const float *logmdct, float *mdct, LastRef last,
                     const float *temp_state, float *temp_state_out, const long long *temp_off, float *npeak, int npeak_stride,
                     float *logmask, int nsel, int mask_stride, float *last_out, int *fit_gate, int floor_index);
int vb200_launch_floor_fit(const vb200_setup *s, cudaStream_t st, int W, int items, int floor_index,
                           const float *logmdct, const float *logmask, int mdct_stride, int mask_stride,
                           const int *gate, int *posts, int *fitted);
int vb200_launch_floor_interp(const vb200_setup *s, cudaStream_t st, int W, int items, int floor_index, const int *A, const int *B,
                              const int *fa, const int *fb, int del, int *out, int *fo);
int vb200_launch_floor_encode(const vb200_setup *s, cudaStream_t st, int W, int items, int floor_index,
                              int *posts, const int *fitted, int *vals, int *ilogmask, int *nonzero);
int vb200_launch_couple(const vb200_setup *s, cudaStream_t st, int W, int blocks, const int *psy, int blobno,
                        const float *mdct, const float *epeak, float *npeak, int npeak_stride, int *iwork, int *nonzero);
int vb200_launch_residue_class(const vb200_setup *s, cudaStream_t st, int W, int blocks, int resnum, int ch, const int *chlist,
                               int tot_ch, const int *in, const int *nonzero, int *partword, int *classified);
int vb200_residue_maxvec(const vb200_setup *s, int resnum);
int vb200_launch_pack(const vb200_setup *s, cudaStream_t st, int W, int blocks, const int *lW, const int *nW,
                      const int *vals, const int *floor_nonzero, int submaps, const int *syms, const int *nsyms,
                      int cap_blocks, int max_syms, unsigned char *pool, unsigned long long pool_cap,
                      unsigned long long *pool_top, long long *pk_off, int *pk_bytes);
int vb200_launch_residue_forward(const vb200_setup *s, cudaStream_t st, int W, int blocks, int resnum, int ch, const int *chlist,
                                 int tot_ch, int *in, const int *nonzero, const int *partword, int *syms, int max_syms, int *nsyms);
#endif
#endif
