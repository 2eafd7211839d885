/* vorbis_b200 -- C ABI of the B200-native aoTuV/Vorbis per-block analysis path.
 *
 * Drop-in scope: everything `mapping0_forward` does per block
 * (/root/reference/lib/mapping0.c:738-1322) -- window, forward MDCT, FFT/dB
 * spectra, noise/tone masking, offset&mix, floor-1 fit/quantise/render, aoTuV
 * couple/quantise/normalise, residue classification + cascade lattice VQ --
 * batched over many blocks, channels and independent streams per launch.
 * Plain pointers and sizes only; no torch types, no C++ types.
 *
 * Every entry point names the reference interface it replaces.  Return value:
 * 0 on success or a negative VB200_E* code (same numbering as libvorbis'
 * OV_E* codes, include/vorbis/codec.h:223-238).  There is NO CPU fallback: with
 * no usable CUDA device every compute call returns VB200_EFAULT.
 *
 * Conventions
 *  - "block"    = one vorbis_block (all channels);  "ch-block" = one channel of it.
 *  - W          = block size flag (0 short / 1 long), n = blocksizes[W], N2 = n/2.
 *  - block_mode = blocktype | (W<<1): 0 impulse, 1 padding, 2 transition, 3 long
 *                 (mapping0.c:769-775).
 *  - Stage calls ("vb200_stage_*") take HOST pointers and copy in/out; they
 *    exist so the reference's own float inputs can be injected stage by stage
 *    (A/B harness).  All items of one stage call share one W.
 */
#ifndef VORBIS_B200_H
#define VORBIS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_MAX_CHANNELS 8
#define VB200_PACKETBLOBS  15      /* codec_internal.h:28 */
#define VB200_MAX_POSTS    65      /* VIF_POSIT+2, backends.h:55 */

#define VB200_OK       0
#define VB200_EFAULT  (-129)       /* OV_EFAULT : CUDA error / no device / internal */
#define VB200_EIMPL   (-130)       /* OV_EIMPL  : configuration outside the built path */
#define VB200_EINVAL  (-131)       /* OV_EINVAL : bad argument */

typedef struct vb200_setup  vb200_setup;   /* immutable per-configuration tables, resident in HBM */
typedef struct vb200_stream vb200_stream;  /* mutable per-stream state (private_state aoTuV fields + bitrate manager) */
typedef struct vb200_batch  vb200_batch;   /* workspace + results of one batched analysis */

/* ---------------------------------------------------------------- library */
const char *vb200_version(void);
/* number of usable CUDA devices (0 when none: every compute call then fails) */
int  vb200_device_count(void);
/* human-readable text of the last error on this thread */
const char *vb200_last_error(void);
/* kernels this library launched since load / since the last reset (bench "gpu_launches") */
long vb200_launch_count(void);
void vb200_launch_count_reset(void);

/* ---------------------------------------------------------------- setup
 * Replaces the table-building half of vorbis_analysis_init / _vds_shared_init
 * (lib/block.c:181-330): consumes the flattened codec_setup_info + look objects
 * (include/vb200_blob.h, written by the reference-side exporter described in
 * INTEGRATION.md) and uploads them to `device`. */
int  vb200_setup_create(const void *blob, size_t bytes, int device, vb200_setup **out);
void vb200_setup_destroy(vb200_setup *s);
/* Front end selection.  The default (on = 1) is K1x, the front end with operation-exact transforms: it performs the
 * reference's own float operations (mdct_forward mdct.c:1799-1869, drft_forward smallft.c:6770 with drftf1/dradf4/dradf2
 * :5652-6170), so that gmdct, logmdct, logfft and local_ampmax -- and through them every later stage and the packets --
 * are bit-identical to the reference's C path.  on = 0 (or VB200_EXACT_TRANSFORM=0 at create time) selects K1, the
 * same transforms in a radix-4 Stockham formulation, within north_star's 1e-5 of the reference but not bit-identical. */
int  vb200_setup_set_exact_transform(vb200_setup *s, int on);
/* host-side queries (valid without a GPU: pass device = -1 to vb200_setup_create) */
int  vb200_setup_channels(const vb200_setup *s);
long vb200_setup_rate(const vb200_setup *s);
int  vb200_setup_blocksize(const vb200_setup *s, int W);
int  vb200_setup_managed(const vb200_setup *s);
int  vb200_setup_posts(const vb200_setup *s, int W);        /* floor posts of the floor used by mode W */
int  vb200_setup_partvals(const vb200_setup *s, int W, int submap);
int  vb200_setup_header(const vb200_setup *s, int i, const unsigned char **data, long *bytes); /* the 3 stream headers */

/* ---------------------------------------------------------------- Ogg framing
   What the encoder's outer loop asks libogg for (examples/encoder_example.c:141-176,224-235:
   ogg_stream_init/packetin/flush/pageout): pages with lacing, sequence numbers and CRC around the packets of one
   logical stream.  Host-only.  The pages accumulate in the muxer; vb200_oggmux_data returns all of them. */
typedef struct vb200_oggmux vb200_oggmux;
int  vb200_oggmux_create(int32_t serialno, vb200_oggmux **out);                 /* ogg_stream_init */
void vb200_oggmux_destroy(vb200_oggmux *m);                                     /* ogg_stream_clear */
int  vb200_oggmux_packet(vb200_oggmux *m, const unsigned char *data, long bytes, int64_t granulepos, int eos);
                                                                                /* ogg_stream_packetin + pageout loop */
int  vb200_oggmux_flush(vb200_oggmux *m);                                       /* ogg_stream_flush loop (after the headers) */
const unsigned char *vb200_oggmux_data(const vb200_oggmux *m, size_t *bytes);
long vb200_oggmux_pages(const vb200_oggmux *m);
uint32_t vb200_ogg_crc(const unsigned char *data, size_t bytes);                /* CRC-32 0x04c11db7, init 0, unreflected */

/* ---------------------------------------------------------------- synthetic PCM (bench / tests) */
uint64_t vb200_splitmix64(uint64_t x);
/* planar out[c*stride + i]; seed 0 selects the survey's default seed */
int  vb200_synth_pcm(uint64_t seed, int channels, long rate, long nsamples, float *out, long stride);

/* ================================================================ stage calls (A/B injection) */

/* K1.  Replaces, per ch-block: _postnoise_detection (psy.c:619-648), _vorbis_apply_window
 * (window.c:2137-2261), mdct_forward (mdct.c:1799-1869), drft_forward (smallft.c:6770) and the dB
 * loops of mapping0.c:848-888 and :933-951.
 *   pcm        [items][n]   un-windowed block PCM (vb->pcm[i])
 *   lW,nW      [items]      window shapes;   block_mode,lW_block_mode [items] for the post-echo detector
 *   gmdct      [items][N2]  MDCT coefficients                    (out)
 *   logfft     [items][N2]  scale_dB+.5*todB(re^2+im^2)+.345     (out)
 *   logmdct    [items][N2]  todB(mdct)+.345                      (out)
 *   local_ampmax [items]    min(max_k logfft,0)                  (out)
 *   poste      [items]      _postnoise_detection result          (out) */
int vb200_stage_transform(vb200_setup *s, int W, int items,
                          const float *pcm, const int32_t *lW, const int32_t *nW,
                          const int32_t *block_mode, const int32_t *lW_block_mode,
                          float *gmdct, float *logfft, float *logmdct,
                          float *local_ampmax, float *poste);

/* K2.  Replaces _vp_noisemask (psy.c:3770-4074) incl. bark_noise_hybridmp and ntfix.
 *   psy [items] index of the vorbis_look_psy (blocktype+(W?2:0)); lowcomp = noise_compand_level.
 *   lastmdct [items][N2] in;  noise, epeak [items][N2] out;  npeak [items][N2/partition] out */
int vb200_stage_noisemask(vb200_setup *s, int W, int items, const int32_t *psy,
                          const int32_t *block_mode, const float *lowcomp, const float *poste,
                          const float *logmdct, const float *lastmdct,
                          float *noise, float *epeak, float *npeak);

/* K3.  Replaces _vp_tonemask (psy.c:4076-4142) incl. seed_loop/seed_curve/seed_chase/max_seeds. */
int vb200_stage_tonemask(vb200_setup *s, int W, int items, const int32_t *psy,
                         const float *logfft, const float *global_ampmax, const float *local_ampmax,
                         float *tone);

/* K4.  Replaces _vp_offset_and_mix (psy.c:4274-4502) incl. set_m3p.
 *   params [items][8] = offset_select, bit_managed, end_block, block_mode, nW, lW_block_mode, lW_no, impadnum
 *   mdct in/out [items][N2]; lastmdct in/out [items][2048]; tempmdct in/out [items][256];
 *   npeak in/out [items][N2/partition]; logmask out [items][N2] */
int vb200_stage_offset_and_mix(vb200_setup *s, int W, int items, const int32_t *psy,
                               const int32_t *params, const float *low_compand,
                               const float *noise, const float *tone, const float *logmdct,
                               float *mdct, float *lastmdct, float *tempmdct, float *npeak,
                               float *logmask);

/* K5a. Replaces floor1_fit (floor1.c:597-750).
 *   posts out [items][VB200_MAX_POSTS]; fitted out [items] (0 == the reference returns NULL) */
int vb200_stage_floor1_fit(vb200_setup *s, int W, int items, int floor_index,
                           const float *logmdct, const float *logmask,
                           int32_t *posts, int32_t *fitted);

/* K5b. Replaces the non-emitting parts of floor1_encode (floor1.c:774-974): post quantisation and
 * prediction (-> `vals`, the numbers the host Huffman-codes) and the integer floor render.
 *   posts in/out [items][VB200_MAX_POSTS] (mutated exactly as the reference mutates `post`);
 *   fitted in [items]; vals out [items][VB200_MAX_POSTS]; ilogmask out [items][N2]; nonzero out [items] */
int vb200_stage_floor1_encode(vb200_setup *s, int W, int items, int floor_index,
                              int32_t *posts, const int32_t *fitted,
                              int32_t *vals, int32_t *ilogmask, int32_t *nonzero);

/* K6.  Replaces _vp_couple_quantize_normalize (psy.c:4858-5142) for `blocks` whole blocks.
 *   mdct,epeak [blocks][ch][N2]; npeak in/out [blocks][ch][N2/partition];
 *   iwork in (ilogmask) / out (quantised residue) [blocks][ch][N2]; nonzero in/out [blocks][ch] */
int vb200_stage_couple(vb200_setup *s, int W, int blocks, const int32_t *psy, int blobno,
                       const float *mdct, const float *epeak, float *npeak,
                       int32_t *iwork, int32_t *nonzero);

/* K7a. Replaces ->class of the residue backend (res1_class/res2_class, res0.c:735-768) for the
 * channels of one submap.  in [blocks][ch][N2]; nonzero [blocks][ch];
 *   partword out [blocks][ch][partvals] (res2 uses row 0 only); classified out [blocks] (0 == NULL) */
int vb200_stage_residue_class(vb200_setup *s, int W, int blocks, int resnum, int ch,
                              const int32_t *in, const int32_t *nonzero,
                              int32_t *partword, int32_t *classified);

/* K7b. Replaces ->forward (res1_forward/res2_forward -> _01forward -> _encodepart ->
 * local_book_besterror, res0.c:316-378,528-640,715-799): the VQ search.  Emits the (book,entry)
 * symbol list in the reference's emission order instead of bits.
 *   in [blocks][ch][N2] in/out (cascade residual left in place for res1; res2 works on the
 *   interleaved copy exactly like the reference, so `in` is unchanged there);
 *   syms out [blocks][max_syms][2]; nsyms out [blocks] */
int vb200_stage_residue_forward(vb200_setup *s, int W, int blocks, int resnum, int ch,
                                int32_t *in, const int32_t *nonzero, const int32_t *partword,
                                int32_t *syms, int32_t max_syms, int32_t *nsyms);

/* ================================================================ whole path */

/* One block handed to the batched path: the fields of vorbis_block / vorbis_block_internal that
 * mapping0_forward reads (codec.h:86-119, codec_internal.h:32-42). */
typedef struct {
  int32_t  stream;       /* index into the stream array given to vb200_batch_run */
  int32_t  W, lW, nW;    /* vb->W, vb->lW, vb->nW */
  int32_t  blocktype;    /* vbi->blocktype (BLOCKTYPE_*) */
  int32_t  eofflag;      /* vb->eofflag */
  int64_t  granulepos;   /* vb->granulepos */
  int64_t  pcm_offset;   /* element offset of this block's channel-0 PCM inside the `pcm` buffer */
  int64_t  pcm_stride;   /* element distance between channels */
} vb200_block;

/* A/B hook for the block-size detector: the windowed 128-point forward MDCT _ve_amp takes of one step
 * (envelope.c:139-186: vec[i] = data[i]*mdct_win[i]; mdct_forward) with the reference's own float operations -- 128
 * samples in, 64 coefficients out, bit-identical to the reference's.  Host code (the device detector K0 runs the same
 * function, csrc/ve_mdct128.cuh). */
int  vb200_stage_detector_mdct(const float *pcm128, float *coef64);
/* Host-side segmentation of a whole stream into analysis blocks (works without a GPU).  Replaces,
 * for offline encoding, vorbis_analysis_buffer/_wrote (block.c:411-553, incl. the LPC head/tail
 * extrapolation of lpc.c), _ve_envelope_search/_ve_envelope_mark (envelope.c:569-706) and the
 * window sequencing of vorbis_analysis_blockout (block.c:557-812).  The returned blocks index the
 * returned timeline buffer ([channels][total], lead-in + PCM + tail), which is what vb200_batch_run
 * takes as `pcm`.  chunk = the vorbis_analysis_wrote granularity to mirror (<=0: 1024). */
typedef struct vb200_segments vb200_segments;
int  vb200_segment_stream(const vb200_setup *s, const float *pcm, long nsamples, long stride, int chunk,
                          int stream_index, vb200_segments **out);
void vb200_segments_free(vb200_segments *g);
int  vb200_segments_count(const vb200_segments *g);
const vb200_block *vb200_segments_blocks(const vb200_segments *g);
const float *vb200_segments_pcm(const vb200_segments *g, long *elems);

/* The same segmentation for many streams at once, with the detector's heavy half (128-point MDCT +
 * band energies every 64 samples; envelope.c:101-217) as CUDA kernels.  The timeline of all streams is
 * uploaded once and stays in HBM: pass vb200_segmenter_device_pcm() to vb200_batch_run with
 * pcm_on_device = 1.  Blocks are stream-major; block.stream = index into pcm[]. */
typedef struct vb200_segmenter vb200_segmenter;
int  vb200_segmenter_create(vb200_setup *s, vb200_segmenter **out);
void vb200_segmenter_destroy(vb200_segmenter *g);
int  vb200_segmenter_run(vb200_segmenter *g, int nstreams, const float *const *pcm, const long *nsamples,
                         const long *stride, int chunk, int threads);
/* the same in two halves, for callers that overlap: _begin enqueues the uploads and the detector kernels and
   returns at once; _finish waits for the marks, sequences the windows and uploads the extrapolated tails */
int  vb200_segmenter_begin(vb200_segmenter *g, int nstreams, const float *const *pcm, const long *nsamples,
                           const long *stride, int chunk, int threads);
/* the same for 16-bit PCM (planar [channels][nsamples], value = sample / 32768.f: the conversion
   examples/encoder_example.c:183-190 performs on the host before vorbis_analysis_wrote).  Half the host-to-device
   bytes; the device converts.  Timeline, blocks and packets are those of vb200_segmenter_begin on the converted
   samples. */
int  vb200_segmenter_begin_s16(vb200_segmenter *g, int nstreams, const short *const *pcm, const long *nsamples,
                               const long *stride, int chunk, int threads);
int  vb200_segmenter_finish(vb200_segmenter *g);
/* order g's uploads behind prev's (both begun before either is finished): the first group's PCM arrives first */
int  vb200_segmenter_after(vb200_segmenter *g, const vb200_segmenter *prev);
int  vb200_segmenter_count(const vb200_segmenter *g);
const vb200_block *vb200_segmenter_blocks(const vb200_segmenter *g);
const float *vb200_segmenter_device_pcm(const vb200_segmenter *g, size_t *elems);
const float *vb200_segmenter_host_pcm(const vb200_segmenter *g);
/* 1 when the last run took the samples straight from the caller's page-locked buffers (every pcm[i] was
   allocated with vb200_pinned_alloc / cudaHostAlloc / cudaHostRegister): the host timeline then holds only the
   head and tail windows the host itself reads.  Pageable input is staged through the segmenter's own pinned
   timeline (0).  Both give the same device timeline and the same blocks. */
int  vb200_segmenter_direct(const vb200_segmenter *g);
/* copy the device timeline of the last run back to host memory (tests, debugging) */
int  vb200_segmenter_read_pcm(const vb200_segmenter *g, float *out, size_t elems);
/* Page-locked host memory for PCM handed to vb200_segmenter_run (the buffer vorbis_analysis_buffer() returns,
   block.c:421-436, would be allocated with this). */
int  vb200_pinned_alloc(size_t bytes, void **out);
void vb200_pinned_free(void *p);

/* Per-stream state: replaces the aoTuV fields of private_state (codec_internal.h:85-92), the
 * ampmax feedback of block.c:649-651 and bitrate_manager_state (bitrate.h:26-38).  The state lives in HBM inside the
 * setup (allocated with its first stream, freed by vb200_setup_destroy): destroy a setup's streams, batches and
 * segmenters before the setup, as vorbis_dsp_clear goes before vorbis_info_clear. */
int  vb200_stream_create(vb200_setup *s, vb200_stream **out);
void vb200_stream_destroy(vb200_stream *st);

int  vb200_batch_create(vb200_setup *s, int max_blocks, vb200_batch **out);
void vb200_batch_destroy(vb200_batch *b);

/* Replaces `vorbis_analysis` + `vorbis_bitrate_addblock` (analysis.c:29-63, bitrate.c:73-227) for
 * `nblocks` blocks of `nstreams` streams in one go.  Blocks of one stream must appear in stream
 * order.  `pcm` is host memory unless pcm_on_device != 0.  Results are packets, fetched below. */
int  vb200_batch_run(vb200_batch *b, vb200_stream **streams, int nstreams,
                     const vb200_block *blocks, int nblocks,
                     const float *pcm, size_t pcm_elems, int pcm_on_device);
/* The same in two halves, so that a caller can overlap the host side of the next batch (segmentation,
 * H2D) with the kernels of this one: submit enqueues everything on the batch's own CUDA stream and
 * returns; collect waits, fetches the packets and (managed mode) runs the bitrate manager. */
int  vb200_batch_submit(vb200_batch *b, vb200_stream **streams, int nstreams,
                        const vb200_block *blocks, int nblocks,
                        const float *pcm, size_t pcm_elems, int pcm_on_device);
int  vb200_batch_collect(vb200_batch *b);
/* order b's kernels behind prev's (both may be submitted back to back): the host half of one job overlaps the device
   half of the previous one without the two sharing the SMs */
int  vb200_batch_after(vb200_batch *b, vb200_batch *prev);
/* Scheduling rank of this batch's CUDA stream among batches that are in flight together: 0 = served first.  Segmenter
   streams always rank above every batch (their detector kernels gate the host's block-size decision).  Call while the
   batch is idle. */
int  vb200_batch_set_rank(vb200_batch *b, int rank);
/* Debug/verification switch: 1 = emit codewords on the host from the device's symbol lists (the
 * reference-order implementation in emit.cu), 0 (default) = pack bits on the device (K8). */
int  vb200_batch_set_host_emit(vb200_batch *b, int on);
/* Replaces `vorbis_bitrate_flushpacket` (bitrate.c:229-252): packet i of the last run, in block
 * order.  Memory is owned by the batch until the next run. */
int  vb200_batch_packet(vb200_batch *b, int i, const unsigned char **data, long *bytes,
                        int64_t *granulepos, int *eos, int *stream);
/* All packets of the last run at once: returns their number; arrays stay owned by the batch. */
int  vb200_batch_packet_table(vb200_batch *b, const unsigned char **base, const size_t **off, const size_t **bytes,
                              const int64_t **granulepos, const int **eos, const int **stream);
/* device time (ms) of the kernels of the last run, measured with CUDA events on the batch's stream */
float vb200_batch_kernel_ms(vb200_batch *b);
/* managed mode: how many times the last run evaluated the masking stages before the set of ch-blocks that write
 * lastmdct / tempmdct (those whose nominal floor fit exists, mapping0.c:1097) agreed with the links between blocks
 * that were derived from it; 1 unless a stream holds digitally silent blocks.  Always 1 in VBR mode. */
int  vb200_batch_link_rounds(vb200_batch *b);
/* Device-side timing of a region that spans several runs / batches: vb200_batch_mark records a CUDA event on the
 * batch's own stream behind everything the batch has been given so far; vb200_batch_mark_elapsed waits for both marks
 * and returns the device time between them (bench.py times its K steps this way, max over ranks). */
int  vb200_batch_mark(vb200_batch *b);
int  vb200_batch_mark_elapsed(vb200_batch *from, vb200_batch *to, float *ms);
/* per-kernel device time of the last run: name[i], ms[i]; returns the number of kernels */
int  vb200_batch_kernel_times(vb200_batch *b, const char **names, float *ms, int cap);
/* intermediate vectors of the last run for the A/B harness (host copies).  what: see VB200_GET_* */
enum { VB200_GET_GMDCT=0, VB200_GET_LOGFFT, VB200_GET_LOGMDCT, VB200_GET_NOISE, VB200_GET_TONE,
       VB200_GET_LOGMASK, VB200_GET_EPEAK, VB200_GET_IWORK, VB200_GET_ILOGMASK, VB200_GET_POSTS,
       VB200_GET_POSTS_LO, VB200_GET_POSTS_HI,   /* managed mode: the low- and the high-rate fit (blobs 0 and 14) */
       VB200_GET_NPEAK, VB200_GET_BLOB_BYTES };  /* npeak after the last blob; bytes of each packet blob (device packer) */
int  vb200_batch_get(vb200_batch *b, int what, int block, int ch, void *out, int count);

#ifdef __cplusplus
}
#endif
#endif
