/* vb200 setup blob -- the flattened, read-only encoder configuration the
 * device side needs.
 *
 * It is what `vorbis_analysis_init` leaves behind in the reference
 * (`codec_setup_info` + the floor/residue/psy look objects,
 * lib/block.c:181-303, lib/codec_internal.h:102-163), serialised as a list of
 * named sections so that the exporter (reference side) and the
 * loader (csrc/setup.cu) only have to agree on names.
 *
 * Layout: vb200_blob_header, then `nsections` x { vb200_blob_section, payload
 * padded to 8 bytes }.
 */
#ifndef VB200_BLOB_H
#define VB200_BLOB_H
#include <stdint.h>

#define VB200_BLOB_MAGIC   0x42323056u  /* "V02B" */
#define VB200_BLOB_VERSION 1u

enum { VB200_F32=0, VB200_I32=1, VB200_U8=2, VB200_F64=3, VB200_U32=4 };

typedef struct {
  uint32_t magic;
  uint32_t version;
  uint32_t nsections;
  uint32_t total_bytes;
} vb200_blob_header;

typedef struct {
  char     name[24];   /* NUL padded, e.g. "psy3.ath" */
  uint32_t dtype;      /* VB200_F32 ... */
  uint32_t count;      /* elements */
} vb200_blob_section;

static inline uint32_t vb200_dtype_size(uint32_t dt){
  return dt==VB200_U8?1u:(dt==VB200_F64?8u:4u);
}

/* index constants for the fixed-position scalar sections ---------------- */
/* "cfg.i32" */
enum { VB200_CFG_CHANNELS=0, VB200_CFG_RATE, VB200_CFG_BS0, VB200_CFG_BS1, VB200_CFG_MODEBITS,
       VB200_CFG_MANAGED, VB200_CFG_LOWPASSR0, VB200_CFG_LOWPASSR1, VB200_CFG_PSYS, VB200_CFG_FLOORS,
       VB200_CFG_RESIDUES, VB200_CFG_BOOKS, VB200_CFG_MAPS, VB200_CFG_MODES, VB200_CFG_EIGHTH_OCTAVE_LINES,
       VB200_CFG_COUNT };
/* "psyK.i32" */
enum { VB200_PSY_N=0, VB200_PSY_BLOCKFLAG, VB200_PSY_FIRSTOC, VB200_PSY_SHIFTOC, VB200_PSY_EIGHTH,
       VB200_PSY_TOTAL_OCTAVE_LINES, VB200_PSY_RATE, VB200_PSY_M3N0, VB200_PSY_M3N1, VB200_PSY_M3N2,
       VB200_PSY_TONECOMP_ENDP, VB200_PSY_MIN_NN_LP, VB200_PSY_TONEFIX_END, VB200_PSY_N25P, VB200_PSY_N33P,
       VB200_PSY_N75P, VB200_PSY_NOISEWINDOWFIXED, VB200_PSY_NORMAL_P, VB200_PSY_NORMAL_START,
       VB200_PSY_NORMAL_PARTITION, VB200_PSY_I32_COUNT };
/* "psyK.f32" */
enum { VB200_PSYF_M_VAL=0, VB200_PSYF_TONECOMP_THRES, VB200_PSYF_ATH_ADJATT, VB200_PSYF_ATH_MAXATT,
       VB200_PSYF_TONE_MASTERATT0, VB200_PSYF_TONE_MASTERATT1, VB200_PSYF_TONE_MASTERATT2,
       VB200_PSYF_TONE_ABS_LIMIT, VB200_PSYF_MAX_CURVE_DB, VB200_PSYF_NOISEMAXSUPP, VB200_PSYF_COUNT };
/* "flrK.i32" */
enum { VB200_FLR_TYPE=0, VB200_FLR_POSTS, VB200_FLR_LOOK_N, VB200_FLR_QUANT_Q, VB200_FLR_MULT, VB200_FLR_INFO_N,
       VB200_FLR_PARTITIONS, VB200_FLR_I32_COUNT };
/* "flrK.f32" */
enum { VB200_FLRF_MAXOVER=0, VB200_FLRF_MAXUNDER, VB200_FLRF_MAXERR, VB200_FLRF_TWOFITWEIGHT, VB200_FLRF_TWOFITATTEN,
       VB200_FLRF_COUNT };
/* "resK.i32" */
enum { VB200_RES_TYPE=0, VB200_RES_BEGIN, VB200_RES_END, VB200_RES_GROUPING, VB200_RES_PARTITIONS, VB200_RES_PARTVALS,
       VB200_RES_GROUPBOOK, VB200_RES_STAGES, VB200_RES_PHRASE_DIM, VB200_RES_I32_COUNT };
/* "mapK.i32" */
enum { VB200_MAP_SUBMAPS=0, VB200_MAP_COUPLING_STEPS, VB200_MAP_I32_COUNT };
/* "bookK.i32" */
enum { VB200_BOOK_DIM=0, VB200_BOOK_ENTRIES, VB200_BOOK_QUANTVALS, VB200_BOOK_MINVAL, VB200_BOOK_DELTA,
       VB200_BOOK_MAPTYPE, VB200_BOOK_I32_COUNT };

#endif
