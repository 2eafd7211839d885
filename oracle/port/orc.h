/* TEST INFRASTRUCTURE -- the CPU oracle: a plain-C restatement of the reference's
 * per-block analysis algorithms (one function per hot-path row of SURVEY.md
 * section 8a), written from the algorithm, not copied, and pinned against the real
 * reference build in oracle/_ref (tests/test_oracle_port.py) and the golden
 * vectors in tests/golden/.  Only tests/, smoke() and bench.py's cpu_baseline
 * leg may call it; the product never links it.
 *
 * All tables are passed in explicitly (filled from the setup blob by Python).
 */
#ifndef VB200_ORC_H
#define VB200_ORC_H
#include <stdint.h>

typedef struct {
  int n, blockflag, firstoc, shiftoc, eighth, total_octave_lines, rate;
  int m3n[3];
  int tonecomp_endp, min_nn_lp, tonefix_end, n25p, n33p, n75p;
  int noisewindowfixed, normal_p, normal_start, normal_partition, partition;
  float m_val, tonecomp_thres, ath_adjatt, ath_maxatt, tone_masteratt[3];
  float tone_abs_limit, max_curve_dB, noisemaxsupp;
  double normal_thresh;
  const float *ath; const int *octave; const int *bark; const float *tonecurves;
  const float *noiseoffset; const float *ntfix; const float *compand; const float *compand_hi;
  const int *stn_compand; const int *freq_bfn128; const int *freq_bfn256;
} orc_psy;

typedef struct {
  int posts, n, quant_q, mult;
  float maxover, maxunder, maxerr, twofitweight, twofitatten;
  const int *postlist, *sorted, *forward, *reverse, *lonb, *hinb;
} orc_floor;

typedef struct {
  int dim, entries, quantvals, minval, delta;
  const uint8_t *len;
} orc_book;

typedef struct {
  int type, begin, end, grouping, partitions, stages, phrase_dim, groupbook;
  const int *secondstages, *classmetric1, *classmetric2, *partbooks; /* partbooks[class*8+stage] */
  const orc_book *books;  /* all books of the setup */
} orc_residue;

typedef struct {
  int coupling_steps; const int *mag, *ang;
  const int *pointlimit; /* [2][15] */
  const int *prepointamp, *postpointamp; /* [15] */
  const double *stereo_th, *stereo_thx;  /* [9] */
  const float *fromdB; /* [256] */
} orc_couple_cfg;

float orc_todB(float x);
void  orc_window(float *d, int n, int ln, int rn, const float *wl, const float *wr);
void  orc_mdct_forward(const float *in, float *out, int n);
void  orc_mdct_forward_exact(const float *in, float *out, int n);   /* the reference's float operations, bit-exact (mdct_exact_port.c) */
void  orc_mdct_tables(int n, float *T, int *bitrev);
void  orc_drft_tables(int n, float *wa, int *ifac);
int   orc_drft_forward_exact(float *data, int n);                   /* smallft.c drft_forward, bit-exact (drft_exact_port.c) */
int   orc_logfft_exact(const float *w, int n, float *logfft, float *local_ampmax);
void  orc_logfft(const float *windowed, int n, float *logfft, float *local_ampmax);
void  orc_logmdct(const float *mdct, int n2, float *logmdct);
float orc_postnoise(const float *pcm, int n, int mode, int lw_mode);
float orc_loudnoise_fix(const orc_psy *p, float level, const float *logmdct, int block_mode, int lW_block_mode);

void  orc_noisemask(const orc_psy *p, float level, const float *logmdct, const float *lastmdct,
                    float *epeak, float *npeak, float *noise, float poste, int block_mode);
void  orc_tonemask(const orc_psy *p, const float *logfft, float *tone, float global_specmax, float local_specmax);
void  orc_offset_and_mix(const orc_psy *p, const float *noise, const float *tone, int sel, int managed,
                         float *logmask, float *mdct, const float *logmdct, float *lastmdct, float *tempmdct,
                         float low_compand, float *npeak, int end_block, int block_mode, int nW,
                         int lW_block_mode, int lW_no, int impadnum);

int   orc_floor1_fit(const orc_floor *f, const float *logmdct, const float *logmask, int *posts);
void  orc_floor1_interpolate(const orc_floor *f, const int *A, const int *B, int del, int *out);
int   orc_floor1_encode(const orc_floor *f, int *post, int fitted, int *vals, int *ilogmask, int n2);

void  orc_couple(const orc_psy *p, const orc_couple_cfg *c, int blobno, float **mdct, float **epeak, float **npeak,
                 int **iwork, int *nonzero, int sliding_lowpass, int ch, int lowpassr);

int   orc_residue_class(const orc_residue *r, int **in, const int *nonzero, int ch, int *partword /*[ch][partvals]*/);
int   orc_residue_forward(const orc_residue *r, int **in, const int *nonzero, int ch, const int *partword,
                          int n2, int *syms /*[max][2]*/, int max_syms);
#endif
