/* vb200_setup: parse the setup blob (include/vb200_blob.h) and make the tables
 * resident in HBM.  Replaces the table-building half of _vds_shared_init
 * (/root/reference/lib/block.c:181-303): mdct_init/drft_init become the
 * twiddle tables computed here in double; _vp_psy_init / floor1_look /
 * res0_look results arrive flattened in the blob. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include "vb200_internal.h"
#include <vector>

static thread_local char g_err[512] = "";
static long g_launches = 0;

extern "C" void vb200_set_error(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char *vb200_last_error(void) { return g_err; }
extern "C" const char *vb200_version(void) { return "vorbis_b200 0.1 (sm_100a)"; }
extern "C" void vb200_count_launch(int n) { __sync_fetch_and_add(&g_launches, (long)n); }
extern "C" long vb200_launch_count(void) { return g_launches; }
extern "C" void vb200_launch_count_reset(void) { g_launches = 0; }

extern "C" int vb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const void *vb200_blob_find(const unsigned char *blob, size_t bytes, const char *name,
                                       uint32_t *dtype, uint32_t *count) {
  const vb200_blob_header *H = (const vb200_blob_header *)blob;
  size_t pos = sizeof(*H);
  for (uint32_t i = 0; i < H->nsections && pos + sizeof(vb200_blob_section) <= bytes; i++) {
    const vb200_blob_section *s = (const vb200_blob_section *)(blob + pos);
    size_t nb = (size_t)s->count * vb200_dtype_size(s->dtype);
    pos += sizeof(*s);
    if (!strncmp(s->name, name, sizeof(s->name))) {
      if (dtype) *dtype = s->dtype;
      if (count) *count = s->count;
      return blob + pos;
    }
    pos += (nb + 7) & ~(size_t)7;
  }
  return NULL;
}

namespace {
struct Loader {
  vb200_setup *s;
  bool ok = true;
  const void *get(const char *name, uint32_t dtype, uint32_t mincount, uint32_t *count = nullptr) {
    uint32_t dt = 0, c = 0;
    const void *p = vb200_blob_find(s->blob, s->blob_bytes, name, &dt, &c);
    if (!p || dt != dtype || c < mincount) {
      if (ok) vb200_set_error("setup blob: section '%s' missing or malformed", name);
      ok = false;
      return nullptr;
    }
    if (count) *count = c;
    return p;
  }
  const void *getk(const char *fmt, int k, uint32_t dtype, uint32_t mincount, uint32_t *count = nullptr) {
    char nm[32]; snprintf(nm, sizeof(nm), fmt, k); return get(nm, dtype, mincount, count);
  }
  /* upload `bytes` to the device (or return nullptr in host-only mode) */
  void *up(const void *src, size_t bytes) {
    if (s->device < 0 || !src || !ok) return nullptr;
    void *d = nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 4) != cudaSuccess) { ok = false; vb200_set_error("cudaMalloc failed in setup"); return nullptr; }
    if (bytes && cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { ok = false; vb200_set_error("cudaMemcpy failed in setup"); }
    if (s->nallocs >= s->capallocs) { s->capallocs = s->capallocs * 2 + 64; s->allocs = (void **)realloc(s->allocs, sizeof(void *) * s->capallocs); }
    s->allocs[s->nallocs++] = d;
    return d;
  }
};

static void copy_i(int *dst, const void *src, int n) { if (src) memcpy(dst, src, sizeof(int) * n); }
}  // namespace

extern "C" int vb200_setup_set_exact_transform(vb200_setup *s, int on) {
  if (!s) { vb200_set_error("vb200_setup_set_exact_transform: bad arguments"); return VB200_EINVAL; }
  s->exact_transform = on ? 1 : 0;
  return VB200_OK;
}

extern "C" void vb200_setup_destroy(vb200_setup *s) {
  if (!s) return;
  if (s->device >= 0) {
    cudaSetDevice(s->device);
    for (int i = 0; i < s->nallocs; i++) cudaFree(s->allocs[i]);
    if (s->state_pool) vb200_state_pool_free(s->state_pool);
    if (s->stream) cudaStreamDestroy((cudaStream_t)s->stream);
  }
  free(s->allocs);
  free(s->blob);
  delete s;
}

extern "C" int vb200_setup_create(const void *blob, size_t bytes, int device, vb200_setup **out) {
  if (!blob || !out || bytes < sizeof(vb200_blob_header)) { vb200_set_error("vb200_setup_create: bad arguments"); return VB200_EINVAL; }
  const vb200_blob_header *H = (const vb200_blob_header *)blob;
  if (H->magic != VB200_BLOB_MAGIC || H->version != VB200_BLOB_VERSION || H->total_bytes != bytes) {
    vb200_set_error("vb200_setup_create: not a vb200 setup blob"); return VB200_EINVAL;
  }
  if (device >= 0) {
    if (device >= vb200_device_count()) { vb200_set_error("vb200_setup_create: CUDA device %d not available (no CPU fallback)", device); return VB200_EFAULT; }
    CUDA_TRY(cudaSetDevice(device));
  }
  vb200_setup *s = new vb200_setup();
  memset(s, 0, sizeof(*s));
  s->device = device;
  { const char *e = getenv("VB200_EXACT_TRANSFORM"); s->exact_transform = (e && *e == '0') ? 0 : 1; }      /* K1x unless switched off */
  s->blob = (unsigned char *)malloc(bytes);
  memcpy(s->blob, blob, bytes);
  s->blob_bytes = bytes;
  Loader L; L.s = s;
  DevSetup &h = s->h;

  const int *cfg = (const int *)L.get("cfg.i32", VB200_I32, VB200_CFG_COUNT);
  if (!cfg) { vb200_setup_destroy(s); return VB200_EINVAL; }
  h.channels = cfg[VB200_CFG_CHANNELS]; h.rate = cfg[VB200_CFG_RATE];
  h.bs[0] = cfg[VB200_CFG_BS0]; h.bs[1] = cfg[VB200_CFG_BS1];
  h.modebits = cfg[VB200_CFG_MODEBITS]; h.managed = cfg[VB200_CFG_MANAGED];
  h.lowpassr[0] = cfg[VB200_CFG_LOWPASSR0]; h.lowpassr[1] = cfg[VB200_CFG_LOWPASSR1];
  h.psys = cfg[VB200_CFG_PSYS]; h.floors = cfg[VB200_CFG_FLOORS]; h.residues = cfg[VB200_CFG_RESIDUES];
  h.books = cfg[VB200_CFG_BOOKS]; h.maps = cfg[VB200_CFG_MAPS]; h.modes = cfg[VB200_CFG_MODES];
  if (h.channels < 1 || h.channels > VB200_MAX_CHANNELS || h.psys > VB200_MAX_PSY || h.floors > VB200_MAX_FLOORS ||
      h.residues > VB200_MAX_RESIDUES || h.maps > VB200_MAX_MAPS || h.books > VB200_MAX_BOOKS || h.modes > 8) {
    vb200_set_error("setup outside the built path (channels/psys/floors/residues/maps/books)"); vb200_setup_destroy(s); return VB200_EIMPL;
  }
  if (h.bs[0] != 256 || h.bs[1] != 2048) {
    vb200_set_error("only blocksizes 256/2048 are built (got %d/%d)", h.bs[0], h.bs[1]); vb200_setup_destroy(s); return VB200_EIMPL;
  }
  { const int *m = (const int *)L.get("cfg.modes", VB200_I32, h.modes * 2);
    for (int i = 0; m && i < h.modes; i++) { h.mode_blockflag[i] = m[i * 2]; h.mode_mapping[i] = m[i * 2 + 1]; } }

  for (int w = 0; w < 2; w++) {
    uint32_t c = 0; const void *p = L.getk("win%d", w, VB200_F32, h.bs[w] / 2, &c);
    h.win[w] = (const float *)L.up(p, sizeof(float) * h.bs[w] / 2);
    /* twiddles in double -> float */
    int n = h.bs[w], N2 = n / 2, q = n / 4;
    float2 *t = (float2 *)malloc(sizeof(float2) * N2);
    for (int m = 0; m < q; m++) { double a = -M_PI * (8.0 * m + 1.0) / (8.0 * N2); t[m].x = (float)cos(a); t[m].y = (float)sin(a); }
    h.mdct_tw[w] = (const float2 *)L.up(t, sizeof(float2) * q);
    for (int k = 0; k < q; k++) { double a = -2.0 * M_PI * k / (double)N2; t[k].x = (float)cos(a); t[k].y = (float)sin(a); }
    h.fft_root[w] = (const float2 *)L.up(t, sizeof(float2) * q);
    {
      /* the same roots laid out per radix-4 pass: level Ns (4, 16, ... <= N2/4) holds w^1[Ns], w^2[Ns], w^3[Ns] with
         w = root[k * N2 / (4 Ns)], at float2 offset Ns - 4; then the twiddles of the one radix-2 pass.  A pass reads its twiddles at unit stride (the strided reads of the root table cost
         up to 16-way shared-memory bank conflicts).  Values are taken from the root table, so results do not change. */
      const int half = N2 / 2;
      std::vector<float2> pt;
      for (int Ns = 4; Ns * 4 <= N2; Ns <<= 2) {
        pt.resize((size_t)(Ns - 4) + 3 * Ns);
        for (int k = 0; k < Ns; k++) {
          const int tt = k * (N2 / (4 * Ns));
          pt[(Ns - 4) + k] = t[tt];
          pt[(Ns - 4) + Ns + k] = t[2 * tt];
          float2 w3; if (3 * tt < half) w3 = t[3 * tt]; else { w3.x = -t[3 * tt - half].x; w3.y = -t[3 * tt - half].y; }
          pt[(Ns - 4) + 2 * Ns + k] = w3;
        }
      }
      h.fft_r2_off[w] = (int)pt.size();
      /* only one of the two transform sizes (N2 and N2/2 points) has an odd log2, i.e. a radix-2 pass */
      int lg = 0; while ((1 << lg) < N2) lg++;
      if (lg & 1) for (int j = 0; j < N2 / 2; j++) pt.push_back(t[j]);         /* N2 points: exp(-2*pi*i*j/N2) */
      else for (int j = 0; j < N2 / 4; j++) pt.push_back(t[2 * j]);            /* N2/2 points */
      h.fft_pass_tw[w] = (const float2 *)L.up(pt.data(), sizeof(float2) * pt.size());
    }
    for (int k = 0; k < N2; k++) { double a = -2.0 * M_PI * k / (double)n; t[k].x = (float)cos(a); t[k].y = (float)sin(a); }
    h.rfft_tw[w] = (const float2 *)L.up(t, sizeof(float2) * N2);
    free(t);
    {
      /* tables of the operation-exact front end (K1x) */
      std::vector<float> xt((size_t)n + n / 4), xw((size_t)2 * n, 0.f);
      std::vector<int> xb((size_t)n / 4);
      vb200_exact_tables_mdct(n, xt.data(), xb.data());
      vb200_exact_tables_drft(n, xw.data());
      h.xm_trig[w] = (const float *)L.up(xt.data(), sizeof(float) * xt.size());
      h.xm_bitrev[w] = (const int *)L.up(xb.data(), sizeof(int) * xb.size());
      h.xf_wa[w] = (const float *)L.up(xw.data(), sizeof(float) * xw.size());
      std::vector<float2> xs((size_t)n / 4); std::vector<int2> xr((size_t)n / 8);
      vb200_exact_tables_k1x(n, xt.data(), xb.data(), xs.data(), xr.data());
      h.xm_stage[w] = (const float2 *)L.up(xs.data(), sizeof(float2) * xs.size());
      h.xm_rev[w] = (const int2 *)L.up(xr.data(), sizeof(int2) * xr.size());
    }
  }

  { const float *f = (const float *)L.get("psyg.f32", VB200_F32, 4);
    if (f) h.ampmax_att_per_sec = f[0];
    copy_i(&h.pointlimit[0][0], L.get("psyg.pointlimit", VB200_I32, 2 * VB200_PACKETBLOBS), 2 * VB200_PACKETBLOBS);
    copy_i(h.prepointamp, L.get("psyg.prepointamp", VB200_I32, VB200_PACKETBLOBS), VB200_PACKETBLOBS);
    copy_i(h.postpointamp, L.get("psyg.postpointamp", VB200_I32, VB200_PACKETBLOBS), VB200_PACKETBLOBS);
    copy_i(&h.sliding_lowpass[0][0], L.get("psyg.sliding_lowpass", VB200_I32, 2 * VB200_PACKETBLOBS), 2 * VB200_PACKETBLOBS);
    copy_i(s->bi_i32, L.get("bi.i32", VB200_I32, 4), 4);
    const double *d = (const double *)L.get("bi.f64", VB200_F64, 2);
    if (d) { s->bi_f64[0] = d[0]; s->bi_f64[1] = d[1]; } }

  for (int k = 0; k < h.psys && L.ok; k++) {
    DevPsy &p = h.psy[k];
    const int *v = (const int *)L.getk("psy%d.i32", k, VB200_I32, VB200_PSY_I32_COUNT);
    const float *f = (const float *)L.getk("psy%d.f32", k, VB200_F32, VB200_PSYF_COUNT);
    const double *d = (const double *)L.getk("psy%d.f64", k, VB200_F64, 1);
    if (!v || !f || !d) break;
    p.n = v[VB200_PSY_N]; p.blockflag = v[VB200_PSY_BLOCKFLAG]; p.firstoc = v[VB200_PSY_FIRSTOC]; p.shiftoc = v[VB200_PSY_SHIFTOC];
    p.eighth = v[VB200_PSY_EIGHTH]; p.total_octave_lines = v[VB200_PSY_TOTAL_OCTAVE_LINES]; p.rate = v[VB200_PSY_RATE];
    p.m3n[0] = v[VB200_PSY_M3N0]; p.m3n[1] = v[VB200_PSY_M3N1]; p.m3n[2] = v[VB200_PSY_M3N2];
    p.tonecomp_endp = v[VB200_PSY_TONECOMP_ENDP]; p.min_nn_lp = v[VB200_PSY_MIN_NN_LP]; p.tonefix_end = v[VB200_PSY_TONEFIX_END];
    p.n25p = v[VB200_PSY_N25P]; p.n33p = v[VB200_PSY_N33P]; p.n75p = v[VB200_PSY_N75P];
    p.noisewindowfixed = v[VB200_PSY_NOISEWINDOWFIXED]; p.normal_p = v[VB200_PSY_NORMAL_P];
    p.normal_start = v[VB200_PSY_NORMAL_START]; p.normal_partition = v[VB200_PSY_NORMAL_PARTITION];
    p.partition = p.normal_p ? p.normal_partition : 16;
    p.m_val = f[VB200_PSYF_M_VAL]; p.tonecomp_thres = f[VB200_PSYF_TONECOMP_THRES]; p.ath_adjatt = f[VB200_PSYF_ATH_ADJATT];
    p.ath_maxatt = f[VB200_PSYF_ATH_MAXATT];
    p.tone_masteratt[0] = f[VB200_PSYF_TONE_MASTERATT0]; p.tone_masteratt[1] = f[VB200_PSYF_TONE_MASTERATT1]; p.tone_masteratt[2] = f[VB200_PSYF_TONE_MASTERATT2];
    p.tone_abs_limit = f[VB200_PSYF_TONE_ABS_LIMIT]; p.max_curve_dB = f[VB200_PSYF_MAX_CURVE_DB]; p.noisemaxsupp = f[VB200_PSYF_NOISEMAXSUPP];
    p.normal_thresh = d[0];
    if (p.n != h.bs[p.blockflag] / 2 || p.total_octave_lines > 1024 || p.eighth != 8) { vb200_set_error("psy %d outside the built path", k); L.ok = false; break; }
    p.ath = (const float *)L.up(L.getk("psy%d.ath", k, VB200_F32, p.n), sizeof(float) * p.n);
    p.octave = (const int *)L.up(L.getk("psy%d.octave", k, VB200_I32, p.n), sizeof(int) * p.n);
    p.bark = (const int *)L.up(L.getk("psy%d.bark", k, VB200_I32, p.n), sizeof(int) * p.n);
    p.tonecurves = (const float *)L.up(L.getk("psy%d.tonecurves", k, VB200_F32, VB200_P_BANDS * VB200_P_LEVELS * VB200_EHMER),
                                       sizeof(float) * VB200_P_BANDS * VB200_P_LEVELS * VB200_EHMER);
    p.noiseoffset = (const float *)L.up(L.getk("psy%d.noiseoffset", k, VB200_F32, 3 * p.n), sizeof(float) * 3 * p.n);
    p.ntfix = (const float *)L.up(L.getk("psy%d.ntfix", k, VB200_F32, p.n), sizeof(float) * p.n);
    const void *c0 = L.getk("psy%d.compand", k, VB200_F32, VB200_COMPAND_LEVELS);
    const void *c1 = L.getk("psy%d.compand_hi", k, VB200_F32, VB200_COMPAND_LEVELS);
    if (c0) memcpy(p.compand, c0, sizeof(p.compand));
    if (c1) memcpy(p.compand_hi, c1, sizeof(p.compand_hi));
    if (!L.ok) break;
    /* static structure of the noise regression windows (bark_noise_hybridmp, psy.c:3540-3637): the
       first bin whose window no longer mirrors below 0 (b1) and the first bin that has no full
       window and re-uses the previous fit (b2), for the bark pass and the fixed-width pass */
    {
      const int *bark = (const int *)L.getk("psy%d.bark", k, VB200_I32, p.n);
      const int n = p.n, fixed = p.noisewindowfixed;
      for (int pass = 0; pass < 2; pass++) {
        int i = 0;
        for (; i < n; i++) {
          int lo, hi;
          if (pass == 0) { lo = bark[i] >> 16; hi = bark[i] & 0xffff; } else { hi = i + fixed / 2; lo = hi - fixed; }
          if (!(hi < n && lo < 0 && (pass == 1 || -lo < n))) break;
        }
        p.nf_b1[pass] = i;
        for (; i < n; i++) {
          int lo, hi;
          if (pass == 0) { lo = bark[i] >> 16; hi = bark[i] & 0xffff; } else { hi = i + fixed / 2; lo = hi - fixed; }
          if (!(hi < n && lo >= 0 && (pass == 1 || lo < n))) break;
        }
        p.nf_b2[pass] = i;
      }
    }
    /* static structure of max_seeds (psy.c:936-1070): which seed span feeds which run of bins */
    {
      const int *oct = (const int *)L.getk("psy%d.octave", k, VB200_I32, p.n);
      const int n = p.n;
      int4 *g = (int4 *)malloc(sizeof(int4) * (n + 1));
      int ng = 0, lin = 0, pos = oct[0] - p.firstoc - (p.eighth >> 1);
      bool bad = pos < 0;
      while (!bad && lin + 1 < n) {
        int p0 = pos, end = ((oct[lin] + oct[lin + 1]) >> 1) - p.firstoc, l0 = lin;
        while (pos + 1 <= end) pos++;
        end = pos + p.firstoc;
        for (; lin < n && oct[lin] <= end; lin++) {}
        if (lin == l0 || pos >= p.total_octave_lines) { bad = true; break; }
        g[ng++] = make_int4(p0, pos, l0, lin);
      }
      if (bad) { vb200_set_error("psy %d: octave table outside the built path", k); L.ok = false; free(g); break; }
      p.n_tone_groups = ng; p.tone_tail = lin;
      {
        int2 *runs = (int2 *)malloc(sizeof(int2) * n);
        int nr = 0;
        for (int i = 0; i < n;) { int e = i; while (e + 1 < n && oct[e + 1] == oct[i]) e++; runs[nr++] = make_int2(i, e); i = e + 1; }
        p.n_tone_runs = nr;
        p.tone_runs = (const int2 *)L.up(runs, sizeof(int2) * nr);
        free(runs);
      }
      p.tone_groups = (const int4 *)L.up(g, sizeof(int4) * (ng ? ng : 1));
      free(g);
    }
  }
  {
    const void *t;
    if ((t = L.get("aotuv.stn_compand", VB200_I32, VB200_COMPAND_LEVELS))) memcpy(h.stn_compand, t, sizeof(h.stn_compand));
    if ((t = L.get("aotuv.freq_bfn128", VB200_I32, 128))) memcpy(h.freq_bfn128, t, sizeof(h.freq_bfn128));
    if ((t = L.get("aotuv.freq_bfn256", VB200_I32, 256))) memcpy(h.freq_bfn256, t, sizeof(h.freq_bfn256));
    if ((t = L.get("aotuv.stereo_th", VB200_F64, 9))) memcpy(h.stereo_th, t, sizeof(h.stereo_th));
    if ((t = L.get("aotuv.stereo_thx", VB200_F64, 9))) memcpy(h.stereo_thx, t, sizeof(h.stereo_thx));
    if ((t = L.get("fromdB", VB200_F32, 256))) memcpy(h.fromdB, t, sizeof(h.fromdB));
  }

  for (int k = 0; k < h.floors && L.ok; k++) {
    DevFloor &fl = h.flr[k];
    const int *v = (const int *)L.getk("flr%d.i32", k, VB200_I32, VB200_FLR_I32_COUNT);
    if (!v) break;
    fl.type = v[VB200_FLR_TYPE];
    if (fl.type != 1) continue;
    const float *f = (const float *)L.getk("flr%d.f32", k, VB200_F32, VB200_FLRF_COUNT);
    if (!f) break;
    fl.posts = v[VB200_FLR_POSTS]; fl.n = v[VB200_FLR_LOOK_N]; fl.quant_q = v[VB200_FLR_QUANT_Q]; fl.mult = v[VB200_FLR_MULT]; fl.info_n = v[VB200_FLR_INFO_N];
    s->hflr[k].partitions = v[VB200_FLR_PARTITIONS];
    fl.maxover = f[VB200_FLRF_MAXOVER]; fl.maxunder = f[VB200_FLRF_MAXUNDER]; fl.maxerr = f[VB200_FLRF_MAXERR];
    fl.twofitweight = f[VB200_FLRF_TWOFITWEIGHT]; fl.twofitatten = f[VB200_FLRF_TWOFITATTEN];
    copy_i(fl.postlist, L.getk("flr%d.postlist", k, VB200_I32, VB200_MAX_POSTS), VB200_MAX_POSTS);
    copy_i(fl.sorted, L.getk("flr%d.sorted", k, VB200_I32, VB200_MAX_POSTS), VB200_MAX_POSTS);
    copy_i(fl.forward, L.getk("flr%d.forward", k, VB200_I32, VB200_MAX_POSTS), VB200_MAX_POSTS);
    copy_i(fl.reverse, L.getk("flr%d.reverse", k, VB200_I32, VB200_MAX_POSTS), VB200_MAX_POSTS);
    copy_i(fl.lonb, L.getk("flr%d.lonb", k, VB200_I32, 63), 63);
    copy_i(fl.hinb, L.getk("flr%d.hinb", k, VB200_I32, 63), 63);
    {
      int lev[VB200_MAX_POSTS] = {0}, nl = 0;
      for (int i = 2; i < fl.posts; i++) {
        const int a = lev[fl.lonb[i - 2]], b = lev[fl.hinb[i - 2]];
        lev[i] = 1 + (a > b ? a : b);
        if (lev[i] > nl) nl = lev[i];
      }
      fl.nlevels = nl;
      int n = 0;
      fl.lv_start[0] = 0; fl.lv_start[1] = 0;
      for (int l = 1; l <= nl; l++) {
        fl.lv_start[l] = n;
        for (int i = 2; i < fl.posts; i++) if (lev[i] == l) fl.lv_order[n++] = i;
      }
      fl.lv_start[nl + 1] = n;
    }
    copy_i(s->hflr[k].partclass, L.getk("flr%d.partclass", k, VB200_I32, 31), 31);
    copy_i(s->hflr[k].class_dim, L.getk("flr%d.class_dim", k, VB200_I32, 16), 16);
    copy_i(s->hflr[k].class_subs, L.getk("flr%d.class_subs", k, VB200_I32, 16), 16);
    copy_i(s->hflr[k].class_book, L.getk("flr%d.class_book", k, VB200_I32, 16), 16);
    copy_i(s->hflr[k].class_subbook, L.getk("flr%d.class_subbook", k, VB200_I32, 128), 128);
    fl.partitions = s->hflr[k].partitions;
    memcpy(fl.partclass, s->hflr[k].partclass, sizeof(fl.partclass));
    memcpy(fl.class_dim, s->hflr[k].class_dim, sizeof(fl.class_dim));
    memcpy(fl.class_subs, s->hflr[k].class_subs, sizeof(fl.class_subs));
    memcpy(fl.class_book, s->hflr[k].class_book, sizeof(fl.class_book));
    memcpy(fl.class_subbook, s->hflr[k].class_subbook, sizeof(fl.class_subbook));
  }

  for (int k = 0; k < h.residues && L.ok; k++) {
    DevResidue &r = h.res[k];
    const int *v = (const int *)L.getk("res%d.i32", k, VB200_I32, VB200_RES_I32_COUNT);
    if (!v) break;
    r.type = v[VB200_RES_TYPE]; r.begin = v[VB200_RES_BEGIN]; r.end = v[VB200_RES_END]; r.grouping = v[VB200_RES_GROUPING];
    r.partitions = v[VB200_RES_PARTITIONS]; r.partvals = v[VB200_RES_PARTVALS]; r.groupbook = v[VB200_RES_GROUPBOOK];
    r.stages = v[VB200_RES_STAGES]; r.phrase_dim = v[VB200_RES_PHRASE_DIM];
    copy_i(r.secondstages, L.getk("res%d.secondstages", k, VB200_I32, 64), 64);
    copy_i(r.classmetric1, L.getk("res%d.classmetric1", k, VB200_I32, 64), 64);
    copy_i(r.classmetric2, L.getk("res%d.classmetric2", k, VB200_I32, 64), 64);
    copy_i(r.partbooks, L.getk("res%d.partbooks", k, VB200_I32, 512), 512);
  }

  for (int k = 0; k < h.maps && L.ok; k++) {
    DevMap &m = h.map[k];
    const int *v = (const int *)L.getk("map%d.i32", k, VB200_I32, VB200_MAP_I32_COUNT);
    if (!v) break;
    m.submaps = v[VB200_MAP_SUBMAPS]; m.coupling_steps = v[VB200_MAP_COUPLING_STEPS];
    if (m.coupling_steps > VB200_MAX_COUPLING || m.submaps > VB200_MAX_SUBMAPS) { vb200_set_error("mapping %d outside the built path", k); L.ok = false; break; }
    copy_i(m.chmux, L.getk("map%d.chmux", k, VB200_I32, h.channels), h.channels);
    copy_i(m.floorsubmap, L.getk("map%d.floorsubmap", k, VB200_I32, 16), 16);
    copy_i(m.residuesubmap, L.getk("map%d.residuesubmap", k, VB200_I32, 16), 16);
    copy_i(m.mag, L.getk("map%d.coupling_mag", k, VB200_I32, m.coupling_steps ? m.coupling_steps : 1), m.coupling_steps);
    copy_i(m.ang, L.getk("map%d.coupling_ang", k, VB200_I32, m.coupling_steps ? m.coupling_steps : 1), m.coupling_steps);
  }

  for (int w = 0; w < 2 && L.ok; w++) {
    const DevMap &m = h.map[h.mode_mapping[w]];
    for (int c = 0; c < h.channels; c++) h.floor_of_ch[w][c] = m.floorsubmap[m.chmux[c]];
  }

  /* codebooks: lattice parameters + the compacted used-entry list with the value each used entry
     decodes to, enumerated the way the exhaustive fallback of res0.c:343-370 walks them */
  DevBook *hb = (DevBook *)calloc(h.books ? h.books : 1, sizeof(DevBook));
  for (int k = 0; k < h.books && L.ok; k++) {
    const int *v = (const int *)L.getk("book%d.i32", k, VB200_I32, VB200_BOOK_I32_COUNT);
    if (!v) break;
    HostBook &b = s->book[k];
    b.dim = v[VB200_BOOK_DIM]; b.entries = v[VB200_BOOK_ENTRIES]; b.quantvals = v[VB200_BOOK_QUANTVALS];
    b.minval = v[VB200_BOOK_MINVAL]; b.delta = v[VB200_BOOK_DELTA]; b.maptype = v[VB200_BOOK_MAPTYPE];
    b.len = (const uint8_t *)L.getk("book%d.len", k, VB200_U8, b.entries);
    b.code = (const uint32_t *)L.getk("book%d.code", k, VB200_U32, b.entries);
    if (!b.len || !b.code) break;
    DevBook &d = hb[k];
    d.dim = b.dim; d.entries = b.entries; d.quantvals = b.quantvals; d.minval = b.minval; d.delta = b.delta;
    d.delta_rcp = (b.delta > 1 && b.delta < 1024) ? (unsigned)((0x100000000ull + (unsigned long long)b.delta - 1) / (unsigned long long)b.delta) : 0u;
    int used = 0;
    for (int i = 0; i < b.entries; i++) if (b.len[i] > 0) used++;
    d.used = used;
    d.len = (const uint8_t *)L.up(b.len, b.entries);
    d.code = (const uint32_t *)L.up(b.code, sizeof(uint32_t) * b.entries);
    if (b.maptype == 1 && b.dim <= 8 && b.dim > 0) {
      int *ul = (int *)malloc(sizeof(int) * (used ? used : 1));
      int *uv = (int *)calloc((size_t)(used ? used : 1) * 8, sizeof(int));
      int e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int maxval = b.minval + b.delta * (b.quantvals - 1), u = 0;
      for (int i = 0; i < b.entries; i++) {
        if (b.len[i] > 0) { ul[u] = i; for (int j = 0; j < 8; j++) uv[u * 8 + j] = e[j]; u++; }
        /* advance the centred odometer: 0,-d,+d,-2d,... per digit, carry when the digit reached maxval */
        int j = 0;
        while (j < 8 && e[j] >= maxval) e[j++] = 0;
        if (j < 8) { if (e[j] >= 0) e[j] += b.delta; e[j] = -e[j]; }
      }
      d.used_list = (const int *)L.up(ul, sizeof(int) * (used ? used : 1));
      d.used_vals = (const int *)L.up(uv, sizeof(int) * 8 * (used ? used : 1));
      if (used > 0 && used < b.entries && used < 32768) {
        /* nearest used entry of every lattice point, first minimum in entry order (res0.c:349-361) */
        short *nn = (short *)malloc(sizeof(short) * b.entries);
        const int qv = b.quantvals, ze = qv >> 1;
        for (int i = 0; i < b.entries; i++) {
          int pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, rem = i;
          for (int j = 0; j < b.dim; j++) {             /* digit j of the index is the zig-zag code of component j */
            const int m = rem % qv; rem /= qv;
            const int v = (m & 1) ? ze - ((m + 1) >> 1) : ze + (m >> 1);
            pt[j] = v * b.delta + b.minval;
          }
          int best = -1, bestu = 0;
          for (int u2 = 0; u2 < used; u2++) {
            int dd = 0;
            for (int j = 0; j < b.dim; j++) { const int df = uv[u2 * 8 + j] - pt[j]; dd += df * df; }
            if (best == -1 || dd < best) { best = dd; bestu = u2; }
          }
          nn[i] = (short)bestu;
        }
        d.nearest = (const short *)L.up(nn, sizeof(short) * b.entries);
        free(nn);
      }
      free(ul); free(uv);
    }
  }
  h.book = (const DevBook *)L.up(hb, sizeof(DevBook) * (h.books ? h.books : 1));
  free(hb);

  for (int i = 0; i < 3; i++) {
    uint32_t c = 0; s->hdr[i] = (const unsigned char *)L.getk("hdr%d", i, VB200_U8, 1, &c); s->hdrlen[i] = c;
  }

  if (L.ok && device >= 0) {
    s->d = (DevSetup *)L.up(&h, sizeof(DevSetup));
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { L.ok = false; vb200_set_error("cudaStreamCreate failed"); }
    else s->stream = st;
  }
  if (!L.ok) { vb200_setup_destroy(s); return VB200_EINVAL; }
  *out = s;
  return VB200_OK;
}

extern "C" int vb200_setup_channels(const vb200_setup *s) { return s ? s->h.channels : VB200_EINVAL; }
extern "C" long vb200_setup_rate(const vb200_setup *s) { return s ? s->h.rate : VB200_EINVAL; }
extern "C" int vb200_setup_blocksize(const vb200_setup *s, int W) { return (s && (W == 0 || W == 1)) ? s->h.bs[W] : VB200_EINVAL; }
extern "C" int vb200_setup_managed(const vb200_setup *s) { return s ? s->h.managed : VB200_EINVAL; }
extern "C" int vb200_setup_posts(const vb200_setup *s, int W) {
  if (!s || (W != 0 && W != 1)) return VB200_EINVAL;
  const DevMap &m = s->h.map[s->h.mode_mapping[W]];
  return s->h.flr[m.floorsubmap[0]].posts;
}
extern "C" int vb200_setup_partvals(const vb200_setup *s, int W, int submap) {
  if (!s || (W != 0 && W != 1)) return VB200_EINVAL;
  const DevMap &m = s->h.map[s->h.mode_mapping[W]];
  if (submap < 0 || submap >= m.submaps) return VB200_EINVAL;
  const DevResidue &r = s->h.res[m.residuesubmap[submap]];
  return (r.end - r.begin) / r.grouping;
}
extern "C" int vb200_setup_header(const vb200_setup *s, int i, const unsigned char **data, long *bytes) {
  if (!s || i < 0 || i > 2 || !data || !bytes) return VB200_EINVAL;
  *data = s->hdr[i]; *bytes = s->hdrlen[i];
  return VB200_OK;
}
