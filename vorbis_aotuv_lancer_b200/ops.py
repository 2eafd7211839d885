"""Host-side mirror of the reference's per-block operator interface, batched.

Each function keeps the reference's name and argument meaning
(/root/reference/lib/psy.h:188-243, codec_internal.h:167-175, backends.h:87-101)
with a leading batch axis, and calls the CUDA path through the C ABI
(include/vorbis_b200.h).  numpy in, numpy out; nothing is computed on the host.
"""
import ctypes as C
import numpy as np

from . import blob as _blob
from .lib import L, check

MAX_POSTS = 65


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Setup:
    """vb200_setup: the flattened codec_setup_info + look objects, resident on `device`
    (device=-1: host-only handle for queries)."""

    def __init__(self, name_or_blob, device=0):
        raw = _blob.load(name_or_blob) if isinstance(name_or_blob, str) else bytes(name_or_blob)
        self.raw = raw
        self.tables = _blob.parse(raw)
        h = C.c_void_p()
        check(L.vb200_setup_create(raw, C.c_size_t(len(raw)), int(device), C.byref(h)), "vb200_setup_create")
        self.h = h
        self.channels = L.vb200_setup_channels(h)
        self.rate = L.vb200_setup_rate(h)
        self.blocksizes = (L.vb200_setup_blocksize(h, 0), L.vb200_setup_blocksize(h, 1))
        self.managed = L.vb200_setup_managed(h)

    def set_exact_transform(self, on=True):
        """K1x instead of K1 for the launches that follow: the reference's own float operations in mdct_forward /
        drft_forward (bit-identical spectra and packets) instead of K1's own formulation (within 1e-5)"""
        check(L.vb200_setup_set_exact_transform(self.h, int(bool(on))), "vb200_setup_set_exact_transform")
        return self

    def close(self):
        if self.h:
            L.vb200_setup_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def n2(self, W):
        return self.blocksizes[W] // 2

    def partition(self, W):
        for k in range(int(self.tables["cfg.i32"][8])):
            t = self.tables["psy%d.i32" % k]
            if t[1] == W:
                return int(t[19]) if t[17] else 16
        return 16

    def residue(self, resnum):
        t = self.tables["res%d.i32" % resnum]
        return dict(type=int(t[0]), begin=int(t[1]), end=int(t[2]), grouping=int(t[3]),
                    partvals=(int(t[2]) - int(t[1])) // int(t[3]))

    def header(self, i):
        d, n = C.POINTER(C.c_ubyte)(), C.c_long()
        check(L.vb200_setup_header(self.h, i, C.byref(d), C.byref(n)))
        return bytes(bytearray(d[:n.value]))


# ------------------------------------------------------------------ K1
def forward_transform(s, W, pcm, lW, nW, block_mode, lW_block_mode):
    """_postnoise_detection + _vorbis_apply_window + mdct_forward + drft_forward + dB loops
    (mapping0.c:822-888, 933-951).  pcm [items, n] -> dict(gmdct, logfft, logmdct, local_ampmax, poste)."""
    pcm = _f(pcm)
    items, n = pcm.shape
    assert n == s.blocksizes[W]
    N2 = n // 2
    out = dict(gmdct=np.empty((items, N2), np.float32), logfft=np.empty((items, N2), np.float32),
               logmdct=np.empty((items, N2), np.float32), local_ampmax=np.empty(items, np.float32),
               poste=np.empty(items, np.float32))
    lW, nW, bm, lbm = _i(lW), _i(nW), _i(block_mode), _i(lW_block_mode)
    check(L.vb200_stage_transform(s.h, W, items, _p(pcm), _p(lW), _p(nW), _p(bm), _p(lbm),
                                  _p(out["gmdct"]), _p(out["logfft"]), _p(out["logmdct"]),
                                  _p(out["local_ampmax"]), _p(out["poste"])), "vb200_stage_transform")
    return out


# ------------------------------------------------------------------ K2
def _vp_noisemask(s, W, psy, noise_compand_level, logmdct, lastmdct, poste, block_mode):
    """psy.c:3770-4074, batched: -> (noise, epeak, npeak)."""
    logmdct, lastmdct = _f(logmdct), _f(lastmdct)
    items, N2 = logmdct.shape
    nps = N2 // s.partition(W)
    noise, epeak, npeak = np.empty((items, N2), np.float32), np.empty((items, N2), np.float32), np.empty((items, nps), np.float32)
    psy, bm, lc, po = _i(psy), _i(block_mode), _f(noise_compand_level), _f(poste)
    check(L.vb200_stage_noisemask(s.h, W, items, _p(psy), _p(bm), _p(lc), _p(po), _p(logmdct), _p(lastmdct),
                                  _p(noise), _p(epeak), _p(npeak)), "vb200_stage_noisemask")
    return noise, epeak, npeak


# ------------------------------------------------------------------ K3
def _vp_tonemask(s, W, psy, logfft, global_specmax, local_specmax):
    """psy.c:4076-4142, batched: -> tone."""
    logfft = _f(logfft)
    items, N2 = logfft.shape
    tone = np.empty((items, N2), np.float32)
    psy, g, l = _i(psy), _f(global_specmax), _f(local_specmax)
    check(L.vb200_stage_tonemask(s.h, W, items, _p(psy), _p(logfft), _p(g), _p(l), _p(tone)), "vb200_stage_tonemask")
    return tone


# ------------------------------------------------------------------ K4
def _vp_offset_and_mix(s, W, psy, noise, tone, params, low_compand, mdct, logmdct, lastmdct, tempmdct, npeak):
    """psy.c:4274-4502, batched.  params [items,8] = offset_select, bit_managed, end_block, block_mode,
    nW_modenumber, lW_block_mode, lW_no, impadnum.  -> (logmask, mdct, lastmdct, tempmdct, npeak)."""
    noise, tone, logmdct = _f(noise), _f(tone), _f(logmdct)
    mdct, last, temp, npk = np.array(mdct, np.float32), np.array(lastmdct, np.float32), np.array(tempmdct, np.float32), np.array(npeak, np.float32)
    items, N2 = noise.shape
    assert last.shape == (items, 2048) and temp.shape == (items, 256)
    mask = np.empty((items, N2), np.float32)
    psy, par, lc = _i(psy), _i(params), _f(low_compand)
    check(L.vb200_stage_offset_and_mix(s.h, W, items, _p(psy), _p(par), _p(lc), _p(noise), _p(tone), _p(logmdct),
                                       _p(mdct), _p(last), _p(temp), _p(npk), _p(mask)), "vb200_stage_offset_and_mix")
    return mask, mdct, last, temp, npk


# ------------------------------------------------------------------ K5
def floor1_fit(s, W, floor_index, logmdct, logmask):
    """floor1.c:597-750, batched: -> (posts [items,65], fitted [items]); fitted==0 is the reference's NULL."""
    logmdct, logmask = _f(logmdct), _f(logmask)
    items = logmdct.shape[0]
    posts, fitted = np.zeros((items, MAX_POSTS), np.int32), np.zeros(items, np.int32)
    check(L.vb200_stage_floor1_fit(s.h, W, items, floor_index, _p(logmdct), _p(logmask), _p(posts), _p(fitted)),
          "vb200_stage_floor1_fit")
    return posts, fitted


def floor1_encode(s, W, floor_index, posts, fitted):
    """floor1.c:774-974 minus bit emission, batched: -> (posts mutated, vals, ilogmask, nonzero)."""
    posts, fitted = np.array(posts, np.int32), _i(fitted)
    items = posts.shape[0]
    N2 = s.n2(W)
    vals, ilog, nz = np.zeros((items, MAX_POSTS), np.int32), np.zeros((items, N2), np.int32), np.zeros(items, np.int32)
    check(L.vb200_stage_floor1_encode(s.h, W, items, floor_index, _p(posts), _p(fitted), _p(vals), _p(ilog), _p(nz)),
          "vb200_stage_floor1_encode")
    return posts, vals, ilog, nz


# ------------------------------------------------------------------ K6
def _vp_couple_quantize_normalize(s, W, psy, blobno, mdct, enpeak, nepeak, iwork, nonzero):
    """psy.c:4858-5142, batched over blocks: mdct/enpeak [blocks,ch,N2], nepeak [blocks,ch,N2/partition],
    iwork [blocks,ch,N2] (ilogmask in), nonzero [blocks,ch] -> (iwork, nonzero, nepeak)."""
    mdct, enpeak = _f(mdct), _f(enpeak)
    npk, iw, nz = np.array(nepeak, np.float32), np.array(iwork, np.int32), np.array(nonzero, np.int32)
    blocks = mdct.shape[0]
    psy = _i(psy)
    check(L.vb200_stage_couple(s.h, W, blocks, _p(psy), int(blobno), _p(mdct), _p(enpeak), _p(npk), _p(iw), _p(nz)),
          "vb200_stage_couple")
    return iw, nz, npk


# ------------------------------------------------------------------ K7
def residue_class(s, W, resnum, rows, nonzero):
    """->class of the residue backend (res0.c:735-768), batched: rows [blocks,ch,N2], nonzero [blocks,ch]
    -> (partword [blocks,ch,partvals], classified [blocks] = number of valid partword rows, 0 == NULL)."""
    rows, nz = _i(rows), _i(nonzero)
    blocks, ch, _ = rows.shape
    pv = s.residue(resnum)["partvals"]
    pw, cl = np.zeros((blocks, ch, pv), np.int32), np.zeros(blocks, np.int32)
    check(L.vb200_stage_residue_class(s.h, W, blocks, resnum, ch, _p(rows), _p(nz), _p(pw), _p(cl)), "vb200_stage_residue_class")
    return pw, cl


def residue_forward(s, W, resnum, rows, nonzero, partword, max_syms=16384):
    """->forward of the residue backend (res0.c:715-799): VQ search, batched.
    -> (syms [blocks] list of [n,2] (book,entry) in emission order, rows after the cascade)."""
    rows, nz, pw = np.array(rows, np.int32), _i(nonzero), _i(partword)
    blocks, ch, _ = rows.shape
    syms, ns = np.zeros((blocks, max_syms, 2), np.int32), np.zeros(blocks, np.int32)
    check(L.vb200_stage_residue_forward(s.h, W, blocks, resnum, ch, _p(rows), _p(nz), _p(pw), _p(syms), max_syms, _p(ns)),
          "vb200_stage_residue_forward")
    return [syms[b, :ns[b]].copy() for b in range(blocks)], rows


# ------------------------------------------------------------------ whole path
class Block(C.Structure):
    """vb200_block (include/vorbis_b200.h): the vorbis_block fields mapping0_forward reads."""
    _fields_ = [("stream", C.c_int32), ("W", C.c_int32), ("lW", C.c_int32), ("nW", C.c_int32),
                ("blocktype", C.c_int32), ("eofflag", C.c_int32), ("granulepos", C.c_int64),
                ("pcm_offset", C.c_int64), ("pcm_stride", C.c_int64)]


class Stream:
    """vb200_stream: per-stream aoTuV history, ampmax feedback and bitrate-manager state."""

    def __init__(self, setup):
        self.setup = setup
        h = C.c_void_p()
        check(L.vb200_stream_create(setup.h, C.byref(h)), "vb200_stream_create")
        self.h = h

    def close(self):
        if self.h:
            L.vb200_stream_destroy(self.h)
            self.h = None


class Batch:
    """vb200_batch: `vorbis_analysis` + `vorbis_bitrate_addblock` + `vorbis_bitrate_flushpacket`
    for many blocks of many streams per call."""

    def __init__(self, setup, max_blocks):
        self.setup = setup
        h = C.c_void_p()
        check(L.vb200_batch_create(setup.h, int(max_blocks), C.byref(h)), "vb200_batch_create")
        self.h = h
        self.n = 0

    def close(self):
        if self.h:
            L.vb200_batch_destroy(self.h)
            self.h = None

    def run(self, streams, blocks, pcm, pcm_on_device=False, pcm_elems=None):
        """blocks: ctypes array of Block (stream order within each stream); pcm: float32 numpy array
        (host) or an integer device address with pcm_elems."""
        sh = (C.c_void_p * len(streams))(*[s.h for s in streams])
        if pcm_on_device:
            ptr, n = C.c_void_p(int(pcm)), int(pcm_elems)
        else:
            pcm = _f(pcm)
            ptr, n = _p(pcm), pcm.size
        check(L.vb200_batch_run(self.h, sh, len(streams), blocks, len(blocks), ptr, C.c_size_t(n), int(pcm_on_device)),
              "vb200_batch_run")
        self.n = len(blocks)

    def submit(self, streams, blocks, pcm, pcm_elems):
        """Asynchronous half of run() for device-resident PCM (vb200_batch_submit)."""
        sh = (C.c_void_p * len(streams))(*[s.h for s in streams])
        check(L.vb200_batch_submit(self.h, sh, len(streams), blocks, len(blocks), C.c_void_p(int(pcm)), C.c_size_t(int(pcm_elems)), 1),
              "vb200_batch_submit")
        self.n = len(blocks)

    def collect(self):
        check(L.vb200_batch_collect(self.h), "vb200_batch_collect")

    def packets(self):
        """[(bytes, granulepos, eos, stream)] of the last run, in block order."""
        out = []
        d, nb, gp, eos, st = C.POINTER(C.c_ubyte)(), C.c_long(), C.c_int64(), C.c_int(), C.c_int()
        for i in range(self.n):
            check(L.vb200_batch_packet(self.h, i, C.byref(d), C.byref(nb), C.byref(gp), C.byref(eos), C.byref(st)))
            out.append((C.string_at(d, nb.value), gp.value, eos.value, st.value))
        return out

    def packet_bytes(self):
        nb = C.c_long()
        tot = 0
        for i in range(self.n):
            L.vb200_batch_packet(self.h, i, None, C.byref(nb), None, None, None)
            tot += nb.value
        return tot

    def kernel_ms(self):
        return float(L.vb200_batch_kernel_ms(self.h))

    def kernel_times(self):
        names = (C.c_char_p * 24)()
        ms = (C.c_float * 24)()
        n = L.vb200_batch_kernel_times(self.h, names, ms, 24)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def get(self, what, block, ch, count, dtype=np.float32):
        out = np.zeros(count, dtype)
        n = check(L.vb200_batch_get(self.h, int(what), int(block), int(ch), _p(out), int(count)), "vb200_batch_get")
        return out[:n]


def blocks_array(descs):
    """descs: iterable of dicts/tuples (stream, W, lW, nW, blocktype, eofflag, granulepos, pcm_offset, pcm_stride)."""
    arr = (Block * len(descs))()
    for i, d in enumerate(descs):
        (arr[i].stream, arr[i].W, arr[i].lW, arr[i].nW, arr[i].blocktype, arr[i].eofflag, arr[i].granulepos,
         arr[i].pcm_offset, arr[i].pcm_stride) = [int(x) for x in d]
    return arr
