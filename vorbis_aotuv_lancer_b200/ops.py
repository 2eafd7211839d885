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

    def partition(self, psy):
        t = self.tables["psy%d.i32" % psy]
        return int(t[19]) if t[17] else 16

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
