"""Generate vorbis_aotuv_lancer_b200/setups/*.vb2s.z from the reference.

Runs the reference's own setup (vorbis_encode_init[_vbr] + vorbis_analysis_init,
vorbisenc.c:722-905, block.c:306-330) inside oracle/_ref and serialises the
resulting tables with oracle/ref_export.c.  Needs /root/reference (this
container only); the blobs are committed so the GPU box does not.
"""
import os
import re
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402
from vorbis_aotuv_lancer_b200 import blob as vblob  # noqa: E402


def _c_array(src, name):
    """Numbers of the initialiser of the file-scope array `name` in C source text."""
    i = src.index(name)
    i = src.index("{", i)
    body = re.sub(r"/\*.*?\*/|//[^\n]*", "", src[i + 1:src.index("};", i)], flags=re.S)
    return [float(x.rstrip("Ff")) for x in re.findall(r"[-+]?[0-9][0-9.]*(?:e[-+]?[0-9]+)?[Ff]?", body)]


def _append(raw, name, dtype, arr):
    """Append one section to a blob (include/vb200_blob.h)."""
    code = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2, np.dtype(np.float64): 3}[np.dtype(dtype)]
    a = np.ascontiguousarray(arr, dtype=dtype)
    payload = a.tobytes()
    payload += b"\0" * ((8 - len(payload) % 8) % 8)
    sec = name.encode().ljust(24, b"\0") + struct.pack("<II", code, a.size) + payload
    magic, version, nsec, total = struct.unpack_from("<IIII", raw, 0)
    raw = raw + sec
    return struct.pack("<IIII", magic, version, nsec + 1, len(raw)) + raw[16:]


def constant_tables(reference="/root/reference"):
    """File-scope constant tables the path needs that are not reachable from the look objects:
    the Vorbis I floor1 inverse-dB table (floor1.c:301-366 == psy.c:4517-4582, normative in
    doc/Vorbis_I_spec) and the aoTuV tuning tables of psy.c:41-89."""
    psy = open(os.path.join(reference, "lib/psy.c")).read()
    flr = open(os.path.join(reference, "lib/floor1.c")).read()
    return [
        ("fromdB", np.float32, _c_array(flr, "FLOOR1_fromdB_LOOKUP[256]")),
        ("aotuv.stereo_th", np.float64, _c_array(psy, "stereo_threshholds[]")),
        ("aotuv.stereo_thx", np.float64, _c_array(psy, "stereo_threshholds_X[]")),
        ("aotuv.freq_bfn128", np.int32, _c_array(psy, "freq_bfn128[128]")),
        ("aotuv.freq_bfn256", np.int32, _c_array(psy, "freq_bfn256[256]")),
        ("aotuv.stn_compand", np.int32, _c_array(psy, "stn_compand[]")),
    ]


def main():
    reflib.build()
    consts = constant_tables()
    assert [len(c[2]) for c in consts] == [256, 9, 9, 128, 256, 40], [len(c[2]) for c in consts]
    R = reflib.RefLib("c")
    os.makedirs(vblob.SETUP_DIR, exist_ok=True)
    for name, (ch, rate, q, abr) in vblob.CATALOGUE.items():
        h = R.open(ch, rate, q, abr)
        raw = R.export_setup(h)
        R.close(h)
        for cname, dt, vals in consts:
            raw = _append(raw, cname, dt, vals)
        z = zlib.compress(raw, 9)
        with open(os.path.join(vblob.SETUP_DIR, name + ".vb2s.z"), "wb") as f:
            f.write(z)
        print("%-16s raw %7d B  packed %7d B  sections %d" % (name, len(raw), len(z), len(vblob.parse(raw))))


if __name__ == "__main__":
    main()
