"""Test helper: run the reference build (oracle/_ref) on the synthetic PCM and hand
back its per-stage inputs/outputs (A/B injection data).  Test infrastructure only."""
import ctypes as C
import functools

import numpy as np

from oracle import reflib
from vorbis_aotuv_lancer_b200 import blob as vblob
from vorbis_aotuv_lancer_b200.lib import L


def synth(channels, rate, seconds, seed=0):
    n = int(rate * seconds)
    a = np.zeros((channels, n), np.float32)
    L.vb200_synth_pcm(C.c_uint64(seed), channels, C.c_long(rate), C.c_long(n), a.ctypes.data_as(C.c_void_p), C.c_long(n))
    return a


def have_ref(flavour="c"):
    return reflib.available(flavour)


@functools.lru_cache(maxsize=4)
def capture(config, seconds=3.0, lo=0, hi=1 << 30, flavour="c", seed=0):
    """Encode `seconds` of synthetic PCM with the reference and return
    dict(pcm, blocks={seq:{tag:[records]}}, packets, headers, blocklist)."""
    ch, rate, q, abr = vblob.CATALOGUE[config]
    R = reflib.RefLib(flavour)
    pcm = synth(ch, rate, seconds, seed)
    h = R.open(ch, rate, q, abr)
    R.tap_all(lo, hi)
    R.write(h, pcm)
    out = dict(pcm=pcm, blocks=reflib.group(R.tap_records()), packets=R.packets(h), headers=R.headers(h),
               blocklist=R.blocks(h))
    R.tap_config([], 0, 0)
    R.close(h)
    return out


def by_size(cap, W):
    """seqs of the captured blocks with size flag W, in order."""
    return [s for s in sorted(cap["blocks"]) if "TAP_BLOCK" in cap["blocks"][s] and cap["blocks"][s]["TAP_BLOCK"][0]["data"][0] == W]


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def load_golden(config):
    """The committed reference vectors of tests/golden (tools/make_golden.py) in capture() layout."""
    import os
    z = np.load(os.path.join(GOLDEN_DIR, config + ".npz"))
    blocks = {}
    for key in z.files:
        if key == "seqs" or key.endswith("/a"):
            continue
        seq, tag, k = key.split("/")
        a = z[key + "/a"]
        blocks.setdefault(int(seq), {}).setdefault(tag, {})[int(k)] = dict(tag=tag, seq=int(seq), a0=int(a[0]), a1=int(a[1]), a2=int(a[2]), a3=int(a[3]), data=z[key])
    for seq in blocks:
        for tag in blocks[seq]:
            d = blocks[seq][tag]
            blocks[seq][tag] = [d[i] for i in sorted(d)]
    return dict(blocks=blocks)
