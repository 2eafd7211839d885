"""Reader for vb200 setup blobs (include/vb200_blob.h) and the catalogue of the
blobs shipped in setups/ (generated from the reference by tools/make_setups.py)."""
import os
import struct
import zlib
import numpy as np

MAGIC = 0x42323056
_DT = {0: np.float32, 1: np.int32, 2: np.uint8, 3: np.float64, 4: np.uint32}
SETUP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "setups")

# name -> (channels, rate, quality, managed nominal bitrate)
CATALOGUE = {
    "q4_44k_2ch": (2, 44100, 0.4, 0),      # BASELINE configs[0], [1]
    "q5_44k_2ch": (2, 44100, 0.5, 0),      # configs[2]
    "q8_48k_6ch": (6, 48000, 0.8, 0),      # configs[3]: 5.1, residue type 2 + LFE residue type 1
    "q10_48k_2ch": (2, 48000, 1.0, 0),     # configs[4a]
    "abr128_44k_2ch": (2, 44100, 0.4, 128000),  # configs[4b]: managed, 15 packet blobs
    "q2_44k_2ch": (2, 44100, 0.2, 0),      # aoTuV M4/M5/M8 + noise normalisation live
    "q4_48k_6ch": (6, 48000, 0.4, 0),      # 5.1 with 4 coupling steps
    "q4_44k_1ch": (1, 44100, 0.4, 0),      # mono, no coupling
    # round 2: the rest of the built scope's corners (parity cases, tests/test_gpu_pipeline.py)
    "q4_32k_2ch": (2, 32000, 0.4, 0),      # the third sampling-rate family (setup_32.h)
    "q0_44k_2ch": (2, 44100, 0.0, 0),      # lowest quality with block sizes 256/2048: strongest point stereo, all aoTuV low-rate paths
    "q7_44k_2ch": (2, 44100, 0.7, 0),      # q6~: the tone-attenuation branch of set_m3p (psy.c:4183), lossless coupling region
    "abr256_48k_6ch": (6, 48000, 0.4, 256000),  # managed 5.1: 15 blobs x two submaps (residue type 2 + LFE type 1)
    "q5_48k_3ch": (3, 48000, 0.5, 0),      # three uncoupled channels
}


def parse(blob):
    magic, version, nsec, total = struct.unpack_from("<IIII", blob, 0)
    if magic != MAGIC or total != len(blob):
        raise ValueError("not a vb200 setup blob")
    pos, out = 16, {}
    for _ in range(nsec):
        name = blob[pos:pos + 24].split(b"\0")[0].decode()
        dtype, count = struct.unpack_from("<II", blob, pos + 24)
        pos += 32
        dt = np.dtype(_DT[dtype])
        nbytes = count * dt.itemsize
        out[name] = np.frombuffer(blob, dtype=dt, count=count, offset=pos).copy()
        pos += (nbytes + 7) & ~7
    return out


def load(name):
    """Raw bytes of a shipped setup blob."""
    with open(os.path.join(SETUP_DIR, name + ".vb2s.z"), "rb") as f:
        return zlib.decompress(f.read())
