"""scratch: does the host/GPU detector formulation (double DCT-IV) ever decide a block differently from the reference's
float MDCT?  CPU only.  usage: python tools/scratch/seg_stress.py [nstreams] [seconds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import refdata
from tests.test_segment import _ref_blocklist
from vorbis_aotuv_lancer_b200 import ops, segment

nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
config = "q4_44k_2ch"
rate, ch = 44100, 2
s = ops.Setup(config, device=-1)
rng = np.random.default_rng(1234)
n = int(rate * seconds)
t = np.arange(n) / rate
bad_streams = 0; total_blocks = 0; bad_blocks = 0
for i in range(nstreams):
    kind = i % 5
    if kind == 0:
        pcm = refdata.synth(ch, rate, seconds, seed=1000 + i)
    elif kind == 1:      # noise with a piecewise-constant random gain
        g = np.repeat(10 ** rng.uniform(-3, 0, n // 2000 + 1), 2000)[:n]
        pcm = (rng.uniform(-1, 1, (ch, n)) * g * 0.5).astype(np.float32)
    elif kind == 2:      # tones with random AM
        pcm = np.zeros((ch, n))
        for _ in range(6):
            f = rng.uniform(60, 12000); am = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.5, 25) * t + rng.uniform(0, 6))
            pcm += 0.15 * am * np.sin(2 * np.pi * f * t + rng.uniform(0, 6))
        pcm = pcm.astype(np.float32)
    elif kind == 3:      # very quiet material around the detector's minimum energy
        pcm = (rng.uniform(-1, 1, (ch, n)) * 10 ** rng.uniform(-5, -3)).astype(np.float32)
        k = rng.integers(0, n - 5000, 20)
        for p in k: pcm[:, p:p + 400] *= rng.uniform(1, 30)
    else:                # decaying noise bursts at random times and levels
        pcm = (rng.uniform(-1, 1, (ch, n)) * 0.002).astype(np.float32)
        for p in rng.integers(0, n - 9000, 40):
            L = rng.integers(200, 8000)
            env = np.exp(-np.arange(L) / rng.uniform(50, 2000)) * 10 ** rng.uniform(-2.5, -0.1)
            pcm[:, p:p + L] += (rng.uniform(-1, 1, (ch, L)) * env).astype(np.float32)
        pcm = np.clip(pcm, -0.99, 0.99)
    pcm = np.ascontiguousarray(pcm, np.float32)
    ref, pk = _ref_blocklist(config, pcm)
    blocks, _ = segment.segment_stream(s, pcm)
    mine = [(b.W, b.lW, b.nW, b.blocktype) for b in blocks]
    refl = [tuple(int(v) for v in r[:4]) for r in ref]
    total_blocks += len(refl)
    if mine != refl:
        bad_streams += 1
        first = next((j for j, (a, b) in enumerate(zip(mine, refl)) if a != b), min(len(mine), len(refl)))
        nb = sum(1 for a, b in zip(mine, refl) if a != b) + abs(len(mine) - len(refl))
        bad_blocks += nb
        print("stream %d kind %d: DIFFERENT block list (len %d vs %d), first at %d: %s vs %s" % (i, kind, len(mine), len(refl), first, mine[first:first + 3], refl[first:first + 3]), flush=True)
print("streams %d, blocks %d, streams that differ %d (blocks %d)" % (nstreams, total_blocks, bad_streams, bad_blocks))
