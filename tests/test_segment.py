"""Host segmentation (csrc/segment.cu) against the reference's own vorbis_analysis_blockout:
same block sequence (W, lW, nW, blocktype, granulepos, eof) and the same block PCM (incl. the
LPC-extrapolated head and tail).  CPU only -- the segmenter is host code."""
import numpy as np
import pytest

from tests import refdata
from vorbis_aotuv_lancer_b200 import ops, segment

pytestmark = pytest.mark.skipif(not refdata.have_ref("c"), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("config,seconds", [("q4_44k_2ch", 6.0), ("q8_48k_6ch", 2.0), ("q4_44k_1ch", 0.3), ("q10_48k_2ch", 3.0)])
def test_blocks_match_reference(config, seconds):
    cap = refdata.capture(config, seconds)
    s = ops.Setup(config, device=-1)
    blocks, timeline = segment.segment_stream(s, cap["pcm"])
    ref = [cap["blocks"][q] for q in sorted(cap["blocks"]) if "TAP_BLOCK" in cap["blocks"][q]]
    assert len(blocks) == len(ref)
    nbad_pcm = 0
    for i, (b, r) in enumerate(zip(blocks, ref)):
        t = r["TAP_BLOCK"][0]["data"]
        gp = (int(t[7]) << 32) | (int(t[6]) & 0xFFFFFFFF)
        assert (b.W, b.lW, b.nW, b.blocktype, b.eofflag, b.granulepos) == (t[0], t[1], t[2], t[3], t[5], gp), i
        n = int(t[4])
        for c in range(s.channels):
            mine = timeline[b.pcm_offset + c * b.pcm_stride: b.pcm_offset + c * b.pcm_stride + n]
            refp = r["TAP_PCM"][c]["data"]
            if not np.array_equal(mine, refp):
                # only the extrapolated head / tail may differ, and only by float rounding of the LPC
                assert np.allclose(mine, refp, rtol=1e-4, atol=1e-5), (i, c, float(np.abs(mine - refp).max()))
                nbad_pcm += 1
    print("%s: %d blocks, %d ch-blocks with rounding-level PCM differences" % (config, len(blocks), nbad_pcm))
    s.close()


def _ref_blocklist(config, pcm):
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob
    ch, rate, q, abr = vblob.CATALOGUE[config]
    R = reflib.RefLib("c")
    h = R.open(ch, rate, q, abr)
    R.write(h, pcm)
    bl, pk = R.blocks(h), R.packets(h)
    R.close(h)
    return bl, pk


@pytest.mark.parametrize("nsamples", [1, 63, 64, 1000, 2047, 2048, 2049, 5000, 9999])
def test_short_and_odd_lengths(nsamples):
    """ragged ends: streams shorter than one block, not a multiple of anything, exactly on block boundaries"""
    config = "q4_44k_2ch"
    pcm = refdata.synth(2, 44100, 1.0, seed=3)[:, :nsamples].copy()
    ref, pk = _ref_blocklist(config, pcm)
    s = ops.Setup(config, device=-1)
    blocks, _ = segment.segment_stream(s, pcm)
    assert len(blocks) == len(ref) == len(pk)
    for b, r, p in zip(blocks, ref, pk):
        assert (b.W, b.lW, b.nW, b.blocktype) == tuple(int(v) for v in r[:4])
        assert b.granulepos == p[1]
    assert blocks[-1].eofflag == 1 and blocks[-1].granulepos == nsamples
    s.close()


@pytest.mark.parametrize("kind", ["silence", "fullscale", "impulses"])
def test_degenerate_signals(kind):
    """digital silence (no transient ever), a clipped square wave, a click train (a transient in every block)"""
    config = "q4_44k_2ch"
    n = 44100
    if kind == "silence":
        pcm = np.zeros((2, n), np.float32)
    elif kind == "fullscale":
        pcm = np.where((np.arange(n) // 50) % 2 == 0, 0.99, -0.99).astype(np.float32)[None, :].repeat(2, 0).copy()
    else:
        pcm = np.zeros((2, n), np.float32)
        pcm[:, ::700] = 0.9
    ref, pk = _ref_blocklist(config, pcm)
    s = ops.Setup(config, device=-1)
    blocks, _ = segment.segment_stream(s, pcm)
    assert [(b.W, b.lW, b.nW, b.blocktype) for b in blocks] == [tuple(int(v) for v in r[:4]) for r in ref]
    assert [b.granulepos for b in blocks] == [p[1] for p in pk]
    s.close()
