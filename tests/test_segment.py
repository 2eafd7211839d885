"""Host segmentation (csrc/segment.cu) against the reference's own vorbis_analysis_blockout:
same block sequence (W, lW, nW, blocktype, granulepos, eof) and the same block PCM (incl. the
LPC-extrapolated head and tail).  CPU only -- the segmenter is host code."""
import numpy as np
import pytest

from tests import refdata
from vorbis_aotuv_lancer_b200 import ops, segment

pytestmark = pytest.mark.skipif(not refdata.have_ref("c"), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("config,seconds", [("q4_44k_2ch", 6.0), ("q8_48k_6ch", 2.0), ("q4_44k_1ch", 0.3), ("q10_48k_2ch", 3.0),
                                            ("q4_32k_2ch", 3.0), ("q0_44k_2ch", 3.0), ("q5_48k_3ch", 2.0)])
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


def test_detector_mdct_is_the_references_own_bit_for_bit():
    """The block-size detector's transform (csrc/ve_mdct128.cuh: window + 128-point forward MDCT with the reference's own
    float operations; host here, the device detector K0 runs the same function) against the reference's envelope state
    itself (ref_envelope_mdct: the handle's mdct_win and mdct_lookup, envelope.c:139-186), and the oracle's
    operation-exact restatement of mdct_forward at n = 128 against the reference's mdct_forward -- all bit for bit."""
    import ctypes as C
    from oracle import port, reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob
    from vorbis_aotuv_lancer_b200.lib import L
    fp = C.POINTER(C.c_float)
    config = "q4_44k_2ch"
    ch, rate, q, abr = vblob.CATALOGUE[config]
    R = reflib.RefLib("c")
    R.lib.ref_envelope_mdct.argtypes = [C.c_void_p, fp, fp]
    R.lib.ref_mdct_forward.argtypes = [C.c_int, fp, fp]
    h = R.open(ch, rate, q, abr)
    OL = C.CDLL(port.LIBPATH)
    OL.orc_mdct_forward_exact.argtypes = [fp, fp, C.c_int]
    OL.orc_mdct_forward_exact.restype = None
    rng = np.random.default_rng(7)
    sig = refdata.synth(ch, rate, 1.0, seed=11)[0]
    cases = [sig[k * 64:k * 64 + 128] for k in range(0, 600, 7)]
    cases += [rng.uniform(-1, 1, 128).astype(np.float32) * np.float32(10.0 ** rng.uniform(-6, 0)) for _ in range(200)]
    cases += [np.zeros(128, np.float32), np.full(128, 0.99, np.float32), -np.full(128, 0.99, np.float32)]
    d = np.zeros(128, np.float32); d[77] = 1.0; cases.append(d)
    for x in cases:
        x = np.ascontiguousarray(x, np.float32)
        ref = np.empty(64, np.float32); mine = np.empty(64, np.float32)
        assert R.lib.ref_envelope_mdct(h, x.ctypes.data_as(fp), ref.ctypes.data_as(fp)) == 128
        assert L.vb200_stage_detector_mdct(x.ctypes.data_as(fp), mine.ctypes.data_as(fp)) == 0
        assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
        # the oracle's restatement of mdct_forward, n = 128, on the windowed samples
        win = (np.sin(np.arange(128) / 127.0 * np.pi).astype(np.float32) ** 2).astype(np.float32)
        w = (x * win).astype(np.float32)
        ref2 = np.empty(64, np.float32)
        R.lib.ref_mdct_forward(128, w.ctypes.data_as(fp), ref2.ctypes.data_as(fp))
        o = np.empty(64, np.float32)
        OL.orc_mdct_forward_exact(w.ctypes.data_as(fp), o.ctypes.data_as(fp), 128)
        assert np.array_equal(o.view(np.uint32), ref2.view(np.uint32))
    R.close(h)
