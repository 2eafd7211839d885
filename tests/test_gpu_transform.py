"""K1 parity: window + MDCT + FFT/dB vs the reference's own outputs on the same PCM."""
import numpy as np
import pytest

from tests import refdata

pytestmark = pytest.mark.gpu


def todB_np(x):
    i = np.asarray(x, np.float32).view(np.uint32) & 0x7FFFFFFF
    return i.astype(np.float32) * np.float32(7.17711438e-7) + np.float32(-764.6161886)


@pytest.mark.parametrize("config", ["q4_44k_2ch", "q8_48k_6ch"])
@pytest.mark.parametrize("W", [1, 0])
def test_transform_matches_reference(config, W):
    """K1 (own radix-4 formulation, opt-in since round 2): within north_star's tolerances of the reference."""
    from vorbis_aotuv_lancer_b200 import ops
    cap = refdata.capture(config, 2.0)
    s = ops.Setup(config).set_exact_transform(False)
    pcm, lW, nW, bm, lbm, ref_m, ref_lf, ref_lm, ref_amp, ref_po = [], [], [], [], [], [], [], [], [], []
    for seq in refdata.by_size(cap, W):
        b = cap["blocks"][seq]
        blk = b["TAP_BLOCK"][0]["data"]
        for c in range(s.channels):
            pcm.append(b["TAP_PCM"][c]["data"]); lW.append(blk[1]); nW.append(blk[2])
            bm.append(b["TAP_PCM"][c]["a1"]); lbm.append(b["TAP_PCM"][c]["a2"])
            ref_m.append(b["TAP_MDCT"][c]["data"]); ref_lf.append(b["TAP_TM_LOGFFT"][c]["data"])
            ref_lm.append(b["TAP_NM_LOGMDCT"][c]["data"]); ref_amp.append(b["TAP_TM_SCAL"][c]["data"][1])
            ref_po.append(b["TAP_POSTE"][c]["data"][0])
    assert len(pcm) > 8
    out = ops.forward_transform(s, W, np.stack(pcm), lW, nW, bm, lbm)
    ref_m, ref_lf, ref_lm = np.stack(ref_m), np.stack(ref_lf), np.stack(ref_lm)
    # MDCT: 1e-5 relative to the block RMS
    rms = np.sqrt((ref_m.astype(np.float64) ** 2).mean(axis=1, keepdims=True)) + 1e-30
    err = np.abs(out["gmdct"] - ref_m) / rms
    assert err.max() < 1e-5, err.max()
    # the dB mapping itself is bit-exact on our own coefficients (scales.h:43-51 + the double .345)
    mine = (todB_np(out["gmdct"]).astype(np.float64) + .345).astype(np.float32)
    assert np.array_equal(mine, out["logmdct"])
    # logmdct: the dB image of the MDCT tolerance.  todB is piecewise linear in the mantissa with a
    # slope of at most 12.05 dB per unit relative error, so |d dB| <= 12.05*(1e-5*rms/|m|) (+1e-4 slack)
    allowed = 12.05 * 1e-5 * rms / np.abs(ref_m).clip(1e-30) + 1e-4
    d = np.abs(out["logmdct"] - ref_lm)
    ok = (d <= allowed) | (np.abs(ref_m) < 1e-4 * rms)
    assert ok.all(), float((d / allowed)[~ok].max())
    # logfft: power spectrum within 1e-5 of the block's peak power -> same dB image (slope 6.03 on power)
    peak = ref_lf.max(axis=1, keepdims=True)
    rel = 10.0 ** ((peak - ref_lf).astype(np.float64) / 10.0)      # peak power / bin power
    allowedf = 6.03 * 1e-5 * rel + 2e-4
    d = np.abs(out["logfft"] - ref_lf)
    ok = (d <= allowedf) | (ref_lf < peak - 90.0)
    assert ok.all(), float((d / allowedf)[~ok].max())
    assert np.allclose(out["local_ampmax"], np.array(ref_amp), rtol=1e-5, atol=1e-4)
    po = np.array(ref_po)
    assert np.array_equal(out["poste"] > 0, po > 0)
    assert np.allclose(out["poste"], po, rtol=1e-5)
    s.close()


@pytest.mark.parametrize("config", ["q4_44k_2ch", "q8_48k_6ch"])
@pytest.mark.parametrize("W", [1, 0])
def test_exact_transform_is_bit_identical_to_the_reference(config, W):
    """K1x (the default front end): the reference's own float operations in mdct_forward / drft_forward,
    one thread per butterfly -- MDCT, logmdct, logfft and the block maximum must equal the reference's bit for bit."""
    from vorbis_aotuv_lancer_b200 import ops
    cap = refdata.capture(config, 2.0)
    s = ops.Setup(config)
    pcm, lW, nW, bm, lbm, ref_m, ref_lf, ref_lm, ref_amp, ref_po = [], [], [], [], [], [], [], [], [], []
    for seq in refdata.by_size(cap, W):
        b = cap["blocks"][seq]
        blk = b["TAP_BLOCK"][0]["data"]
        for c in range(s.channels):
            pcm.append(b["TAP_PCM"][c]["data"]); lW.append(blk[1]); nW.append(blk[2])
            bm.append(b["TAP_PCM"][c]["a1"]); lbm.append(b["TAP_PCM"][c]["a2"])
            ref_m.append(b["TAP_MDCT"][c]["data"]); ref_lf.append(b["TAP_TM_LOGFFT"][c]["data"])
            ref_lm.append(b["TAP_NM_LOGMDCT"][c]["data"]); ref_amp.append(b["TAP_TM_SCAL"][c]["data"][1])
            ref_po.append(b["TAP_POSTE"][c]["data"][0])
    assert len(pcm) > 8
    out = ops.forward_transform(s, W, np.stack(pcm), lW, nW, bm, lbm)
    bad = int((out["gmdct"] != np.stack(ref_m)).sum())
    assert bad == 0, "MDCT: %d of %d coefficients differ" % (bad, out["gmdct"].size)
    assert np.array_equal(out["logmdct"], np.stack(ref_lm))
    bad = int((out["logfft"] != np.stack(ref_lf)).sum())
    assert bad == 0, "logfft: %d of %d bins differ" % (bad, out["logfft"].size)
    assert np.array_equal(out["local_ampmax"], np.array(ref_amp, np.float32))
    assert np.allclose(out["poste"], np.array(ref_po), rtol=1e-5)
    s.close()
