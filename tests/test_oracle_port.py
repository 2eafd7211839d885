"""Pins the CPU oracle (oracle/port/*.c) against the real reference (oracle/_ref, scalar C
flavour, -O2 -fno-fast-math) stage by stage on the same synthetic PCM.  CPU only."""
import numpy as np
import pytest

from oracle import port as oport
from tests import refdata
from vorbis_aotuv_lancer_b200 import blob as vblob

CONFIGS = ["q4_44k_2ch", "q2_44k_2ch", "q8_48k_6ch", "abr128_44k_2ch", "q4_48k_6ch",
           "q4_32k_2ch", "q0_44k_2ch", "q7_44k_2ch", "abr256_48k_6ch", "q5_48k_3ch"]      # second row: round 2
SECONDS = {"q4_44k_2ch": 3.0, "q2_44k_2ch": 2.0, "q8_48k_6ch": 1.5, "abr128_44k_2ch": 1.5, "q4_48k_6ch": 1.5,
           "q4_32k_2ch": 1.5, "q0_44k_2ch": 1.5, "q7_44k_2ch": 1.5, "abr256_48k_6ch": 1.0, "q5_48k_3ch": 1.5}
GOLDEN = ["q4_44k_2ch", "q2_44k_2ch", "q8_48k_6ch", "abr128_44k_2ch"]


@pytest.fixture(scope="module", params=[c + "@live" for c in CONFIGS] + [c + "@golden" for c in GOLDEN])
def ctx(request):
    """live: the reference build runs here (oracle/_ref); golden: the committed vectors of tests/golden."""
    name, src = request.param.split("@")
    if src == "live":
        if not refdata.have_ref("c"):
            pytest.skip("oracle/_ref not built (needs /root/reference)")
        cap = refdata.capture(name, SECONDS[name])
    else:
        cap = refdata.load_golden(name)
    orc = oport.Oracle(vblob.parse(vblob.load(name)))
    return name, cap, orc


def blocks(cap):
    for seq in sorted(cap["blocks"]):
        b = cap["blocks"][seq]
        if "TAP_BLOCK" in b:
            yield seq, b, b["TAP_BLOCK"][0]["data"]


def test_front_end(ctx):
    name, cap, orc = ctx
    worst = 0.0
    for seq, b, blk in blocks(cap):
        W, lW, nW = int(blk[0]), int(blk[1]), int(blk[2])
        for c in range(orc.channels):
            pcm = b["TAP_PCM"][c]
            w = orc.window(pcm["data"], W, lW, nW)
            assert np.array_equal(w, b["TAP_WINDOWED"][c]["data"])
            ref = b["TAP_MDCT"][c]["data"]
            m = orc.mdct_forward(w)
            rms = np.sqrt((ref.astype(np.float64) ** 2).mean()) + 1e-30
            worst = max(worst, float(np.abs(m - ref).max() / rms))
            assert np.array_equal(orc.mdct_forward_exact(w), ref)      # the operation-by-operation restatement
            assert np.array_equal(orc.logmdct(ref), b["TAP_NM_LOGMDCT"][c]["data"])
            lf, amp = orc.logfft(w)
            rlf = b["TAP_TM_LOGFFT"][c]["data"]
            good = rlf > rlf.max() - 60
            assert np.abs(lf - rlf)[good].max() < 2e-3
            assert abs(amp - b["TAP_TM_SCAL"][c]["data"][1]) < 2e-3
            lfx, ampx = orc.logfft_exact(w)                            # the operation-by-operation restatement
            assert np.array_equal(lfx, rlf) and np.float32(ampx) == b["TAP_TM_SCAL"][c]["data"][1]
            assert orc.postnoise(pcm["data"], pcm["a1"], pcm["a2"]) == pytest.approx(b["TAP_POSTE"][c]["data"][0], rel=1e-6)
            lf_ = b["TAP_LOUDFIX"][c]
            assert orc.loudnoise_fix(b["TAP_NM_LOGMDCT"][c]["a1"], lf_["data"][0], b["TAP_NM_LOGMDCT"][c]["data"],
                                     lf_["a1"], lf_["a2"]) == lf_["data"][1]
    assert worst < 1e-5, worst   # the reference float MDCT itself is ~6e-6 of the block RMS off the exact transform


def test_noisemask_bit_exact(ctx):
    name, cap, orc = ctx
    n = 0
    for seq, b, blk in blocks(cap):
        for c in range(orc.channels):
            lm = b["TAP_NM_LOGMDCT"][c]
            sc = b["TAP_NM_SCAL"][c]["data"]
            no, ep, npk = orc.noisemask(lm["a1"], sc[0], lm["data"], b["TAP_NM_LASTMDCT"][c]["data"], sc[1], lm["a2"])
            assert np.array_equal(no, b["TAP_NM_NOISE"][c]["data"]), (seq, c)
            assert np.array_equal(ep, b["TAP_NM_EPEAK"][c]["data"]), (seq, c)
            assert np.array_equal(npk, b["TAP_NM_NPEAK"][c]["data"]), (seq, c)
            n += 1
    assert n > 0


def test_tonemask_bit_exact(ctx):
    name, cap, orc = ctx
    for seq, b, blk in blocks(cap):
        for c in range(orc.channels):
            lf = b["TAP_TM_LOGFFT"][c]
            sc = b["TAP_TM_SCAL"][c]["data"]
            tone = orc.tonemask(lf["a1"], lf["data"], sc[0], sc[1])
            assert np.array_equal(tone, b["TAP_TM_TONE"][c]["data"]), (seq, c)


def test_offset_and_mix_bit_exact(ctx):
    name, cap, orc = ctx
    for seq, b, blk in blocks(cap):
        for k, iv in enumerate(b["TAP_MIX_I32"]):
            v = iv["data"]
            g = lambda t: b[t][k]["data"]  # noqa: E731
            mask, mdct, last, temp, npk = orc.offset_and_mix(
                iv["a1"], g("TAP_MIX_NOISE"), g("TAP_MIX_TONE"), int(v[0]), int(v[1]), g("TAP_MIX_MDCT_IN"),
                g("TAP_MIX_LOGMDCT"), g("TAP_MIX_LAST_IN"), g("TAP_MIX_TEMP_IN"), float(g("TAP_MIX_F32")[0]),
                g("TAP_MIX_NPEAK_IN"), int(v[2]), int(v[3]), int(v[4]), int(v[5]), int(v[6]), int(v[7]))
            assert np.array_equal(mask, g("TAP_MIX_LOGMASK")), (seq, k)
            assert np.array_equal(mdct, g("TAP_MIX_MDCT_OUT")), (seq, k)
            assert np.array_equal(last, g("TAP_MIX_LAST_OUT")), (seq, k)
            assert np.array_equal(temp, g("TAP_MIX_TEMP_OUT")), (seq, k)
            assert np.array_equal(npk, g("TAP_MIX_NPEAK_OUT")), (seq, k)


def test_floor1_bit_exact(ctx):
    name, cap, orc = ctx
    nfit = 0
    for seq, b, blk in blocks(cap):
        for k, lm in enumerate(b["TAP_FIT_LOGMDCT"]):
            posts = orc.floor1_fit(lm["a1"], lm["data"], b["TAP_FIT_LOGMASK"][k]["data"])
            ref = b["TAP_FIT_POSTS"][k]["data"]
            if len(ref) == 0:
                assert posts is None
            else:
                assert posts is not None and np.array_equal(posts, ref), (seq, k)
            nfit += 1
        for k, pin in enumerate(b["TAP_FENC_IN"]):
            posts = pin["data"] if len(pin["data"]) else None
            nz, post, vals, ilog = orc.floor1_encode(pin["a1"], posts, int(blk[4]) // 2)
            assert nz == b["TAP_FENC_RET"][k]["data"][0]
            assert np.array_equal(ilog, b["TAP_FENC_ILOG"][k]["data"]), (seq, k)
            if posts is not None:
                assert np.array_equal(post, b["TAP_FENC_POST"][k]["data"])
    assert nfit > 0


def test_couple_bit_exact(ctx):
    name, cap, orc = ctx
    for seq, b, blk in blocks(cap):
        W = int(blk[0])
        mapping = int(orc.modes[W][1])
        for k, iv in enumerate(b["TAP_CPL_I32"]):
            v = iv["data"]
            ch = int(v[2])
            g = lambda t, k=k: b[t][k]["data"]  # noqa: E731
            iw, nz, npk = orc.couple(int(v[4]), mapping, int(v[0]), g("TAP_CPL_MDCT").reshape(ch, -1),
                                     g("TAP_CPL_EPEAK").reshape(ch, -1), g("TAP_CPL_NPEAK_IN").reshape(ch, -1),
                                     g("TAP_CPL_IWORK_IN").reshape(ch, -1), g("TAP_CPL_NZ_IN"), int(v[1]), int(v[3]))
            assert np.array_equal(iw, g("TAP_CPL_IWORK_OUT").reshape(ch, -1)), (seq, k)
            assert np.array_equal(nz, g("TAP_CPL_NZ_OUT"))
            assert np.array_equal(npk, g("TAP_CPL_NPEAK_OUT").reshape(ch, -1))


def test_residue_bit_exact(ctx):
    name, cap, orc = ctx
    nsym = 0
    for seq, b, blk in blocks(cap):
        for k, iv in enumerate(b["TAP_RES_I32"]):
            v = iv["data"]
            resnum, ch = int(v[0]), int(v[2])
            rows = b["TAP_RES_IN"][k]["data"].reshape(ch, -1)
            nz = b["TAP_RES_NZ"][k]["data"]
            pw = orc.residue_class(resnum, rows, nz)
            refpw = b["TAP_RES_PARTWORD"][k]
            if refpw["a2"] == 0:
                assert pw is None
                assert len(b["TAP_RES_SYMS"][k]["data"]) == 0
                continue
            assert np.array_equal(pw, refpw["data"].reshape(refpw["a2"], -1)), (seq, k)
            syms, left = orc.residue_forward(resnum, rows, nz, pw)
            assert np.array_equal(syms.ravel(), b["TAP_RES_SYMS"][k]["data"]), (seq, k)
            if int(v[1]) == 1:
                assert np.array_equal(left, b["TAP_RES_OUT"][k]["data"].reshape(ch, -1))
            nsym += len(syms)
    assert nsym > 100
