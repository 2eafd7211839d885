"""A/B injection parity of every CUDA stage against the reference build (oracle/_ref, scalar C,
-O2 -fno-fast-math): each stage is fed the reference's own inputs for the same block and must
return the reference's outputs -- bit-exact for the integer stages (floor posts, ilogmask,
quantised residue, classes, VQ entries) and for the max/min/add-only tone mask; the noise mask
and the mix are float pipelines that we replay op for op, so they are compared bit-exact too
(the stated tolerance is 1e-5 relative; exact is stronger)."""
import numpy as np
import pytest

from tests import refdata

pytestmark = pytest.mark.gpu

CONFIGS = ["q4_44k_2ch", "q2_44k_2ch", "q8_48k_6ch", "abr128_44k_2ch", "q4_48k_6ch"]
SECONDS = {"q4_44k_2ch": 3.0, "q2_44k_2ch": 2.0, "q8_48k_6ch": 1.5, "abr128_44k_2ch": 1.5, "q4_48k_6ch": 1.5}


GOLDEN = ["q4_44k_2ch", "q2_44k_2ch", "q8_48k_6ch", "abr128_44k_2ch"]


@pytest.fixture(scope="module", params=[c + "@live" for c in CONFIGS] + [c + "@golden" for c in GOLDEN])
def ctx(request):
    """live: the reference build (oracle/_ref) runs on this box; golden: the committed tests/golden vectors."""
    from vorbis_aotuv_lancer_b200 import ops
    name, src = request.param.split("@")
    if src == "live":
        if not refdata.have_ref("c"):
            pytest.skip("oracle/_ref not present")
        cap = refdata.capture(name, SECONDS[name])
    else:
        cap = refdata.load_golden(name)
    s = ops.Setup(name)
    yield name, cap, s, ops
    s.close()


def blocks(cap, W):
    for seq in refdata.by_size(cap, W):
        b = cap["blocks"][seq]
        yield seq, b, b["TAP_BLOCK"][0]["data"]


def report(name, got, ref):
    """first mismatch, for the log"""
    got, ref = np.asarray(got), np.asarray(ref)
    bad = np.argwhere(got != ref)
    return "%s: %d mismatches, first at %s got %s ref %s" % (name, len(bad), bad[0], got[tuple(bad[0])], ref[tuple(bad[0])])


@pytest.mark.parametrize("W", [1, 0])
def test_noisemask(ctx, W):
    name, cap, s, ops = ctx
    psy, bm, lc, po, lm, last, r_no, r_ep, r_np = [], [], [], [], [], [], [], [], []
    for seq, b, blk in blocks(cap, W):
        for c in range(s.channels):
            t = b["TAP_NM_LOGMDCT"][c]
            psy.append(t["a1"]); bm.append(t["a2"]); lm.append(t["data"]); last.append(b["TAP_NM_LASTMDCT"][c]["data"])
            sc = b["TAP_NM_SCAL"][c]["data"]; lc.append(sc[0]); po.append(sc[1])
            r_no.append(b["TAP_NM_NOISE"][c]["data"]); r_ep.append(b["TAP_NM_EPEAK"][c]["data"]); r_np.append(b["TAP_NM_NPEAK"][c]["data"])
    assert len(lm) > 0
    no, ep, npk = ops._vp_noisemask(s, W, psy, lc, np.stack(lm), np.stack(last), po, bm)
    assert np.array_equal(no, np.stack(r_no)), report("noise", no, np.stack(r_no))
    assert np.array_equal(ep, np.stack(r_ep)), report("epeak", ep, np.stack(r_ep))
    assert np.array_equal(npk, np.stack(r_np)), report("npeak", npk, np.stack(r_np))


@pytest.mark.parametrize("W", [1, 0])
def test_tonemask(ctx, W):
    name, cap, s, ops = ctx
    psy, lf, g, l, ref = [], [], [], [], []
    for seq, b, blk in blocks(cap, W):
        for c in range(s.channels):
            t = b["TAP_TM_LOGFFT"][c]
            psy.append(t["a1"]); lf.append(t["data"])
            sc = b["TAP_TM_SCAL"][c]["data"]; g.append(sc[0]); l.append(sc[1])
            ref.append(b["TAP_TM_TONE"][c]["data"])
    tone = ops._vp_tonemask(s, W, psy, np.stack(lf), g, l)
    assert np.array_equal(tone, np.stack(ref)), report("tone", tone, np.stack(ref))


@pytest.mark.parametrize("W", [1, 0])
def test_offset_and_mix(ctx, W):
    name, cap, s, ops = ctx
    psy, par, lc, A = [], [], [], {k: [] for k in ("NOISE", "TONE", "MDCT_IN", "LOGMDCT", "LAST_IN", "TEMP_IN", "NPEAK_IN",
                                                  "LOGMASK", "MDCT_OUT", "LAST_OUT", "TEMP_OUT", "NPEAK_OUT")}
    for seq, b, blk in blocks(cap, W):
        for k, iv in enumerate(b["TAP_MIX_I32"]):
            psy.append(iv["a1"]); par.append(iv["data"]); lc.append(b["TAP_MIX_F32"][k]["data"][0])
            for key in A:
                A[key].append(b["TAP_MIX_" + key][k]["data"])
    A = {k: np.stack(v) for k, v in A.items()}
    mask, mdct, last, temp, npk = ops._vp_offset_and_mix(s, W, psy, A["NOISE"], A["TONE"], np.stack(par), lc,
                                                         A["MDCT_IN"], A["LOGMDCT"], A["LAST_IN"], A["TEMP_IN"], A["NPEAK_IN"])
    assert np.array_equal(mask, A["LOGMASK"]), report("logmask", mask, A["LOGMASK"])
    assert np.array_equal(mdct, A["MDCT_OUT"]), report("mdct", mdct, A["MDCT_OUT"])
    assert np.array_equal(npk, A["NPEAK_OUT"]), report("npeak", npk, A["NPEAK_OUT"])
    assert np.array_equal(temp, A["TEMP_OUT"]), report("tempmdct", temp, A["TEMP_OUT"])
    assert np.array_equal(last, A["LAST_OUT"]), report("lastmdct", last, A["LAST_OUT"])


@pytest.mark.parametrize("W", [1, 0])
def test_floor1(ctx, W):
    name, cap, s, ops = ctx
    byfloor = {}
    for seq, b, blk in blocks(cap, W):
        for k, t in enumerate(b["TAP_FIT_LOGMDCT"]):
            d = byfloor.setdefault(t["a1"], dict(lm=[], mk=[], ref=[]))
            d["lm"].append(t["data"]); d["mk"].append(b["TAP_FIT_LOGMASK"][k]["data"]); d["ref"].append(b["TAP_FIT_POSTS"][k]["data"])
    assert byfloor
    for fi, d in byfloor.items():
        posts, fitted = ops.floor1_fit(s, W, fi, np.stack(d["lm"]), np.stack(d["mk"]))
        for i, ref in enumerate(d["ref"]):
            if len(ref) == 0:
                assert fitted[i] == 0
            else:
                assert fitted[i] == 1 and np.array_equal(posts[i, :len(ref)], ref), (fi, i, posts[i, :len(ref)], ref)
    # encode: quantise/predict + render
    byfloor = {}
    for seq, b, blk in blocks(cap, W):
        for k, t in enumerate(b["TAP_FENC_IN"]):
            d = byfloor.setdefault(t["a1"], dict(pin=[], fit=[], post=[], ilog=[], ret=[]))
            p = np.zeros(65, np.int32); p[:len(t["data"])] = t["data"]
            d["pin"].append(p); d["fit"].append(int(len(t["data"]) > 0))
            d["post"].append(b["TAP_FENC_POST"][k]["data"]); d["ilog"].append(b["TAP_FENC_ILOG"][k]["data"]); d["ret"].append(b["TAP_FENC_RET"][k]["data"][0])
    for fi, d in byfloor.items():
        posts, vals, ilog, nz = ops.floor1_encode(s, W, fi, np.stack(d["pin"]), d["fit"])
        assert np.array_equal(nz, np.array(d["ret"]))
        assert np.array_equal(ilog, np.stack(d["ilog"])), report("ilogmask", ilog, np.stack(d["ilog"]))
        for i, ref in enumerate(d["post"]):
            assert np.array_equal(posts[i, :len(ref)], ref)


@pytest.mark.parametrize("W", [1, 0])
def test_couple(ctx, W):
    name, cap, s, ops = ctx
    byblob = {}
    for seq, b, blk in blocks(cap, W):
        for k, iv in enumerate(b["TAP_CPL_I32"]):
            v = iv["data"]; ch = int(v[2])
            d = byblob.setdefault(int(v[0]), {x: [] for x in ("psy", "md", "ep", "npi", "iwi", "nzi", "iwo", "nzo", "npo", "slp", "lpr")})
            g = lambda t: b[t][k]["data"]  # noqa: E731
            d["psy"].append(int(v[4])); d["slp"].append(int(v[1])); d["lpr"].append(int(v[3]))
            d["md"].append(g("TAP_CPL_MDCT").reshape(ch, -1)); d["ep"].append(g("TAP_CPL_EPEAK").reshape(ch, -1))
            d["npi"].append(g("TAP_CPL_NPEAK_IN").reshape(ch, -1)); d["iwi"].append(g("TAP_CPL_IWORK_IN").reshape(ch, -1))
            d["nzi"].append(g("TAP_CPL_NZ_IN")); d["iwo"].append(g("TAP_CPL_IWORK_OUT").reshape(ch, -1))
            d["nzo"].append(g("TAP_CPL_NZ_OUT")); d["npo"].append(g("TAP_CPL_NPEAK_OUT").reshape(ch, -1))
    assert byblob
    for blob, d in byblob.items():
        # the library derives sliding_lowpass / lowpassr from the setup exactly as mapping0.c:778-781,1255 does
        iw, nz, npk = ops._vp_couple_quantize_normalize(s, W, d["psy"], blob, np.stack(d["md"]), np.stack(d["ep"]),
                                                        np.stack(d["npi"]), np.stack(d["iwi"]), np.stack(d["nzi"]))
        assert np.array_equal(iw, np.stack(d["iwo"])), report("iwork blob %d" % blob, iw, np.stack(d["iwo"]))
        assert np.array_equal(nz, np.stack(d["nzo"]))
        assert np.array_equal(npk, np.stack(d["npo"])), report("nepeak", npk, np.stack(d["npo"]))


@pytest.mark.parametrize("W", [1, 0])
def test_residue(ctx, W):
    name, cap, s, ops = ctx
    byres = {}
    for seq, b, blk in blocks(cap, W):
        for k, iv in enumerate(b["TAP_RES_I32"]):
            v = iv["data"]; resnum, ch = int(v[0]), int(v[2])
            d = byres.setdefault((resnum, ch, int(v[1])), dict(rows=[], nz=[], pw=[], syms=[], out=[]))
            d["rows"].append(b["TAP_RES_IN"][k]["data"].reshape(ch, -1)); d["nz"].append(b["TAP_RES_NZ"][k]["data"])
            d["pw"].append(b["TAP_RES_PARTWORD"][k]); d["syms"].append(b["TAP_RES_SYMS"][k]["data"])
            d["out"].append(b["TAP_RES_OUT"][k]["data"].reshape(ch, -1))
    assert byres
    total = 0
    for (resnum, ch, rtype), d in byres.items():
        rows, nz = np.stack(d["rows"]), np.stack(d["nz"])
        pw, cl = ops.residue_class(s, W, resnum, rows, nz)
        for i, ref in enumerate(d["pw"]):
            assert cl[i] == ref["a2"], (resnum, i, cl[i], ref["a2"])
            if ref["a2"]:
                assert np.array_equal(pw[i, :ref["a2"]].ravel(), ref["data"]), (resnum, i)
        syms, left = ops.residue_forward(s, W, resnum, rows, nz, pw)
        for i, ref in enumerate(d["syms"]):
            assert np.array_equal(syms[i].ravel(), ref), (resnum, i, len(syms[i]), len(ref) // 2)
            total += len(syms[i])
            if rtype == 1:
                assert np.array_equal(left[i], d["out"][i])
    assert total > 50
