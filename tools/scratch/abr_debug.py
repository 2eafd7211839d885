"""scratch: where does the managed-mode (ABR) chain first leave the reference?  run under gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import refdata
from tests.test_gpu_pipeline import build_blocks
from vorbis_aotuv_lancer_b200 import ops

G = dict(GMDCT=0, LOGFFT=1, LOGMDCT=2, NOISE=3, TONE=4, LOGMASK=5, EPEAK=6, IWORK=7, ILOGMASK=8, POSTS=9, POSTS_LO=10, POSTS_HI=11, NPEAK=12, BLOB_BYTES=13)
config = sys.argv[1] if len(sys.argv) > 1 else "abr128_44k_2ch"
cap = refdata.capture(config, 3.0)
s = ops.Setup(config)
blocks, pcm = build_blocks(cap, s.channels)
st = ops.Stream(s)
bt = ops.Batch(s, len(blocks))
bt.run([st], (ops.Block * len(blocks))(*blocks), pcm)
pk = bt.packets()
ref = cap["packets"]
seqs = [q for q in sorted(cap["blocks"]) if "TAP_BLOCK" in cap["blocks"][q]]
nshown = 0
for bi, (a, r, seq) in enumerate(zip(pk, ref, seqs)):
    b = cap["blocks"][seq]
    blk = b["TAP_BLOCK"][0]["data"]
    W = int(blk[0]); N2 = 1024 if W else 128
    bb = bt.get(G["BLOB_BYTES"], bi, 0, 15, np.int32)
    rb = np.asarray(b["TAP_BLOBBYTES"][0]["data"]) if "TAP_BLOBBYTES" in b else None
    same_pk = a[0] == r[0]
    blob_same = rb is not None and np.array_equal(bb, rb)
    if same_pk and blob_same:
        continue
    nshown += 1
    if nshown > 6:
        break
    print("block %d seq %d W=%d type=%d lW=%d nW=%d hist=%s packet_same=%s" % (bi, seq, W, blk[3], blk[1], blk[2], list(blk[8:12]), same_pk))
    print("   blob bytes ours", list(bb)); print("   blob bytes ref ", None if rb is None else list(rb))
    for c in range(s.channels):
        def cmp(name, ours, theirs):
            theirs = np.asarray(theirs)
            ours = np.asarray(ours)[:len(theirs)]
            if theirs.dtype.kind == 'f':
                bad = int((ours.view(np.uint32) != theirs.astype(np.float32).view(np.uint32)).sum())
            else:
                bad = int((ours != theirs).sum())
            print("   ch%d %-14s %s" % (c, name, "ok" if bad == 0 else "DIFF %d/%d first at %s" % (bad, len(theirs), np.flatnonzero(ours != theirs)[:5])))
        cmp("logmdct", bt.get(G["LOGMDCT"], bi, c, N2), b["TAP_NM_LOGMDCT"][c]["data"])
        cmp("lastmdct(NM)", np.zeros(0), np.zeros(0))
        cmp("noise", bt.get(G["NOISE"], bi, c, N2), b["TAP_NM_NOISE"][c]["data"])
        cmp("epeak", bt.get(G["EPEAK"], bi, c, N2), b["TAP_NM_EPEAK"][c]["data"])
        cmp("tone", bt.get(G["TONE"], bi, c, N2), b["TAP_TM_TONE"][c]["data"])
        lm = bt.get(G["LOGMASK"], bi, c, 3 * N2)
        mixes = [m for m in b["TAP_MIX_LOGMASK"] if m["a0"] == c]
        for m in mixes:
            sel = m["a2"]; slot = {1: 0, 2: 1, 0: 2}[sel]
            cmp("logmask sel%d" % sel, lm[slot * N2:(slot + 1) * N2], m["data"])
        mo = [m for m in b["TAP_MIX_MDCT_OUT"] if m["a0"] == c]
        cmp("mdct after mix", bt.get(G["GMDCT"], bi, c, N2), mo[-1]["data"])
        fits = [m for m in b["TAP_FIT_POSTS"] if m["a0"] == c]
        names = ["POSTS", "POSTS_HI", "POSTS_LO"]
        for k, f in enumerate(fits):
            ours = bt.get(G[names[k]], bi, c, 65, np.int32)
            cmp("fit#%d" % k, ours, f["data"])
        npk_out = [m for m in b["TAP_MIX_NPEAK_OUT"] if m["a0"] == c]
        lin = [m for m in b["TAP_MIX_LAST_IN"] if m["a0"] == c]
        print("      mix calls: sels", [m["a2"] for m in mixes])
print("differing blocks shown:", nshown)
