"""Whole-path parity: the batched CUDA path is handed the reference's own block descriptors and
block PCM (what `vorbis_analysis_blockout` produced) and must produce the reference's stream:
 (a) with the default front end (K1x, the reference's own float operations) EVERY packet byte-identical to the
     reference's, on all eight configurations incl. managed bitrate;
 (b) with the K1 front end (own MDCT/FFT formulation, north_star's 1e-5 tolerance) byte-identical for the vast
     majority of blocks (the only permitted source of difference is that tolerance flipping a quantiser decision),
     bitrate within 0.5 % of the reference stream, decodable by the reference decoder, decoded-PCM SNR against the
     reference-encoded stream stated per config."""
import numpy as np
import pytest

from tests import refdata

pytestmark = pytest.mark.gpu

CASES = [("q4_44k_2ch", 6.0), ("q5_44k_2ch", 3.0), ("q10_48k_2ch", 3.0), ("q8_48k_6ch", 2.0), ("q2_44k_2ch", 3.0),
         ("q4_48k_6ch", 2.0), ("q4_44k_1ch", 3.0), ("abr128_44k_2ch", 3.0)]


def build_blocks(cap, ch, stream=0):
    from vorbis_aotuv_lancer_b200 import ops
    descs, pcm, off = [], [], 0
    for seq in sorted(cap["blocks"]):
        b = cap["blocks"][seq]
        if "TAP_BLOCK" not in b:
            continue
        blk = b["TAP_BLOCK"][0]["data"]
        n = int(blk[4])
        gp = (int(blk[7]) << 32) | (int(blk[6]) & 0xFFFFFFFF)
        descs.append((stream, blk[0], blk[1], blk[2], blk[3], blk[5], gp, off, n))
        for c in range(ch):
            pcm.append(b["TAP_PCM"][c]["data"])
        off += n * ch
    return ops.blocks_array(descs), np.concatenate(pcm)


@pytest.mark.parametrize("config,seconds", CASES)
@pytest.mark.parametrize("front_end", ["K1x", "K1"])
def test_stream_matches_reference(config, seconds, front_end):
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import ops
    cap = refdata.capture(config, seconds)
    s = ops.Setup(config)
    if front_end == "K1":
        s.set_exact_transform(False)
    blocks, pcm = build_blocks(cap, s.channels)
    st = ops.Stream(s)
    bt = ops.Batch(s, len(blocks))
    # two batches, to exercise the cross-batch stream state
    half = len(blocks) // 2
    bt.run([st], (ops.Block * half)(*blocks[:half]), pcm)
    pk = bt.packets()
    bt.run([st], (ops.Block * (len(blocks) - half))(*blocks[half:]), pcm)
    pk += bt.packets()
    ref = cap["packets"]
    assert len(pk) == len(ref)
    same = sum(1 for a, r in zip(pk, ref) if a[0] == r[0])
    gbytes, rbytes = sum(len(a[0]) for a in pk), sum(len(r[0]) for r in ref)
    frac = same / len(ref)
    # decode both with the reference decoder
    R = reflib.RefLib("c")
    n = cap["pcm"].shape[1]
    dec_g, _ = R.decode(cap["headers"], [(a[0], a[1], a[2]) for a in pk], n + 8192)
    dec_r, _ = R.decode(cap["headers"], ref, n + 8192)
    m = min(dec_g.shape[1], dec_r.shape[1])
    assert m > n - 4096
    err = (dec_g[:, :m] - dec_r[:, :m]).astype(np.float64)
    snr = 10 * np.log10((dec_r[:, :m].astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-30))
    print("\n%s %s: %d blocks, identical packets %.1f%%, bytes gpu %d ref %d (%.3f%%), SNR vs ref-encoded %.1f dB"
          % (config, front_end, len(ref), 100 * frac, gbytes, rbytes, 100 * (gbytes - rbytes) / rbytes, snr))
    assert [a[1] for a in pk] == [r[1] for r in ref]          # granule positions
    if front_end == "K1x":
        assert same == len(ref), "%d of %d packets differ from the reference's" % (len(ref) - same, len(ref))
    assert abs(gbytes - rbytes) / rbytes < 0.005
    assert frac > 0.80
    assert snr > 30.0
    # the device bit packer (K8) and the host reference-order emitter must agree byte for byte
    from vorbis_aotuv_lancer_b200.lib import L
    st2 = ops.Stream(s)
    L.vb200_batch_set_host_emit(bt.h, 1)
    bt.run([st2], (ops.Block * half)(*blocks[:half]), pcm)
    pk2 = bt.packets()
    bt.run([st2], (ops.Block * (len(blocks) - half))(*blocks[half:]), pcm)
    pk2 += bt.packets()
    assert [a[0] for a in pk2] == [a[0] for a in pk]
    st.close(); st2.close(); bt.close(); s.close()


def test_device_bitrate_manager_equals_host_manager_over_many_streams():
    """in managed mode the reservoirs are walked on the device, one thread per stream, and only
    the chosen packets come back.  The host restatement of bitrate.c:73-227 (host-emission debug path) must give
    the same packets, stream by stream, across two batches (the reservoirs carry over)."""
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops
    from vorbis_aotuv_lancer_b200.lib import L
    config = "abr128_44k_2ch"
    ch, rate, _, _ = vblob.CATALOGUE[config]
    s = ops.Setup(config)
    assert s.managed
    pcms = [refdata.synth(ch, rate, 2.0 + 0.5 * i, seed=31 + i) for i in range(5)]
    enc = encoder.Encoder(s, 4096, threads=4, chunks=1)
    blocks, timeline, counts = enc.segment(pcms)
    view = np.frombuffer(blocks, dtype=encoder.BLOCK_DTYPE)
    first = np.zeros(len(view), bool)                       # first half of every stream goes into batch one
    pos = 0
    for c in counts:
        first[pos:pos + c // 2] = True
        pos += c
    parts = [np.ascontiguousarray(view[first]), np.ascontiguousarray(view[~first])]

    def run(host_emit):
        streams = [ops.Stream(s) for _ in pcms]
        bt = ops.Batch(s, len(view))
        L.vb200_batch_set_host_emit(bt.h, 1 if host_emit else 0)
        out = []
        for part in parts:
            arr = (ops.Block * len(part)).from_buffer_copy(part.tobytes())
            bt.run(streams, arr, timeline)
            tab = encoder.table(bt)
            out.append([encoder.packets_of(tab, i) for i in range(len(pcms))])
        for st in streams:
            st.close()
        bt.close()
        return out
    dev, host = run(False), run(True)
    for half in range(2):
        for i in range(len(pcms)):
            assert len(dev[half][i]) > 10
            assert dev[half][i] == host[half][i], (half, i)
    # nominal 128 kbit/s: the managed stream must sit near it (the reservoir keeps it there)
    secs = sum(p.shape[1] for p in pcms) / rate
    kbps = sum(len(p[0]) for h in dev for st in h for p in st) * 8 / secs / 1000
    assert 100 < kbps < 150, kbps
    enc.close(); s.close()


# single batch, every configuration: all packets byte-identical (measured on a B200, round 2)
EXACT_CASES = [("q4_44k_2ch", 4.0, 1.0), ("q8_48k_6ch", 2.0, 1.0), ("q2_44k_2ch", 3.0, 1.0), ("q10_48k_2ch", 3.0, 1.0),
               ("q4_48k_6ch", 2.0, 1.0), ("q4_44k_1ch", 3.0, 1.0), ("abr128_44k_2ch", 3.0, 1.0), ("q5_44k_2ch", 3.0, 1.0),
               # round 2: 32 kHz, lowest quality, q6~ branch of M3, managed 5.1, three uncoupled channels
               ("q4_32k_2ch", 3.0, 1.0), ("q0_44k_2ch", 3.0, 1.0), ("q7_44k_2ch", 3.0, 1.0), ("abr256_48k_6ch", 1.5, 1.0),
               ("q5_48k_3ch", 2.0, 1.0)]


@pytest.mark.parametrize("config,seconds,min_identical", EXACT_CASES)
def test_stream_with_exact_transform(config, seconds, min_identical):
    """Whole path with K1x (operation-exact MDCT / real FFT): every later stage is bit-exact on reference inputs
    (test_gpu_stages), so the packets ARE the reference's: every packet byte-identical."""
    from vorbis_aotuv_lancer_b200 import ops
    cap = refdata.capture(config, seconds)
    s = ops.Setup(config).set_exact_transform(True)
    blocks, pcm = build_blocks(cap, s.channels)
    st = ops.Stream(s)
    bt = ops.Batch(s, len(blocks))
    bt.run([st], (ops.Block * len(blocks))(*blocks), pcm)
    pk = bt.packets()
    ref = cap["packets"]
    assert len(pk) == len(ref)
    same = sum(1 for a, r in zip(pk, ref) if a[0] == r[0])
    gbytes, rbytes = sum(len(a[0]) for a in pk), sum(len(r[0]) for r in ref)
    print("\n%s with the exact transform: %d blocks, identical packets %.2f%% (%d differ), bytes gpu %d ref %d"
          % (config, len(ref), 100.0 * same / len(ref), len(ref) - same, gbytes, rbytes))
    assert [a[1] for a in pk] == [r[1] for r in ref]
    assert same / len(ref) >= min_identical
    st.close(); bt.close(); s.close()


def _capture_pcm(config, pcm):
    """like refdata.capture, for a given signal: the reference's blocks (taps) and packets"""
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob
    ch, rate, q, abr = vblob.CATALOGUE[config]
    R = reflib.RefLib("c")
    h = R.open(ch, rate, q, abr)
    R.tap_all(0, 1 << 30)
    R.write(h, pcm)
    out = dict(pcm=pcm, blocks=reflib.group(R.tap_records()), packets=R.packets(h))
    R.tap_config([], 0, 0)
    R.close(h)
    return out


@pytest.mark.parametrize("config", ["abr128_44k_2ch", "q4_44k_2ch"])
def test_time_split_batches_with_digital_silence(config):
    """A stream with stretches of digital silence analysed in FIVE consecutive batches (what the drop-in does with
    its block queue): lastmdct / tempmdct and, in managed mode, the floor-fit gates that decide who writes them
    (mapping0.c:1097) must carry over the batch boundaries exactly -- every packet byte-identical to the reference's."""
    from vorbis_aotuv_lancer_b200 import blob as vblob, ops
    from vorbis_aotuv_lancer_b200.lib import L
    ch, rate, q, abr = vblob.CATALOGUE[config]
    sig = refdata.synth(ch, rate, 4.0, seed=5)
    pcm = sig.copy()
    pcm[:, :20000] = 0.0
    pcm[:, 60000:90000] = 0.0
    for k in range(4):                                  # bursts out of silence: short blocks next to silent ones
        p = 62000 + k * 6000
        pcm[:, p:p + 250] = sig[:, p:p + 250]
    pcm[:, 150000:] = 0.0
    cap = _capture_pcm(config, pcm)
    s = ops.Setup(config)
    blocks, flat = build_blocks(cap, s.channels)
    ref = cap["packets"]
    st = ops.Stream(s)
    bt = ops.Batch(s, len(blocks))
    pk, rounds = [], []
    cuts = [0, len(blocks) // 7, len(blocks) // 3, len(blocks) // 2, (3 * len(blocks)) // 4, len(blocks)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        bt.run([st], (ops.Block * (b - a))(*blocks[a:b]), flat)
        pk += bt.packets()
        rounds.append(int(L.vb200_batch_link_rounds(bt.h)))
    assert len(pk) == len(ref)
    diff = [i for i, (x, r) in enumerate(zip(pk, ref)) if x[0] != r[0]]
    print("\n%s in 5 batches: %d packets, %d differ, link rounds per batch %s" % (config, len(ref), len(diff), rounds))
    assert not diff, diff[:10]
    if s.managed:
        assert max(rounds) > 1          # the silent blocks did close gates: the links were rebuilt
    else:
        assert rounds == [1] * 5
    st.close(); bt.close(); s.close()


def test_bit_packer_second_pass():
    """K8 packs mono/stereo blocks into a 4 KB shared-memory image first (more CTAs per SM); a block that does not
    fit is marked and packed again by k_pack_retry with the 16 KB image.  VB200_PACK_LIMIT=64 lets the first pass
    fill 256 bytes only, so nearly every long block takes the second pass -- the stream must still be the
    reference's, byte for byte (the variable is read once per process, hence the child process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VB200_PACK_LIMIT="64")
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    assert "100.0% byte-identical" in r.stdout, r.stdout[-1000:]


def test_general_stream_scan_gives_the_same_stream():
    """Setups whose loud-noise fix is off (every quality >= q3) take k_stream_scan_ampmax (chunked shared-memory replay
    of the ampmax recurrence); q <= 2 takes k_stream_scan, which also carries the loud-noise level.  VB200_SCAN_GENERAL
    sends a q4 stream through the general kernel: it must produce the reference's stream as well (the q2 / q0 cases
    above pin the general kernel where it is the only choice)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VB200_SCAN_GENERAL="1")
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    assert "100.0% byte-identical" in r.stdout, r.stdout[-1000:]
