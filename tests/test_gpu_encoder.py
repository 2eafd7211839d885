"""The public call (Encoder.encode): GPU-assisted segmentation must equal the host segmenter (and through
tests/test_segment.py the reference's blockout) exactly, and the stream must decode."""
import numpy as np
import pytest

from tests import refdata

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,seconds", [("q4_44k_2ch", 4.0), ("q8_48k_6ch", 1.5), ("q4_44k_1ch", 0.4)])
def test_gpu_segmenter_equals_host_segmenter(config, seconds):
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops, segment
    ch, rate, _, _ = vblob.CATALOGUE[config]
    s = ops.Setup(config)
    pcms = [refdata.synth(ch, rate, seconds, seed=11 + i) for i in range(3)] + [refdata.synth(ch, rate, 0.05, seed=5)]
    g = encoder.Segmenter(s)
    blocks, dptr, elems = g.run(pcms, threads=4)
    tl = g.host_timeline(elems)
    pos, base = 0, 0
    for i, p in enumerate(pcms):
        hb, ht = segment.segment_stream(s, p, stream_index=i)
        for k in range(len(hb)):
            a, b = blocks[pos + k], hb[k]
            assert (a.stream, a.W, a.lW, a.nW, a.blocktype, a.eofflag, a.granulepos, a.pcm_offset - base, a.pcm_stride) == \
                   (b.stream, b.W, b.lW, b.nW, b.blocktype, b.eofflag, b.granulepos, b.pcm_offset, b.pcm_stride), (i, k)
        assert np.array_equal(tl[base:base + ht.size], ht)
        pos += len(hb)
        base += ht.size
    assert pos == len(blocks)
    g.close(); s.close()


@pytest.mark.parametrize("config,seconds", [("q4_44k_2ch", 4.0), ("q8_48k_6ch", 1.5), ("q4_44k_1ch", 0.4)])
def test_pinned_input_goes_up_directly_and_gives_the_same_timeline(config, seconds):
    """PCM in page-locked memory skips the staging copy; device timeline and blocks must not change."""
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops
    ch, rate, _, _ = vblob.CATALOGUE[config]
    s = ops.Setup(config)
    pcms = [refdata.synth(ch, rate, seconds, seed=21 + i) for i in range(3)] + [refdata.synth(ch, rate, 0.05, seed=6),
                                                                                 refdata.synth(ch, rate, 0.2, seed=7)]
    pins = [encoder.PinnedArray(*p.shape) for p in pcms]
    for a, p in zip(pins, pcms):
        a.array[...] = p

    def timeline(g, inputs):
        blocks, dptr, elems = g.run(inputs, threads=4)
        return np.frombuffer(blocks, dtype=encoder.BLOCK_DTYPE).copy(), g.device_timeline(elems)
    g0, g1 = encoder.Segmenter(s), encoder.Segmenter(s)
    b0, t0 = timeline(g0, pcms)
    b1, t1 = timeline(g1, [a.array for a in pins])
    assert not g0.direct and g1.direct
    assert np.array_equal(b0, b1)
    assert np.array_equal(t0, t1)
    g0.close(); g1.close(); s.close()
    for a in pins:
        a.close()


def test_encode_decodes_and_matches_reference_bitrate():
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import encoder, ops
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    cap = refdata.capture("q4_44k_2ch", 6.0)
    s = ops.Setup("q4_44k_2ch")
    enc = encoder.Encoder(s, 4096, threads=4)
    tab = enc.encode([cap["pcm"], cap["pcm"][:, :44100]])
    pk = encoder.packets_of(tab, 0)
    ref = cap["packets"]
    assert len(pk) == len(ref)
    assert [p[0] for p in pk] == [r[0] for r in ref]          # the reference's packets, byte for byte
    R = reflib.RefLib("c")
    dec, _ = R.decode(cap["headers"], pk, cap["pcm"].shape[1] + 8192)
    decr, _ = R.decode(cap["headers"], ref, cap["pcm"].shape[1] + 8192)
    assert np.array_equal(dec, decr)
    assert len(encoder.packets_of(tab, 1)) > 20
    enc.close(); s.close()


@pytest.mark.parametrize("config,seconds", [("q4_44k_2ch", 6.0), ("q8_48k_6ch", 2.0), ("abr128_44k_2ch", 4.0)])
def test_gpu_stream_is_accepted_by_an_independent_decoder(config, seconds):
    """north_star: "the GPU-encoded stream decodes with stock libvorbis".  Beyond the reference's own decoder
    (above), the stream goes through csrc/oggframe.c into an .ogg file and FFmpeg -- its own Ogg demultiplexer
    (CRC-checking) and its own Vorbis decoder, tests/ffcheck.py -- must read it: same packets, the input's length,
    and PCM equal to the reference decoder's to float rounding.  Headers come from the product (Setup.header)."""
    from oracle import reflib
    from tests import ffcheck
    from vorbis_aotuv_lancer_b200 import encoder, ogg, ops
    if not ffcheck.available():
        pytest.skip("no FFmpeg libraries in this image")
    ch, rate = refdata.vblob.CATALOGUE[config][:2]
    pcm = refdata.synth(ch, rate, seconds, seed=3)
    s = ops.Setup(config)
    enc = encoder.Encoder(s, 2048, threads=2, chunks=1)
    pk = encoder.packets_of(enc.encode([pcm]), 0)
    headers = [s.header(i) for i in range(3)]
    got = ffcheck.read_ogg(ogg.mux(headers, pk), ch)
    assert [g[0] for g in got["packets"]] == [p[0] for p in pk]
    assert got["pcm"].shape == pcm.shape
    order = {6: [0, 2, 1, 5, 3, 4]}.get(ch, list(range(ch)))
    if reflib.available("c"):
        ref, _ = reflib.RefLib("c").decode(headers, pk, pcm.shape[1] + 8192)
        assert ref.shape == pcm.shape
        assert np.abs(got["pcm"] - ref[order]).max() < 2e-6
    # and it is the input, lossy-coded: SNR against the PCM that went in, per configuration
    sig = pcm[order].astype(np.float64)
    snr = 10 * np.log10((sig ** 2).sum() / ((got["pcm"] - sig) ** 2).sum())
    print("%s: FFmpeg-decoded GPU stream vs input PCM: %.1f dB" % (config, snr))
    assert snr > 8.0
    enc.close(); s.close()


def test_a_group_that_outgrows_its_batch_gets_a_larger_one():
    """Batch capacities are sized from the expected blocks per call; a group with more blocks than that (uneven streams,
    many groups over few streams) must get a larger batch, not an error -- and the same packets."""
    from vorbis_aotuv_lancer_b200 import encoder, ops
    pcm = refdata.synth(2, 44100, 2.0, seed=5)
    s = ops.Setup("q4_44k_2ch")
    one = encoder.Encoder(s, 4096, threads=2, chunks=1)
    want = encoder.packets_of(one.encode([pcm]), 0)
    one.close()
    enc = encoder.Encoder(s, 48, threads=2, chunks=3)          # 48 blocks expected in all: every group overflows
    tabs = enc.encode([pcm, pcm, pcm])
    for k in range(3):
        got = encoder.packets_of(tabs, k)
        assert [p[0] for p in got] == [p[0] for p in want]
    tabs = enc.encode([pcm, pcm, pcm])                          # and the grown batches serve the next call
    assert [p[0] for p in encoder.packets_of(tabs, 2)] == [p[0] for p in want]
    enc.close(); s.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_sixteen_bit_pcm_gives_the_stream_of_its_float_conversion(pinned):
    """vb200_segmenter_begin_s16: 16-bit PCM goes up as it is (half the bytes) and the device converts, sample / 32768.f
    -- what examples/encoder_example.c:183-190 does on the host.  Device timeline, blocks and packets must be exactly
    those of the float entry point fed with the converted samples; from page-locked memory (direct upload + conversion
    kernel) and from pageable memory (staged through the host timeline)."""
    from vorbis_aotuv_lancer_b200 import encoder, ops
    rng = np.random.default_rng(11)
    s = ops.Setup("q4_44k_2ch")
    streams16 = []
    for k, secs in enumerate((3.0, 1.3, 0.01)):
        f = refdata.synth(2, 44100, secs, seed=20 + k)
        q = np.clip(np.rint(f * 32767.0), -32768, 32767).astype(np.int16)
        if k == 1:
            q[:, 100:140] = rng.integers(-32768, 32768, (2, 40), dtype=np.int16)       # full-scale values incl. -32768
        streams16.append(q)
    as_float = [(q.astype(np.float32) / np.float32(32768.0)) for q in streams16]
    holders = []
    if pinned:
        ins = []
        for q in streams16:
            h = encoder.PinnedArray(q.shape[0], q.shape[1], np.int16)
            h.array[:] = q
            holders.append(h); ins.append(h.array)
    else:
        ins = streams16
    enc = encoder.Encoder(s, 4096, threads=2, chunks=2)
    want = enc.encode(as_float)
    got = enc.encode(ins)
    if pinned:
        assert enc.segmenters[0].direct
    for k in range(len(streams16)):
        a, b = encoder.packets_of(got, k), encoder.packets_of(want, k)
        assert len(a) == len(b) and [p[0] for p in a] == [p[0] for p in b] and [p[1] for p in a] == [p[1] for p in b]
    enc.close(); s.close()
    for h in holders:
        h.close()


def test_ragged_batch_of_degenerate_streams():
    """One call, streams of very different kinds and lengths: digital silence (every floor unused), one sample, less
    than a short block, a clipped square wave, a click train (a transient in every block), ordinary signal.  Each must
    come out as the reference encodes it alone: same packet count and granule positions, silence byte for byte."""
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    config = "q4_44k_2ch"
    ch, rate, q, abr = vblob.CATALOGUE[config]
    n = rate
    clicks = np.zeros((ch, n), np.float32)
    clicks[:, ::700] = 0.9
    square = np.where((np.arange(n) // 50) % 2 == 0, 0.99, -0.99).astype(np.float32)[None, :].repeat(ch, 0).copy()
    normal = refdata.synth(ch, rate, 2.0, seed=9)
    pcms = [np.zeros((ch, n), np.float32), normal[:, :1].copy(), normal[:, :63].copy(), square, clicks, normal, normal[:, :2049].copy()]
    R = reflib.RefLib("c")
    refs = []
    for p in pcms:
        h = R.open(ch, rate, q, abr)
        R.write(h, p)
        refs.append(R.packets(h))
        R.close(h)
    s = ops.Setup(config)
    enc = encoder.Encoder(s, 2048, threads=4, chunks=3)
    tabs = enc.encode(pcms)
    for i, ref in enumerate(refs):
        pk = encoder.packets_of(tabs, i)
        assert len(pk) == len(ref), i
        assert [p[1] for p in pk] == [r[1] for r in ref], i
        assert [p[2] for p in pk] == [r[2] for r in ref], i
        gb, rb = sum(len(p[0]) for p in pk), sum(len(r[0]) for r in ref)
        same = sum(1 for a, r in zip(pk, ref) if a[0] == r[0]) / len(ref)
        print("stream %d: %d packets, identical %.0f%%, bytes %d vs %d" % (i, len(ref), 100 * same, gb, rb))
        assert abs(gb - rb) <= max(8, 0.005 * rb), (i, gb, rb)
        # through the GPU segmenter and the default front end (K1x) every stream IS the reference's stream -- also the
        # clipped square wave (i == 3), whose coefficients sit on quantiser thresholds and which K1's 1e-5 MDCT
        # tolerance used to leave at 50-60 % identical packets
        assert same == 1.0 and gb == rb, (i, same, gb, rb)
    enc.close(); s.close()


def test_full_length_streams_properties():
    """BASELINE-size streams (60 s, q4 44.1 kHz stereo, the bench's synthetic signal): the result must not depend on
    how the streams are grouped into batches (one group vs three), a second run must reproduce it bit for bit, and
    the streams must BE the reference's streams: packet count, granule positions, every packet byte-identical,
    decoding to the same PCM."""
    import ctypes as C
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops
    from vorbis_aotuv_lancer_b200.lib import L
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    config = "q4_44k_2ch"
    ch, rate, q, abr = vblob.CATALOGUE[config]
    n, S = 60 * rate, 6
    pcms = []
    for i in range(S):
        a = np.empty((ch, n), np.float32)
        seed = L.vb200_splitmix64(C.c_uint64(88172645463325252 + i))
        L.vb200_synth_pcm(C.c_uint64(seed), ch, C.c_long(rate), C.c_long(n), a.ctypes.data_as(C.c_void_p), C.c_long(n))
        pcms.append(a)
    s = ops.Setup(config)

    def run(chunks):
        enc = encoder.Encoder(s, S * 4000, threads=4, chunks=chunks)
        tabs = enc.encode(pcms)
        out = [encoder.packets_of(tabs, i) for i in range(S)]
        enc.close()
        return out
    one, three, again = run(1), run(3), run(3)
    assert one == three, "grouping into batches changed the stream"
    assert three == again, "not reproducible"
    R = reflib.RefLib("c")
    for i in (0, S - 1):
        h = R.open(ch, rate, q, abr)
        R.write(h, pcms[i])
        ref, hdr = R.packets(h), R.headers(h)
        R.close(h)
        pk = one[i]
        assert len(pk) == len(ref) and [p[1] for p in pk] == [r[1] for r in ref]
        gb, rb = sum(len(p[0]) for p in pk), sum(len(r[0]) for r in ref)
        same = sum(1 for a, r in zip(pk, ref) if a[0] == r[0]) / len(ref)
        assert gb == rb and same == 1.0, (i, gb, rb, same)
        dec, _ = R.decode(hdr, pk, n + 8192)
        decr, _ = R.decode(hdr, ref, n + 8192)
        m = min(dec.shape[1], decr.shape[1])
        assert m > n - 4096
        assert np.array_equal(dec[:, :m], decr[:, :m])          # identical packets decode to identical PCM
        print("stream %d: %d packets, identical %.1f%%, bytes %d vs %d" % (i, len(ref), 100 * same, gb, rb))
    s.close()


def _silence_streams(ch, rate):
    """signals with stretches of DIGITAL silence: in managed mode a block whose nominal floor fit is NULL skips the
    sel = 2 mix (mapping0.c:1097), so lastmdct / tempmdct keep what an OLDER block left"""
    sig = refdata.synth(ch, rate, 4.0, seed=5)
    n = sig.shape[1]
    a = sig.copy(); a[:, :rate] = 0.0                                   # leading silence
    b = sig.copy(); b[:, rate + 3000:2 * rate + 1234] = 0.0             # a gap in the middle of the track
    c = sig.copy(); c[:, 3 * rate:] = 0.0                               # trailing silence
    d = sig.copy()                                                      # clicks out of silence, silence between short blocks
    d[:, :] = 0.0
    for k in range(6):
        p = 9000 + k * 17000
        d[:, p:p + 300] = sig[:, p:p + 300]
    e = np.zeros((ch, 3 * rate), np.float32)                            # silence, one burst, silence
    e[:, rate:rate + 20000] = sig[:, :20000]
    return [a, b, c, d, e]


@pytest.mark.parametrize("config", ["abr128_44k_2ch", "q4_44k_2ch"])
def test_streams_with_digital_silence_equal_the_reference(config):
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    ch, rate, q, abr = vblob.CATALOGUE[config]
    pcms = _silence_streams(ch, rate)
    R = reflib.RefLib("c")
    refs = []
    for p in pcms:
        h = R.open(ch, rate, q, abr)
        R.write(h, p)
        refs.append(R.packets(h))
        R.close(h)
    s = ops.Setup(config)
    for chunks in (1, 2):
        enc = encoder.Encoder(s, 4096, threads=4, chunks=chunks)
        tabs = enc.encode(pcms)
        for i, ref in enumerate(refs):
            pk = encoder.packets_of(tabs, i)
            assert len(pk) == len(ref), i
            diff = [j for j, (a, r) in enumerate(zip(pk, ref)) if a[0] != r[0]]
            print("%s stream %d (%d groups): %d packets, %d differ%s" % (config, i, chunks, len(ref), len(diff), (" first at %d" % diff[0]) if diff else ""))
            assert not diff, (config, i, chunks, len(diff), diff[:8])
        enc.close()
    s.close()
