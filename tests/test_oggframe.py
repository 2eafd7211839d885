"""Ogg framing (csrc/oggframe.c).  libogg is not in this image and the reference pins nothing
here, so the check is structural: an independent demultiplexer (written from RFC 3533, bitwise CRC) must get the
same packets back, and the paging rules the module states must hold.  On top of that FFmpeg's Ogg demultiplexer
and Vorbis decoder -- third-party implementations that share no code with libogg/libvorbis or with this repository,
present in this image inside the opencv wheel (tests/ffcheck.py) -- must accept the pages (they verify every CRC),
return the same packets and timestamps, and decode them to the reference decoder's PCM."""
import struct

import numpy as np
import pytest


def crc_bitwise(data):
    r = 0
    for b in data:
        r ^= b << 24
        for _ in range(8):
            r = ((r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if r & 0x80000000 else (r << 1) & 0xFFFFFFFF
    return r


def demux(stream):
    """-> pages [(flags, granulepos, serial, pageno, lacing, body)], packets [(bytes, page index where it ended)]"""
    pages, packets, cur, pos = [], [], b"", 0
    open_packet = False
    while pos < len(stream):
        assert stream[pos:pos + 4] == b"OggS" and stream[pos + 4] == 0
        flags = stream[pos + 5]
        gp, serial, pageno, crc, nseg = struct.unpack_from("<qIIIB", stream, pos + 6)
        lacing = list(stream[pos + 27:pos + 27 + nseg])
        blen = sum(lacing)
        page = bytearray(stream[pos:pos + 27 + nseg + blen])
        page[22:26] = b"\0\0\0\0"
        assert crc_bitwise(page) == crc, "page %d CRC" % pageno
        body = stream[pos + 27 + nseg:pos + 27 + nseg + blen]
        assert bool(flags & 1) == open_packet, "continued flag of page %d" % pageno
        o = 0
        for v in lacing:
            cur += body[o:o + v]
            o += v
            open_packet = v == 255
            if v < 255:
                packets.append((cur, len(pages)))
                cur = b""
        pages.append((flags, gp, serial, pageno, lacing, body))
        pos += 27 + nseg + blen
    assert not open_packet
    return pages, packets


def test_crc_known_answers():
    from vorbis_aotuv_lancer_b200 import ogg
    rng = np.random.default_rng(5)
    for n in (0, 1, 27, 255, 4096):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ogg.crc(d) == crc_bitwise(d)
    assert ogg.crc(b"\x00" * 8) == 0                       # initial value 0, no final xor
    assert ogg.crc(b"\x01") == 0x04C11DB7                  # the polynomial itself
    # published check value: the CRC catalogue lists CRC-32/CKSUM (same polynomial, init 0, unreflected, final
    # xor 0xFFFFFFFF) with check("123456789") = 0x765E7680; Ogg's variant has no final xor
    assert ogg.crc(b"123456789") == 0x765E7680 ^ 0xFFFFFFFF == 0x89A1897F


def test_pages_round_trip_and_follow_the_paging_rules():
    from vorbis_aotuv_lancer_b200 import ogg
    rng = np.random.default_rng(7)
    headers = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (30, 45, 3900)]
    sizes = list(rng.integers(1, 700, 400)) + [0, 255, 510, 254, 256, 70000, 5, 65025, 1]
    gp, packets = 0, []
    for i, n in enumerate(sizes):
        gp += 1024 if i % 3 else 576
        packets.append((rng.integers(0, 256, int(n), dtype=np.uint8).tobytes(), gp, int(i == len(sizes) - 1)))
    stream = ogg.mux(headers, packets, serial=0x1234)
    pages, got = demux(stream)
    assert [g[0] for g in got] == headers + [p[0] for p in packets]
    assert [p[3] for p in pages] == list(range(len(pages)))            # page sequence numbers
    assert all(p[2] == 0x1234 for p in pages)
    # first page: BOS, the identification header alone, granule 0
    assert pages[0][0] & 2 and pages[0][4] == [30] and pages[0][1] == 0
    assert not any(p[0] & 2 for p in pages[1:])
    # the other two headers are flushed before audio: audio packet 0 starts a page
    first_audio_page = got[3][1] if len(packets[0][0]) < 255 * 255 else None
    assert got[2][1] < got[3][1] or got[3][1] == first_audio_page
    hdr_last = got[2][1]
    assert all(idx > hdr_last for _, idx in got[3:])
    # EOS on the last page only
    assert pages[-1][0] & 4 and not any(p[0] & 4 for p in pages[:-1])
    # granule position of a page = that of the last packet finished on it, -1 if none finished
    ends = {}
    for k, (_, idx) in enumerate(got):
        ends[idx] = k
    for idx, p in enumerate(pages):
        if idx in ends and ends[idx] >= 3:
            assert p[1] == packets[ends[idx] - 3][1]
        elif idx not in ends:
            assert p[1] == -1
    # paging policy: an audio page that is not the last one is full (255 segments) or holds > 4096 bytes and >= 4
    # finished packets
    for idx, p in enumerate(pages[hdr_last + 1:-1], hdr_last + 1):
        done = sum(1 for v in p[4] if v < 255)
        assert len(p[4]) == 255 or (len(p[5]) > 4096 and done >= 4), (idx, len(p[4]), len(p[5]), done)
    assert all(len(p[4]) <= 255 for p in pages)


def test_reference_stream_muxes():
    """real headers + packets from the reference encoder (oracle/_ref) -> .ogg; the demultiplexer gives them back and
    the decoder input is unchanged"""
    import pytest
    from oracle import reflib
    from tests import refdata
    from vorbis_aotuv_lancer_b200 import ogg
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    cap = refdata.capture("q4_44k_2ch", 1.0)
    headers, packets = [bytes(h) for h in cap["headers"]], [(bytes(p[0]), p[1], p[2]) for p in cap["packets"]]
    assert headers[0][:7] == b"\x01vorbis" and headers[1][:7] == b"\x03vorbis" and headers[2][:7] == b"\x05vorbis"
    stream = ogg.mux(headers, packets)
    pages, got = demux(stream)
    assert [x[0] for x in got] == headers + [p[0] for p in packets]
    assert pages[-1][1] == packets[-1][1] and pages[-1][0] & 4
    assert len(pages) >= 4


@pytest.mark.parametrize("config,seconds", [("q4_44k_2ch", 3.0), ("q8_48k_6ch", 1.0), ("q4_44k_1ch", 1.0),
                                            ("abr128_44k_2ch", 2.0)])
def test_ffmpeg_reads_the_pages_and_decodes_the_stream(config, seconds):
    """csrc/oggframe.c against an independent reader: FFmpeg's demultiplexer (CRC-checking) returns the packets that
    went in, packet k starts where packet k-1's granule position ends, the stream ends at the input length, and
    FFmpeg's own Vorbis decoder agrees with the reference's decoder (vorbis_synthesis) to float rounding."""
    from oracle import reflib
    from tests import ffcheck, refdata
    from vorbis_aotuv_lancer_b200 import ogg
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    if not ffcheck.available():
        pytest.skip("no FFmpeg libraries in this image")
    cap = refdata.capture(config, seconds)
    headers, packets = [bytes(h) for h in cap["headers"]], [(bytes(p[0]), p[1], p[2]) for p in cap["packets"]]
    ch, n = cap["pcm"].shape
    got = ffcheck.read_ogg(ogg.mux(headers, packets), ch)
    assert [g[0] for g in got["packets"]] == [p[0] for p in packets]
    # timestamps: FFmpeg derives the pts of the first packet of a page from THAT page's granule position (minus the
    # durations of the page's packets); it must be the granule position the previous packet ended on.  (Inside a
    # page FFmpeg interpolates with its own duration rule, which is off by design at a long->short transition.)
    _, mine = demux(ogg.mux(headers, packets))
    first = [k for k in range(1, len(packets)) if mine[k + 3][1] != mine[k + 2][1]]
    assert len(first) >= 2
    assert [got["packets"][k][1] for k in first] == [packets[k - 1][1] for k in first]
    same = sum(1 for k in range(1, len(packets)) if got["packets"][k][1] == packets[k - 1][1])
    assert same >= 0.9 * (len(packets) - 1)
    ref, _ = reflib.RefLib("c").decode(headers, packets, n + 8192)
    assert got["pcm"].shape == ref.shape == (ch, n)        # the last page's granule position trims the tail
    # FFmpeg hands 5.1 out in its own channel order (FL FR FC LFE BL BR), Vorbis order is FL FC FR BL BR LFE
    order = {6: [0, 2, 1, 5, 3, 4]}.get(ch, list(range(ch)))
    err = np.abs(got["pcm"] - ref[order]).max()
    assert err < 2e-6, err


def test_ffmpeg_rejects_a_damaged_page():
    """the check above has teeth: one flipped body bit and FFmpeg no longer returns the same packets"""
    from oracle import reflib
    from tests import ffcheck, refdata
    from vorbis_aotuv_lancer_b200 import ogg
    if not reflib.available("c"):
        pytest.skip("oracle/_ref not present")
    if not ffcheck.available():
        pytest.skip("no FFmpeg libraries in this image")
    cap = refdata.capture("q4_44k_2ch", 3.0)
    headers, packets = [bytes(h) for h in cap["headers"]], [(bytes(p[0]), p[1], p[2]) for p in cap["packets"]]
    stream = bytearray(ogg.mux(headers, packets))
    stream[len(stream) // 2] ^= 0x10
    try:
        got = ffcheck.read_ogg(bytes(stream), 2, decode=False)
    except AssertionError:
        return
    assert [g[0] for g in got["packets"]] != [p[0] for p in packets]
