"""The drop-in boundary, live: the reference's own encoder loop --
vorbis_analysis_buffer/_wrote/_blockout, vorbis_analysis, vorbis_bitrate_addblock/_flushpacket, all unmodified
reference objects -- with _mapping_P[0]->forward bound to libvorbis_b200.so (oracle/ref_dropin.c).  The stream it
returns must be what our own public call produces for the same PCM, byte for byte, and must relate to the stock
reference stream like every other whole-path result does."""
import numpy as np
import pytest

from tests import refdata

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,seconds,batch", [("q4_44k_2ch", 6.0, 64), ("q8_48k_6ch", 2.0, 37), ("abr128_44k_2ch", 3.0, 256),
                                                  ("q4_32k_2ch", 3.0, 100), ("abr256_48k_6ch", 1.5, 50), ("q0_44k_2ch", 3.0, 4096)])
def test_reference_encoder_loop_over_the_b200_mapping(config, seconds, batch):
    from oracle import reflib
    from vorbis_aotuv_lancer_b200 import blob as vblob, encoder, ops
    if not reflib.available("drop"):
        pytest.skip("oracle/_ref/libvorbis_ref_drop.so not present")
    ch, rate, q, abr = vblob.CATALOGUE[config]
    cap = refdata.capture(config, seconds)                 # stock reference: PCM, headers, packets
    pcm = cap["pcm"]
    R = reflib.RefLib("drop")
    R.dropin_constants(vblob.load("q4_44k_2ch"), batch_blocks=batch)
    h = R.open(ch, rate, q, abr)
    R.write(h, pcm, eos=True)
    assert R.dropin_error() == ""
    pk = R.packets(h)
    R.close(h)
    ref = cap["packets"]
    assert len(pk) == len(ref)
    assert [p[1] for p in pk] == [r[1] for r in ref]       # granule positions
    assert pk[-1][2] == 1 and all(p[2] == 0 for p in pk[:-1])
    gb, rb = sum(len(p[0]) for p in pk), sum(len(r[0]) for r in ref)
    same = sum(1 for a, r in zip(pk, ref) if a[0] == r[0]) / len(ref)
    print("\n%s through the reference API: identical packets %.1f%%, bytes %d vs %d" % (config, 100 * same, gb, rb))
    assert gb == rb and same == 1.0      # the default front end (K1x) performs the reference's own float operations
    # our own public call on the same PCM
    s = ops.Setup(config)
    enc = encoder.Encoder(s, 4096, threads=2, chunks=1)
    mine = encoder.packets_of(enc.encode([pcm]), 0)
    assert [p[0] for p in mine] == [p[0] for p in pk]
    # and the reference decoder takes it
    Rc = reflib.RefLib("c")
    dec, _ = Rc.decode(cap["headers"], pk, pcm.shape[1] + 8192)
    assert dec.shape[1] > pcm.shape[1] - 4096
    enc.close(); s.close()
