"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/vorbis_b200.h
declares, answers host-side queries, and REFUSES compute without a GPU (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest

from vorbis_aotuv_lancer_b200 import blob as vblob
from vorbis_aotuv_lancer_b200 import lib, ops


def test_every_declared_symbol_is_exported():
    names = lib.declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib.L, n)]
    assert not missing, missing


@pytest.mark.parametrize("name", sorted(vblob.CATALOGUE))
def test_setup_blobs_load_host_only(name):
    ch, rate, q, abr = vblob.CATALOGUE[name]
    s = ops.Setup(name, device=-1)
    assert (s.channels, s.rate, s.blocksizes) == (ch, rate, (256, 2048))
    assert s.managed == (1 if abr else 0)
    for i in range(3):
        assert len(s.header(i)) > 0
    assert s.header(0)[:7] == b"\x01vorbis"
    s.close()


def test_malformed_blob_is_rejected():
    h = C.c_void_p()
    assert lib.L.vb200_setup_create(b"\0" * 64, C.c_size_t(64), -1, C.byref(h)) == -131
    raw = bytearray(vblob.load("q4_44k_2ch"))
    raw[0] ^= 0xFF
    assert lib.L.vb200_setup_create(bytes(raw), C.c_size_t(len(raw)), -1, C.byref(h)) == -131


def test_compute_fails_loudly_without_a_gpu():
    if lib.L.vb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.VB200Error):
        ops.Setup("q4_44k_2ch", device=0)
    s = ops.Setup("q4_44k_2ch", device=-1)
    with pytest.raises(lib.VB200Error) as e:
        ops.forward_transform(s, 1, np.zeros((1, 2048), np.float32), [1], [1], [3], [3])
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(lib.VB200Error):
        ops.Stream(s)
    with pytest.raises(lib.VB200Error):
        ops.Batch(s, 16)
    s.close()


def test_synthetic_pcm_is_deterministic():
    from tests import refdata
    a, b = refdata.synth(2, 44100, 0.5), refdata.synth(2, 44100, 0.5)
    assert np.array_equal(a, b) and np.abs(a).max() <= 0.99 and a.std() > 0.05
    assert not np.array_equal(a, refdata.synth(2, 44100, 0.5, seed=7))


@pytest.mark.parametrize("n", [256, 2048])
def test_exact_transform_tables_equal_the_oracle_tables(n):
    """the tables K1x reads (product code, k_transform_exact.cu) against the oracle's restatement of mdct_init /
    drfti1, which tests/test_oracle_port.py pins bit-exactly to the reference through the transforms built on them"""
    import ctypes as C
    import numpy as np
    from vorbis_aotuv_lancer_b200.lib import L
    from oracle import port
    port.build()
    O = C.CDLL(port.LIBPATH) if hasattr(port, "LIBPATH") else C.CDLL(os.path.join(os.path.dirname(port.__file__), "liboracle.so"))
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    T1, T2 = np.zeros(n + n // 4, np.float32), np.zeros(n + n // 4, np.float32)
    B1, B2 = np.zeros(n // 4, np.int32), np.zeros(n // 4, np.int32)
    L.vb200_exact_tables_mdct.restype = None
    L.vb200_exact_tables_mdct(n, T1.ctypes.data_as(fp), B1.ctypes.data_as(ip))
    O.orc_mdct_tables(n, T2.ctypes.data_as(fp), B2.ctypes.data_as(ip))
    assert np.array_equal(T1.view(np.uint32), T2.view(np.uint32)) and np.array_equal(B1, B2)
    W1, W2 = np.zeros(2 * n, np.float32), np.zeros(2 * n, np.float32)
    fac = np.zeros(32, np.int32)
    L.vb200_exact_tables_drft.restype = None
    L.vb200_exact_tables_drft(n, W1.ctypes.data_as(fp))
    O.orc_drft_tables(n, W2.ctypes.data_as(fp), fac.ctypes.data_as(ip))
    assert np.array_equal(W1.view(np.uint32), W2.view(np.uint32))
    assert list(fac[:2]) == [n, 6 if n == 2048 else 4] and list(fac[2:2 + fac[1]]) == ([2, 4, 4, 4, 4, 4] if n == 2048 else [4, 4, 4, 4])
