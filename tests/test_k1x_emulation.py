"""K1x (the operation-exact front end, csrc/k_transform_exact.cu) run on the CPU with one OS thread per CUDA thread of
a block and a pthread barrier for __syncthreads (tests/emu/k1x_emu.cpp) -- against the reference's own MDCT, logmdct,
logfft and block maximum, bit for bit.  Covers the kernel's index arithmetic, thread-to-item mappings, stage order,
barrier placement, tables and buffer reuse without a GPU; the `-m gpu` tests (test_gpu_transform.py) cover the real
launch."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import refdata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "emu", "libk1x_emu.so")
fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "k1x_emu.cpp")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O1", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-w", "-pthread",
                           "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "vorbis_aotuv_lancer_b200", "csrc"), "-o", SO, src])
    return C.CDLL(SO)


@pytest.mark.parametrize("config", ["q4_44k_2ch", "q8_48k_6ch"])
def test_emulated_kernel_equals_the_reference(emu, config):
    cap = refdata.capture(config, 2.0)
    seen = {0: 0, 1: 0}
    for seq in sorted(cap["blocks"]):
        b = cap["blocks"][seq]
        if "TAP_BLOCK" not in b:
            continue
        W = int(b["TAP_BLOCK"][0]["data"][0])
        if seen[W] >= 12:
            continue
        seen[W] += 1
        for c in range(len(b["TAP_WINDOWED"])):
            w = np.ascontiguousarray(b["TAP_WINDOWED"][c]["data"], np.float32)
            n = len(w)
            gm, lf, lm = (np.empty(n // 2, np.float32) for _ in range(3))
            amp = C.c_float()
            assert emu.k1x_emu(n, w.ctypes.data_as(fp), gm.ctypes.data_as(fp), lf.ctypes.data_as(fp), lm.ctypes.data_as(fp), C.byref(amp)) == 0
            assert np.array_equal(gm, b["TAP_MDCT"][c]["data"]), (seq, c, "mdct")
            assert np.array_equal(lm, b["TAP_NM_LOGMDCT"][c]["data"]), (seq, c, "logmdct")
            assert np.array_equal(lf, b["TAP_TM_LOGFFT"][c]["data"]), (seq, c, "logfft")
            assert np.float32(amp.value) == b["TAP_TM_SCAL"][c]["data"][1], (seq, c, "ampmax")
    assert seen[0] > 0 and seen[1] > 0


def test_emulated_kernel_has_no_data_race(tmp_path):
    """The same emulation under ThreadSanitizer: every shared-memory hand-over between the kernel's stages must be
    ordered by a barrier."""
    exe = str(tmp_path / "k1x_tsan")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-O1", "-g", "-fsanitize=thread", "-fno-fast-math", "-ffp-contract=off", "-std=c++17", "-w", "-pthread",
                        "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "vorbis_aotuv_lancer_b200", "csrc"),
                        "-o", exe, os.path.join(ROOT, "tests", "emu", "k1x_tsan_main.cpp"), os.path.join(ROOT, "tests", "emu", "k1x_emu.cpp")],
                       capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in (r.stderr or "").lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "ThreadSanitizer" not in run.stderr, run.stderr[-3000:]
    assert "n=2048" in run.stdout and "n=256" in run.stdout
