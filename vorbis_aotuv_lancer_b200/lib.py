"""ctypes loader of the C-ABI library (include/vorbis_b200.h).

There is no CPU fallback: if libvorbis_b200.so is missing this module raises on
import, and every compute call returns an error when no CUDA device is usable.
"""
import ctypes as C
import os
import re
import subprocess

# The encoder keeps a dozen CUDA streams busy (one per segmenter and per batch of every stream group); with the default 8
# hardware queues they would alias and serialise each other.  Read by the CUDA driver when the context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libvorbis_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "vorbis_b200.h")


def build(verbose=False):
    """Compile every CUDA source for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-j8", "-C", os.path.join(HERE, "csrc")], stdout=out)
    return LIBPATH


def declared_symbols():
    """Every function the public header declares."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vb200_[a-z0-9_]+)\s*\(", src)))


class VB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIBPATH):
        raise ImportError("libvorbis_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                          "there is no CPU fallback")
    L = C.CDLL(LIBPATH)
    L.vb200_version.restype = C.c_char_p
    L.vb200_last_error.restype = C.c_char_p
    L.vb200_launch_count.restype = C.c_long
    L.vb200_setup_rate.restype = C.c_long
    L.vb200_splitmix64.restype = C.c_uint64
    L.vb200_splitmix64.argtypes = [C.c_uint64]
    if hasattr(L, "vb200_batch_kernel_ms"):
        L.vb200_batch_kernel_ms.restype = C.c_float
    return L


L = _load()


def check(ret, what=""):
    if ret < 0:
        raise VB200Error("%s failed (%d): %s" % (what, ret, L.vb200_last_error().decode()))
    return ret
