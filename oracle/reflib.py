"""TEST INFRASTRUCTURE -- ctypes binding of the reference build in oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It drives the UNMODIFIED reference
(compiled from /root/reference/lib/*.c by oracle/Makefile) through the driver in
oracle/ref_driver.c and parses the A/B tap log of oracle/ref_taps.c.
"""
import ctypes as C
import os
import re
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")


def _tap_enum():
    src = open(os.path.join(HERE, "ref_taps.h")).read()
    body = src[src.index("enum {"):src.index("TAP_COUNT")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b(TAP_[A-Z0-9_]+)", body)
    out, v = {}, 1
    for n in names:
        out[n] = v
        v += 1
    out["TAP_COUNT"] = v
    return out


TAP = _tap_enum()
TAPNAME = {v: k for k, v in TAP.items()}


def build(reference="/root/reference"):
    """(Re)build oracle/_ref from the reference sources when they are present."""
    if not os.path.isdir(reference):
        return False
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "REF=" + reference], stdout=subprocess.DEVNULL)
    return True


def build_dropin(reference="/root/reference"):
    """Build the drop-in flavour (the reference bound to libvorbis_b200.so, ref_dropin.c); needs the product
    library to exist already."""
    if not os.path.isdir(reference):
        return False
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "REF=" + reference, "drop"], stdout=subprocess.DEVNULL)
    return True


def available(flavour="c"):
    return os.path.exists(os.path.join(REFDIR, "libvorbis_ref_%s.so" % flavour))


class RefLib:
    def __init__(self, flavour="c"):
        path = os.path.join(REFDIR, "libvorbis_ref_%s.so" % flavour)
        self.lib = L = C.CDLL(path)
        self.flavour = flavour
        if flavour == "drop":
            L.ref_dropin_constants.argtypes = [C.c_char_p, C.c_long]
            L.ref_dropin_batch_blocks.argtypes = [C.c_int]
            L.ref_dropin_error.restype = C.c_char_p
        L.ref_enc_open.restype = C.c_void_p
        L.ref_enc_open.argtypes = [C.c_int, C.c_long, C.c_float, C.c_long]
        L.ref_enc_close.argtypes = [C.c_void_p]
        L.ref_enc_header.restype = C.POINTER(C.c_ubyte)
        L.ref_enc_header.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_long)]
        L.ref_enc_write.restype = C.c_long
        L.ref_enc_write.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int]
        L.ref_enc_npackets.restype = C.c_long
        L.ref_enc_npackets.argtypes = [C.c_void_p]
        L.ref_enc_nblocks.restype = C.c_long
        L.ref_enc_nblocks.argtypes = [C.c_void_p]
        L.ref_enc_blocks.restype = C.POINTER(C.c_int)
        L.ref_enc_blocks.argtypes = [C.c_void_p]
        L.ref_enc_packets.restype = C.POINTER(C.c_ubyte)
        L.ref_enc_packets.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4 + [C.POINTER(C.c_long)]
        L.ref_decode.restype = C.c_long
        L.ref_export_setup.restype = C.POINTER(C.c_ubyte)
        L.ref_export_setup.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_tap_config.argtypes = [C.c_char_p, C.c_long, C.c_long]
        L.ref_tap_data.restype = C.POINTER(C.c_ubyte)
        L.ref_tap_data.argtypes = [C.POINTER(C.c_long)]

    # ---------------------------------------------------------------- encode
    def dropin_constants(self, raw_blob, batch_blocks=256):
        """drop flavour only: hand over a shipped setup blob (for the configuration-independent constant tables)
        and the number of blocks queued per vb200_batch_run."""
        self.lib.ref_dropin_constants(raw_blob, len(raw_blob))
        self.lib.ref_dropin_batch_blocks(int(batch_blocks))

    def dropin_error(self):
        return self.lib.ref_dropin_error().decode()

    def open(self, channels, rate, quality=0.4, managed_bitrate=0):
        h = self.lib.ref_enc_open(channels, rate, quality, managed_bitrate)
        if not h:
            raise RuntimeError("reference rejected the encode setup")
        return h

    def close(self, h):
        self.lib.ref_enc_close(h)

    def headers(self, h):
        out = []
        for i in range(3):
            n = C.c_long()
            p = self.lib.ref_enc_header(h, i, C.byref(n))
            out.append(bytes(bytearray(p[:n.value])))
        return out

    def write(self, h, pcm, eos=True, chunk=1024):
        """pcm: float32 [channels, nsamples] planar."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self.lib.ref_enc_write(h, pcm.ctypes.data, pcm.shape[1], pcm.shape[1], chunk, int(eos))

    def blocks(self, h):
        n = self.lib.ref_enc_nblocks(h)
        p = self.lib.ref_enc_blocks(h)
        return np.ctypeslib.as_array(p, shape=(n, 4)).copy()  # W lW nW blocktype

    def packets(self, h):
        off, nb, gp, eos = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        tot = C.c_long()
        p = self.lib.ref_enc_packets(h, C.byref(off), C.byref(nb), C.byref(gp), C.byref(eos), C.byref(tot))
        n = self.lib.ref_enc_npackets(h)
        if n == 0:
            return []
        data = bytes(bytearray(np.ctypeslib.as_array(p, shape=(tot.value,))))
        off = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_long)), shape=(n,)).copy()
        nb = np.ctypeslib.as_array(C.cast(nb, C.POINTER(C.c_long)), shape=(n,)).copy()
        gp = np.ctypeslib.as_array(C.cast(gp, C.POINTER(C.c_longlong)), shape=(n,)).copy()
        eos = np.ctypeslib.as_array(C.cast(eos, C.POINTER(C.c_int)), shape=(n,)).copy()
        return [(data[off[i]:off[i] + nb[i]], int(gp[i]), int(eos[i])) for i in range(n)]

    def export_setup(self, h):
        n = C.c_long()
        p = self.lib.ref_export_setup(h, C.byref(n))
        blob = bytes(bytearray(p[:n.value]))
        self.lib.ref_free(p)
        return blob

    # ---------------------------------------------------------------- decode
    def decode(self, headers, packets, max_samples):
        """packets: list of (bytes, granulepos, eos).  Returns float32 [channels, n]."""
        data = b"".join(p[0] for p in packets)
        nb = np.array([len(p[0]) for p in packets], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(nb)[:-1]]).astype(np.int64)
        gp = np.array([p[1] for p in packets], dtype=np.int64)
        ch, rate = C.c_int(), C.c_long()
        cap = int(max_samples)
        out = np.zeros((8, cap), dtype=np.float32)
        n = self.lib.ref_decode(headers[0], C.c_long(len(headers[0])), headers[1], C.c_long(len(headers[1])),
                                headers[2], C.c_long(len(headers[2])), data,
                                off.ctypes.data_as(C.c_void_p), nb.ctypes.data_as(C.c_void_p),
                                gp.ctypes.data_as(C.c_void_p), C.c_long(len(packets)),
                                out.ctypes.data_as(C.c_void_p), C.c_long(cap), C.byref(ch), C.byref(rate))
        if n < 0:
            raise RuntimeError("reference decoder failed: %d" % n)
        return out[:ch.value, :n].copy(), rate.value

    # ---------------------------------------------------------------- taps
    def tap_config(self, tags, lo=0, hi=1 << 60):
        en = bytearray(TAP["TAP_COUNT"])
        for t in tags:
            en[TAP[t] if isinstance(t, str) else t] = 1
        self.lib.ref_tap_config(bytes(en), lo, hi)
        self.lib.ref_tap_reset()

    def tap_all(self, lo=0, hi=1 << 60, exclude=()):
        self.tap_config([k for k in TAP if k != "TAP_COUNT" and k not in exclude], lo, hi)

    def tap_records(self):
        """Parse the tap log -> list of dicts(tag, seq, a0..a3, data)."""
        n = C.c_long()
        p = self.lib.ref_tap_data(C.byref(n))
        if n.value == 0:
            return []
        raw = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        out, pos = [], 0
        while pos < len(raw):
            hdr = raw[pos:pos + 32].view(np.int32)
            tag, seq, a0, a1, a2, dtype, count, a3 = [int(x) for x in hdr]
            pos += 32
            if dtype == 2:
                nbytes = (count + 3) & ~3
                data = raw[pos:pos + count].copy()
            else:
                nbytes = count * 4
                data = raw[pos:pos + nbytes].view(np.float32 if dtype == 0 else np.int32).copy()
            pos += nbytes
            out.append(dict(tag=TAPNAME[tag], seq=seq, a0=a0, a1=a1, a2=a2, a3=a3, data=data))
        return out


def group(records):
    """records -> {seq: {tag: [record, ...]}} preserving call order."""
    out = {}
    for r in records:
        out.setdefault(r["seq"], {}).setdefault(r["tag"], []).append(r)
    return out
