"""The call a user makes: PCM of many independent streams in, Vorbis audio packets out.

Host side (north_star): stream segmentation (envelope / block-size decision, csrc/segment.cu) and,
in managed mode, the bitrate manager; device side: everything mapping0_forward does, batched over
all blocks of all streams (csrc/pipeline.cu), including the bit packing."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import ops, segment
from .lib import L, check


BLOCK_DTYPE = np.dtype([("stream", "<i4"), ("W", "<i4"), ("lW", "<i4"), ("nW", "<i4"), ("blocktype", "<i4"), ("eofflag", "<i4"),
                        ("granulepos", "<i8"), ("pcm_offset", "<i8"), ("pcm_stride", "<i8")])
assert BLOCK_DTYPE.itemsize == C.sizeof(ops.Block)


L.vb200_segmenter_blocks.restype = C.POINTER(ops.Block)
L.vb200_segmenter_device_pcm.restype = C.c_void_p
L.vb200_segmenter_host_pcm.restype = C.POINTER(C.c_float)


class Segmenter:
    """vb200_segmenter: segmentation of many streams with the detector's heavy half on the GPU; leaves the
    PCM timeline of all streams in HBM for Batch.run(pcm_on_device=True)."""

    def __init__(self, setup):
        self.setup = setup
        h = C.c_void_p()
        check(L.vb200_segmenter_create(setup.h, C.byref(h)), "vb200_segmenter_create")
        self.h = h

    def close(self):
        if self.h:
            L.vb200_segmenter_destroy(self.h)
            self.h = None

    def run(self, pcms, chunk=1024, threads=8):
        """pcms: list of float32 [channels, nsamples].  -> (Block array view, device address, elems)"""
        pcms = [np.ascontiguousarray(p, np.float32) for p in pcms]
        n = len(pcms)
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pcms])
        ns = (C.c_long * n)(*[p.shape[1] for p in pcms])
        check(L.vb200_segmenter_run(self.h, n, ptrs, ns, ns, int(chunk), int(threads)), "vb200_segmenter_run")
        nb = L.vb200_segmenter_count(self.h)
        blocks = C.cast(L.vb200_segmenter_blocks(self.h), C.POINTER(ops.Block * nb)).contents
        elems = C.c_size_t()
        dptr = L.vb200_segmenter_device_pcm(self.h, C.byref(elems))
        return blocks, dptr, elems.value

    def host_timeline(self, elems):
        return np.ctypeslib.as_array(L.vb200_segmenter_host_pcm(self.h), shape=(elems,))


class Encoder:
    def __init__(self, setup, max_blocks, threads=8):
        self.setup = setup
        self.batch = ops.Batch(setup, max_blocks)
        self.threads = threads
        self.pool = ThreadPoolExecutor(max_workers=threads)
        self.segmenter = Segmenter(setup)

    def close(self):
        self.batch.close()
        self.segmenter.close()
        self.pool.shutdown()

    def segment(self, pcms):
        """pcms: list of float32 [channels, nsamples] -> (Block array, one timeline buffer, blocks per stream)."""
        s = self.setup
        res = list(self.pool.map(lambda a: segment.segment_stream(s, a[1], stream_index=a[0]), enumerate(pcms)))
        n = sum(len(r[0]) for r in res)
        blocks = (ops.Block * n)()
        timeline = np.empty(sum(r[1].size for r in res), np.float32)
        view = np.frombuffer(blocks, dtype=BLOCK_DTYPE)
        pos, base, counts = 0, 0, []
        for blk, tl in res:
            C.memmove(C.byref(blocks, pos * C.sizeof(ops.Block)), blk, C.sizeof(ops.Block) * len(blk))
            view[pos:pos + len(blk)]["pcm_offset"] += base
            timeline[base:base + tl.size] = tl
            pos += len(blk)
            base += tl.size
            counts.append(len(blk))
        return blocks, timeline, counts

    def encode(self, pcms, streams=None):
        """Host PCM of many streams -> packet table.  Segmentation (detector on the GPU, LPC + window
        sequencing on the host) uploads the PCM once; the analysis batch then reads it in place."""
        blocks, dptr, elems = self.segmenter.run(pcms, threads=self.threads)
        own = streams is None
        if own:
            streams = [ops.Stream(self.setup) for _ in pcms]
        self.batch.run(streams, blocks, dptr, pcm_on_device=True, pcm_elems=elems)
        tab = table(self.batch)
        if own:
            for st in streams:
                st.close()
        return tab


def table(batch):
    """All packets of the last run: dict(data=bytes-like base, off, bytes, granulepos, eos, stream) as numpy arrays."""
    base = C.POINTER(C.c_ubyte)()
    off, nb, gp, eos, st = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = check(L.vb200_batch_packet_table(batch.h, C.byref(base), C.byref(off), C.byref(nb), C.byref(gp), C.byref(eos), C.byref(st)))
    if n == 0:
        return dict(n=0, total=0)
    as_arr = lambda p, t: np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(n,))  # noqa: E731
    o, b_ = as_arr(off, C.c_size_t), as_arr(nb, C.c_size_t)
    return dict(n=n, base=base, off=o, bytes=b_, granulepos=as_arr(gp, C.c_int64), eos=as_arr(eos, C.c_int),
                stream=as_arr(st, C.c_int), total=int(b_.sum()))


def packets_of(tab, stream):
    """[(bytes, granulepos, eos)] of one stream from a packet table."""
    out = []
    for i in np.nonzero(tab["stream"] == stream)[0]:
        out.append((C.string_at(C.addressof(tab["base"].contents) + int(tab["off"][i]), int(tab["bytes"][i])),
                    int(tab["granulepos"][i]), int(tab["eos"][i])))
    return out
