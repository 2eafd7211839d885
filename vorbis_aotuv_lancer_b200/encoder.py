"""The call a user makes: PCM of many independent streams in, Vorbis audio packets out.

Host side (north_star): stream segmentation (envelope / block-size decision, csrc/segment.cu) and,
in managed mode, the bitrate manager; device side: everything mapping0_forward does, batched over
all blocks of all streams (csrc/pipeline.cu), including the bit packing."""
import ctypes as C
import os
import sys
import time
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


class PinnedArray:
    """float32 (or int16) [channels, nsamples] in page-locked host memory (vb200_pinned_alloc): PCM held in one of
    these goes to the device straight from where it lies."""

    def __init__(self, channels, nsamples, dtype=np.float32):
        p = C.c_void_p()
        dt = np.dtype(dtype)
        check(L.vb200_pinned_alloc(C.c_size_t(dt.itemsize * channels * max(1, nsamples)), C.byref(p)), "vb200_pinned_alloc")
        self.p = p
        ctype = C.c_float if dt == np.float32 else C.c_int16
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(channels, nsamples))

    def close(self):
        if self.p:
            self.array = None
            L.vb200_pinned_free(self.p)
            self.p = None


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

    def begin(self, pcms, chunk=1024, threads=8):
        """Enqueue the uploads and the detector kernels of these streams; returns at once.  pcms: float32
        [channels, nsamples] each -- or int16 each (16-bit PCM, value = sample / 32768: half the upload)."""
        s16 = len(pcms) > 0 and all(np.asarray(p).dtype == np.int16 for p in pcms)
        self._pcms = [np.ascontiguousarray(p, np.int16 if s16 else np.float32) for p in pcms]       # kept alive until finish()
        n = len(self._pcms)
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in self._pcms])
        ns = (C.c_long * n)(*[p.shape[1] for p in self._pcms])
        if s16:
            check(L.vb200_segmenter_begin_s16(self.h, n, ptrs, ns, ns, int(chunk), int(threads)), "vb200_segmenter_begin_s16")
        else:
            check(L.vb200_segmenter_begin(self.h, n, ptrs, ns, ns, int(chunk), int(threads)), "vb200_segmenter_begin")

    def finish(self):
        """-> (Block array view, device address, elems)"""
        check(L.vb200_segmenter_finish(self.h), "vb200_segmenter_finish")
        self._pcms = None
        nb = L.vb200_segmenter_count(self.h)
        blocks = C.cast(L.vb200_segmenter_blocks(self.h), C.POINTER(ops.Block * nb)).contents
        elems = C.c_size_t()
        dptr = L.vb200_segmenter_device_pcm(self.h, C.byref(elems))
        return blocks, dptr, elems.value

    def run(self, pcms, chunk=1024, threads=8):
        """pcms: list of float32 [channels, nsamples].  -> (Block array view, device address, elems)"""
        self.begin(pcms, chunk, threads)
        return self.finish()

    @property
    def direct(self):
        return bool(L.vb200_segmenter_direct(self.h))

    def device_timeline(self, elems):
        out = np.empty(elems, np.float32)
        check(L.vb200_segmenter_read_pcm(self.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(elems)), "vb200_segmenter_read_pcm")
        return out

    def host_timeline(self, elems):
        return np.ctypeslib.as_array(L.vb200_segmenter_host_pcm(self.h), shape=(elems,))


def group_weights(chunks):
    """Relative sizes of the stream groups one encode() call pipelines: 1, 2, 4, 6, then equal shares of 8 -- a short
    ramp so that the device has work after one stream's upload, then groups small enough that little is left to do
    when the last upload ends (measured on a B200, 64 streams: 8 groups 83.6 k audio-s/s, 10 groups 86.5 k; with the
    round-2 kernels 8 / 10 / 12 / 16 / 20 / 24 / 32 groups: 86.7 / 87.0 / 88.7 / 88.9 / 89.4 / 89.7 / 87.2 k)."""
    ramp = [1.0, 2.0, 4.0, 6.0]
    return [ramp[g] if g < len(ramp) else 8.0 for g in range(max(1, int(chunks)))]


def group_cuts(n, weights):
    """Split streams 0..n-1 into consecutive groups with sizes proportional to `weights` (every group non-empty;
    fewer groups than weights when there are fewer streams).  -> [(first, last+1)]"""
    k = min(len(weights), n)
    w = list(weights[:k])
    acc, cuts = 0.0, [0]
    for g in range(k):
        acc += w[g]
        cuts.append(max(cuts[-1] + 1, min(n - (k - 1 - g), int(round(n * acc / sum(w))))))
    if k:
        cuts[-1] = n
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


class Encoder:
    """PCM of many independent streams in, packets out.  The streams are processed in `chunks` groups, each with
    its own segmenter and batch (own CUDA streams): while the kernels of group g run, the host segments group
    g+1 and its PCM goes up -- the host stage and the device stage overlap."""

    def __init__(self, setup, max_blocks, threads=8, chunks=6, nstreams=None):
        """max_blocks: blocks of one encode() call over all streams; nstreams: how many streams a call will bring, if
        known -- the groups' batch capacities then follow the actual group sizes (with many groups and few streams a
        group's share of the streams is far from its share of the weights).  A group that still outgrows its batch
        gets a larger one on the spot (encode)."""
        self.setup = setup
        self.threads = threads
        self.chunks = max(1, int(chunks))
        # group sizes ramp up (1, 2, 4, 6, then equal shares): the device gets its first work after one stream's upload
        # and never starves while the rest arrives; 25 % slack for uneven stream lengths
        env = os.environ.get("VB200_GROUP_WEIGHTS")
        self.weights = [float(x) for x in env.split(",")] if env else group_weights(self.chunks)
        self.chunks = len(self.weights)
        share = max(self.weights) / sum(self.weights)
        if nstreams:
            cuts = group_cuts(int(nstreams), self.weights)
            share = max(b - a for a, b in cuts) / float(nstreams)
        per = int(1.25 * int(max_blocks) * share) + 64 if self.chunks > 1 else int(max_blocks)
        self.capacity = [per] * self.chunks
        self.batches = [ops.Batch(setup, per) for _ in range(self.chunks)]
        for rank, b in enumerate(self.batches):      # earlier groups are served first, so their packets are out first
            check(L.vb200_batch_set_rank(b.h, rank), "vb200_batch_set_rank")
        self.segmenters = [Segmenter(setup) for _ in range(self.chunks)]
        for a, b in zip(self.segmenters[:-1], self.segmenters[1:]):
            check(L.vb200_segmenter_after(b.h, a.h), "vb200_segmenter_after")
        self.batch = self.batches[0]
        self.segmenter = self.segmenters[0]
        self.pool = ThreadPoolExecutor(max_workers=threads)
        self.begin_pool = ThreadPoolExecutor(max_workers=1)      # one thread: the groups' uploads must be queued in order
        self.submit_pool = ThreadPoolExecutor(max_workers=1)     # one thread: batches are launched in group order

    def close(self):
        for b in self.batches:
            b.close()
        for g in self.segmenters:
            g.close()
        self.pool.shutdown()
        self.begin_pool.shutdown()
        self.submit_pool.shutdown()

    def segment(self, pcms):
        """Host-only segmentation: list of float32 [channels, nsamples] -> (Block array, one timeline buffer, counts)."""
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
        """Host PCM of many streams -> list of (packet table, first stream index), one per group."""
        n = len(pcms)
        own = streams is None
        if own:
            streams = [ops.Stream(self.setup) for _ in pcms]
        groups = group_cuts(n, self.weights)
        trace = os.environ.get("VB200_TRACE")
        t0 = time.perf_counter()
        marks = []
        # all uploads and detector kernels are queued up front, in group order, so the copy engine never waits for the
        # host stage of a group -- by a helper thread, so that the first group's host stage does not wait for the
        # enqueueing of the last one either
        begun = [self.begin_pool.submit(self.segmenters[g].begin, pcms[a:b], 1024, self.threads) for g, (a, b) in enumerate(groups)]
        # the host stage of a group is two halves -- finish (window sequencing, tails) and submit (descriptor pre-pass,
        # launches); this thread does the finishes, a second one the submits, so group g+1 is sequenced while group g
        # is being launched
        subs = []
        for g, (a, b) in enumerate(groups):
            begun[g].result()
            blocks, dptr, elems = self.segmenters[g].finish()
            if len(blocks) > self.capacity[g]:          # rare: an uneven group; never silently truncate
                self.batches[g].close()
                self.capacity[g] = int(1.25 * len(blocks)) + 64
                self.batches[g] = ops.Batch(self.setup, self.capacity[g])
                check(L.vb200_batch_set_rank(self.batches[g].h, g), "vb200_batch_set_rank")
                if g == 0:
                    self.batch = self.batches[0]
            t1 = time.perf_counter()
            subs.append(self.submit_pool.submit(self.batches[g].submit, streams[a:b], blocks, dptr, elems))
            marks.append((t1 - t0,))
        for f in subs:
            f.result()
        marks.append((time.perf_counter() - t0,))
        out = []
        for g, (a, b) in enumerate(groups):
            self.batches[g].collect()
            marks.append((time.perf_counter() - t0,))
            out.append((table(self.batches[g]), a))
        if trace:
            print("[vb200 encode] ms since start: " + " ".join("/".join("%.1f" % (x * 1e3) for x in m) for m in marks), file=sys.stderr)
        if own:
            for st in streams:
                st.close()
        return out


def total_bytes(tabs):
    return sum(t["total"] for t, _ in tabs)


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


def packets_of(tabs, stream):
    """[(bytes, granulepos, eos)] of one stream from the result of Encoder.encode (or a single table)."""
    if isinstance(tabs, dict):
        tabs = [(tabs, 0)]
    tab = None
    for t, first in sorted(tabs, key=lambda x: x[1]):
        if first <= stream:
            tab, rel = t, stream - first
    if tab is None or not tab["n"]:
        return []
    stream = rel
    out = []
    for i in np.nonzero(tab["stream"] == stream)[0]:
        out.append((C.string_at(C.addressof(tab["base"].contents) + int(tab["off"][i]), int(tab["bytes"][i])),
                    int(tab["granulepos"][i]), int(tab["eos"][i])))
    return out
