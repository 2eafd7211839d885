"""Host-side stream segmentation (csrc/segment.cu): PCM of a whole stream -> analysis blocks.
Mirrors vorbis_analysis_wrote + vorbis_analysis_blockout of the reference (block.c:482-812)."""
import ctypes as C

import numpy as np

from .lib import L, check
from .ops import Block

L.vb200_segments_blocks.restype = C.POINTER(Block)
L.vb200_segments_pcm.restype = C.POINTER(C.c_float)


def segment_stream(setup, pcm, stream_index=0, chunk=1024):
    """pcm float32 [channels, nsamples] -> (blocks: ctypes array of Block, timeline float32 [channels*total])."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    g = C.c_void_p()
    check(L.vb200_segment_stream(setup.h, pcm.ctypes.data_as(C.c_void_p), C.c_long(pcm.shape[1]), C.c_long(pcm.shape[1]),
                                 int(chunk), int(stream_index), C.byref(g)), "vb200_segment_stream")
    try:
        n = L.vb200_segments_count(g)
        src = L.vb200_segments_blocks(g)
        blocks = (Block * n)()
        C.memmove(blocks, src, C.sizeof(Block) * n)
        elems = C.c_long()
        p = L.vb200_segments_pcm(g, C.byref(elems))
        timeline = np.ctypeslib.as_array(p, shape=(elems.value,)).copy()
    finally:
        L.vb200_segments_free(g)
    return blocks, timeline
