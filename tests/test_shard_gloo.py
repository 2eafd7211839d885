"""The N>1 path on CPU: two gloo ranks shard the streams `i mod G`, each runs the host side of the
path (segmentation) on its shard, and the gathered result equals the single-process one.  No
data-path collective exists, so this is the whole multi-rank logic."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import refdata
    from vorbis_aotuv_lancer_b200 import ops, segment, shard
    s = ops.Setup("q4_44k_2ch", device=-1)
    mine = shard.shard_streams(n_streams, rank, world)
    local = {}
    for i in mine:
        pcm = refdata.synth(2, 44100, 1.0, seed=100 + i)
        blocks, timeline = segment.segment_stream(s, pcm, stream_index=i)
        local[i] = (len(blocks), int(sum(b.W for b in blocks)), float(np.abs(timeline).sum()))
    allr = shard.gather_stats(local, world)
    dist.barrier()
    if rank == 0:
        q.put(allr)
    dist.destroy_process_group()


def test_two_ranks_cover_all_streams_once():
    from tests import refdata
    from vorbis_aotuv_lancer_b200 import ops, segment, shard
    n_streams, world = 5, 2
    assert sorted(shard.shard_streams(n_streams, 0, world) + shard.shard_streams(n_streams, 1, world)) == list(range(n_streams))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    allr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for d in allr:
        assert not (set(d) & set(merged))
        merged.update(d)
    assert sorted(merged) == list(range(n_streams))
    s = ops.Setup("q4_44k_2ch", device=-1)
    for i in range(n_streams):
        blocks, timeline = segment.segment_stream(s, refdata.synth(2, 44100, 1.0, seed=100 + i), stream_index=i)
        assert merged[i] == (len(blocks), int(sum(b.W for b in blocks)), float(np.abs(timeline).sum()))


def test_encoder_stream_groups_cover_every_stream_once():
    """Host logic of Encoder.encode: the ramped stream groups (1, 2, 4, 6, then equal shares) partition the streams."""
    from vorbis_aotuv_lancer_b200 import encoder
    for n in (1, 2, 3, 7, 8, 16, 63, 64, 65, 1024):
        for chunks in (1, 2, 6, 8, 11):
            groups = encoder.group_cuts(n, encoder.group_weights(chunks))
            assert groups[0][0] == 0 and groups[-1][1] == n
            assert all(a < b for a, b in groups)
            assert all(x[1] == y[0] for x, y in zip(groups[:-1], groups[1:]))
            assert len(groups) == min(chunks, n)
    sizes = [b - a for a, b in encoder.group_cuts(64, encoder.group_weights(8))]
    assert sizes[0] == 1 and sizes == sorted(sizes[:4]) + sizes[4:] and max(sizes) <= 14
