"""Multi-GPU sharding of the path: by stream, `i mod G`, no data-path collective.

All mutable encoder state is per stream (codec_internal.h:60-93), so ranks never exchange data on
the hot path; the only cross-rank traffic is gathering packets / statistics on the host, which is
plain torch.distributed object gathering (gloo on CPU, nccl-backed processes on the GPU box)."""


def shard_streams(n_streams, rank, world):
    """Indices of the streams rank `rank` of `world` encodes."""
    return list(range(rank, n_streams, world))


def gather_stats(local, world):
    """Gather one small python object per rank on every rank (packet counts, byte totals, timings)."""
    if world <= 1:
        return [local]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, local)
    return out
