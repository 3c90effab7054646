"""Frame sharding across ranks (one process per GPU).

Frames are independent packets (state is reset at `last`,
src/viterbiDecoderButterflyk1.c:259), so the decode/encode path needs no
inter-GPU traffic: rank r of W owns a contiguous frame range.  The only
collective is the sum of BER counters (torch.distributed all_reduce; NCCL on
GPUs, gloo in the CPU tests).
"""


def shard_range(n_frames, rank, world):
    """Contiguous [lo, hi) of frames owned by `rank`; sizes differ by at most 1."""
    base, extra = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_counts(counters):
    """Sum a tensor of int64 counters over all ranks (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters
