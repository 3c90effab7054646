"""world_size-2 gloo test of the only multi-rank logic on the path: contiguous
frame sharding (no data-path collective) and the all-reduce of BER counters."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import oracle
    from convolutionalencdec_b200.sharding import allreduce_counts, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = oracle.port()
    n_frames, bits = 37, 256
    rng = np.random.default_rng(123)                      # same data on every rank
    msgs = rng.integers(0, 256, (n_frames, bits // 8), dtype=np.uint8)
    segs = P.encode_batch(7, oracle.K7_G, msgs)
    flips = rng.random(segs.shape + (2,)) < 0.09
    noisy = segs ^ (flips[..., 0].astype(np.uint8) | (flips[..., 1].astype(np.uint8) << 1))
    lo, hi = shard_range(n_frames, rank, world)
    dec = P.decode_batch(7, oracle.K7_G, noisy[lo:hi], bits + 6)   # stand-in for the per-rank GPU decode
    errs = int(np.unpackbits(dec ^ msgs[lo:hi]).sum())
    counts = torch.tensor([errs, (hi - lo) * bits], dtype=torch.int64)
    allreduce_counts(counts)
    full = P.decode_batch(7, oracle.K7_G, noisy, bits + 6)
    want = [int(np.unpackbits(full ^ msgs).sum()), n_frames * bits]
    assert counts.tolist() == want, (counts.tolist(), want)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ber_counts_sum_to_the_unsharded_result(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
