"""Single-process multi-GPU C ABI (ced_multi_*, include/ced_abi.h): shard arithmetic on the CPU,
and -- on however many GPUs the box has -- host batches sharded by the C library against the oracle, the NCCL
counter all-reduce, and the plain-C example."""
import os
import subprocess

import numpy as np
import pytest

import convolutionalencdec_b200 as ced
import oracle
from conftest import ROOT, bsc

K7 = oracle.K7_G


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 64, 65537, 1 << 22):
        for g in (1, 2, 3, 4, 8, 16):
            ranges = [ced.shard_range(n, g, i) for i in range(g)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
            for (f0, c0), (f1, _) in zip(ranges, ranges[1:]):
                assert f1 == f0 + c0                      # contiguous, in rank order
            counts = [c for _, c in ranges]
            assert max(counts) - min(counts) <= 1         # balanced
    assert ced.shard_range(10, 4, 9) == (0, 0) and ced.shard_range(10, 0, 0) == (0, 0)
    # the ranks of a torchrun launch own the same ranges (sharding.py)
    from convolutionalencdec_b200.sharding import shard_range as rank_range
    for n, g in ((1 << 22, 8), (1000, 3)):
        for r in range(g):
            f0, cnt = ced.shard_range(n, g, r)
            assert rank_range(n, r, g) == (f0, f0 + cnt)


@pytest.fixture(scope="module")
def multi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m = ced.MultiContext()
    yield m
    m.close()


@pytest.mark.gpu
def test_host_batch_sharded_over_all_gpus(multi, port):
    import torch
    rng = np.random.default_rng(8)
    frames, bits = 70001, 256                 # not a multiple of the GPU count, the lane count or 32
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    segs = np.zeros((frames, 272), dtype=np.uint8)
    multi.encode_batch_host(ced.K7_DEFAULT, msgs, segs)
    assert np.array_equal(segs[:, :T], port.encode_batch(7, K7, msgs))
    noisy = np.zeros_like(segs)
    noisy[:, :T] = bsc(rng, segs[:, :T], 0.03)
    out = np.zeros((frames, bits // 8), dtype=np.uint8)
    multi.decode_batch_host(ced.K7_DEFAULT, noisy, bits, out)
    sample = rng.choice(frames, 3000, replace=False)
    assert np.array_equal(out[sample], port.decode_batch(7, K7, noisy[sample, :T], T))
    # pinned buffers take the other staging path
    h_in, h_out = torch.from_numpy(noisy).pin_memory(), torch.zeros((frames, bits // 8), dtype=torch.uint8).pin_memory()
    multi.decode_batch_host(ced.K7_DEFAULT, h_in, bits, h_out)
    assert np.array_equal(h_out.numpy(), out)
    up, down = multi.probe_copy_ceiling(64 << 20, 2)
    assert up > 1e9 and down > 1e9


@pytest.mark.gpu
def test_ber_counters_are_summed_over_nccl(multi):
    import torch
    counters, want = [], np.zeros(4, dtype=np.int64)
    for g in range(multi.n_devices):
        c = multi.ctx(g)
        with torch.cuda.device(c.device):
            msgs = torch.empty((4096, 32), dtype=torch.uint8, device="cuda")
            c.random_bytes(msgs, seed=3, first_frame=4096 * g)
            segs = torch.zeros((4096, 272), dtype=torch.uint8, device="cuda")
            c.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
            cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
            c.bsc_channel(segs, 262, 2, 0.06, seed=4, first_frame=4096 * g, counters=cnt[:2])
            dec = c.decode_batch(ced.K7_DEFAULT, segs, 256)
            c.ber_count(dec, msgs, cnt[2:])
            c.sync()
            torch.cuda.synchronize()
            counters.append(cnt)
            want += cnt.cpu().numpy()
    assert want[1] == multi.n_devices * 4096 * 262 * 2 and want[3] == multi.n_devices * 4096 * 256 and want[2] > 0
    multi.ber_allreduce(counters)
    for cnt in counters:
        assert np.array_equal(cnt.cpu().numpy(), want)
    assert ced.load_abi().ced_nccl_version() > 0


@pytest.mark.gpu
def test_plain_c_multi_gpu_example():
    exe = os.path.join(ROOT, "examples", "_bin", "multi_gpu_roundtrip")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "examples"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([exe, str(1 << 15), "1024"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 wrong bytes" in r.stdout and "every GPU holds the sum: yes" in r.stdout
