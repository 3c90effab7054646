"""Rate-k/n codes with k > 1 on the GPU (ced_encode_batch_k / ced_decode_batch_k, csrc/radix_k.cu) against the CPU
restatement oracle/ced_oracle_k.c, which tests/test_oracle_k.py pins to the unmodified reference built with k = 2
parameters (encoder, trellis labels, per-step path metrics)."""
import os

import numpy as np
import pytest

import convolutionalencdec_b200 as ced
from conftest import ROOT

pytestmark = pytest.mark.gpu

CODES = [(3, 2, (0o27, 0o75, 0o72)), (4, 2, (0o236, 0o155, 0o337)), (2, 2, (0o17, 0o06, 0o15)), (3, 2, (0o53, 0o75)),
         (5, 2, (0o1236, 0o0155, 0o1337)), (4, 2, (0o321, 0o256, 0o177)),
         (5, 2, (0o1236, 0o0155, 0o1337, 0o1701)), (3, 4, (0o7531, 0o6427, 0o5173, 0o3355, 0o1777)), (2, 4, (0o357, 0o261, 0o173, 0o225, 0o316))]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


@pytest.fixture(scope="module")
def ctx(torch_cuda):
    c = ced.Context(0)
    yield c
    c.close()


def noisy_k(rng, segs, n, p):
    flips = rng.random(segs.shape + (n,)) < p
    out = segs.copy()
    for j in range(n):
        out ^= (flips[..., j].astype(np.uint8) << j)
    return out


@pytest.mark.parametrize("K,k,g", CODES)
def test_k_encode_and_decode_match_the_restatement(torch_cuda, ctx, port, K, k, g):
    torch = torch_cuda
    rng = np.random.default_rng(K * 100 + k)
    code = ced.Code(K, g)
    n = len(g)
    for nbytes, frames, pad in ((1, 3, 0), (3, 37, 5), (64, 300, 3), (512, 70, 16)):
        msgs = rng.integers(0, 256, (frames, nbytes), dtype=np.uint8)
        want_segs = port.encode_batch_k(K, k, g, msgs)
        T = want_segs.shape[1]
        assert T == 8 * nbytes // k + K - 1
        d_msgs = torch.from_numpy(msgs).cuda()
        d_segs = torch.full((frames, T + pad), 0xEE, dtype=torch.uint8, device="cuda")
        ctx.encode_batch_k(code, k, d_msgs, out=d_segs)
        ctx.sync()
        assert np.array_equal(d_segs[:, :T].cpu().numpy(), want_segs), (K, k, nbytes)
        if pad:
            assert bool((d_segs[:, T:] == 0xEE).all())
        for p in (0.0, 0.03, 0.5):
            noisy = noisy_k(rng, want_segs, n, p)
            noisy |= rng.integers(0, 2, noisy.shape, dtype=np.uint8) << 7 if n < 8 else 0   # bits above n are ignored
            want = port.decode_batch_k(K, k, g, noisy & ((1 << n) - 1), T)
            d_noisy = torch.full((frames, T + pad), 0xEE, dtype=torch.uint8, device="cuda")
            d_noisy[:, :T] = torch.from_numpy(noisy).cuda()
            before = ctx.launches
            out = ctx.decode_batch_k(code, k, d_noisy, 8 * nbytes)
            ctx.sync()
            assert ctx.launches - before == 2
            assert np.array_equal(out.cpu().numpy(), want), (K, k, nbytes, p)
            if p == 0.0:
                assert np.array_equal(want, msgs)


def test_k_golden_vectors(torch_cuda, ctx):
    """tests/golden/k2_vectors.npz: encoder output of the UNMODIFIED reference built with k = 2 parameters, decoded bytes
    of the restatement whose add-compare-select was checked against the reference when the file was made."""
    torch = torch_cuda
    z = np.load(os.path.join(ROOT, "tests", "golden", "k2_vectors.npz"))
    for name, (K, k, g) in {"k2K3n3": (3, 2, (0o27, 0o75, 0o72)), "k2K4n3": (4, 2, (0o236, 0o155, 0o337))}.items():
        code = ced.Code(K, g)
        segs = ctx.encode_batch_k(code, k, torch.from_numpy(z[name + "_msgs"]).cuda())
        out = ctx.decode_batch_k(code, k, torch.from_numpy(z[name + "_noisy"]).cuda(), 8 * z[name + "_msgs"].shape[1])
        ctx.sync()
        assert np.array_equal(segs.cpu().numpy(), z[name + "_segs"])
        assert np.array_equal(out.cpu().numpy(), z[name + "_decoded"])


def test_k_waves_and_argument_checks(torch_cuda, ctx, port, monkeypatch):
    torch = torch_cuda
    K, k, g = 4, 2, (0o236, 0o155, 0o337)
    code = ced.Code(K, g)
    rng = np.random.default_rng(5)
    msgs = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    segs = noisy_k(rng, port.encode_batch_k(K, k, g, msgs), 3, 0.04)
    want = port.decode_batch_k(K, k, g, segs, segs.shape[1])
    monkeypatch.setenv("CED_MAX_WAVE_FRAMES", "256")
    c2 = ced.Context(0)
    try:
        before = c2.launches
        out = c2.decode_batch_k(code, k, torch.from_numpy(segs).cuda(), 256)
        c2.sync()
        assert c2.launches - before == 8          # four waves of 256 frames
        assert np.array_equal(out.cpu().numpy(), want)
    finally:
        c2.close()
    d = torch.zeros((4, 200), dtype=torch.uint8, device="cuda")
    with pytest.raises(ced.CedError):
        ctx.decode_batch_k(code, 3, d, 64)        # 8 % k != 0
    with pytest.raises(ced.CedError):
        ctx.decode_batch_k(ced.Code(6, g), 2, d, 64)   # 2^(k(K-1)) = 1024 states
    with pytest.raises(ced.CedError):
        ctx.decode_batch_k(code, 2, torch.zeros((4, 20), dtype=torch.uint8, device="cuda"), 64)   # stride shorter than a frame
    # k = 1 forwards to the k = 1 entry points
    m1 = rng.integers(0, 256, (9, 8), dtype=np.uint8)
    s1 = ctx.encode_batch_k(ced.K7_DEFAULT, 1, torch.from_numpy(m1).cuda())
    o1 = ctx.decode_batch_k(ced.K7_DEFAULT, 1, s1, 64)
    ctx.sync()
    assert np.array_equal(o1.cpu().numpy(), m1)
