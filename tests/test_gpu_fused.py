"""k7FusedKernel (decode_fused.cuh): forward ACS with the traceback inside the kernel on a ring of the newest 168
steps of decisions.  Its output must be the reference's full traceback bit for bit at every noise level -- frames
whose in-kernel traceback cannot be proven exact are handed to the two-kernel path inside the same call -- so every
case is compared with the oracle / the unmodified reference AND with the two-kernel path (CED_FUSED=0)."""
import os

import numpy as np
import pytest

import convolutionalencdec_b200 as ced
import oracle
from conftest import bsc

pytestmark = pytest.mark.gpu
K7 = oracle.K7_G


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture()
def ctx(torch_cuda):
    c = ced.Context(0)
    yield c
    c.close()
    for k in ("CED_FUSED", "CED_FUSED_MIN_FRAMES", "CED_FUSED_GEOM"):
        os.environ.pop(k, None)


def both_paths(ctx, code, segs, bits, packed=False, mode="1"):
    """(fused output, frames handed back, two-kernel output); mode 1: traceback inside the forward warps
    (k7FusedKernel), 2: on its own warp with bulk-copy streaming (k7FusedWsKernel; default code, byte format, aligned
    rows -- anything else falls back to mode 1)"""
    os.environ["CED_FUSED"] = mode
    os.environ["CED_FUSED_MIN_FRAMES"] = "1"
    dec = ctx.decode_batch_packed if packed else ctx.decode_batch
    a = dec(code, segs, bits).clone()
    ctx.sync()
    back = ctx.last_fallback_frames()
    os.environ["CED_FUSED"] = "0"
    b = dec(code, segs, bits).clone()
    ctx.sync()
    return a, back, b


@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("bits,frames,pad,offset", [
    (8, 1, 0, 0), (40, 33, 0, 0), (88, 127, 10, 0), (96, 128, 2, 0), (184, 129, 6, 0), (192, 100, 0, 0), (280, 64, 0, 0),
    (376, 40, 8, 0), (392, 40, 8, 0), (1000 // 8 * 8, 200, 2, 0), (2048, 257, 16 - (2054 % 16), 0), (4096, 300, 10, 0),
    (4096, 70, 0, 3), (16384, 9, 10, 0),
])
def test_fused_matches_oracle_shapes_and_alignment(torch_cuda, ctx, port, bits, frames, pad, offset, mode):
    """Frame lengths around the 96-step segment and the 24-step block, frame counts off the 32 / 128 grid, aligned and
    misaligned rows, the reference's maximum packet length; clean, useful, heavy and pure noise."""
    torch = torch_cuda
    rng = np.random.default_rng(bits * 31 + frames)
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    clean = port.encode_batch(7, K7, msgs)
    for p in (0.0, 0.04, 0.10, 0.5):
        noisy = bsc(rng, clean, p, junk_upper_bits=True)
        buf = torch.zeros(frames * (T + pad) + 16, dtype=torch.uint8, device="cuda")
        d = buf[offset:offset + frames * (T + pad)].view(frames, T + pad)
        d[:, :T] = torch.from_numpy(noisy).cuda()
        a, back, b = both_paths(ctx, ced.K7_DEFAULT, d, bits, mode=mode)
        want = port.decode_batch(7, K7, noisy, T)
        assert np.array_equal(a.cpu().numpy(), want), (p, back)
        assert torch.equal(a, b)
        if p == 0.0:
            assert back == 0 and np.array_equal(want, msgs)


@pytest.mark.parametrize("code,g,mode", [(ced.K7_DEFAULT, K7, "1"), (ced.K7_TEXTBOOK, (0o133, 0o171), "1"),
                                         (ced.K7_DEFAULT, K7, "2")])
def test_fused_config2_shape_all_noise_levels(torch_cuda, ctx, port, ref, code, g, mode):
    """2^16 frames x 4096 bits: identical to the two-kernel path on every frame; a sample against the unmodified
    reference; the number of frames handed back grows with the noise and is ~0 where the code is useful."""
    torch = torch_cuda
    frames, bits = 1 << 16, 4096
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=5)
    clean = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(code, msgs, out=clean)
    handed = {}
    for p in (0.0377, 0.06, 0.12, 0.5):
        segs = clean.clone()
        ctx.bsc_channel(segs, T, 2, p, seed=int(p * 1e4))
        a, back, b = both_paths(ctx, code, segs, bits, mode=mode)
        assert torch.equal(a, b), p
        handed[p] = back
        sample = np.arange(0, frames, 701)
        noisy = segs[torch.from_numpy(sample).cuda()][:, :T].cpu().numpy()
        want = ref.decode_batch(noisy, T) if g == K7 else port.decode_batch(7, g, noisy, T)
        assert np.array_equal(a.cpu().numpy()[sample], want), p
    assert handed[0.0377] < frames // 1000 and handed[0.06] < frames // 50
    assert handed[0.12] > frames // 4 and handed[0.5] > frames // 2


def test_fused_packed_format_and_waves(torch_cuda, ctx, port, monkeypatch):
    """Packed symbols (two 96-step segments per staged tile) and a batch cut into several waves."""
    torch = torch_cuda
    rng = np.random.default_rng(12)
    frames, bits = 5000, 1000 // 8 * 8
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    noisy = bsc(rng, port.encode_batch(7, K7, msgs), 0.05)
    want = port.decode_batch(7, K7, noisy, T)
    d = torch.from_numpy(noisy).cuda()
    packed = ctx.pack_symbols(d, T, packed_stride=((T + 3) // 4 + 15) // 16 * 16)
    a, back, b = both_paths(ctx, ced.K7_DEFAULT, packed, bits, packed=True)
    assert np.array_equal(a.cpu().numpy(), want) and torch.equal(a, b)
    monkeypatch.setenv("CED_MAX_WAVE_FRAMES", "1024")
    c2 = ced.Context(0)
    try:
        a2, _, _ = both_paths(c2, ced.K7_DEFAULT, d, bits)
        assert np.array_equal(a2.cpu().numpy(), want)
    finally:
        c2.close()


def test_fused_call_launch_count(torch_cuda, ctx):
    """ced_decode_batch of a GPU-filling batch with CED_FUSED=1 = 1 fused launch + 4 (empty) hand-back launches; small
    batches keep the two-kernel path."""
    torch = torch_cuda
    os.environ["CED_FUSED"] = "1"
    segs = torch.zeros((1 << 15, 272), dtype=torch.uint8, device="cuda")
    l0 = ctx.launches
    ctx.decode_batch(ced.K7_DEFAULT, segs, 256)
    ctx.sync()
    assert ctx.launches - l0 == 5 and ctx.last_fallback_frames() == 0
    l0 = ctx.launches
    ctx.decode_batch(ced.K7_DEFAULT, segs[:4096], 256)
    ctx.sync()
    assert ctx.launches - l0 == 2
