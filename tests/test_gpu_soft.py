"""Soft-decision decoding on the GPU (ced_decode_batch_soft, k7SoftForwardKernel + k7TracebackKernel<Lanes16>)
against the soft oracle (oracle/ced_oracle.c:orc_dec_step_soft) and, through the constant-magnitude
reduction, against the reference's own hard decoder.  Bit-exact: all arithmetic is 16-bit integer."""
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


@pytest.fixture(scope="module")
def ctx(torch_cuda):
    c = ced.Context(0)
    yield c
    c.close()


def soft_rows(torch, soft, T):
    """int8 [frames, 2T] -> CUDA tensor with 16-byte aligned rows (stride rounded up, padding = junk)."""
    frames = soft.shape[0]
    stride = (2 * T + 15) // 16 * 16
    buf = np.full((frames, stride), 55, dtype=np.int8)
    buf[:, :2 * T] = soft
    return torch.from_numpy(buf).cuda()


def awgn(rng, segs, T, amp, sigma, lo=-127):
    bits = np.stack([segs[:, :T] & 1, (segs[:, :T] >> 1) & 1], axis=-1).reshape(segs.shape[0], 2 * T)
    return np.clip(np.round(amp * ((1.0 - 2.0 * bits) + sigma * rng.standard_normal(bits.shape))), lo, 127).astype(np.int8)


@pytest.mark.parametrize("bits,frames", [(8, 1), (16, 33), (40, 31), (96, 127), (104, 128), (192, 129), (1000 // 8 * 8, 200),
                                         (2048, 257), (4096, 300), (16384, 9)])
def test_soft_decode_matches_soft_oracle(torch_cuda, ctx, port, bits, frames):
    """Ragged frame counts (off the 32 / 128 grid), frame lengths off the 48-step tile and the 6-step body,
    the reference's maximum packet length; moderate noise, heavy noise, clipping to -128."""
    rng = np.random.default_rng(bits * 7 + frames)
    T = bits + 6
    segs = port.encode_batch(7, K7, rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8))
    for amp, sigma, lo in ((32, 0.7, -127), (100, 1.3, -128), (3, 1.0, -127)):
        soft = awgn(rng, segs, T, amp, sigma, lo)
        out = ctx.decode_batch_soft(ced.K7_DEFAULT, soft_rows(torch_cuda, soft, T), bits)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), port.decode_soft_batch(7, K7, soft, T)), (amp, sigma)


def test_soft_decode_tie_heavy_inputs(torch_cuda, ctx, port):
    """All-zero (every comparison ties), extremes only, tiny magnitudes, uniform noise over the full int8 range."""
    rng = np.random.default_rng(3)
    bits, frames = 1024, 96
    T = bits + 6
    shape = (frames, 2 * T)
    for soft in (np.zeros(shape, dtype=np.int8), rng.choice([-128, 127], shape).astype(np.int8),
                 rng.integers(-2, 3, shape).astype(np.int8), rng.integers(-128, 128, shape).astype(np.int8)):
        out = ctx.decode_batch_soft(ced.K7_DEFAULT, soft_rows(torch_cuda, soft, T), bits)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), port.decode_soft_batch(7, K7, soft, T))


@pytest.mark.parametrize("code,g", [(ced.K7_DEFAULT, K7), (ced.K7_TEXTBOOK, (0o133, 0o171))])
def test_constant_magnitude_soft_equals_hard_path_config2_shape(torch_cuda, ctx, port, ref, code, g):
    """The pin to the reference (src/viterbiDecoderButterflyk1.c:104-140): soft inputs saturated to +-A must give
    the hard decoder's output bit for bit -- all 2^16 frames x 4096 bits against the GPU hard path, a sample of
    them against the unmodified reference decoder (default generators) / the oracle (0133/0171)."""
    torch = torch_cuda
    frames, bits = 1 << 16, 4096
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=99)
    segs = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(code, msgs, out=segs)
    ctx.bsc_channel(segs, T, 2, 0.05, seed=5)
    hard = ctx.decode_batch(code, segs, bits)
    for A in (1, 127):
        s3 = segs[:, :T].to(torch.int16)
        soft = torch.zeros((frames, 8208), dtype=torch.int8, device="cuda")
        soft[:, 0:2 * T:2] = torch.where((s3 & 1) == 1, -A, A).to(torch.int8)
        soft[:, 1:2 * T:2] = torch.where((s3 & 2) == 2, -A, A).to(torch.int8)
        out = ctx.decode_batch_soft(code, soft, bits)
        ctx.sync()
        assert torch.equal(out, hard), A
    sample = np.arange(0, frames, 509)
    noisy = segs[torch.from_numpy(sample).cuda()][:, :T].cpu().numpy()
    want = ref.decode_batch(noisy, T) if g == K7 else port.decode_batch(7, g, noisy, T)
    assert np.array_equal(hard.cpu().numpy()[sample], want)


def test_soft_decode_config2_shape_vs_soft_oracle(torch_cuda, ctx, port):
    """2^16 frames x 4096 bits through the device AWGN channel at Eb/N0 = 2 dB; a sample of frames against the
    soft oracle, and the whole batch through encode -> channel -> decode -> compare (BER must be far below the
    hard decoder's on the sliced symbols)."""
    torch = torch_cuda
    frames, bits = 1 << 16, 4096
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=7)
    segs = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
    cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
    soft = ctx.awgn_channel(segs, T, 2.0, seed=11, counters=cnt[:2])
    out = ctx.decode_batch_soft(ced.K7_DEFAULT, soft, bits)
    hard_syms = ctx.slice_soft_to_bytes(soft, T, seg_stride=4112)
    out_hard = ctx.decode_batch(ced.K7_DEFAULT, hard_syms, bits)
    ctx.ber_count(out, msgs, cnt[2:4])
    ctx.ber_count(out_hard, msgs, cnt[4:6])
    ctx.sync()
    sample = np.arange(0, frames, 257)
    idx = torch.from_numpy(sample).cuda()
    s_np = soft[idx][:, :2 * T].cpu().numpy()
    assert np.array_equal(out[idx].cpu().numpy(), port.decode_soft_batch(7, K7, s_np, T))
    # slicing keeps the sign: the hard symbols are the signs of the soft ones, and decode to the oracle's bytes
    h_np = hard_syms[idx][:, :T].cpu().numpy()
    assert np.array_equal(h_np, ((s_np[:, 0::2] < 0) | ((s_np[:, 1::2] < 0) << 1)).astype(np.uint8))
    assert np.array_equal(out_hard[idx].cpu().numpy(), port.decode_batch(7, K7, h_np, T))
    flips, coded, soft_err, n1, hard_err, n2 = [int(v) for v in cnt.cpu()]
    assert coded == frames * T * 2 and n1 == n2 == frames * bits
    p = flips / coded
    assert abs(p - 0.1040) < 0.002       # Q(sqrt(Eb/N0)) at 2 dB, SURVEY 8(d) BER sweep table
    assert hard_err > 10 * max(soft_err, 1)   # ~2 dB of soft-decision gain (measured here: 0.106 vs 5.6e-3)


def test_awgn_channel_is_independent_of_sharding(torch_cuda, ctx):
    torch = torch_cuda
    frames, T = 512, 262
    segs = torch.randint(0, 4, (frames, T), dtype=torch.uint8, device="cuda")
    whole = ctx.awgn_channel(segs, T, 3.0, seed=4)
    a = ctx.awgn_channel(segs[:200], T, 3.0, seed=4)
    b = ctx.awgn_channel(segs[200:], T, 3.0, seed=4, first_frame=200)
    ctx.sync()
    assert torch.equal(whole[:, :2 * T], torch.cat([a, b])[:, :2 * T])


def test_soft_argument_checks(torch_cuda, ctx):
    torch = torch_cuda
    soft = torch.zeros((4, 2 * 70 + 4), dtype=torch.int8, device="cuda")   # stride 144: fine; 64 bits -> T = 70
    ctx.decode_batch_soft(ced.K7_DEFAULT, soft, 64)
    with pytest.raises(ced.CedError):
        ctx.decode_batch_soft(ced.K7_DEFAULT, soft[:, 1:], 64)               # misaligned base
    with pytest.raises(ced.CedError):
        ctx.decode_batch_soft(ced.Code(7, (0o117, 0o155)), soft, 64)         # run-time code: not on the soft path
    with pytest.raises(ced.CedError):
        ctx.decode_batch_soft(ced.K7_DEFAULT, soft, 128)                     # stride shorter than a frame
    ctx.sync()
