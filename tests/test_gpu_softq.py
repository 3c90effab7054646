"""3-bit soft-decision decoder (ced_decode_batch_softq, csrc/softq_decode.cuh): byte path metrics at the hard kernel's
speed.  Definition: a level x in 0..7 is the reliability s = 7 - 2x of the int8 soft decoder, so the decoded bytes must
equal oracle.decode_soft_batch on those int8 values bit for bit (that restatement is pinned to the reference by reduction,
tests/test_oracle.py)."""
import numpy as np
import pytest

import convolutionalencdec_b200 as ced
import oracle

pytestmark = pytest.mark.gpu
K7 = oracle.K7_G


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


def to_int8_pairs(syms, T):
    """x0 | x1 << 3 -> the int8 reliabilities (s0, s1) = (7 - 2 x0, 7 - 2 x1) per segment"""
    x0 = (syms[:, :T] & 7).astype(np.int16)
    x1 = ((syms[:, :T] >> 3) & 7).astype(np.int16)
    soft = np.empty((syms.shape[0], 2 * T), dtype=np.int8)
    soft[:, 0::2] = 7 - 2 * x0
    soft[:, 1::2] = 7 - 2 * x1
    return soft


@pytest.mark.parametrize("bits,frames,pad,off", [(8, 3, 0, 0), (96, 40, 10, 0), (104, 70, 5, 3), (1000 // 8 * 8, 200, 0, 1),
                                                 (4096, 300, 10, 0), (4096, 64, 5, 7)])
def test_softq_matches_the_int8_soft_decoder_on_quantised_inputs(torch_cuda, ctx, port, bits, frames, pad, off):
    torch = torch_cuda
    rng = np.random.default_rng(bits + frames)
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    clean = port.encode_batch(7, K7, msgs)
    for kind in ("clean", "3dB", "noise", "ties"):
        if kind == "clean":
            x = np.where((clean[..., None] >> np.arange(2)) & 1, 7, 0)
        elif kind == "3dB":
            bit = ((clean[..., None] >> np.arange(2)) & 1).astype(np.float64)
            y = (1 - 2 * bit) + rng.normal(0, 10 ** (-3 / 20), bit.shape)
            x = np.clip(3 - np.floor(y / 0.35).astype(np.int64), 0, 7)
        elif kind == "noise":
            x = rng.integers(0, 8, clean.shape + (2,))
        else:                                            # only the two middle levels: nearly every comparison ties
            x = rng.integers(3, 5, clean.shape + (2,))
        syms = (x[..., 0] | (x[..., 1] << 3)).astype(np.uint8)
        junk = syms | (rng.integers(0, 4, syms.shape, dtype=np.uint8) << 6)      # bits 6-7 are ignored
        want = port.decode_soft_batch(7, K7, to_int8_pairs(syms, T), T)
        flat = torch.full((frames * (T + pad) + 64,), 0xFF, dtype=torch.uint8, device="cuda")
        view = flat[off:off + frames * (T + pad)].view(frames, T + pad)
        view[:, :T] = torch.from_numpy(junk).cuda()
        before = ctx.launches
        out = ctx.decode_batch_softq(ced.K7_DEFAULT, view, bits)
        ctx.sync()
        assert ctx.launches - before == 2
        assert np.array_equal(out.cpu().numpy(), want), (bits, frames, kind)
        if kind == "clean":
            assert np.array_equal(want, msgs)


def test_softq_config2_shape_second_code_and_waves(torch_cuda, ctx, port, monkeypatch):
    """2^16 x 4096-bit shape through channel -> quantiser -> decoder on the device; a sample of frames against the int8
    soft oracle; the textbook generators; several waves."""
    torch = torch_cuda
    frames, bits, T = 1 << 16, 4096, 4102
    for code, g in ((ced.K7_DEFAULT, K7), (ced.K7_TEXTBOOK, [0o133, 0o171])):
        msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
        ctx.random_bytes(msgs, seed=9)
        segs = ctx.encode_batch(code, msgs, seg_stride=4112)
        soft = ctx.awgn_channel(segs, T, 3.0, seed=77)
        sigma_i8 = 32.0 * 10 ** (-3.0 / 20)
        syms = ctx.quantize_soft(soft, T, 0.6 * sigma_i8, sym_stride=4112)
        out = ctx.decode_batch_softq(code, syms, bits)
        hard = ctx.decode_batch(code, ctx.slice_soft_to_bytes(soft, T, seg_stride=4112), bits)
        ctx.sync()
        sample = np.arange(0, frames, 997)
        want = port.decode_soft_batch(7, g, to_int8_pairs(syms[sample].cpu().numpy(), T), T)
        assert np.array_equal(out[sample].cpu().numpy(), want)
        soft_err = int((out != msgs).sum())
        hard_err = int((hard != msgs).sum())
        assert soft_err * 20 < hard_err, (soft_err, hard_err)      # 3 dB: hard BER 2.8e-2, 3-bit soft around 6e-4
    monkeypatch.setenv("CED_MAX_WAVE_FRAMES", "512")
    c2 = ced.Context(0)
    try:
        small = syms[:1500].contiguous()
        got = c2.decode_batch_softq(ced.K7_TEXTBOOK, small, bits)
        c2.sync()
        assert torch.equal(got, out[:1500])
    finally:
        c2.close()


def test_softq_host_buffers_match_device_path(torch_cuda, ctx, port):
    """ced_decode_batch_softq_host: pageable and page-locked host arrays through the chunked pipeline (20000 frames =
    three chunks) give the device path's bytes."""
    torch = torch_cuda
    rng = np.random.default_rng(12)
    frames, bits, T = 20000, 512, 518
    syms = rng.integers(0, 64, (frames, T + 2), dtype=np.uint8)
    want = ctx.decode_batch_softq(ced.K7_DEFAULT, torch.from_numpy(syms).cuda(), bits)
    ctx.sync()
    out = np.zeros((frames, bits // 8), dtype=np.uint8)
    ctx.decode_batch_softq_host(ced.K7_DEFAULT, syms, bits, out)                 # pageable numpy arrays
    assert np.array_equal(out, want.cpu().numpy())
    h_in = torch.from_numpy(syms).pin_memory()
    h_out = torch.zeros((frames, bits // 8), dtype=torch.uint8).pin_memory()
    ctx.decode_batch_softq_host(ced.K7_DEFAULT, h_in, bits, h_out)               # page-locked
    assert torch.equal(h_out, want.cpu())
    sample = np.arange(0, frames, 401)
    assert np.array_equal(out[sample], port.decode_soft_batch(7, K7, to_int8_pairs(syms[sample], T), T))


@pytest.mark.parametrize("n_streams,total,call,depth,kind", [(3, 96 * 3 + 14, 96, 24, "3dB"), (70, 1000 // 8 * 8 + 6, 192, 48, "3dB"),
                                                              (40, 2054, 480, 48, "noise"), (33, 2054, 96, 96, "ties"),
                                                              (10, 4102, 960, 48, "3dB"), (5, 518, 4800, 48, "3dB")])
def test_softq_window_decode_matches_oracle_definition(torch_cuda, ctx, port, n_streams, total, call, depth, kind):
    """ced_decode_window_batch_softq: continuous streams of 3-bit soft symbols, slice by slice, against
    orc_decode_window_soft on the reliabilities s = 7 - 2x (the window procedure of orc_decode_window around the soft
    recursion); with depth >= stream length it is the full-frame soft decode."""
    torch = torch_cuda
    rng = np.random.default_rng(total + call + depth)
    bits = total - 6
    msgs = rng.integers(0, 256, (n_streams, bits // 8), dtype=np.uint8)
    clean = port.encode_batch(7, K7, msgs)
    if kind == "3dB":
        bit = ((clean[..., None] >> np.arange(2)) & 1).astype(np.float64)
        y = (1 - 2 * bit) + rng.normal(0, 10 ** (-3 / 20), bit.shape)
        x = np.clip(3 - np.floor(y / 0.35).astype(np.int64), 0, 7)
    elif kind == "noise":
        x = rng.integers(0, 8, clean.shape + (2,))
    else:
        x = rng.integers(3, 5, clean.shape + (2,))
    syms = (x[..., 0] | (x[..., 1] << 3)).astype(np.uint8)
    soft = to_int8_pairs(syms, total)
    want = np.stack([port.decode_window_soft(7, K7, soft[i], call, depth) for i in range(n_streams)])
    d = torch.from_numpy(syms).cuda()
    wd = ctx.window_decoder(ced.K7_DEFAULT, n_streams, depth=depth, softq=True)
    parts = [wd.push(d[:, a:min(a + call, total)], last=a + call >= total).clone() for a in range(0, total, call)]
    ctx.sync()
    got = torch.cat(parts, dim=1).cpu().numpy()
    assert np.array_equal(got, want), (n_streams, total, call, depth, kind)
    if call >= total:      # one slice: the full-frame soft decode
        assert np.array_equal(got, port.decode_soft_batch(7, K7, soft, total))


def test_softq_argument_checks(torch_cuda, ctx):
    torch = torch_cuda
    d = torch.zeros((4, 64), dtype=torch.uint8, device="cuda")
    with pytest.raises(ced.CedError):
        ctx.decode_batch_softq(ced.Code(7, (0o117, 0o155)), d, 48)
    with pytest.raises(ced.CedError):
        ctx.decode_batch_softq(ced.K7_DEFAULT, d, 64)          # 70 segments do not fit a 64-byte row
    with pytest.raises(ced.CedError):
        ctx.decode_batch_softq(ced.K7_DEFAULT, d, 12)


def test_plain_c_soft_decisions_example():
    """examples/soft_decisions.c: a C program (no CUDA headers) sends one batch through the AWGN channel and decodes it
    with hard, int8-soft and 3-bit-soft decisions; the error rates must order as expected."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "examples", "_bin", "soft_decisions")
    if not os.path.exists(exe):
        pytest.skip("examples/_bin/soft_decisions not built")
    r = subprocess.run([exe, "3.0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3-bit soft" in r.stdout
