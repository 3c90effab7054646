"""Parity tests proper: the CUDA path, called through the C ABI (include/ced_abi.h)
and through the reference-named host C API, against the oracle, the committed
reference fixtures and size-independent properties.  Bit-exact everywhere: all
arithmetic on this path is 8-bit integer."""
import os
import subprocess
import sys

import numpy as np
import pytest

import convolutionalencdec_b200 as ced
import oracle
from conftest import ROOT, bsc

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


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ batch decode
@pytest.mark.parametrize("bits", [8, 64, 256, 2048, 4096])
def test_decode_batch_matches_reference_fixtures(torch_cuda, ctx, golden, bits):
    for tag in ("p0", "p02", "p06", "p50"):
        noisy = golden["noisy_%d_%s" % (bits, tag)]
        out = ctx.decode_batch(ced.K7_DEFAULT, dev(torch_cuda, noisy), bits)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), golden["dec_%d_%s" % (bits, tag)]), (bits, tag)


@pytest.mark.parametrize("bits,frames,stride_pad,offset", [
    (8, 1, 0, 0), (16, 3, 0, 0), (40, 33, 0, 0), (96, 127, 10, 0), (104, 128, 2, 0), (192, 129, 6, 0),
    (1000 // 8 * 8, 200, 0, 0), (2048, 257, 16 - (2054 % 16), 0), (4096, 300, 10, 0), (4096, 64, 0, 0),
    (4096, 70, 0, 3), (264, 500, 5, 1), (16384, 9, 10, 0),
])
def test_decode_batch_vs_oracle_shapes_and_alignment(torch_cuda, ctx, port, bits, frames, stride_pad, offset):
    """Ragged sizes: frame counts off the 32/128 grid, rows 16-byte aligned (fast staging path) and
    not (generic path), base pointer misaligned, max packet length, junk in the unused symbol bits."""
    torch = torch_cuda
    rng = np.random.default_rng(bits * 131 + frames)
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    clean = port.encode_batch(7, K7, msgs)
    for p in (0.0, 0.05, 0.5):
        noisy = bsc(rng, clean, p, junk_upper_bits=True)
        want = port.decode_batch(7, K7, noisy, T)
        stride = T + stride_pad
        flat = torch.zeros(frames * stride + offset + 64, dtype=torch.uint8, device="cuda")
        view = flat[offset:offset + frames * stride].view(frames, stride)
        view[:, :T] = dev(torch, noisy)
        view[:, T:] = 0xFF  # padding must never influence the result
        out = ctx.decode_batch(ced.K7_DEFAULT, view, bits)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), want), (bits, frames, p)
        if p == 0.0:
            assert np.array_equal(want, msgs)


def test_decode_batch_textbook_generators(torch_cuda, ctx, port):
    g = (0o133, 0o171)
    rng = np.random.default_rng(9)
    msgs = rng.integers(0, 256, (150, 64), dtype=np.uint8)
    noisy = bsc(rng, port.encode_batch(7, g, msgs), 0.06)
    out = ctx.decode_batch(ced.K7_TEXTBOOK, dev(torch_cuda, noisy), 512)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), port.decode_batch(7, g, noisy, 518))


@pytest.mark.parametrize("K,g,bits,frames", [
    (3, (0b111, 0b110), 8, 5), (3, (0b111, 0b110), 256, 300), (3, (0b111, 0b101), 1000 // 8 * 8, 40),
    (5, (0o23, 0o35), 512, 200), (7, (0o117, 0o155), 2048, 150), (7, (0o133, 0o145, 0o175), 256, 100),
    (9, (0o561, 0o753), 512, 150), (9, (0o557, 0o663, 0o711), 4096, 33), (2, (0b11, 0b10), 64, 17),
])
def test_decode_batch_other_code_parameters(torch_cuda, ctx, port, K, g, bits, frames):
    """any k=1 code with K <= 9 decodes through ced_decode_batch (generic one-warp/CTA-per-frame
    kernel), incl. the hand-traced K=3 code whose generators lack the symmetry the reference's butterfly
    needs; oracle = general butterflies (generic decoder's ACS, src/viterbiDecoder.c:95-128)."""
    torch = torch_cuda
    rng = np.random.default_rng(K * 1000 + bits)
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    clean = port.encode_batch(K, g, msgs)
    n = len(g)
    T = bits + K - 1
    for p in (0.0, 0.03, 0.3):
        flips = rng.random(clean.shape + (n,)) < p
        noisy = clean.copy()
        for j in range(n):
            noisy ^= (flips[..., j].astype(np.uint8) << j)
        want = port.decode_batch(K, g, noisy, T, symmetric=False)
        view = torch.full((frames, T + 5), 0xFF, dtype=torch.uint8, device="cuda")
        view[:, :T] = dev(torch, noisy)
        out = ctx.decode_batch(ced.Code(K, g), view, bits)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), want), (K, g, p)
        if p == 0.0:
            assert np.array_equal(want, msgs)


@pytest.mark.parametrize("K,g", [(3, (0b111, 0b110)), (3, (0b111, 0b101, 0b011)), (4, (0o15, 0o17)), (4, (0o13, 0o15, 0o17)),
                                 (5, (0o23, 0o35)), (5, (0o25, 0o33, 0o37)), (7, (0o133, 0o170)), (7, (0o066, 0o171)),
                                 (7, (0o133, 0o145, 0o174)), (9, (0o561, 0o753)), (9, (0o460, 0o353)),
                                 (6, (0o53, 0o75)), (6, (0o47, 0o53, 0o75)), (8, (0o247, 0o371)), (8, (0o300, 0o073)),
                                 (9, (0o557, 0o663, 0o711))])
def test_decode_batch_generic_codes_run_the_table_driven_swar_kernels(torch_cuda, ctx, port, K, g):
    """SURVEY 8(f)3 / VERDICT r1 missing 3: K = 3, 4, 5, 7, 9 with 2 or 3 generators of ANY shape (non-symmetric ones
    included) take the thread-per-frame SIMD-in-word kernels of swar_generic.cu -- two launches per call -- over
    several 32-frame groups and work units, aligned and misaligned rows, clean / noisy / pure-noise channels."""
    torch = torch_cuda
    rng = np.random.default_rng(K * 977 + sum(g))
    n, S = len(g), K - 1
    for bits, frames, pad, off in ((8, 3, 0, 0), (104, 70, 11, 0), (1000, 1300, 5, 3), (4096, 260, 16 - S, 0)):
        T = bits + S
        msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
        clean = port.encode_batch(K, list(g), msgs)
        for p in (0.0, 0.04, 0.5):
            flips = rng.random(clean.shape + (n,)) < p
            noisy = clean.copy()
            for j in range(n):
                noisy ^= (flips[..., j].astype(np.uint8) << j)
            noisy |= rng.integers(0, 32, noisy.shape, dtype=np.uint8) << 3
            want = port.decode_batch(K, list(g), noisy, T, symmetric=False)
            flat = torch.full((frames * (T + pad) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
            view = flat[off:off + frames * (T + pad)].view(frames, T + pad)
            view[:, :T] = dev(torch, noisy)
            before = ctx.launches
            out = ctx.decode_batch(ced.Code(K, g), view, bits)
            ctx.sync()
            assert ctx.launches - before == 2, "not the swar_generic kernels"
            assert np.array_equal(out.cpu().numpy(), want), (K, g, bits, frames, p)
            if p == 0.0:
                assert np.array_equal(want, msgs)


@pytest.mark.parametrize("g", [(0o171, 0o133), (0o117, 0o155), (0o135, 0o163), (0o145, 0o175), (0o101, 0o177),
                               (0o133, 0o171, 0o165), (0o133, 0o145, 0o175), (0o175, 0o133, 0o171)])
def test_decode_batch_any_symmetric_k7_code_runs_the_swar_kernel(torch_cuda, ctx, port, g):
    """K=7 generators (n = 2 or 3) known only at run time -- the SWAR forward kernel driven by a step
    table (RuntimeK7<n>) -- byte and packed symbols, aligned and misaligned rows, and the windowed decoder."""
    torch = torch_cuda
    rng = np.random.default_rng(g[0] * 1000 + g[1])
    code = ced.Code(7, g)
    frames, bits = 333, 4096
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    enc = ctx.encode_batch(code, dev(torch, msgs))
    ctx.sync()
    clean = port.encode_batch(7, list(g), msgs)
    assert np.array_equal(enc.cpu().numpy()[:, :T], clean)
    for p in (0.0, 0.05, 0.5):
        noisy = clean.copy()
        flips = rng.random(clean.shape + (len(g),)) < p
        for j in range(len(g)):
            noisy ^= (flips[..., j].astype(np.uint8) << j)
        noisy |= (rng.integers(0, 2, clean.shape, dtype=np.uint8) << len(g))   # junk above the n coded bits
        want = port.decode_batch(7, list(g), noisy, T)
        aligned = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
        aligned[:, :T] = dev(torch, noisy)
        ragged = dev(torch, noisy)                       # stride 4102: the alignment-agnostic staging
        outs = [ctx.decode_batch(code, aligned, bits), ctx.decode_batch(code, ragged, bits)]
        if len(g) == 2:
            outs.append(ctx.decode_batch_packed(code, ctx.pack_symbols(aligned, T), bits))
        ctx.sync()
        for out in outs:
            assert np.array_equal(out.cpu().numpy(), want), (g, p)
    wd = ctx.window_decoder(code, frames, depth=48)
    pieces = [wd.push(aligned[:, a:min(a + 960, T)], last=a + 960 >= T).clone() for a in range(0, T, 960)]
    ctx.sync()
    got = torch.cat(pieces, dim=1).cpu().numpy()
    for i in range(0, frames, 37):
        assert np.array_equal(got[i], port.decode_window(7, list(g), noisy[i], 960, 48))


def test_decode_batch_rejects_what_it_cannot_do(torch_cuda, ctx):
    torch = torch_cuda
    segs = torch.zeros((4, 70), dtype=torch.uint8, device="cuda")
    with pytest.raises(ced.CedError):
        ctx.decode_batch(ced.Code(10, (0o1167, 0o1545)), segs, 56)   # K > 9
    with pytest.raises(ced.CedError):
        ctx.decode_batch_packed(ced.Code(5, (0o23, 0o35)), segs, 64)  # packed format: K=7 codes only
    with pytest.raises(ced.CedError):
        ctx.decode_batch(ced.K7_DEFAULT, segs, 60)                    # not a multiple of 8
    with pytest.raises(ced.CedError):
        ctx.decode_batch(ced.K7_DEFAULT, segs, 128)                   # stride shorter than a frame
    out = ctx.decode_batch(ced.K7_DEFAULT, segs, 64, n_frames=0)      # empty batch is a no-op
    assert out.shape[0] == 0


# ------------------------------------------------------------------ packed wire format
def np_pack(segs, T):
    """numpy model of the packed format: segment t -> bits 2*(t%4).. of byte t/4."""
    nf = segs.shape[0]
    padded = np.zeros((nf, (T + 3) // 4 * 4), dtype=np.uint8)
    padded[:, :T] = segs[:, :T] & 3
    q = padded.reshape(nf, -1, 4)
    return (q[..., 0] | (q[..., 1] << 2) | (q[..., 2] << 4) | (q[..., 3] << 6)).astype(np.uint8)


@pytest.mark.parametrize("bits,frames,stride_pad,offset", [
    (8, 1, 0, 0), (16, 5, 0, 0), (40, 33, 1, 0), (184, 127, 3, 0), (192, 128, 0, 0), (376, 129, 2, 1),
    (2048, 257, 16 - (514 % 16), 0), (4096, 300, 16 - (1026 % 16), 0), (4096, 70, 0, 0), (16384, 9, 7, 0),
])
def test_packed_format_pack_and_decode(torch_cuda, ctx, port, bits, frames, stride_pad, offset):
    """No reference oracle exists for this format; it is pinned by: device pack == numpy pack of the
    byte symbols, and decode(packed) == decode(bytes) == oracle, on aligned and unaligned rows."""
    torch = torch_cuda
    rng = np.random.default_rng(bits + 7 * frames)
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    noisy = bsc(rng, port.encode_batch(7, K7, msgs), 0.06, junk_upper_bits=True)
    want = port.decode_batch(7, K7, noisy, T)
    d_noisy = dev(torch, noisy)
    pb = (T + 3) // 4
    packed = ctx.pack_symbols(d_noisy, T)
    ctx.sync()
    assert np.array_equal(packed.cpu().numpy(), np_pack(noisy, T))
    stride = pb + stride_pad
    flat = torch.full((frames * stride + offset + 64,), 0xFF, dtype=torch.uint8, device="cuda")
    view = flat[offset:offset + frames * stride].view(frames, stride)
    view[:, :pb] = packed
    out = ctx.decode_batch_packed(ced.K7_DEFAULT, view, bits)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("nbytes,frames,pad", [(1, 3, 0), (5, 40, 1), (32, 129, 2), (512, 200, 14), (2048, 5, 0)])
def test_packed_encoder_output(torch_cuda, ctx, port, nbytes, frames, pad):
    torch = torch_cuda
    rng = np.random.default_rng(nbytes * 3 + frames)
    msgs = rng.integers(0, 256, (frames, nbytes), dtype=np.uint8)
    T = 8 * nbytes + 6
    pb = (T + 3) // 4
    out = torch.full((frames, pb + pad), 0xEE, dtype=torch.uint8, device="cuda")
    ctx.encode_batch_packed(ced.K7_DEFAULT, dev(torch, msgs), out=out)
    ctx.sync()
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :pb], np_pack(port.encode_batch(7, K7, msgs), T))
    assert (got[:, pb:] == 0xEE).all()
    k3 = ctx.encode_batch_packed(ced.Code(3, (0b111, 0b110)), dev(torch, msgs))
    ctx.sync()
    assert np.array_equal(k3.cpu().numpy(), np_pack(port.encode_batch(3, (0b111, 0b110), msgs), 8 * nbytes + 2))


@pytest.mark.parametrize("bits,frames,pad", [(8, 2, 0), (40, 33, 3), (256, 129, 4), (4096, 100, 28)])
def test_soft_symbols_are_hard_sliced(torch_cuda, ctx, port, bits, frames, pad):
    """Soft input has no reference oracle; it is pinned by 'hard-slicing the soft symbols must reproduce
    the hard path bit-exactly' (SURVEY 8c): sign = coded bit, any magnitude (0 slices to bit 0)."""
    torch = torch_cuda
    rng = np.random.default_rng(bits + frames)
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    noisy = bsc(rng, port.encode_batch(7, K7, msgs), 0.07)
    mag = rng.integers(0, 128, (frames, T, 2), dtype=np.int16)
    bit = np.stack([noisy & 1, (noisy >> 1) & 1], axis=-1).astype(np.int16)
    soft = np.where(bit == 1, -np.maximum(mag, 1), mag).astype(np.int8)       # bit 1 -> negative, bit 0 -> >= 0
    buf = np.zeros((frames, 2 * T + pad), dtype=np.int8)
    buf[:, :2 * T] = soft.reshape(frames, 2 * T)
    packed = ctx.slice_soft_symbols(dev(torch, buf), T)
    ctx.sync()
    assert np.array_equal(packed.cpu().numpy(), np_pack(noisy, T))
    out = ctx.decode_batch_packed(ced.K7_DEFAULT, packed, bits)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), port.decode_batch(7, K7, noisy, T))


def test_packed_host_pipeline(torch_cuda, ctx, port):
    torch = torch_cuda
    rng = np.random.default_rng(33)
    frames, bits = 40000, 256
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    noisy = bsc(rng, port.encode_batch(7, K7, msgs), 0.04)
    packed = np.zeros((frames, 80), dtype=np.uint8)
    packed[:, :(T + 3) // 4] = np_pack(noisy, T)
    out = torch.empty((frames, bits // 8), dtype=torch.uint8).pin_memory()
    ctx.decode_batch_packed_host(ced.K7_DEFAULT, torch.from_numpy(packed).pin_memory(), bits, out)
    sample = rng.choice(frames, 300, replace=False)
    assert np.array_equal(out.numpy()[sample], port.decode_batch(7, K7, noisy[sample], T))


# ------------------------------------------------------------------ batch encode
@pytest.mark.parametrize("nbytes,frames,stride_pad", [(1, 1, 0), (2, 5, 0), (3, 40, 3), (32, 100, 10), (256, 257, 10),
                                                      (512, 300, 10), (512, 64, 0), (2048, 7, 10)])
def test_encode_batch_vs_oracle(torch_cuda, ctx, port, golden, nbytes, frames, stride_pad):
    torch = torch_cuda
    rng = np.random.default_rng(nbytes + frames)
    msgs = rng.integers(0, 256, (frames, nbytes), dtype=np.uint8)
    T = 8 * nbytes + 6
    out = torch.full((frames, T + stride_pad), 0xEE, dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, dev(torch, msgs), out=out)
    ctx.sync()
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :T], port.encode_batch(7, K7, msgs))
    assert (got[:, T:] == 0xEE).all()  # padding untouched


@pytest.mark.parametrize("K,g", [(7, (0o113, 0o171)), (7, (0o133, 0o171)), (9, (0o561, 0o753)), (3, (0b111, 0b110)),
                                 (5, (0o23, 0o35))])
def test_encode_batch_table_kernel(torch_cuda, ctx, port, K, g):
    """16-byte aligned segment rows, n = 2: encodeBatchLutKernel (a warp per frame, 512 bytes per warp store, table
    spread).  Frame lengths around the 64-byte iteration boundaries, many frames per warp, canaries."""
    torch = torch_cuda
    code = ced.Code(K, g)
    rng = np.random.default_rng(K)
    for nbytes, frames in ((1, 5), (3, 40), (4, 3), (8, 1000), (33, 50), (62, 9), (64, 70), (66, 9), (124, 33), (128, 70),
                           (132, 9), (512, 2500), (1000, 17), (2048, 5)):
        msgs = rng.integers(0, 256, (frames, nbytes), dtype=np.uint8)
        msgs[0] = 0xFF
        T = 8 * nbytes + K - 1
        stride = (T + 15) // 16 * 16 + 16
        out = torch.full((frames, stride), 0xEE, dtype=torch.uint8, device="cuda")
        ctx.encode_batch(code, dev(torch, msgs), out=out)
        ctx.sync()
        got = out.cpu().numpy()
        assert np.array_equal(got[:, :T], port.encode_batch(K, g, msgs)), (K, g, nbytes)
        assert (got[:, T:] == 0xEE).all(), (K, g, nbytes)
    # message rows embedded in a wider, 4-byte aligned array
    wide = rng.integers(0, 256, (300, 72), dtype=np.uint8)
    d_wide = dev(torch, wide)
    out = torch.full((300, 8 * 64 + 16), 0xEE, dtype=torch.uint8, device="cuda")
    ctx.encode_batch(code, d_wide[:, 4:68], out=out)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy()[:, :8 * 64 + K - 1], port.encode_batch(K, g, wide[:, 4:68]))


def test_encode_batch_matches_reference_fixtures_and_other_codes(torch_cuda, ctx, port, golden):
    torch = torch_cuda
    for bits in (8, 64, 256, 2048, 4096):
        out = ctx.encode_batch(ced.K7_DEFAULT, dev(torch, golden["msg_%d" % bits]))
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), golden["segs_%d" % bits])
    rng = np.random.default_rng(4)
    msgs = rng.integers(0, 256, (50, 33), dtype=np.uint8)
    for K, g in ((3, (0b111, 0b110)), (7, (0o133, 0o171)), (9, (0o561, 0o753)), (7, (0o133, 0o145, 0o175))):
        out = ctx.encode_batch(ced.Code(K, g), dev(torch, msgs))
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), port.encode_batch(K, g, msgs)), (K, g)


# ------------------------------------------------------------------ host-buffer (e2e) entry points
def test_host_buffer_pipeline_roundtrip(torch_cuda, ctx, port):
    torch = torch_cuda
    rng = np.random.default_rng(21)
    frames, bits = 40000, 256     # > 2 pipeline chunks of 16384 frames
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    T = bits + 6
    segs = torch.empty((frames, T + 10), dtype=torch.uint8).pin_memory()
    ctx.encode_batch_host(ced.K7_DEFAULT, torch.from_numpy(msgs).pin_memory(), segs)
    sample = rng.choice(frames, 300, replace=False)
    assert np.array_equal(segs.numpy()[sample, :T], port.encode_batch(7, K7, msgs[sample]))
    noisy = segs.numpy().copy()
    noisy[:, :T] = bsc(rng, noisy[:, :T], 0.04)
    out = np.zeros((frames, bits // 8), dtype=np.uint8)       # pageable host memory also works
    ctx.decode_batch_host(ced.K7_DEFAULT, noisy, bits, out)
    assert np.array_equal(out[sample], port.decode_batch(7, K7, noisy[sample, :T], T))


def test_host_transfer_compression_modes_agree(torch_cuda, ctx, port, monkeypatch):
    """ced_decode_batch_host on page-locked buffers: raw copies, every chunk packed by the host threads, and the
    adaptive mix (the default) must return the same bytes, equal to the oracle's on a sample."""
    torch = torch_cuda
    rng = np.random.default_rng(77)
    frames, bits = 70000, 512      # 9 pipeline chunks of 8192 frames, ragged last one
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    noisy = np.full((frames, T + 7), 0xEC, dtype=np.uint8)     # rows not 16-byte aligned, upper bits set
    noisy[:, :T] = bsc(rng, port.encode_batch(7, K7, msgs), 0.05) | 0xA0
    h_in = torch.from_numpy(noisy).pin_memory()
    sample = rng.choice(frames, 200, replace=False)
    want = port.decode_batch(7, K7, noisy[sample, :T] & 3, T)
    outs = []
    for mode, look in (("0", "1"), ("1", "1"), ("2", "1"), ("2", "3"), (None, None)):
        if mode is None:
            monkeypatch.delenv("CED_HOST_PACK", raising=False)
            monkeypatch.delenv("CED_HOST_PACK_LOOKBACK", raising=False)
        else:
            monkeypatch.setenv("CED_HOST_PACK", mode)
            monkeypatch.setenv("CED_HOST_PACK_LOOKBACK", look)
        out = torch.zeros((frames, bits // 8), dtype=torch.uint8).pin_memory()
        ctx.decode_batch_host(ced.K7_DEFAULT, h_in, bits, out)
        assert np.array_equal(out.numpy()[sample], want), mode
        outs.append(out.numpy().copy())
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    # a run-time K=7 code takes the same route
    code = ced.Code(7, (0o117, 0o155))
    noisy2 = bsc(rng, port.encode_batch(7, (0o117, 0o155), msgs[:20000]), 0.03)
    h2 = torch.from_numpy(noisy2).pin_memory()
    out2 = torch.zeros((20000, bits // 8), dtype=torch.uint8).pin_memory()
    monkeypatch.setenv("CED_HOST_PACK", "2")
    monkeypatch.setenv("CED_HOST_PACK_LOOKBACK", "1")
    ctx.decode_batch_host(code, h2, bits, out2)
    assert np.array_equal(out2.numpy()[:100], port.decode_batch(7, (0o117, 0o155), noisy2[:100], T))


# ------------------------------------------------------------------ full-size properties
def test_full_size_roundtrip_and_sampled_parity(torch_cuda, ctx, port):
    """BASELINE config 2 shape: 2^16 frames x 4096 bits.  encode -> decode must be the identity;
    encode -> BSC(5 dB) -> decode must equal the oracle on a random sample of frames, and the
    on-device error counter must equal a host recount."""
    torch = torch_cuda
    frames, bits, stride = 1 << 16, 4096, 4112
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=314)
    segs = torch.zeros((frames, stride), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
    dec = ctx.decode_batch(ced.K7_DEFAULT, segs, bits)
    ctx.sync()
    assert torch.equal(dec, msgs)
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    ctx.bsc_channel(segs, bits + 6, 2, 0.0377, seed=7, counters=counters[:2])
    dec = ctx.decode_batch(ced.K7_DEFAULT, segs, bits)
    ctx.ber_count(dec, msgs, counters[2:])
    ctx.sync()
    c = counters.cpu().numpy()
    assert c[1] == frames * (bits + 6) * 2 and c[3] == frames * bits
    assert abs(c[0] / c[1] - 0.0377) < 2e-4
    assert c[2] == int(np.bitwise_count((dec ^ msgs).cpu().numpy()).sum())
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(frames, 256, replace=False))
    noisy = segs[torch.from_numpy(sample).cuda()].cpu().numpy()[:, :bits + 6]
    assert np.array_equal(dec[torch.from_numpy(sample).cuda()].cpu().numpy(),
                          port.decode_batch(7, K7, noisy, bits + 6))
    # and EVERY one of the 65,536 frames against the unmodified reference decoder, when it was built
    R = oracle.ref()
    if R is not None:
        want = R.decode_batch_mt(segs.cpu().numpy(), bits + 6)
        got = dec.cpu().numpy()
        assert np.array_equal(got, want), "frames differing from the reference: %d" % int((got != want).any(axis=1).sum())


def test_channel_is_independent_of_sharding(torch_cuda, ctx):
    torch = torch_cuda
    frames, T = 1000, 262
    whole = torch.zeros((frames, T), dtype=torch.uint8, device="cuda")
    ctx.bsc_channel(whole, T, 2, 0.1, seed=5)
    parts = torch.zeros((frames, T), dtype=torch.uint8, device="cuda")
    for lo, hi in ((0, 333), (333, 334), (334, 1000)):
        ctx.bsc_channel(parts[lo:hi], T, 2, 0.1, seed=5, first_frame=lo)
    ctx.sync()
    assert torch.equal(whole, parts)
    assert 0.09 < (whole & 1).float().mean().item() < 0.11 and int(whole.max()) <= 3


# ------------------------------------------------------------------ reference-named per-frame API
def test_handtraced_sequence_through_the_dropin_api():
    """handTracedTest/handTraced.c:29-111 step for step (K=3 testParams)."""
    api = ced.RefApi("k3")
    enc = api.encoder()
    enc.resetConvEncoder()
    enc.initConvEncoder()
    assert enc.convEnc(np.array([0b01101000], dtype=np.uint8), True).tolist() == [0, 3, 0, 2, 2, 3, 1, 0, 0, 0]
    dec = api.decoder()
    dec.VITERBI_RESET()
    dec.VITERBI_INIT()
    api.viterbiConfigCheck()
    corrupted = np.array([1, 3, 1, 2, 2, 3, 1, 0, 0, 0], dtype=np.uint8)
    assert dec.VITERBI_DECODER_HARD(corrupted, True).tolist() == [0b01101000]
    dec.VITERBI_RESET()
    assert dec.nodeMetricsCur().tolist() == [0, 5, 5, 5]
    for i, want in enumerate([[1, 1, 6, 5], [3, 1, 1, 3], [1, 3, 2, 2], [2, 2, 2, 4]]):
        assert dec.VITERBI_DECODER_HARD(corrupted[i:i + 1], False).size == 0
        assert dec.nodeMetricsCur().tolist() == want


def test_streaming_api_matches_reference_metrics_and_chunking(golden, port):
    api = ced.RefApi("k7")
    dec = api.decoder()
    dec.VITERBI_RESET()
    dec.VITERBI_INIT()
    noise, want_dec, want_metrics = golden["stream_noise"], golden["stream_dec"], golden["stream_metrics"]
    for c in range(want_metrics.shape[0]):
        assert dec.VITERBI_DECODER_HARD(noise[64 * c:64 * c + 64], False).size == 0
        assert np.array_equal(dec.nodeMetricsCur(), want_metrics[c]), c     # exact 121-step renorm schedule
    assert np.array_equal(dec.VITERBI_DECODER_HARD(noise[:0], True), want_dec)  # last with segmentsIn == 0
    assert dec.nodeMetricsCur().tolist() == [0] + [65] * 63                   # reset after last
    assert np.array_equal(dec.VITERBI_DECODER_HARD(noise, True), want_dec)     # one-shot == chunked
    # ragged chunking, maximum packet length
    rng = np.random.default_rng(8)
    big = rng.integers(0, 4, 16384 + 6, dtype=np.uint8)
    want = port.decode_batch(7, K7, big[None, :], big.size)[0]
    pos = 0
    for step in (1, 5, 121, 4097, 3000, 9166):
        assert dec.VITERBI_DECODER_HARD(big[pos:pos + step], False).size == 0
        pos += step
    assert pos == big.size
    assert np.array_equal(dec.VITERBI_DECODER_HARD(big[:0], True, max_bytes=4096), want)


@pytest.mark.parametrize("split", ["default", "0", "134", "16"])
@pytest.mark.parametrize("T", [13, 14, 33, 38, 39, 65, 70, 129, 130, 133, 134, 262, 1030, 2054, 2055, 4102, 16390])
def test_one_shot_packets_take_the_frame_parallel_kernels(port, T, split, monkeypatch):
    """A whole K=7 packet in ONE VITERBI_DECODER_HARD(last=true) call runs two parallel kernels in one graph launch:
    wsBlockKernel / wsJoinKernel (csrc/warp_split.cu: the packet cut into blocks that start from guessed metrics, the
    hand-overs checked -- the default from 134 segments on when the length is whole bytes + 6) or fpBlockKernel /
    fpSelectKernel (csrc/frame_parallel.cuh; CED_STREAM_SPLIT=0 and every other length).  Same bytes as the sequential
    decoder for clean, noisy, all-zero and pure-noise packets, for lengths that leave a short last block or a partial
    last byte, and the decoder is usable for chunked packets afterwards."""
    monkeypatch.setenv("CED_STREAM_SERVER", "0")
    monkeypatch.delenv("CED_WARP_SPLIT", raising=False)
    if split != "default":
        monkeypatch.setenv("CED_STREAM_SPLIT", split)
    api = ced.RefApi("k7")
    dec = api.decoder()
    dec.VITERBI_RESET()
    dec.VITERBI_INIT()
    rng = np.random.default_rng(T)
    L = T - 6
    lib = ced.load_abi()
    count = lambda: int(lib.ced_launch_count(lib.ced_default_ctx()))
    for rep, p in enumerate((0.0, 0.04, 0.12, 0.5, None, "zero")):
        msg = rng.integers(0, 256, (1, (L + 7) // 8), dtype=np.uint8)
        segs = port.encode_batch(7, K7, msg)[:, :T].copy()
        if p is None:
            segs[:] = rng.integers(0, 256, segs.shape)       # only the low two bits of a byte count
        elif p == "zero":
            segs[:] = 0
        else:
            segs = bsc(rng, segs, p)
        want = port.decode_batch(7, K7, segs, T)[0][:(L - 1) // 8 + 1]
        launches = count()
        got = dec.VITERBI_DECODER_HARD(segs[0], True, max_bytes=4096)
        assert count() - launches == 2, "expected the two frame-parallel kernels"
        assert np.array_equal(got, want), (T, rep)
        assert dec.nodeMetricsCur().tolist() == [0] + [65] * 63
    # the same packet in two calls goes through the sequential kernel and gives the same bytes
    assert dec.VITERBI_DECODER_HARD(segs[0, :7], False).size == 0
    assert np.array_equal(dec.VITERBI_DECODER_HARD(segs[0, 7:], True, max_bytes=4096), want)


def server_stats():
    import ctypes as C
    lib = ced.load_abi()
    req, lau = C.c_uint64(0), C.c_uint64(0)
    on = lib.ced_stream_server_stats(C.byref(req), C.byref(lau))
    return on, req.value, lau.value


@pytest.mark.parametrize("T", [13, 14, 33, 38, 39, 65, 70, 129, 130, 133, 134, 262, 1030, 2054, 2055, 4102, 16390])
def test_one_shot_packets_through_the_resident_kernel(port, T, monkeypatch):
    """The same packets through fpServerKernel (csrc/frame_server.cuh, CED_STREAM_SERVER=1): the packet goes into the
    mailbox, the resident kernel answers; no launch per call.  Same bytes as the sequential decoder."""
    monkeypatch.setenv("CED_STREAM_SERVER", "1")
    api = ced.RefApi("k7")
    dec = api.decoder()
    dec.VITERBI_RESET()
    dec.VITERBI_INIT()
    rng = np.random.default_rng(1000 + T)
    L = T - 6
    for rep, p in enumerate((0.0, 0.04, 0.12, 0.5, None, "zero")):
        msg = rng.integers(0, 256, (1, (L + 7) // 8), dtype=np.uint8)
        segs = port.encode_batch(7, K7, msg)[:, :T].copy()
        if p is None:
            segs[:] = rng.integers(0, 256, segs.shape)
        elif p == "zero":
            segs[:] = 0
        else:
            segs = bsc(rng, segs, p)
        want = port.decode_batch(7, K7, segs, T)[0][:(L - 1) // 8 + 1]
        _, req0, _ = server_stats()
        got = dec.VITERBI_DECODER_HARD(segs[0], True, max_bytes=4096)
        on, req1, _ = server_stats()
        assert on == 1 and req1 - req0 == 1, "expected the resident packet decoder to answer"
        assert np.array_equal(got, want), (T, rep)
        assert dec.nodeMetricsCur().tolist() == [0] + [65] * 63
    assert dec.VITERBI_DECODER_HARD(segs[0, :7], False).size == 0
    assert np.array_equal(dec.VITERBI_DECODER_HARD(segs[0, 7:], True, max_bytes=4096), want)


def test_resident_kernel_coexists_with_the_rest_of_the_library(torch_cuda, ctx, port, monkeypatch):
    """Packets through the resident kernel interleaved with per-frame encodes, chunked decodes, batch decodes on another
    context, idle periods longer than its time-out, and device-wide synchronisation: it steps aside and comes back."""
    import time
    monkeypatch.setenv("CED_STREAM_SERVER", "1")
    torch = torch_cuda
    api = ced.RefApi("k7")
    dec, enc = api.decoder(), api.encoder()
    dec.VITERBI_RESET(); dec.VITERBI_INIT()
    enc.resetConvEncoder(); enc.initConvEncoder()
    rng = np.random.default_rng(99)
    msgs = rng.integers(0, 256, (40, 256), dtype=np.uint8)
    clean = port.encode_batch(7, K7, msgs)
    noisy = bsc(rng, clean, 0.03)
    want = port.decode_batch(7, K7, noisy, 2054)
    d_noisy = torch.from_numpy(noisy).cuda()
    _, req0, lau0 = server_stats()
    for i in range(40):
        if i % 5 == 1:
            assert np.array_equal(enc.convEnc(msgs[i], True), clean[i])
        if i % 7 == 2:
            time.sleep(0.01)                                   # longer than the idle time-out: it has left
        if i % 9 == 3:
            out = ctx.decode_batch(ced.K7_DEFAULT, d_noisy, 2048)
            ctx.sync()
            assert np.array_equal(out.cpu().numpy(), want)
        if i % 11 == 4:
            torch.cuda.synchronize()
        if i % 6 == 5:                                         # a chunked packet in between
            assert dec.VITERBI_DECODER_HARD(noisy[i, :1000], False).size == 0
            assert np.array_equal(dec.VITERBI_DECODER_HARD(noisy[i, 1000:], True, max_bytes=4096), want[i])
        assert np.array_equal(dec.VITERBI_DECODER_HARD(noisy[i], True, max_bytes=4096), want[i]), i
    on, req1, lau1 = server_stats()
    assert on == 1 and req1 - req0 == 40
    assert 2 <= lau1 - lau0 <= 30


def test_streaming_encoder_chunks_and_kat(golden, port):
    api = ced.RefApi("k7")
    enc = api.encoder()
    enc.resetConvEncoder()
    enc.initConvEncoder()
    assert np.array_equal(enc.convEnc(golden["kat_msg"], True), golden["kat_segs"])
    rng = np.random.default_rng(2)
    msg = rng.integers(0, 256, 1000, dtype=np.uint8)
    parts, pos = [], 0
    for step in (1, 2, 3, 250, 744):
        parts.append(enc.convEnc(msg[pos:pos + step], pos + step == msg.size))
        pos += step
    assert np.array_equal(np.concatenate(parts), port.encode(7, K7, msg)[0])
    # encoder state was reset by last=true: a second packet encodes from state 0
    assert np.array_equal(enc.convEnc(golden["kat_msg"], True), golden["kat_segs"])


def test_bertest_first_packets_through_the_dropin_api(golden):
    """The first 64 packets of berTestK7's first configuration (same rand() stream as the reference
    binary) decoded through VITERBI_DECODER_HARD: same decoded-error count as the reference."""
    api = ced.RefApi("k7")
    dec = api.decoder()
    dec.VITERBI_RESET()
    dec.VITERBI_INIT()
    noisy, msgs, counts = golden["ber64_noisy"], golden["ber64_msgs"], golden["ber64_counts"]
    errs = 0
    for f in range(noisy.shape[0]):
        out = dec.VITERBI_DECODER_HARD(noisy[f], True)
        errs += int(np.unpackbits(out ^ msgs[f]).sum())
    assert errs == int(counts[2])


def test_bertestk7_golden_integers_through_the_batch_path(torch_cuda, ctx, port):
    """berTestK7's three configurations (berTestK7/berTestK7.c:95-165, srand(9865)): the very packets the
    reference driver generates (same rand() stream, reproduced by the oracle's driver restatement) decoded
    10000 at a time by ced_decode_batch must give the integers the reference binary prints:
    92418 / 9655 / 655 decoded bit errors of 20,480,000."""
    torch = torch_cuda
    golden = [(2296339, 92418), (1525431, 9655), (928843, 655)]
    port.lib.orc_srand(9865)
    for (flips, errors), p in zip(golden, (5.585640e-02, 3.716174e-02, 2.262231e-02)):
        counts, noisy, msgs = port.bertest(7, K7, 10000, 256, p, want_data=True)
        assert counts.tolist() == [flips, 41080000, errors, 20480000]
        d_noisy, d_msgs = dev(torch, noisy), dev(torch, msgs)
        dec = ctx.decode_batch(ced.K7_DEFAULT, d_noisy, 2048)
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        ctx.ber_count(dec, d_msgs, cnt)
        ctx.sync()
        assert cnt.tolist() == [errors, 20480000]
        packed = ctx.pack_symbols(d_noisy, 2054)
        dec_p = ctx.decode_batch_packed(ced.K7_DEFAULT, packed, 2048)
        ctx.sync()
        assert torch.equal(dec, dec_p)


def test_batches_in_flight_on_separate_contexts(torch_cuda, port):
    """What bench.py does for `value`: several decodes in flight, one ced_ctx + CUDA stream each, plus two
    calls on ONE context from different streams (serialised by the context's event).  Every output must
    still equal the oracle's."""
    torch = torch_cuda
    rng = np.random.default_rng(5)
    bits, frames = 2048, 3000
    T = bits + 6
    ctxs = [ced.Context(0) for _ in range(3)]
    streams = [torch.cuda.Stream() for _ in range(4)]
    inputs, outs = [], []
    for i in range(4):
        msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
        noisy = bsc(rng, port.encode_batch(7, K7, msgs), 0.05)
        inputs.append(noisy)
    d_in = [dev(torch, x) for x in inputs]
    torch.cuda.synchronize()
    for rep in range(3):
        outs = []
        for i in range(4):
            c = ctxs[min(i, 2)]                      # lanes 2 and 3 share a context on different streams
            outs.append(c.decode_batch(ced.K7_DEFAULT, d_in[i], bits, stream=streams[i]))
        torch.cuda.synchronize()
        sample = rng.choice(frames, 64, replace=False)
        for i in range(4):
            assert np.array_equal(outs[i].cpu().numpy()[sample], port.decode_batch(7, K7, inputs[i][sample], T)), (rep, i)
    for c in ctxs:
        c.close()


def test_concurrent_host_threads(torch_cuda, port):
    """The reference library is re-entrant per state struct (SURVEY 8b 'Threading'); here every call funnels
    into one GPU context, so calls from several host threads are serialised by the context's lock and must
    still each get their own answer: per-frame calls on private state structs next to batch calls."""
    import threading
    torch = torch_cuda
    rng = np.random.default_rng(17)
    api = ced.RefApi("k7")
    ctx = ced.Context(0)
    frames = [bsc(rng, port.encode_batch(7, K7, rng.integers(0, 256, (1, 64), dtype=np.uint8)), 0.05)[0]
              for _ in range(6)]
    want = [port.decode_batch(7, K7, f[None, :], 518)[0] for f in frames]
    batch_in = bsc(rng, port.encode_batch(7, K7, rng.integers(0, 256, (500, 64), dtype=np.uint8)), 0.05)
    batch_want = port.decode_batch(7, K7, batch_in, 518)
    errors = []

    def per_frame(i):
        try:
            dec = api.decoder()
            dec.VITERBI_RESET()
            dec.VITERBI_INIT()
            for rep in range(25):
                a = dec.VITERBI_DECODER_HARD(frames[i][:200], False)
                b = dec.VITERBI_DECODER_HARD(frames[i][200:], True)
                if a.size or not np.array_equal(b, want[i]):
                    errors.append(("frame", i, rep))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def batch():
        try:
            torch.cuda.set_device(0)
            d_in = torch.from_numpy(batch_in).cuda()
            for rep in range(25):
                out = ctx.decode_batch(ced.K7_DEFAULT, d_in, 512)
                ctx.sync()
                if not np.array_equal(out.cpu().numpy(), batch_want):
                    errors.append(("batch", rep))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=per_frame, args=(i,)) for i in range(6)] + [threading.Thread(target=batch)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    ctx.close()
    assert not errors, errors[:5]


def test_ber_sweep_subset_identical_to_reference_decoder(torch_cuda):
    """BASELINE config 4 in miniature: per Eb/N0 point the GPU's decoded bytes equal the reference
    decoder's on the same hard symbols, so the BER curves are identical."""
    import ber_sweep
    res = ber_sweep.main(["--frames-per-gpu", "4096", "--subset", "512", "--points", "0,3,5,8"])
    ps = [r["bsc_p"] for r in res["points"]]
    assert abs(ps[0] - 0.1587) < 2e-4 and abs(ps[2] - 0.0377) < 2e-4
    for r in res["points"]:
        chk = r["subset_check"]
        assert chk["bytes_identical"] and chk["reference_decoded_errors"] == chk["gpu_decoded_errors"]
        assert abs(r["channel_ber"] - r["bsc_p"]) < 0.02 * r["bsc_p"] + 1e-4
    bers = [r["decoded_ber"] for r in res["points"]]
    assert bers[0] > bers[1] > bers[2] > bers[3]


def test_multi_wave_batches(torch_cuda, port):
    """Batches larger than one wave of survivor scratch are decoded wave by wave (forced small here)."""
    code = ("import os, sys, numpy as np, torch\n"
            "sys.path.insert(0, %r)\n"
            "import convolutionalencdec_b200 as ced\n"
            "ctx = ced.Context(0)\n"
            "msgs = torch.empty((1000, 32), dtype=torch.uint8, device='cuda'); ctx.random_bytes(msgs, seed=3)\n"
            "segs = ctx.encode_batch(ced.K7_DEFAULT, msgs, seg_stride=272)\n"
            "ctx.bsc_channel(segs, 262, 2, 0.05, seed=4)\n"
            "dec = ctx.decode_batch(ced.K7_DEFAULT, segs, 256); ctx.sync()\n"
            "np.save(sys.argv[1], dec.cpu().numpy()); np.save(sys.argv[2], segs.cpu().numpy())\n" % ROOT)
    import tempfile
    outs = []
    for wave in ("128", "1048576"):
        with tempfile.TemporaryDirectory() as d:
            a, b = os.path.join(d, "dec.npy"), os.path.join(d, "segs.npy")
            env = dict(os.environ, CED_MAX_WAVE_FRAMES=wave)
            subprocess.run([sys.executable, "-c", code, a, b], check=True, env=env, timeout=300)
            outs.append((np.load(a), np.load(b)))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][0], port.decode_batch(7, K7, outs[0][1][:, :262], 262))


def test_exactly_sized_buffers_and_awkward_shapes():
    """tools/sanitize_cases.py: every kernel on buffers with no slack behind them, misaligned bases, generic
    codes, soft input, chunked per-frame calls (the script a memcheck run would use; compute-sanitizer is
    closed on this pool, so it runs bare here)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_cases.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_plain_c_user_of_the_batch_abi():
    """examples/batch_roundtrip.c: a C program (no CUDA headers) encodes, corrupts and decodes 20000 frames
    through ced_encode_batch_host / ced_decode_batch_host."""
    path = os.path.join(ROOT, "examples", "_bin", "batch_roundtrip")
    if not os.path.exists(path):
        pytest.skip("examples/_bin/batch_roundtrip not built")
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 wrong bytes" in r.stdout, r.stdout + r.stderr


def test_packet_queue_keeps_the_per_packet_loop_shape():
    """examples/speed_queued.c: the reference's one-packet-per-call decode loop with Submit/Flush of
    include/viterbiDecoderQueue.h; the program checks every decoded packet itself."""
    import re
    path = os.path.join(ROOT, "examples", "_bin", "speed_queued")
    if not os.path.exists(path):
        pytest.skip("examples/_bin/speed_queued not built")
    for per_batch in ("8192", "1000"):        # 1000 does not divide the 16384 packets: partial batches at Flush
        r = subprocess.run([path, "1.0", per_batch], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Success!" in r.stdout, r.stdout + r.stderr
        rate = float(re.search(r"^Rate: ([0-9.]+) Mbps", r.stdout, re.M).group(1))
        enc_rate = float(re.search(r"Encode rate: ([0-9.]+) Mbps", r.stdout).group(1))
        assert rate > 500.0 and enc_rate > 500.0, r.stdout   # the synchronous per-packet calls: ~8 / ~400 Mbps


# ------------------------------------------------------------------ the reference's own drivers, unchanged
def _driver(name):
    path = os.path.join(ROOT, "drivers", "_bin", name)
    if not os.path.exists(path):
        pytest.skip("drivers/_bin/%s not built" % name)
    return path


def test_driver_handtraced_passes():
    r = subprocess.run([_driver("handTraced")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "++++ Test Passed! ++++" in r.stdout and "Decoded 0x68, Expected 0x68" in r.stdout


def test_driver_bertestk7_reproduces_golden_integers():
    r = subprocess.run([_driver("berTestK7")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "Success!" in r.stdout
    rows = [l for l in r.stdout.splitlines() if "|" in l and l.strip()[0] in "-0123456789"]
    got = [[int(x) for x in l.replace("|", " ").split() if x.isdigit()] for l in rows]
    assert got == [[2296339, 41080000, 92418, 20480000], [1525431, 41080000, 9655, 20480000],
                   [928843, 41080000, 655, 20480000]], r.stdout[-3000:]


@pytest.mark.parametrize("name,word", [("speedDecode", "Decoded"), ("speedEncode", "Encoded")])
def test_speed_drivers_run(name, word):
    path = _driver(name)
    try:
        r = subprocess.run(["timeout", "-s", "INT", "9", "stdbuf", "-oL", path], capture_output=True, text=True, timeout=60)
    except subprocess.TimeoutExpired:
        pytest.fail("%s did not stop" % name)
    assert "Could not" not in r.stdout, r.stdout
    assert "Rate:" in r.stdout and "Mbps" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]


def test_large_batch_is_decoded_as_waves_in_flight(torch_cuda, ctx, port):
    """ced_decode_batch of several GPU-fills runs as waves of 2^16 frames on three internal streams
    (decodeBatchPipelined): same bytes as the same frames decoded in single-wave calls, a sample against the
    oracle, ragged last wave, caller's stream semantics kept (the result is complete when the caller's stream is)."""
    torch = torch_cuda
    frames, bits = 3 * 65536 + 777, 64
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=21)
    segs = torch.zeros((frames, 80), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
    ctx.bsc_channel(segs, T, 2, 0.05, seed=3)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    l0 = ctx.launches
    out = torch.zeros((frames, bits // 8), dtype=torch.uint8, device="cuda")
    stream.wait_stream(torch.cuda.current_stream())
    ctx.decode_batch(ced.K7_DEFAULT, segs, bits, out=out, stream=stream)
    stream.synchronize()                       # only the caller's stream: the internal streams were joined into it
    assert ctx.launches - l0 == 2 * 4          # four waves, forward + traceback each
    pieces = [ctx.decode_batch(ced.K7_DEFAULT, segs[a:a + 100000], bits).clone() for a in range(0, frames, 100000)]
    ctx.sync()
    assert torch.equal(out, torch.cat(pieces))
    sample = np.arange(0, frames, 997)
    noisy = segs[torch.from_numpy(sample).cuda()][:, :T].cpu().numpy()
    assert np.array_equal(out.cpu().numpy()[sample], port.decode_batch(7, K7, noisy, T))


def test_randomised_parity_soak():
    """tools/fuzz_parity.py for 25 s: random codes (K = 3..9, k = 1 and 2, n = 2..4, random generators), frame shapes,
    strides, base offsets and channels through ced_decode_batch / _packed / _soft / _softq / _k and the windowed decoder,
    every result compared with the oracle (a 240 s run with 49,275 cases is kept in profiles/fuzz_parity_r2.txt)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "25", "7"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
