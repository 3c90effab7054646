"""Continuous streams with windowed traceback through the C ABI, against the oracle's definition
(oracle/ced_oracle.c orc_decode_window), bit-exact.  The reference's own windowed decoder does not run at HEAD;
the definition is tied to the reference three ways: the forward recursion is shared with ced_decode_batch, a depth
longer than the stream must equal the reference's full traceback, and -- statistically -- the definition with
one-step slices and depth 35 reproduces the MATLAB vitdec(tblen = 35) expectations held in berTestK7.c:98
(tests/test_oracle.py::test_window_definition_reproduces_matlab_tblen_expectations), while the GPU decoder at its
own slice shapes reproduces the full-traceback expectations of berTestK7.c:96-97 (last test of this file)."""
import numpy as np
import pytest

import convolutionalencdec_b200 as ced
import oracle
from conftest import bsc

pytestmark = pytest.mark.gpu
K7 = oracle.K7_G


@pytest.fixture(scope="module")
def ctx():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = ced.Context(0)
    yield c
    c.close()


def make_streams(port, rng, n_streams, total_segs, p):
    bits = total_segs - 6
    assert bits % 8 == 0
    msgs = rng.integers(0, 256, (n_streams, bits // 8), dtype=np.uint8)
    clean = port.encode_batch(7, K7, msgs)
    return msgs, bsc(rng, clean, p)


def run_window(ctx, noisy, call_segs, depth, offset=0):
    import torch
    n_streams, total = noisy.shape
    stride = (total + 15) // 16 * 16 + 16
    buf = torch.zeros(n_streams * stride + 16, dtype=torch.uint8, device="cuda")
    d = buf[offset:offset + n_streams * stride].view(n_streams, stride)
    d[:, :total] = torch.from_numpy(noisy).cuda()
    dec = ctx.window_decoder(ced.K7_DEFAULT, n_streams, depth)
    pieces, pos = [], 0
    while pos < total:
        n = total - pos if total - pos <= call_segs else call_segs
        pieces.append(dec.push(d[:, pos:pos + n], last=(pos + n == total)).clone())
        pos += n
    ctx.sync()
    return torch.cat(pieces, dim=1).cpu().numpy()


@pytest.mark.parametrize("n_streams,total,call,depth,p,offset", [
    (100, 4902, 960, 48, 0.04, 0),
    (100, 4902, 960, 24, 0.08, 0),
    (33, 96 * 30 + 6, 96, 192, 0.06, 0),      # depth longer than a slice: the first calls emit nothing
    (64, 96 * 20 + 14, 192, 96, 0.05, 1),     # misaligned base pointer
    (5, 96 * 12 + 38, 288, 48, 0.10, 0),
    (1, 96 + 6, 96, 24, 0.02, 0),
    (40, 1926, 96 * 40, 48, 0.05, 0),         # one call: first and last at once
])
def test_window_decode_matches_oracle_definition(ctx, port, n_streams, total, call, depth, p, offset):
    rng = np.random.default_rng(total * 7 + depth)
    msgs, noisy = make_streams(port, rng, n_streams, total, p)
    got = run_window(ctx, noisy, call, depth, offset)
    assert got.shape == msgs.shape
    for i in range(n_streams):
        want = port.decode_window(7, K7, noisy[i], call, depth)
        assert np.array_equal(got[i], want), (i, int(np.bitwise_count(got[i] ^ want).sum()))


def test_window_longer_than_stream_equals_full_traceback(ctx, port, ref):
    """depth >= stream length: every bit is decided by the final traceback from state 0, i.e. the reference's
    full-frame decoder (src/viterbiDecoderButterflyk1.c:200-260) on the same symbols."""
    rng = np.random.default_rng(5)
    total = 96 * 8 + 6
    msgs, noisy = make_streams(port, rng, 70, total, 0.07)
    got = run_window(ctx, noisy, 96 * 2, 96 * 9)
    assert np.array_equal(got, ref.decode_batch(noisy, total))


def test_window_long_stream_beyond_reference_packet_limit(ctx, port):
    """A stream of 2^17 bits (8x the reference's MAX_PKT_LEN_UNCODED_BITS) with 12 KB of survivors per stream in
    flight; noise-free symbols must give back the message, noisy ones stay within the full decoder's error count."""
    import torch
    rng = np.random.default_rng(11)
    total = (1 << 17) + 6
    assert (total - 6) % 96 != 0    # the last slice is ragged
    msgs, noisy = make_streams(port, rng, 32, total, 0.0)
    assert np.array_equal(run_window(ctx, noisy, 96 * 16, 48), msgs)
    noisy = bsc(rng, noisy, 0.03)
    got = run_window(ctx, noisy, 96 * 16, 48)
    want = np.stack([port.decode_window(7, K7, noisy[i], 96 * 16, 48) for i in range(4)])
    assert np.array_equal(got[:4], want)
    errs = int(np.bitwise_count(got ^ msgs).sum())
    assert errs < 1e-3 * msgs.size * 8, errs


def test_window_argument_checks(ctx):
    import torch
    segs = torch.zeros((4, 192), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ctx.window_decoder(ced.K7_DEFAULT, 4, 35)          # depth must be a multiple of 24
    dec = ctx.window_decoder(ced.K7_DEFAULT, 4, 48)
    with pytest.raises(ced.CedError):
        dec.push(segs[:, :100])                            # slices are multiples of 96
    with pytest.raises(ced.CedError):
        dec.push(segs[:, :96 + 7], last=True)              # stream must end on a byte boundary + tail
    with pytest.raises(ValueError):
        ctx.window_decoder(ced.Code(9, [0o561, 0o753]), 4, 48)   # K <= 7 only (the carry block has no size for it)
    with pytest.raises(ced.CedError):
        ctx.window_decoder(ced.Code(3, [7, 6]), 4, 48, packed=True).push(segs[:, :96])   # packed format: K=7 family only


def test_outputs_stay_inside_their_rows(ctx, port):
    """compute-sanitizer is not available on the GPU pool: canary bytes around and between the output rows of the
    batch decoder, the encoder and the windowed decoder (and around its carry block) must survive."""
    import torch
    rng = np.random.default_rng(3)
    frames, bits = 77, 96 * 6 + 88
    T = bits + 6
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    d_msgs = torch.from_numpy(msgs).cuda()

    def guarded(rows, row_bytes, stride):
        flat = torch.full((rows * stride + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        return flat, flat[32:32 + rows * stride].view(rows, stride)

    def intact(flat, view, row_bytes):
        v = view.clone()
        v[:, :row_bytes] = 0xA5
        return bool((flat[:32] == 0xA5).all() and (flat[-32:] == 0xA5).all() and (v == 0xA5).all())

    for code in (ced.K7_DEFAULT, ced.Code(7, (0o171, 0o133)), ced.Code(7, (0o133, 0o171, 0o165)), ced.Code(3, (7, 6))):
        Tc = bits + code.S
        sflat, segs = guarded(frames, Tc, Tc + 9)
        ctx.encode_batch(code, d_msgs, out=segs)
        oflat, out = guarded(frames, bits // 8, bits // 8 + 5)
        ctx.decode_batch(code, segs, bits, out=out)
        ctx.sync()
        assert intact(sflat, segs, Tc) and intact(oflat, out, bits // 8), code.g
        assert torch.equal(out[:, :bits // 8], d_msgs)
    segs = ctx.encode_batch(ced.K7_DEFAULT, d_msgs)
    wd = ctx.window_decoder(ced.K7_DEFAULT, frames, depth=48)
    carry_bytes = wd.carry.numel()
    cflat = torch.full((carry_bytes + 64,), 0xA5, dtype=torch.uint8, device="cuda")
    wd.carry = cflat[32:32 + carry_bytes]
    got = []
    for a in range(0, T, 192):
        n, last = min(192, T - a), a + 192 >= T
        rows = (n + 48) // 8
        oflat, out = guarded(frames, rows, rows + 3)
        piece = wd.push(segs[:, a:a + n], last=last, out=out)
        ctx.sync()
        assert intact(oflat, out, piece.shape[1])
        got.append(piece.clone())
    assert bool((cflat[:32] == 0xA5).all() and (cflat[-32:] == 0xA5).all())
    assert torch.equal(torch.cat(got, dim=1), d_msgs)


def run_window_packed(ctx, code, noisy, call_segs, depth):
    """the same streams in the packed wire format (4 segments per byte), slices of call_segs (a multiple of 192)"""
    import torch
    n_streams, total = noisy.shape
    d = ctx.pack_symbols(torch.from_numpy(noisy).cuda(), total, packed_stride=((total + 3) // 4 + 15) // 16 * 16)
    dec = ctx.window_decoder(code, n_streams, depth, packed=True)
    pieces, pos = [], 0
    while pos < total:
        n = total - pos if total - pos <= call_segs else call_segs
        pieces.append(dec.push(d[:, pos // 4:], last=(pos + n == total), n_segments=n).clone())
        pos += n
    ctx.sync()
    return torch.cat(pieces, dim=1).cpu().numpy()


@pytest.mark.parametrize("n_streams,total,call,depth,p", [
    (100, 4902, 960, 48, 0.04), (33, 192 * 15 + 6, 192, 384, 0.06), (5, 192 * 6 + 38, 576, 48, 0.10),
    (1, 192 + 6, 192, 24, 0.02), (40, 1926, 192 * 20, 48, 0.05),
])
def test_packed_window_decode_matches_oracle_definition(ctx, port, n_streams, total, call, depth, p):
    rng = np.random.default_rng(total * 3 + depth)
    msgs, noisy = make_streams(port, rng, n_streams, total, p)
    got = run_window_packed(ctx, ced.K7_DEFAULT, noisy, call, depth)
    for i in range(n_streams):
        assert np.array_equal(got[i], port.decode_window(7, K7, noisy[i], call, depth)), i
    with pytest.raises(ced.CedError):      # packed slices are multiples of 192 segments
        run_window_packed(ctx, ced.K7_DEFAULT, noisy, 96, depth)


def test_window_ber_meets_the_reference_held_expectations(ctx, port):
    """berTestK7's three BSC points (berTestK7.c:95-100) through ced_decode_window_batch with the generators the MATLAB
    expectations were made with (133/171, scripts/matlab/viterbiBEREstimate.m:11): slices of 96 segments at depth 48
    decide every bit with at least 48 >= 5K steps of traceback, so the BER must meet the full-traceback expectations
    of berTestK7.c:96-97 under the reference's own +-10 % rule (:167-172).  The same symbols in the packed format give
    the same bytes; a sample of packets equals the oracle's definition bit for bit."""
    import torch
    code, g = ced.K7_TEXTBOOK, (0o133, 0o171)
    bits, T = 2048, 2054
    for p, want, pkts in ((5.585640e-02, 4.765898e-03, 1 << 13), (3.716174e-02, 5.184082e-04, 1 << 15),
                          (2.262231e-02, 3.499023e-05, 1 << 17)):
        msgs = torch.empty((pkts, bits // 8), dtype=torch.uint8, device="cuda")
        ctx.random_bytes(msgs, seed=int(p * 1e6))
        segs = torch.zeros((pkts, 2064), dtype=torch.uint8, device="cuda")
        ctx.encode_batch(code, msgs, out=segs)
        ctx.bsc_channel(segs, T, 2, p, seed=17)
        dec = ctx.window_decoder(code, pkts, 48)
        pieces = [dec.push(segs[:, a:min(a + 96, T)], last=a + 96 >= T).clone() for a in range(0, T, 96)]
        out = torch.cat(pieces, dim=1).contiguous()
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        ctx.ber_count(out, msgs, cnt)
        ctx.sync()
        ber = int(cnt[0]) / int(cnt[1])
        assert int(cnt[1]) == pkts * bits and abs(ber - want) / want < 0.10, (p, ber, want)
        noisy = segs[:64, :T].cpu().numpy()
        for i in range(64):
            assert np.array_equal(out[i].cpu().numpy(), port.decode_window(7, g, noisy[i], 96, 48))
        if pkts == 1 << 13:
            assert np.array_equal(run_window_packed(ctx, code, segs[:512, :T].cpu().numpy(), 192, 48),
                                  np.stack([port.decode_window(7, g, r, 192, 48) for r in segs[:512, :T].cpu().numpy()]))


@pytest.mark.parametrize("K,g", [(3, (0b111, 0b110)), (3, (0b111, 0b101, 0b011)), (4, (0o15, 0o17)), (5, (0o23, 0o35)),
                                 (5, (0o25, 0o33, 0o37)), (7, (0o133, 0o170)), (7, (0o133, 0o145, 0o174))])
def test_window_decode_other_code_parameters(ctx, port, K, g):
    """ced_decode_window_batch for codes outside the symmetric K=7 family (table-driven kernels, K <= 7, 2 or 3 generators
    of any shape): bit-exact against the window definition with the general branch costs; with depth >= stream length it
    is the full-frame decode of the same code."""
    import torch
    rng = np.random.default_rng(K * 31 + sum(g))
    n, S = len(g), K - 1
    code = ced.Code(K, g)
    for n_streams, bits, call, depth, p in ((3, 96 * 3 + 40, 96, 24, 0.0), (70, 1000 // 8 * 8, 192, 48, 0.03),
                                            (40, 2048, 480, 96, 0.08), (33, 512, 96, 48, 0.5), (5, 512, 4800, 48, 0.03)):
        total = bits + S
        msgs = rng.integers(0, 256, (n_streams, bits // 8), dtype=np.uint8)
        clean = port.encode_batch(K, list(g), msgs)
        flips = rng.random(clean.shape + (n,)) < p
        noisy = clean.copy()
        for j in range(n):
            noisy ^= (flips[..., j].astype(np.uint8) << j)
        want = np.stack([port.decode_window(K, list(g), noisy[i], call, depth, symmetric=False) for i in range(n_streams)])
        d = torch.from_numpy(noisy).cuda()
        wd = ctx.window_decoder(code, n_streams, depth=depth)
        parts = [wd.push(d[:, a:min(a + call, total)], last=a + call >= total).clone() for a in range(0, total, call)]
        ctx.sync()
        got = torch.cat(parts, dim=1).cpu().numpy()
        assert np.array_equal(got, want), (K, g, n_streams, bits, call, depth, p)
        if call >= total:
            assert np.array_equal(got, port.decode_batch(K, list(g), noisy, total, symmetric=False))
        if p == 0.0:
            assert np.array_equal(got, msgs)
