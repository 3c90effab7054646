"""Small batches through ced_decode_batch: the warp-per-frame kernel (csrc/warp_frame.cu -- states over the lanes, radix-4
steps for 64 states, decisions in shared memory, warp-parallel exact traceback) against the CPU oracle, bit for bit."""
import os

import numpy as np
import pytest

import convolutionalencdec_b200 as ced

pytestmark = pytest.mark.gpu

K7 = [0o113, 0o171]


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    c = ced.Context(0)
    yield c
    c.close()


@pytest.fixture
def env():
    keys = ("CED_WARP_FRAME_MAX", "CED_WARP_FRAME_RADIX", "CED_WARP_FRAME_ANY_K", "CED_WARP_SPLIT", "CED_WARP_SPLIT_WARMUP",
            "CED_WARP_SPLIT_LEN")
    saved = {k: os.environ.get(k) for k in keys}
    os.environ["CED_WARP_SPLIT"] = "0"    # the tests of the one-warp-per-frame kernel; the split tests switch it on
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def noisy(rng, clean, n, p, junk=False):
    flips = rng.random(clean.shape + (n,)) < p
    out = clean.copy()
    for j in range(n):
        out ^= (flips[..., j].astype(np.uint8) << j)
    if junk and n == 2:
        out |= (rng.integers(0, 64, out.shape, dtype=np.uint8) << 2)
    return out


def place(arr, pad, off):
    import torch
    frames, T = arr.shape
    flat = torch.full((frames * (T + pad) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    view = flat[off:off + frames * (T + pad)].view(frames, T + pad)
    view[:, :T] = torch.from_numpy(arr).cuda()
    return view


@pytest.mark.parametrize("radix", [4, 2])
def test_small_batches_of_the_default_code(ctx, port, env, radix):
    """K=7 {0113, 0171}: 1 .. 600 frames of 8 .. 16384 bits, clean / noisy / pure-noise channels (where nearly every
    comparison ties and survivor paths merge slowly: the hand-over check of the parallel traceback has to repeat),
    misaligned rows, junk in the unused symbol bits; one launch per call."""
    env["CED_WARP_FRAME_RADIX"] = str(radix)
    env.pop("CED_WARP_FRAME_MAX", None)
    rng = np.random.default_rng(100 + radix)
    for frames, bits, p, pad, off in ((1, 8, 0.0, 0, 0), (1, 2048, 0.03, 0, 0), (16, 2048, 0.0377, 0, 0), (16, 2048, 0.5, 3, 5),
                                      (33, 4096, 0.06, 16, 0), (100, 200, 0.5, 1, 13), (5, 16384, 0.03, 0, 7),
                                      (7, 16384, 0.5, 5, 1), (590, 96, 0.1, 0, 0), (64, 1000 // 8 * 8, 0.2, 2, 3),
                                      (3, 24, 0.5, 0, 9), (40, 4096, 0.5, 0, 0)):
        msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
        rx = noisy(rng, port.encode_batch(7, K7, msgs), 2, p, junk=True)
        want = port.decode_batch(7, K7, rx & 3, bits + 6)
        before = ctx.launches
        got = ctx.decode_batch(ced.K7_DEFAULT, place(rx, pad, off), bits)
        ctx.sync()
        # the radix-2 form keeps 4 bytes of table offset per step: frames of 16384 bits do not fit its shared memory
        assert ctx.launches - before == (2 if radix == 2 and bits > 8192 else 1), "not the warp-per-frame kernel"
        assert np.array_equal(got.cpu().numpy(), want), (frames, bits, p, pad, off)
        if p == 0.0:
            assert np.array_equal(want, msgs)


@pytest.mark.parametrize("K,g", [(3, (0b111, 0b110)), (3, (0b111, 0b101, 0b011)), (4, (0o15, 0o17)), (5, (0o23, 0o35)),
                                 (5, (0o25, 0o33, 0o37)), (6, (0o53, 0o75)), (7, (0o133, 0o171)), (7, (0o133, 0o170)),
                                 (7, (0o133, 0o145, 0o174)), (7, (0o171, 0o133)), (7, (0o1, 0o100))])
def test_small_batches_of_other_codes(ctx, port, env, K, g):
    """Any k = 1 code with <= 64 states and 2 or 3 generators of any shape (general butterflies,
    src/viterbiDecoder.c:95-128), both step forms where there are two."""
    rng = np.random.default_rng(K * 131 + sum(g))
    n = len(g)
    code = ced.Code(K, g)
    env["CED_WARP_FRAME_ANY_K"] = "1"   # by default only K = 6 and 7 come here (fewer states leave most lanes idle)
    for radix in ((4, 2) if K == 7 and n == 2 else (2,)):
        env["CED_WARP_FRAME_RADIX"] = str(radix)
        for frames, bits, p in ((1, 16, 0.0), (20, 2048, 0.03), (70, 512, 0.5), (9, 4096, 0.1)):  # noqa
            msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
            rx = noisy(rng, port.encode_batch(K, list(g), msgs), n, p)
            want = port.decode_batch(K, list(g), rx, bits + K - 1, symmetric=False)
            before = ctx.launches
            got = ctx.decode_batch(code, place(rx, int(rng.integers(0, 9)), int(rng.integers(0, 16))), bits)
            ctx.sync()
            assert ctx.launches - before == 1
            assert np.array_equal(got.cpu().numpy(), want), (K, g, radix, frames, bits, p)


def test_small_and_large_batches_agree(ctx, env):
    """The same frames through the warp-per-frame kernel and through the thread-per-frame kernels (CED_WARP_FRAME_MAX=0)."""
    import torch
    rng = np.random.default_rng(5)
    frames, bits = 500, 4096
    msgs = torch.from_numpy(rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)).cuda()
    segs = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
    ctx.bsc_channel(segs, bits + 6, 2, 0.06, seed=3)
    env["CED_WARP_FRAME_MAX"] = "0"
    l0 = ctx.launches
    big = ctx.decode_batch(ced.K7_DEFAULT, segs, bits).clone()
    ctx.sync()
    assert ctx.launches - l0 == 2
    env.pop("CED_WARP_FRAME_MAX")
    l0 = ctx.launches
    small = ctx.decode_batch(ced.K7_DEFAULT, segs, bits)
    ctx.sync()
    assert ctx.launches - l0 == 1
    assert torch.equal(big, small)


def test_small_host_batches_take_the_direct_route(ctx, port, env):
    """ced_decode_batch_host with a small batch (speedDecode's 16 packets of 2048 bits, speedDecode.c:18-19): one copy in,
    one launch, one copy out -- pageable and page-locked buffers, padded rows."""
    import torch
    rng = np.random.default_rng(9)
    for frames, bits, pad, pinned in ((16, 2048, 0, False), (16, 2048, 10, True), (1, 4096, 0, False), (300, 512, 3, False)):
        msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
        rx = noisy(rng, port.encode_batch(7, K7, msgs), 2, 0.04)
        T = bits + 6
        h_in = np.full((frames, T + pad), 0xEE, dtype=np.uint8)
        h_out = np.zeros((frames, bits // 8), dtype=np.uint8)
        if pinned:
            h_in = torch.from_numpy(h_in).pin_memory().numpy()
            h_out = torch.from_numpy(h_out).pin_memory().numpy()
        h_in[:, :T] = rx
        before = ctx.launches
        ctx.decode_batch_host(ced.K7_DEFAULT, h_in, bits, h_out)
        assert ctx.launches - before == 1
        assert np.array_equal(h_out, port.decode_batch(7, K7, rx, T)), (frames, bits, pad, pinned)


def test_which_batches_take_the_warp_per_frame_kernel(ctx, env):
    """Default selection: K = 6 and 7 only, a few rounds of resident CTAs at most; CED_WARP_FRAME_MAX overrides."""
    import torch
    for k in ("CED_WARP_FRAME_MAX", "CED_WARP_FRAME_RADIX", "CED_WARP_FRAME_ANY_K", "CED_WARP_SPLIT"):
        env.pop(k, None)

    def launches(code, frames, bits):
        segs = torch.zeros((frames, bits + 16), dtype=torch.uint8, device="cuda")
        before = ctx.launches
        ctx.decode_batch(code, segs, bits)
        ctx.sync()
        return ctx.launches - before

    assert launches(ced.K7_DEFAULT, 16, 2048) == 2            # cut in time as well: block kernel + join kernel
    assert launches(ced.K7_DEFAULT, 200, 4096) == 2
    assert launches(ced.K7_DEFAULT, 400, 2048) == 1           # one warp per frame
    assert launches(ced.K7_DEFAULT, 2048, 2048) == 1
    assert launches(ced.K7_DEFAULT, 8192, 2048) == 2          # forward + traceback of the thread-per-frame path
    assert launches(ced.Code(7, (0o133, 0o145, 0o175)), 64, 1024) == 1
    assert launches(ced.Code(6, (0o53, 0o75)), 64, 1024) == 1
    assert launches(ced.Code(5, (0o23, 0o35)), 64, 1024) == 2
    assert launches(ced.Code(3, (7, 6)), 64, 1024) == 2
    env["CED_WARP_FRAME_MAX"] = "0"
    assert launches(ced.K7_DEFAULT, 16, 2048) == 2


@pytest.mark.parametrize("warmup,length", [(None, None), (8, 64), (96, 8), (24, 200), (4096, 64)])
def test_frames_cut_in_time(ctx, port, env, warmup, length):
    """warp_split.cu: every (frame, block) on its own warp from guessed metrics, hand-overs checked, blocks whose guess
    was wrong run again.  Short warm-ups and pure-noise frames make wrong guesses the rule, a warm-up longer than the
    frame makes every block start from step 0; the result is the oracle's bit for bit either way."""
    env.pop("CED_WARP_SPLIT")
    if warmup:
        env["CED_WARP_SPLIT_WARMUP"], env["CED_WARP_SPLIT_LEN"] = str(warmup), str(length)
    rng = np.random.default_rng(77 + (warmup or 0))
    for g in (K7, [0o133, 0o171], [0o133, 0o170]):
        code = ced.Code(7, g)
        for frames, bits, p, pad, off in ((1, 2048, 0.0, 0, 0), (1, 2048, 0.04, 0, 3), (16, 2048, 0.0377, 0, 0), (16, 2048, 0.5, 3, 5),
                                          (3, 8192, 0.08, 1, 9), (40, 512, 0.5, 0, 0), (7, 128, 0.1, 2, 1), (2, 16384, 0.5, 0, 0),
                                          (100, 4096, 0.03, 16, 0)):
            msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
            rx = noisy(rng, port.encode_batch(7, g, msgs), 2, p, junk=True)
            want = port.decode_batch(7, g, rx & 3, bits + 6, symmetric=False)
            before = ctx.launches
            got = ctx.decode_batch(code, place(rx, pad, off), bits)
            ctx.sync()
            if bits + 6 > (length or 64):          # a frame of one block is not cut
                assert ctx.launches - before == 2, "not the block + join kernels"
            assert np.array_equal(got.cpu().numpy(), want), (g, frames, bits, p, pad, off)
            if p == 0.0:
                assert np.array_equal(want, msgs)


@pytest.mark.parametrize("split", ["0", "1"])
def test_small_packed_batches(ctx, port, env, split):
    """ced_decode_batch_packed (four 2-bit symbols per byte) with few frames: the same small-batch kernels, a nibble of
    the packed row being a pair of symbols; other code families are refused as at any batch size."""
    import torch
    env["CED_WARP_SPLIT"] = split
    rng = np.random.default_rng(40 + int(split))
    for g in (K7, [0o133, 0o171], [0o117, 0o155]):
        code = ced.Code(7, g)
        for frames, bits, p in ((1, 8, 0.0), (1, 2048, 0.04), (16, 2048, 0.0377), (50, 1000 // 8 * 8, 0.5), (3, 16384, 0.06),
                                (290, 512, 0.1)):
            T = bits + 6
            msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
            rx = noisy(rng, port.encode_batch(7, g, msgs), 2, p)
            want = port.decode_batch(7, g, rx, T, symmetric=False)
            packed = ctx.pack_symbols(torch.from_numpy(rx).cuda(), T, packed_stride=(T + 3) // 4 + int(rng.integers(0, 7)))
            before = ctx.launches
            got = ctx.decode_batch_packed(code, packed, bits)
            ctx.sync()
            assert ctx.launches - before == (2 if split == "1" and T > 64 else 1)
            assert np.array_equal(got.cpu().numpy(), want), (g, frames, bits, p, split)
    with pytest.raises(ced.CedError):
        ctx.decode_batch_packed(ced.Code(7, [0o133, 0o170]), torch.zeros((4, 520), dtype=torch.uint8, device="cuda"), 2048)
