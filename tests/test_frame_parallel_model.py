"""The block decomposition behind the one-packet latency path (csrc/frame_parallel.cuh), as a numpy
model (tests/frame_parallel_model.py), against the oracle: same bytes for clean, noisy, all-zero and
pure-noise packets (ties everywhere) and for lengths that leave a short last block."""
import numpy as np
import pytest

import oracle
import frame_parallel_model as fp


@pytest.mark.parametrize("T", [13, 14, 33, 38, 65, 70, 129, 133, 134, 262, 1030])
def test_block_decomposition_equals_the_sequential_decoder(T):
    P = oracle.port()
    rng = np.random.default_rng(T)
    for rep, p in enumerate((0.0, 0.04, 0.12, 0.5, None, "zero")):
        L = T - 6
        msg = rng.integers(0, 256, (1, (L + 7) // 8), dtype=np.uint8)
        segs = P.encode_batch(7, oracle.K7_G, msg)[:, :T].copy()
        if p is None:
            segs[:] = rng.integers(0, 4, segs.shape)
        elif p == "zero":
            segs[:] = 0
        else:
            flips = rng.random((1, T, 2)) < p
            segs ^= flips[:, :, 0].astype(np.uint8) | (flips[:, :, 1].astype(np.uint8) << 1)
        want = P.decode_batch(7, oracle.K7_G, segs, T)[0]
        assert np.array_equal(fp.decode(segs[0], T), want[:(L - 1) // 8 + 1]), (T, rep)


def test_other_starting_metrics():
    """v_0 is whatever the state struct holds (reset values in the drivers); the model takes any."""
    P = oracle.port()
    rng = np.random.default_rng(3)
    segs = rng.integers(0, 4, (1, 102), dtype=np.uint8)
    want = P.decode_batch(7, oracle.K7_G, segs, 102)[0]
    assert np.array_equal(fp.decode(segs[0], 102, init_metrics=[0] + [65] * 63), want)


@pytest.mark.parametrize("block", [8, 32, 64, 128, 256])
def test_any_block_length_gives_the_same_bytes(block):
    """The decomposition does not depend on where the packet is cut (the kernel uses 128-step blocks)."""
    P = oracle.port()
    rng = np.random.default_rng(block)
    for T in (block + 3, 3 * block + 7, 518):
        segs = rng.integers(0, 4, (1, T), dtype=np.uint8)      # pure noise: ties everywhere
        want = P.decode_batch(7, oracle.K7_G, segs, T)[0][:(T - 7) // 8 + 1]
        assert np.array_equal(fp.decode(segs[0], T, block=block), want), (block, T)
