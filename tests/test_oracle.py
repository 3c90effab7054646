"""Pins the oracle (oracle/ced_oracle.c) against every golden vector the
reference's own tests hold for the path, against fixtures generated
from the unmodified reference (tests/golden/make_golden.py), and -- when it was
built here -- against oracle/_ref itself."""
import numpy as np
import pytest

import oracle
import os

from conftest import ROOT, bsc

K7, K3 = oracle.K7_G, oracle.K3_G


def test_handtraced_encoder_vector(port):
    # handTracedTest/handTraced.c:29,38
    segs, _ = port.encode(3, K3, np.array([0b01101000], dtype=np.uint8))
    assert segs.tolist() == [0b00, 0b11, 0b00, 0b10, 0b10, 0b11, 0b01, 0b00, 0b00, 0b00]


def test_handtraced_decode_and_metrics(port):
    # handTracedTest/handTraced.c:55,66,72-111
    corrupted = np.array([0b01, 0b11, 0b01, 0b10, 0b10, 0b11, 0b01, 0, 0, 0], dtype=np.uint8)
    d = port.decoder(3, K3, symmetric=False)
    assert d.step(corrupted, True).tolist() == [0b01101000]
    assert d.metrics().tolist() == [0, 5, 5, 5]
    expected = [[1, 1, 6, 5], [3, 1, 1, 3], [1, 3, 2, 2], [2, 2, 2, 4]]
    for i, want in enumerate(expected):
        d.step(corrupted[i:i + 1], False)
        assert d.metrics().tolist() == want


def test_k7_tables(port, golden):
    assert port.taps(7, K7) == [0x69, 0x4F] == golden["polys"].tolist()
    want = [0, 2, 2, 0, 3, 1, 1, 3, 0, 2, 2, 0, 3, 1, 1, 3, 1, 3, 3, 1, 2, 0, 0, 2, 1, 3, 3, 1, 2, 0, 0, 2]
    d = port.decoder(7, K7)
    assert d.edge_symm().tolist() == want == golden["edge_symm"].tolist()
    # symmetric and general butterflies label the same trellis
    e = d.edge()
    for j in range(32):
        assert e[0][j] == want[j] and e[1][j] == 3 - want[j]
        assert e[0][j + 32] == 3 - want[j] and e[1][j + 32] == want[j]


def test_k7_kat_vector(port, golden):
    segs, _ = port.encode(7, K7, golden["kat_msg"])
    want = [3, 2, 1, 1, 2, 1, 1, 0, 2, 2, 1, 2, 2, 3, 3, 0, 2, 3, 1, 3, 0, 1, 2, 2, 1, 3, 1, 2, 2, 3, 0, 3, 2, 2, 3,
            0, 1, 3]
    assert segs.tolist() == want == golden["kat_segs"].tolist()
    assert port.decode_batch(7, K7, segs[None, :], 38)[0].tolist() == [0xA5, 0x3C, 0xFF, 0x01]


@pytest.mark.parametrize("bits", [8, 64, 256, 2048, 4096])
def test_against_reference_fixtures(port, golden, bits):
    msgs = golden["msg_%d" % bits]
    assert np.array_equal(port.encode_batch(7, K7, msgs), golden["segs_%d" % bits])
    for tag in ("p0", "p02", "p06", "p50"):
        got = port.decode_batch(7, K7, golden["noisy_%d_%s" % (bits, tag)], bits + 6)
        assert np.array_equal(got, golden["dec_%d_%s" % (bits, tag)]), tag
        # the general-butterfly formulation makes identical decisions
        got2 = port.decode_batch(7, K7, golden["noisy_%d_%s" % (bits, tag)], bits + 6, symmetric=False)
        assert np.array_equal(got2, got)
    assert np.array_equal(golden["dec_%d_p0" % bits], msgs)


def test_streaming_matches_reference_metrics(port, golden):
    noise, want_dec, want_metrics = golden["stream_noise"], golden["stream_dec"], golden["stream_metrics"]
    d = port.decoder(7, K7)
    for c in range(want_metrics.shape[0]):
        assert d.step(noise[64 * c:64 * c + 64], False).size == 0
        assert np.array_equal(d.metrics(), want_metrics[c]), c
    assert np.array_equal(d.step(noise[:0], True), want_dec)
    # one-shot == chunked
    assert np.array_equal(port.decode_batch(7, K7, noise[None, :], 2054)[0], want_dec)


def test_bertest_golden_counts(port, golden):
    # berTestK7/berTestK7.c with srand(9865): integers printed by the reference binary
    want = [[2296339, 41080000, 92418, 20480000], [1525431, 41080000, 9655, 20480000],
            [928843, 41080000, 655, 20480000]]
    assert golden["ber_counts_full"].tolist() == want
    port.lib.orc_srand(9865)
    for row, p in zip(want, (5.585640e-02, 3.716174e-02, 2.262231e-02)):
        assert port.bertest(7, K7, 10000, 256, p)[0].tolist() == row


def test_metric_bound_holds(port):
    # with the reference's 121-step renorm metrics stay <= 135 (no uint8 wrap)
    rng = np.random.default_rng(5)
    d = port.decoder(7, K7)
    worst = 0
    for chunk in range(60):
        d.step(rng.integers(0, 4, 61, dtype=np.uint8), False)
        worst = max(worst, int(d.metrics().max()))
    assert worst <= 135


def test_against_live_reference(port, ref):
    rng = np.random.default_rng(11)
    assert ref.polys().tolist() == [0x69, 0x4F]
    for bits in (8, 40, 1000 // 8 * 8, 4096):
        msgs = rng.integers(0, 256, (40, bits // 8), dtype=np.uint8)
        segs = port.encode_batch(7, K7, msgs)
        assert np.array_equal(segs, ref.encode_batch(msgs))
        for p in (0.0, 0.03, 0.08, 0.5):
            noisy = bsc(rng, segs, p, junk_upper_bits=True)
            assert np.array_equal(port.decode_batch(7, K7, noisy, bits + 6), ref.decode_batch(noisy, bits + 6))
    # chunked encode with carried shift register
    msg = rng.integers(0, 256, 100, dtype=np.uint8)
    reg, parts = 0, []
    for lo in range(0, 100, 7):
        s, reg = port.encode(7, K7, msg[lo:lo + 7], last=(lo + 7 >= 100), reg=reg)
        parts.append(s)
    assert np.array_equal(np.concatenate(parts), ref.encode(msg, chunk=7))


# ------------------------------------------------------------------ soft decisions
def _bpsk_soft(rng, segs, T, amp, sigma, lo=-127):
    """int8 soft symbols [frames, 2T] of byte-per-segment symbols: amp * (+-1 + sigma * N(0,1)), rounded, clamped."""
    bits = np.stack([segs[:, :T] & 1, (segs[:, :T] >> 1) & 1], axis=-1).reshape(segs.shape[0], 2 * T)
    y = amp * ((1.0 - 2.0 * bits) + sigma * rng.standard_normal(bits.shape))
    return np.clip(np.round(y), lo, 127).astype(np.int8)


def test_soft_oracle_reduces_to_the_reference_for_constant_magnitudes(port, ref):
    """The pin of orc_dec_step_soft: with every |s| equal, its costs are A * calcHammingDist
    (src/viterbiDecoder.c:260-285), so its output must be the UNMODIFIED reference decoder's on the
    sliced symbols -- pure-noise frames included, where nearly every comparison is a tie."""
    rng = np.random.default_rng(21)
    for bits in (8, 96, 2048):
        msgs = rng.integers(0, 256, (24, bits // 8), dtype=np.uint8)
        segs = port.encode_batch(7, K7, msgs)
        T = bits + 6
        for p in (0.0, 0.04, 0.5):
            noisy = bsc(rng, segs, p)
            want = ref.decode_batch(noisy, T)
            assert np.array_equal(want, port.decode_batch(7, K7, noisy, T))
            for A in (1, 7, 64, 127):
                hard = np.stack([noisy & 1, (noisy >> 1) & 1], axis=-1).reshape(24, 2 * T).astype(np.int16)
                soft = np.where(hard == 1, -A, A).astype(np.int8)
                assert np.array_equal(port.decode_soft_batch(7, K7, soft, T), want), (bits, p, A)


def test_soft_lane_arithmetic_matches_soft_oracle(port):
    """trellis_swar16.cuh (what k7SoftForwardKernel runs per thread: correlation-form branch words, 16-bit
    guard-bit compare, Lanes16 survivor layout and traceback) on the host against orc_dec_step_soft."""
    import ctypes as C
    import subprocess
    subprocess.run(["make", "-C", ROOT, "hostsim"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libswar_sim.so"))
    rng = np.random.default_rng(5)
    worst = 0
    for bits in (8, 48, 96, 512, 4096):
        T = bits + 6
        segs = port.encode_batch(7, K7, rng.integers(0, 256, (5, bits // 8), dtype=np.uint8))
        shape = (5, 2 * T)
        cases = [_bpsk_soft(rng, segs, T, 32, 0.8), _bpsk_soft(rng, segs, T, 90, 1.5, lo=-128),
                 rng.integers(-128, 128, shape).astype(np.int8),             # pure noise, full range
                 rng.choice([-128, 127], shape).astype(np.int8),             # extremes only
                 np.zeros(shape, dtype=np.int8),                             # all erasures: every compare ties
                 rng.integers(-2, 3, shape).astype(np.int8)]                 # tiny magnitudes, many ties
        for soft in cases:
            want = port.decode_soft_batch(7, K7, soft, T)
            for f in range(5):
                out, mx = np.zeros(bits // 8, dtype=np.uint8), C.c_uint32(0)
                row = np.ascontiguousarray(soft[f])
                lib.swar_sim_decode_soft(row.ctypes.data_as(C.c_void_p), T, out.ctypes.data_as(C.c_void_p), C.byref(mx))
                assert np.array_equal(out, want[f])
                worst = max(worst, mx.value)
    assert worst < 15872   # the bound the 16-bit guard-bit compare relies on (trellis_swar16.cuh)


# ------------------------------------------------------------------ windowed traceback
def test_window_definition_reproduces_matlab_tblen_expectations(port):
    """The statistical pin of orc_decode_window.  The reference's own windowed decoder aborts at HEAD, but
    berTestK7.c:98 holds MATLAB's vitdec(..., tblen = 5K = 35, 'term', 'hard') expectations (generators 133/171,
    scripts/matlab/viterbiBEREstimate.m:11,17,99).  orc_decode_window with one-step slices and depth 35 is that
    decoder; its BER on berTestK7's BSC points must meet them under the reference's own +-10 % rule (:167-172).
    The same definition with the depth beyond the packet length is the full traceback and meets the full-traceback
    expectations of :96-97 instead -- the two sets differ by 4-11 %, so the test tells the two decoders apart."""
    g = (0o133, 0o171)
    points = ((5.585640e-02, 5.295410e-03, 4.765898e-03, 4000), (3.716174e-02, 5.421997e-04, 5.184082e-04, 20000),
              (2.262231e-02, 3.385010e-05, 3.499023e-05, 100000))
    excess = []
    for p, want_tblen, want_full, pkts in points:
        c = port.window_ber(7, g, pkts, 256, p, 1, 35, seed=1)
        assert abs(c[0] / c[1] - p) < 0.01 * p
        ber = c[2] / c[3]
        assert abs(ber - want_tblen) / want_tblen < 0.10, (p, ber, want_tblen)
        full = port.window_ber(7, g, pkts // 2, 256, p, 4096, 4096, seed=1)
        assert abs(full[2] / full[3] - want_full) / want_full < 0.10, (p, full[2] / full[3], want_full)
        excess.append(ber / (full[2] / full[3]))
    assert excess[0] > 1.03 and excess[1] > 1.0     # truncating the traceback at 35 steps costs errors, as in MATLAB's numbers


def test_window_soft_definition_anchors(port):
    """orc_decode_window_soft (semantics of ced_decode_window_batch_softq): with depth >= stream length it is the full-frame
    soft decode, and with reliabilities of one constant magnitude it is orc_decode_window on the sliced symbols -- the
    same two anchors that tie the hard window definition to the pinned decoder."""
    rng = np.random.default_rng(5)
    T = 96 * 5 + 38
    msgs = rng.integers(0, 256, (6, (T - 6) // 8), dtype=np.uint8)
    clean = port.encode_batch(7, oracle.K7_G, msgs)
    soft = rng.integers(-7, 8, (6, 2 * T)).astype(np.int8)
    full = port.decode_soft_batch(7, oracle.K7_G, soft, T)
    for i in range(6):
        assert np.array_equal(port.decode_window_soft(7, oracle.K7_G, soft[i], 96, 10000), full[i])
    flips = rng.random(clean.shape + (2,)) < 0.08
    noisy = clean ^ (flips[..., 0].astype(np.uint8) | (flips[..., 1].astype(np.uint8) << 1))
    const = np.empty((6, 2 * T), dtype=np.int8)
    const[:, 0::2] = np.where(noisy & 1, -5, 5)
    const[:, 1::2] = np.where(noisy & 2, -5, 5)
    for call, depth in ((96, 24), (192, 48), (480, 96)):
        for i in range(6):
            assert np.array_equal(port.decode_window_soft(7, oracle.K7_G, const[i], call, depth),
                                  port.decode_window(7, oracle.K7_G, noisy[i], call, depth)), (call, depth, i)
