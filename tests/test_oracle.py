"""Pins the oracle (oracle/ced_oracle.c) against every golden vector the
reference's own tests hold for the path, against fixtures generated
from the unmodified reference (tests/golden/make_golden.py), and -- when it was
built here -- against oracle/_ref itself."""
import numpy as np
import pytest

import oracle
from conftest import bsc

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
