"""The BER-theory helpers against the numbers the reference quotes from MATLAB (berTestK7/berTestK7.c:88-96) and
textbook distance spectra; on the GPU, the simulated BER of the default code against its union bound."""
import math

import pytest

from convolutionalencdec_b200 import ber_theory as bt


def test_distance_spectrum_textbook_codes():
    s = bt.distance_spectrum(7, [0o133, 0o171], 10)    # the code of scripts/matlab/berCurveCoded.m:11-15
    assert s.dfree == 10
    assert s.event == [11, 0, 38, 0, 193, 0, 1331, 0, 7275, 0]
    assert s.weight == [36, 0, 211, 0, 1404, 0, 11633, 0, 77433, 0]
    s = bt.distance_spectrum(3, [7, 5], 5)
    assert (s.dfree, s.event, s.weight) == (5, [1, 2, 4, 8, 16], [1, 4, 12, 32, 80])
    s = bt.distance_spectrum(9, [0o561, 0o753], 1)
    assert (s.dfree, s.event, s.weight) == (12, [11], [33])


def test_hard_decision_bound_matches_the_matlab_numbers_in_the_reference():
    # berTestK7.c:88-94: bercoding(...'hard'...) for generators 0133/0171 at these channel BERs,
    # with distspec's default single term and with 10 terms
    ps = [3.716174e-02, 2.262231e-02, 1.232962e-02]
    one = bt.distance_spectrum(7, [0o133, 0o171], 1)
    ten = bt.distance_spectrum(7, [0o133, 0o171], 10)
    for p, want1, want10 in zip(ps, [2.835189e-04, 2.490713e-05, 1.240189e-06],
                                [1.104553e-03, 5.016878e-05, 1.711085e-06]):
        assert bt.hard_decision_ber(p, one) == pytest.approx(want1, rel=2e-6)
        assert bt.hard_decision_ber(p, ten) == pytest.approx(want10, rel=2e-6)


def test_channel_ber_matches_bertest_points():
    # berTestK7.c:95-96: SNR -5/-4/-3 dB at 4 samples per symbol -> Es/N0 = SNR + 10 log10(4); BPSK
    for snr, want in zip([-5, -4, -3], [5.585640e-02, 3.716174e-02, 2.262231e-02]):
        esn0 = snr + 10.0 * math.log10(4.0)
        assert bt.bpsk_ber(esn0) == pytest.approx(want, rel=1e-6)
        assert bt.coded_channel_ber(esn0 + 10.0 * math.log10(2.0), rate=0.5) == pytest.approx(want, rel=1e-6)


def test_default_code_spectrum():
    # the library's default generators (src/defaultParams/convCodeParams.c:6) are 0113/0171, not MATLAB's 0133/0171
    s = bt.distance_spectrum(7, [0o113, 0o171], 4)
    assert s.dfree == 9 and s.event == [3, 5, 8, 20] and s.weight == [7, 22, 46, 114]
    assert bt.soft_decision_ber(6.0, s) < bt.hard_decision_ber(bt.coded_channel_ber(6.0), s)


@pytest.mark.gpu
def test_simulated_ber_approaches_the_union_bound():
    """At high Eb/N0 the union bound is tight: the GPU chain's decoded BER must sit just below it."""
    import torch
    import convolutionalencdec_b200 as ced
    ctx, k7 = ced.Context(0), ced.K7_DEFAULT
    frames, bits = 1 << 18, 2048
    T = bits + 6
    spectrum = bt.distance_spectrum(7, [0o113, 0o171], 20)
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    segs = torch.zeros((frames, (T + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=77)
    for ebn0, lo in ((6.0, 0.55), (7.0, 0.7)):
        p = bt.coded_channel_ber(ebn0)
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        ctx.encode_batch(k7, msgs, out=segs)
        ctx.bsc_channel(segs, T, 2, p, seed=1000 + int(ebn0), counters=counters[:2])
        dec = ctx.decode_batch(k7, segs, bits)
        ctx.ber_count(dec, msgs, counters[2:])
        ctx.sync()
        flips, coded, errs, total = (int(v) for v in counters.cpu())
        assert flips / coded == pytest.approx(p, rel=0.01)
        bound = bt.hard_decision_ber(p, spectrum)
        ber = errs / total
        sigma = math.sqrt(max(errs, 1)) / total * 3.0   # errors come in bursts: generous 3x
        assert lo * bound < ber < bound + 3.0 * sigma, (ebn0, ber, bound)
    ctx.close()
