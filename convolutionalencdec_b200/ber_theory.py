"""Expected bit-error rates of a rate-1/n convolutional code under Viterbi decoding (pure host-side math).

Python form of what the reference's MATLAB scripts ask the Communications Toolbox for
(scripts/matlab/berCurveCoded.m:15,46-53,107-126: poly2trellis -> distspec -> bercoding / berawgn), so the
BER sweep (BASELINE config 4) can be compared with theory without MATLAB:

  distance_spectrum(K, gens, terms)   ~ distspec(poly2trellis(K, gens), terms)
  hard_decision_ber(p, spectrum)      ~ bercoding(EbN0, 'conv', 'hard', rate, spectrum, 'psk', 2, 'nondiff')
  soft_decision_ber(ebn0_db, ...)     ~ bercoding(EbN0, 'conv', 'soft', ...)
  bpsk_ber(ebn0_db)                   ~ berawgn(EbN0, 'psk', 2, 'nondiff')

The numbers the reference's author obtained from those calls are quoted in berTestK7/berTestK7.c:88-94 and pin
this module (tests/test_ber_theory.py).
"""
import math
from collections import namedtuple

Spectrum = namedtuple("Spectrum", "dfree weight event")  # weight[i], event[i] belong to distance dfree + i


def q_function(x):
    return 0.5 * math.erfc(x / math.sqrt(2.0))


def bpsk_ber(ebn0_db):
    """Uncoded coherent BPSK over AWGN: Q(sqrt(2 Eb/N0))."""
    return q_function(math.sqrt(2.0 * 10.0 ** (ebn0_db / 10.0)))


def coded_channel_ber(ebn0_db, rate=0.5):
    """Crossover probability the decoder sees: hard-sliced BPSK at Es/N0 = rate * Eb/N0."""
    return q_function(math.sqrt(2.0 * rate * 10.0 ** (ebn0_db / 10.0)))


def _edge(K, gens, state, bit):
    """state = the last K-1 input bits, newest in bit 0 (src/viterbiDecoder.c:44-46); generators are octal
    numbers whose MSb taps the newest bit (src/convEncode.c:13-17 reverses them onto the shift register)."""
    window = (state << 1) | bit  # K bits, newest in bit 0
    w = 0
    for g in gens:
        taps = int(format(g, "0%db" % K)[::-1], 2)  # bit i of taps = tap on the input i steps ago
        w += bin(window & taps).count("1") & 1
    return ((state << 1) | bit) & ((1 << (K - 1)) - 1), w


def distance_spectrum(K, gens, terms=10, max_steps=10000):
    """First `terms` entries of the distance spectrum: for d = dfree .. dfree+terms-1 the number of error events
    (paths that leave the all-zero state and first return to it) of output weight d, and the total number of
    information-bit errors on them."""
    n_states = 1 << (K - 1)
    table = [[_edge(K, gens, s, b) for b in (0, 1)] for s in range(n_states)]
    cap = None  # largest distance of interest, known once dfree is
    events, weights = {}, {}
    s1, w1 = table[0][1]
    live = {(s1, w1): (1, 1)}  # (state, output weight) -> (paths, sum of input weights)
    for _ in range(max_steps):
        if not live:
            break
        nxt = {}
        for (s, w), (cnt, bits) in live.items():
            for b in (0, 1):
                s2, dw = table[s][b]
                w2 = w + dw
                if cap is not None and w2 > cap:
                    continue
                if s2 == 0:
                    events[w2] = events.get(w2, 0) + cnt
                    weights[w2] = weights.get(w2, 0) + bits + b * cnt
                    if cap is None or min(events) + terms - 1 < cap:
                        cap = min(events) + terms - 1
                    continue
                c0, b0 = nxt.get((s2, w2), (0, 0))
                nxt[(s2, w2)] = (c0 + cnt, b0 + bits + b * cnt)
        if cap is None and len(nxt) > 64 * n_states * K:
            raise ValueError("no finite free distance found (catastrophic code?)")
        live = {k: v for k, v in nxt.items() if cap is None or k[1] <= cap}
    else:
        raise ValueError("distance spectrum did not terminate (catastrophic code?)")
    dfree = min(events)
    ds = range(dfree, dfree + terms)
    return Spectrum(dfree, [weights.get(d, 0) for d in ds], [events.get(d, 0) for d in ds])


def pairwise_error_hard(d, p):
    """Probability that a path at Hamming distance d beats the sent one on a BSC(p); ties split evenly."""
    q = 1.0 - p
    if d % 2:
        return sum(math.comb(d, k) * p ** k * q ** (d - k) for k in range((d + 1) // 2, d + 1))
    return (0.5 * math.comb(d, d // 2) * (p * q) ** (d // 2)
            + sum(math.comb(d, k) * p ** k * q ** (d - k) for k in range(d // 2 + 1, d + 1)))


def hard_decision_ber(p, spectrum):
    """Union bound on the decoded BER for hard decisions: sum_d c_d P_d (k = 1 input bit per step)."""
    return sum(c * pairwise_error_hard(spectrum.dfree + i, p) for i, c in enumerate(spectrum.weight) if c)


def soft_decision_ber(ebn0_db, spectrum, rate=0.5):
    """Union bound for unquantised soft decisions: sum_d c_d Q(sqrt(2 d R Eb/N0))."""
    g = rate * 10.0 ** (ebn0_db / 10.0)
    return sum(c * q_function(math.sqrt(2.0 * (spectrum.dfree + i) * g))
               for i, c in enumerate(spectrum.weight) if c)
