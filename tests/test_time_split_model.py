"""CPU model of csrc/warp_split.cu (no GPU): a frame cut into blocks that start from GUESSED path metrics after a warm-up,
hand-overs compared as metric vectors minus their minimum, blocks whose guess was wrong run again from the true vector.
The claim the kernels rest on -- the decisions are the sequential decoder's, whatever the guesses were worth -- is checked
here step by step against a plain sequential add-compare-select, and the walked-back bytes against the oracle."""
import numpy as np
import pytest

K7 = [0o113, 0o171]


def edge_labels(K, g):
    """edge[b][st]: coded segment of the branch that leaves state st with input bit b (src/viterbiDecoder.c:32-50)."""
    n, N = len(g), 1 << (K - 1)
    taps = [int(format(x, "0%db" % K)[::-1], 2) for x in g]
    e = np.zeros((2, N), dtype=np.int64)
    for b in range(2):
        for st in range(N):
            reg = ((st << 1) | b) & ((1 << K) - 1)
            for i in range(n):
                e[b, st] |= (bin(reg & taps[i]).count("1") & 1) << i
    return e


def acs_run(m, rx, edge, n):
    """steps over rx from metrics m: returns the metrics after them and the decisions [step][state] (1 = upper predecessor);
    the lower predecessor wins a tie (src/viterbiDecoderButterflyk1.c:129-130)."""
    N, H = m.size, m.size // 2
    j = np.arange(H)
    dec = np.zeros((rx.size, N), dtype=np.uint8)
    hd = lambda lab, r: ((lab ^ r) & 1) + (((lab ^ r) >> 1) & 1) + (((lab ^ r) >> 2) & 1)   # n <= 3 coded bits
    for t, r in enumerate(rx):
        r = int(r) & ((1 << n) - 1)
        new = np.empty_like(m)
        for b in range(2):
            a0 = m[j] + hd(edge[b, j], r)
            a1 = m[j + H] + hd(edge[b, j + H], r)
            d = a0 > a1
            new[2 * j + b] = np.where(d, a1, a0)
            dec[t, 2 * j + b] = d
        m = new
    return m, dec


def walk_back(dec, S):
    """state 0 at the last step, S unrecorded tail steps, bits MSb first (:200-256)."""
    T = dec.shape[0]
    s, bits = 0, np.zeros(T - S, dtype=np.uint8)
    for t in range(T - 1, -1, -1):
        if t < T - S:
            bits[t] = s & 1
        s = (s >> 1) | (int(dec[t, s]) << (S - 1))
    return np.packbits(bits)


@pytest.mark.parametrize("D,length", [(96, 64), (8, 16), (24, 40), (4096, 64)])
def test_blocks_from_guessed_metrics_reproduce_the_sequential_decisions(port, D, length):
    rng = np.random.default_rng(D + length)
    edge = edge_labels(7, K7)
    reruns = {}
    for bits, p in ((256, 0.0), (512, 0.04), (512, 0.12), (384, 0.5)):
        T = bits + 6
        msg = rng.integers(0, 256, (1, bits // 8), dtype=np.uint8)
        rx = port.encode_batch(7, K7, msg)[0, :T].copy()
        flips = rng.random((T, 2)) < p
        rx ^= flips[:, 0].astype(np.uint8) | (flips[:, 1].astype(np.uint8) << 1)
        start = np.full(64, 65, dtype=np.int64)
        start[0] = 0
        _, want = acs_run(start, rx, edge, 2)
        # every block on its own, from all-equal metrics D steps earlier (from the true start where that reaches step 0)
        B = (T + length - 1) // length
        dec = np.zeros_like(want)
        v_start, v_end = [None] * B, [None] * B
        for c in range(B):
            s, e, lo = c * length, min(T, (c + 1) * length), max(0, c * length - D)
            m = start.copy() if lo == 0 else np.zeros(64, dtype=np.int64)
            m, _ = acs_run(m, rx[lo:s], edge, 2)
            v_start[c] = m - m.min()
            m, dec[s:e] = acs_run(m, rx[s:e], edge, 2)
            v_end[c] = m - m.min()
        # the join: compare hand-overs, run wrong guesses again from the true vector
        n_rerun = 0
        for c in range(1, B):
            if c * length - D > 0 and not np.array_equal(v_start[c], v_end[c - 1]):
                s, e = c * length, min(T, (c + 1) * length)
                m, dec[s:e] = acs_run(v_end[c - 1].copy(), rx[s:e], edge, 2)
                v_end[c] = m - m.min()
                n_rerun += 1
        reruns[p] = n_rerun
        assert np.array_equal(dec, want), (D, length, bits, p)
        assert np.array_equal(walk_back(dec, 6), port.decode_batch(7, K7, rx[None, :], T)[0])
        if p == 0.0:
            assert np.array_equal(walk_back(dec, 6), msg[0])
    if D == 96:
        assert reruns[0.0] == 0 and reruns[0.04] == 0          # useful noise levels: every guess holds after 96 steps
    if D == 8:
        assert reruns[0.5] > 0                                   # pure noise, short warm-up: guesses fail and are repaired
    if D == 4096:
        assert sum(reruns.values()) == 0                         # every block reaches back to step 0: nothing to guess


@pytest.mark.parametrize("p", [0.0, 0.05, 0.5])
def test_segmented_traceback_with_handover_check_equals_the_sequential_walk(port, p):
    """warp_frame.cuh wfTraceback: 32 lanes, lane i walks segment i after a warm-up of one segment from state 0 above it;
    a lane that entered its segment in another state than the lane above left in walks again, until every hand-over
    agrees.  The top lane starts in state 0 at the last step, so the loop ends and the states -- hence the bits -- are the
    sequential walk's."""
    rng = np.random.default_rng(int(p * 100) + 3)
    edge = edge_labels(7, K7)
    for bits in (8, 64, 1000 // 8 * 8, 2048):
        T, S = bits + 6, 6
        msg = rng.integers(0, 256, (1, bits // 8), dtype=np.uint8)
        rx = port.encode_batch(7, K7, msg)[0, :T].copy()
        flips = rng.random((T, 2)) < p
        rx ^= flips[:, 0].astype(np.uint8) | (flips[:, 1].astype(np.uint8) << 1)
        start = np.full(64, 65, dtype=np.int64)
        start[0] = 0
        _, dec = acs_run(start, rx, edge, 2)

        def walk(s, hi, lo):                     # states entered while walking steps hi-1 .. lo; returns (state after, bits)
            out = {}
            for t in range(hi - 1, lo - 1, -1):
                out[t] = s & 1
                s = (s >> 1) | (int(dec[t, s]) << (S - 1))
            return s, out

        seg = ((T + 31) // 32 + 7) // 8 * 8
        top = (T - 1) // seg
        s_in, s_leave, got = [0] * 32, [0] * 32, {}
        for lane in range(top + 1):
            lo, hi = lane * seg, min(T, lane * seg + seg)
            if lane < top:
                s_in[lane], _ = walk(0, min(T, hi + seg), hi)
            s_leave[lane], bits_l = walk(s_in[lane], hi, lo)
            got.update(bits_l)
        rounds = 0
        while True:
            redo = [lane for lane in range(top) if s_leave[lane + 1] != s_in[lane]]
            if not redo:
                break
            above = list(s_leave)
            for lane in redo:
                s_in[lane] = above[lane + 1]
                s_leave[lane], bits_l = walk(s_in[lane], min(T, lane * seg + seg), lane * seg)
                got.update(bits_l)
            rounds += 1
            assert rounds <= 32
        want_state, want = walk(0, T, 0)
        assert s_leave[0] == want_state and all(got[t] == want[t] for t in range(T)), (bits, p, rounds)
