"""The lane mapping of fpBlockKernel (csrc/frame_parallel.cuh), emulated with numpy arrays of 32 lanes:
rotating labels (period 5 over the lane bits), ONE xor-shuffle per register and step, the PRMT selectors
0x5410 / 0x3276 for (lower, upper) predecessor, packed u16x2 candidates, and the `a <= b` predicate of
VIMNMX.U16x2 as the tie rule.  One pass per start state must give exactly the costs and survivor bits of the
plain per-state recursion in tests/frame_parallel_model.py (which is pinned to the oracle)."""
import numpy as np
import pytest

import frame_parallel_model as fp

LANES = np.arange(32)
UNREACH = 0x1000


def role(b, r):                       # fpLaneBitRole: state bit held by lane bit b in phase r
    return ((b + r) % 5) + 1


def byte_perm(x, y, sel):             # __byte_perm on arrays of uint32
    src = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    out = np.zeros_like(x)
    for i in range(4):
        idx = (sel >> (4 * i)) & 7
        out |= np.choose(idx, src) << (8 * i)
    return out


def dist_table(edge):
    """sDist[phase][rx][lane] -> (d00 | d0h << 16, d10 | d1h << 16), as built at the top of fpBlockKernel."""
    tab = np.zeros((5, 4, 32, 2), dtype=np.int64)
    hd = lambda e, rx: bin((e ^ rx) & 3).count("1")
    for r in range(5):
        q = 4 - r
        for l in range(32):
            j = (l >> q) & 1
            for b in range(5):
                if b != q:
                    j |= ((l >> b) & 1) << role(b, r)
            for rx in range(4):
                tab[r, rx, l, 0] = hd(edge[0][j], rx) | hd(edge[0][j + 32], rx) << 16
                tab[r, rx, l, 1] = hd(edge[1][j], rx) | hd(edge[1][j + 32], rx) << 16
    return tab


def one_pass(tab, rx_seq, s):
    M = np.where(2 * LANES == s, 0, UNREACH) | (np.where(2 * LANES + 1 == s, 0, UNREACH) << 16)
    p0 = np.zeros(32, dtype=object)
    p1 = np.zeros(32, dtype=object)
    for t, rx in enumerate(rx_seq):
        r, q = t % 5, 4 - (t % 5)
        up = ((LANES >> q) & 1).astype(bool)
        partner = LANES ^ (1 << q)
        send = np.where(up, p0, p1)
        recvM, recvP = M[partner], send[partner]
        sel = np.where(up, 0x3276, 0x5410)
        LH = byte_perm(M, recvM, sel)
        pLo, pHi = np.where(up, recvP, p0), np.where(up, p1, recvP)
        d = tab[r, int(rx) & 3]
        A, B = LH + d[:, 0], LH + d[:, 1]
        a_lo = (A & 0xFFFF) <= (A >> 16)            # VIMNMX.U16x2 predicate of the low half: a0 <= a1
        b_lo = (B & 0xFFFF) <= (B >> 16)
        M = np.minimum(A & 0xFFFF, A >> 16) | (np.minimum(B & 0xFFFF, B >> 16) << 16)
        p0 = np.where(a_lo, pLo, pHi)
        p1 = np.where(b_lo, pLo, pHi) | (1 << t)
    r = len(rx_seq) % 5
    e = np.zeros(32, dtype=np.int64)
    for b in range(5):
        e |= ((LANES >> b) & 1) << role(b, r)
    cost, bits = {}, {}
    for l in range(32):
        cost[int(e[l])], bits[int(e[l])] = int(M[l] & 0xFFFF), int(p0[l])
        cost[int(e[l]) + 1], bits[int(e[l]) + 1] = int(M[l] >> 16), int(p1[l])
    return cost, bits


@pytest.mark.parametrize("steps", [6, 7, 31, 64, 128])
def test_lane_mapping_reproduces_the_per_state_recursion(steps):
    sym = fp.edge_symm()
    edge = [[0] * 64, [0] * 64]                 # edge[b][state] as labelEdges() builds it for symmetric generators
    for j in range(32):
        edge[0][j], edge[0][j + 32], edge[1][j], edge[1][j + 32] = sym[j], sym[j] ^ 3, sym[j] ^ 3, sym[j]
    tab = dist_table(edge)
    rng = np.random.default_rng(steps)
    rx = rng.integers(0, 4, steps)
    want_cost, want_bits = fp.block_transfer(sym, rx, block=max(steps, 64))
    for s in range(0, 64, 7):
        cost, bits = one_pass(tab, rx, s)
        for e in range(64):
            if want_cost[s][e] < fp.INF:
                w = sum(int(want_bits[k][s][e]) << (64 * k) for k in range(want_bits.shape[0]))
                assert cost[e] == want_cost[s][e] and bits[e] == w, (steps, s, e)
            else:
                assert cost[e] >= UNREACH
