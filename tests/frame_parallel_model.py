"""TEST INFRASTRUCTURE -- numpy model of the frame-parallel single-packet decoder
(convolutionalencdec_b200/csrc/frame_parallel.cuh): the same block decomposition, keys and
tie-break, written with plain arrays so the theory can be checked against the oracle without a GPU.

Why it is exact.  The reference's add-compare-select keeps the path from the LOWER predecessor j on
equal metrics (strict `>`, src/viterbiDecoderButterflyk1.c:129-130).  The two predecessors j and j+32
of a state differ only in their oldest input bit, so the survivor of every (time, state) is the
minimum, over all paths into that state, of the pair (cost, U) with U = sum u_t 2^t (later input
bits more significant, 0 beats 1).  That order is compatible with cutting a packet into blocks:
    best path into e at the end of block c  =  min over start states s of
        ( v_c[s] + cost_c[s][e],  bits_c[s][e],  rev6(s) )
where cost_c[s][e] / bits_c[s][e] are the cost and the input bits of the best path from s to e inside
the block (found by a forward pass that starts with metric 0 in s only, same local tie rule), and
rev6(s) compares the six input bits that precede the block, newest first.
"""
import numpy as np

BLOCK = 128
INF = 1 << 20


def edge_symm(g=(0o113, 0o171), K=7):
    """edgeCodedBitsSymm[j] (src/viterbiDecoderButterflyk1.c:24-29): output of state j, input 0."""
    def rev(x):
        return int(format(x, "0%db" % K)[::-1], 2)
    polys = [rev(x) for x in g]
    out = np.zeros(1 << (K - 2), dtype=np.int64)
    for j in range(out.size):
        td = j << 1
        out[j] = sum((bin(td & p).count("1") & 1) << i for i, p in enumerate(polys))
    return out


def block_transfer(sym, rx, block=BLOCK):
    """cost[s][e], bits[w][s][e] (64-bit words, w = 0 lowest) of the best in-block path from s to e:
    64 single-start passes."""
    N = 64
    m = np.full((N, N), INF, dtype=np.int64)
    m[np.arange(N), np.arange(N)] = 0
    bits = np.zeros(((block + 63) // 64, N, N), dtype=np.uint64)
    for t, r in enumerate(rx):
        x = sym ^ (int(r) & 3)
        d = (x & 1) + (x >> 1)                      # calcHammingDist(.., n = 2)
        lo, hi = m[:, :32], m[:, 32:]
        blo, bhi = bits[:, :, :32], bits[:, :, 32:]
        a0, a1 = lo + d, hi + (2 - d)
        b0, b1 = lo + (2 - d), hi + d
        da, db = a0 > a1, b0 > b1
        nm = np.empty_like(m)
        nb = np.empty_like(bits)
        nm[:, 0::2] = np.where(da, a1, a0)
        nm[:, 1::2] = np.where(db, b1, b0)
        nb[:, :, 0::2] = np.where(da, bhi, blo)
        nb[:, :, 1::2] = np.where(db, bhi, blo)
        nb[t // 64, :, 1::2] |= np.uint64(1 << (t % 64))
        m, bits = nm, nb
    return m, bits


def rev6(s):
    return int(format(s, "06b")[::-1], 2)


def decode(segs, T, init_metrics=None, g=(0o113, 0o171), block=BLOCK):
    """Decoded bytes of one K=7 n=2 packet of T segments (L = T - 6 information bits)."""
    sym = edge_symm(g)
    v = np.full(64, 65, dtype=np.int64)
    v[0] = 0
    if init_metrics is not None:
        v = np.asarray(init_metrics, dtype=np.int64)
    nb = (T + block - 1) // block
    costs, bitss, vs = [], [], [v]
    for c in range(nb):
        cost, bits = block_transfer(sym, segs[c * block:min(T, (c + 1) * block)], block)
        costs.append(cost)
        bitss.append(bits)
        v = (v[:, None] + cost).min(axis=0)
        vs.append(v)
    e = 0
    u = np.zeros(nb * block, dtype=np.uint8)
    for c in range(nb - 1, -1, -1):
        def bits_of(s):
            return sum(int(bitss[c][w][s][e]) << (64 * w) for w in range((block + 63) // 64))
        keys = [(int(vs[c][s] + costs[c][s][e]), bits_of(s), rev6(s)) for s in range(64)]
        s = min(range(64), key=lambda i: keys[i])
        w = bits_of(s)
        for t in range(block):
            u[c * block + t] = (w >> t) & 1
        e = s
    L = T - 6
    u[L:] = 0
    return np.packbits(u[:((L - 1) // 8 + 1) * 8])
