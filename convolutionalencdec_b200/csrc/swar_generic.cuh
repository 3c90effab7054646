/*
 * swar_generic.cuh -- the SIMD-in-word forward pass of trellis_swar.cuh for ANY rate-1/n, k = 1 code with 2^S states,
 * S = K - 1 in 2 .. 8 and n = 2 or 3: generators that do NOT tap both ends included
 * (the reference's headers advertise generic K / n, src/convEncode.h:8-18, src/viterbiDecoder.h:47-62,
 * and its own handTracedTest code g = {7, 6} is one of them).  K = 7 codes whose generators tap the newest and the oldest
 * bit keep the hand-scheduled kernel (decode_batch.cuh); everything else used to run on a one-warp-per-frame kernel
 * at 3-14 Gbit/s.
 *
 * Same construction: one thread = one frame, metrics are bytes packed four to a register (2^S / 4 registers), the
 * state labelling rotates so that butterflies are in place (position p holds state rotl_S(p, ph) before a step of
 * phase ph = t mod S; partners differ in position bit q = S - 1 - ph: a register bit for q >= 2, a lane bit else).
 * What is different: nothing about the code is compiled in.  The host builds a step table once per code --
 *   register phases, per register pair and received symbol:  Xa0 Xa1 Xb0 Xb1 | Ea Eb      (32-byte entry)
 *       a0 = lo + Xa0, a1 = hi + Xa1 -> successor 2j      b0 = lo + Xb0, b1 = hi + Xb1 -> successor 2j+1
 *       Ea = Xa1 - Xa0 + 0x80.., Eb = Xb1 - Xb0 + 0x80..  (guard-bit compare, trellis_swar.cuh acsStep)
 *   lane phases, per register and received symbol:            Xself Xcross E               (16-byte entry)
 * -- every word holds the four lanes' own branch costs HD(rx, label of that lane's edge), so symmetric and general
 * butterflies (src/viterbiDecoder.c:95-128) are the same code.  Decisions and tie rule are the reference's (strict '>'
 * keeps the lower predecessor, src/viterbiDecoderButterflyk1.c:129-130).
 *
 * Exactness of 8-bit metrics: the start metrics are the reference's, 0 for state 0 and (uint8_t)(NUM_STATES + 1) for
 * the rest (src/viterbiDecoderButterflyk1.c:59-67), capped at n*S + 1 where that changes no decision (genInitMetrics);
 * for K = 9 the reference's uint8_t METRIC_TYPE turns 257 into 1, which this kernel reproduces.  Any state is reachable
 * from state 0 within S steps at cost <= n*S, so the largest value seen is <= 2 n*S + 1 in the first S steps, the
 * spread is <= n*S afterwards and the minimum grows by <= n per step: with a renormalisation every <= 24 steps
 * candidates stay <= n*S + 24*n + n <= 99 < 128 for every supported (S, n).  tests/hostsim asserts the largest metric.
 *
 * Survivors: W = max(1, 2^S / 32) decision words per step and frame, [group][step][lane] rows of W words (4, 8 or 32
 * bytes per lane): position p = (register r, lane l) -> word r >> 3, bit 8 l + (r & 7).
 */
#pragma once
#include "trellis_swar.cuh"

namespace ced {

struct GenCode {
    int S, n;            /* state bits (K - 1), coded bits per segment */
    uint32_t tap[3];     /* generators, bit 0 on the newest input bit (src/convEncode.c:163-175) */
};

CED_HD uint32_t rotlS(uint32_t x, int r, int S)
{
    const uint32_t m = (1u << S) - 1u;
    r %= S;
    return r == 0 ? (x & m) : (((x << r) | (x >> (S - r))) & m);
}

/* coded segment on the edge leaving state s with input bit b (src/viterbiDecoder.c:44-46, src/convEncode.c:132-161) */
CED_HD uint32_t genEdgeLabel(const GenCode &c, uint32_t s, uint32_t b)
{
    const uint32_t reg = ((s << 1) | b) & ((2u << c.S) - 1u);
    uint32_t v = 0;
    for (int i = 0; i < c.n; i++)
        v |= parity32(reg & c.tap[i]) << i;
    return v;
}

CED_HD uint32_t hdN(uint32_t a, uint32_t b, int n)
{
    uint32_t x = (a ^ b) & ((1u << n) - 1u), d = 0;
    for (; x; x >>= 1)
        d += x & 1u;
    return d;
}

template <int S>
struct GenGeom {
    static_assert(S >= 2 && S <= 8, "4 .. 256 states");
    static constexpr int kStates = 1 << S;
    static constexpr int kRegs = kStates >= 4 ? kStates / 4 : 1;
    static constexpr int kRegPhases = S - 2;                     /* phases whose pair bit is a register bit */
    static constexpr int kPairs = kRegs / 2;
    static constexpr int kWords = kRegs >= 8 ? kRegs / 8 : 1;    /* decision words per step */
    static constexpr int kRenorm = 24 / S * S;   /* whole label rotations: 24, or 20 / 21 for S = 5 / 7 */
    static constexpr int kChunk = 96 / S * S;    /* trellis steps per staged tile: 96, or 95 / 91 for S = 5 / 7 */
    static constexpr int kPhases = S;
    static constexpr int kTail = S;          /* tail segments */
    static constexpr int kStepBits = 1;      /* decoded bits per trellis step */
    /* byte offset of the received symbol's variant inside an entry row; `off` = rx * 32 as staged in the tile */
    CED_HD static constexpr uint32_t rxOffset(int ph, uint32_t off) { return ph < kRegPhases ? off : off >> 1; }
    CED_HD static constexpr int entryStride(int ph, int V) { return (ph < kRegPhases ? 32 : 16) * V; }
    /* bytes of step table for V received symbols: 32 per (register phase, pair, rx), 16 per (lane phase, register, rx) */
    CED_HD static constexpr int tableBytes(int V) { return (kRegPhases * kPairs * 32 + 2 * kRegs * 16) * V; }
    CED_HD static constexpr int phaseBase(int ph, int V)               /* byte offset of phase ph's entries */
    {
        return ph < kRegPhases ? ph * kPairs * 32 * V : (kRegPhases * kPairs * 32 + (ph - kRegPhases) * kRegs * 16) * V;
    }
};

struct GenPairEntry { uint32_t xa0, xa1, xb0, xb1, ea, eb, pad0, pad1; };   /* 32 bytes */
struct GenLaneEntry { uint32_t xself, xcross, e, pad; };                    /* 16 bytes */

/* host: fill `table` (GenGeom<S>::tableBytes(1 << n) bytes) for code c */
template <int S>
inline void buildGenTable(const GenCode &c, uint8_t *table)
{
    using G = GenGeom<S>;
    const int V = 1 << c.n;
    const uint32_t H = 1u << (S - 1);
    for (int ph = 0; ph < S; ph++) {
        const int q = S - 1 - ph;
        uint8_t *base = table + G::phaseBase(ph, V);
        if (q >= 2) {
            const int rb = q - 2;
            int idx = 0;
            for (int r = 0; r < G::kRegs; r++) {
                if ((r >> rb) & 1)
                    continue;
                for (int rx = 0; rx < V; rx++) {
                    GenPairEntry e = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (uint32_t l = 0; l < 4; l++) {
                        const uint32_t j = rotlS(4u * (uint32_t)r + l, ph, S);     /* lower state of this lane's butterfly */
                        e.xa0 |= hdN(rx, genEdgeLabel(c, j, 0), c.n) << (8 * l);
                        e.xa1 |= hdN(rx, genEdgeLabel(c, j + H, 0), c.n) << (8 * l);
                        e.xb0 |= hdN(rx, genEdgeLabel(c, j, 1), c.n) << (8 * l);
                        e.xb1 |= hdN(rx, genEdgeLabel(c, j + H, 1), c.n) << (8 * l);
                    }
                    e.ea = e.xa1 - e.xa0 + 0x80808080u;
                    e.eb = e.xb1 - e.xb0 + 0x80808080u;
                    *reinterpret_cast<GenPairEntry *>(base + ((size_t)idx * V + rx) * 32) = e;
                }
                idx++;
            }
        } else {
            for (int r = 0; r < G::kRegs; r++)
                for (int rx = 0; rx < V; rx++) {
                    GenLaneEntry e = {0, 0, 0, 0};
                    uint32_t guard = 0;
                    for (uint32_t l = 0; l < 4; l++) {
                        const uint32_t s = rotlS(4u * (uint32_t)r + l, ph, S);     /* this lane's state */
                        const bool upper = (l >> q) & 1u;                          /* holds j + H: its new value is 2j + 1 */
                        const uint32_t j = s & (H - 1u);
                        const uint32_t self = upper ? genEdgeLabel(c, j + H, 1) : genEdgeLabel(c, j, 0);
                        const uint32_t cross = upper ? genEdgeLabel(c, j, 1) : genEdgeLabel(c, j + H, 0);
                        e.xself |= hdN(rx, self, c.n) << (8 * l);
                        e.xcross |= hdN(rx, cross, c.n) << (8 * l);
                        guard |= (upper ? 0x7Fu : 0x80u) << (8 * l);               /* a tie goes to the lower predecessor */
                    }
                    e.e = e.xcross - e.xself + guard;
                    *reinterpret_cast<GenLaneEntry *>(base + ((size_t)r * V + rx) * 16) = e;
                }
        }
    }
}

template <int S>
CED_HD void genInitMetrics(uint32_t (&R)[GenGeom<S>::kRegs], int n)
{
    /* METRIC_TYPE forceNot = NUM_STATES + 1 (:59-60) as a uint8_t.  Where that exceeds n S + 1 it is a true "never wins" (a
     * path from another start state costs more than the path from state 0 with the same inputs, which merges with it after
     * S steps at <= n S) and n S + 1 gives the same decisions while leaving the guard bit free (K = 8: 129 would not);
     * where it does not (K = 3 with n = 3: 5; K = 9: 257 wraps to 1) the reference's value is kept as it is. */
    const uint32_t ref = (uint32_t)((GenGeom<S>::kStates + 1) & 0xFF), cap = (uint32_t)(n * S + 1);
    const uint32_t never = (ref < cap ? ref : cap) * 0x01010101u;
#pragma unroll
    for (int r = 0; r < GenGeom<S>::kRegs; r++)
        R[r] = never;
    R[0] &= 0xFFFFFF00u; /* state 0 sits at position 0 in every phase */
}

/*
 * One trellis step of compile-time phase PH.  `tab` = the entries of (PH, rx): entry i at tab + i * stride bytes
 * (stride = entry size * number of received symbols).  T[w] receives the decision word w.
 */
template <int S, int PH>
CED_HD void genStep(uint32_t (&R)[GenGeom<S>::kRegs], const uint8_t *tab, int stride, uint32_t minusOne,
                    uint32_t (&T)[GenGeom<S>::kWords])
{
    using G = GenGeom<S>;
    constexpr int q = S - 1 - PH;
#pragma unroll
    for (int w = 0; w < G::kWords; w++)
        T[w] = 0;
    /* Table entries are fetched in batches of eight before they are used.  (Written this way after ncu showed
     * short_scoreboard as the top stall, profiles/r2_gen_*_ncu.txt; ptxas schedules the loads of the one-pair-at-a-time
     * form the same way -- 83.6 vs 83.8 Gbit/s for non-symmetric K=7 -- so the stall is the latency of the first
     * entry of a step with three warps per sub-partition, not the order of the loads.) */
    if constexpr (q >= 2) {
        constexpr int rb = q - 2;
        constexpr int kBatch = G::kPairs < 8 ? G::kPairs : 8;
        static_assert(G::kPairs % kBatch == 0, "whole batches");
#pragma unroll
        for (int b0 = 0; b0 < G::kPairs; b0 += kBatch) {
            uint4 x[kBatch];
            uint2 e[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; i++) {
                x[i] = *reinterpret_cast<const uint4 *>(tab + (size_t)(b0 + i) * stride);
                e[i] = *reinterpret_cast<const uint2 *>(tab + (size_t)(b0 + i) * stride + 16);
            }
#pragma unroll
            for (int i = 0; i < kBatch; i++) {
                /* pair number idx -> its lower register: the idx-th register index with bit rb clear */
                constexpr int low = (1 << rb) - 1;
                const int idx = b0 + i;
                const int r = ((idx & ~low) << 1) | (idx & low), rh = r | (1 << rb);
                const uint32_t lo = R[r], hi = R[rh];
                const uint32_t a0 = lo + x[i].x, a1 = hi + x[i].y;
                const uint32_t b0v = lo + x[i].z, b1 = hi + x[i].w;
                const uint32_t delta = subOnFma(hi, lo, minusOne);
                const uint32_t ma = signMask(delta + e[i].x);      /* FF: keep the lower predecessor */
                const uint32_t mb = signMask(delta + e[i].y);
                R[r] = sel(ma, a0, a1);
                R[rh] = sel(mb, b0v, b1);
                T[r >> 3] |= ~ma & (0x01010101u << (r & 7));
                T[rh >> 3] |= ~mb & (0x01010101u << (rh & 7));
            }
        }
    } else {
        constexpr uint32_t swapSel = (q == 1) ? 0x1032u : 0x2301u;
        constexpr uint32_t upper = (q == 1) ? 0xFFFF0000u : 0xFF00FF00u;
        constexpr int kBatch = G::kRegs < 8 ? G::kRegs : 8;
#pragma unroll
        for (int b0 = 0; b0 < G::kRegs; b0 += kBatch) {
            uint4 e[kBatch];   /* xself xcross e pad */
#pragma unroll
            for (int i = 0; i < kBatch; i++)
                e[i] = *reinterpret_cast<const uint4 *>(tab + (size_t)(b0 + i) * stride);
#pragma unroll
            for (int i = 0; i < kBatch; i++) {
                const int r = b0 + i;
                const uint32_t self = R[r] + e[i].x;
                const uint32_t swapped = prmt(R[r], 0u, swapSel);
                const uint32_t cross = swapped + e[i].y;
                const uint32_t m = signMask(subOnFma(swapped, R[r], minusOne) + e[i].z);
                R[r] = sel(m, self, cross);
                T[r >> 3] |= ~m & (0x01010101u << (r & 7));
            }
        }
        constexpr uint32_t used = G::kRegs >= 8 ? 0xFFFFFFFFu : (0x01010101u * ((1u << G::kRegs) - 1u));
#pragma unroll
        for (int w = 0; w < G::kWords; w++)
            T[w] ^= upper & used;
    }
}

template <int S>
CED_HD void genRenorm(uint32_t (&R)[GenGeom<S>::kRegs])
{
    constexpr int N = GenGeom<S>::kRegs;
    uint32_t v = R[0];
#pragma unroll
    for (int r = 1; r < N; r++)
        v = byteMin(v, R[r]);
    v = byteMin(v, prmt(v, 0u, 0x1032u));
    v = byteMin(v, prmt(v, 0u, 0x2301u));
#pragma unroll
    for (int r = 0; r < N; r++)
        R[r] -= v;
}

/* one backward step through trellis step t: p = position after step t; returns the decoded bit of step t (the state's
 * newest bit, src/viterbiDecoderButterflyk1.c:244-249) and moves p to the predecessor (:252) */
template <int S>
CED_HD uint32_t genTracebackStep(uint32_t &p, const uint32_t *words, int t)
{
    const int q = S - 1 - (t % S);
    const uint32_t r = p >> 2, l = p & 3u;
    const uint32_t dec = (words[GenGeom<S>::kWords > 1 ? (r >> 3) : 0] >> (8u * l + (r & 7u))) & 1u;
    const uint32_t bit = (p >> q) & 1u;
    p = (p & ~(1u << q)) | (dec << q);
    return bit;
}

/* words[idx] of N words without indexing the (register) array dynamically for N <= 2 */
template <int N>
CED_HD uint32_t pickWord(const uint32_t *words, uint32_t idx)
{
    if (N == 1)
        return words[0];
    if (N == 2)
        return idx ? words[1] : words[0];
    return words[idx];
}

/* the same step with the phase known at compile time (shift counts become immediates) */
template <int S, int PH>
CED_HD uint32_t genTracebackStepC(uint32_t &p, const uint32_t *words)
{
    constexpr int q = S - 1 - PH;
    const uint32_t r = p >> 2, l = p & 3u;
    const uint32_t dec = (pickWord<GenGeom<S>::kWords>(words, r >> 3) >> (8u * l + (r & 7u))) & 1u;
    const uint32_t bit = (p >> q) & 1u;
    p = (p & ~(1u << q)) | (dec << q);
    return bit;
}

/* what genForwardKernel / genTracebackKernel (swar_generic.cu) need to know about a trellis */
template <int S>
struct GenPolicy : GenGeom<S> {
    using G = GenGeom<S>;
    template <int PH>
    CED_HD static void step(uint32_t (&R)[G::kRegs], const uint8_t *tab, int V, uint32_t minusOne, uint32_t (&T)[G::kWords])
    {
        genStep<S, PH>(R, tab, G::entryStride(PH, V), minusOne, T);
    }
    CED_HD static void init(uint32_t (&R)[G::kRegs], int n) { genInitMetrics<S>(R, n); }
    CED_HD static void renorm(uint32_t (&R)[G::kRegs]) { genRenorm<S>(R); }
    CED_HD static uint32_t tbStep(uint32_t &p, const uint32_t *words, int t) { return genTracebackStep<S>(p, words, t); }
    template <int PH>
    CED_HD static uint32_t tbStepC(uint32_t &p, const uint32_t *words) { return genTracebackStepC<S, PH>(p, words); }
};

} // namespace ced
