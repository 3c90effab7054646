/*
 * swar_radix4.cuh -- SIMD-in-word add-compare-select for rate-2/n codes (k = 2): the reference's ONE shift
 * register of k*K bits that takes two message bits per coded segment (src/convEncode.h:8-18, src/convEncode.c:46-130),
 * decoded on a trellis of 4^S states (S = K - 1) with four branches into every state (src/viterbiDecoder.c:95-128):
 *   destination d: edgeOut = d mod 4, sources d / 4 + edgeIn * 4^(S-1), the LOWEST edgeIn keeps a tie
 *   (argminPathMetrics' '<=' trees, src/convHelpers.h:40-63).
 *
 * Construction of swar_generic.cuh carried over to radix 4: one thread = one frame, byte metrics four to a register
 * (4^S / 4 registers), and the state labelling rotates by TWO bits per step so that butterflies are in place: position
 * p holds state rotl(p, 2 ph) before a step of phase ph = t mod S.  A butterfly is the four states that differ in the
 * two oldest bits; they sit in position bits (q, q+1), q = 2 (S - 1 - ph):
 *   q >= 2  four REGISTERS r_0..r_3 (register-index bits (q-2, q-1) = e); lanes are four different butterflies
 *           new r_o = min_e (r_e + X[e][o]) -- 16 adds, and per output a two-level tournament in edgeIn order:
 *           (0 vs 1), (2 vs 3) with the differences r_1 - r_0 and r_3 - r_2 shared by the four outputs (compare =
 *           difference + E, E = X1 - X0 + 0x80.. from the table), then winner vs winner on the selected values
 *   q == 0  the four LANES of one register are the butterfly (last phase): source e is lane e broadcast to all lanes
 *           (one PRMT), destination o is lane o, same tournament on explicit differences
 * The winner of every comparison is the lower edgeIn unless the other is strictly smaller, which is the guard-bit
 * compare of trellis_swar.cuh with guard 0x80 in every lane.  Decisions are 2 bits per state: bit 1 = the (2, 3) pair
 * won, bit 0 = the odd member of the winning pair won; a step stores W = 2 * max(1, registers / 8) words: first the
 * bit-0 words, then the bit-1 words, position (register r, lane l) -> word r >> 3, bit 8 l + (r & 7).
 *
 * Nothing about the code is compiled in: the host builds a step table (buildR4Table) --
 *   register phases, per quad and received symbol: X[e][o] (16 words), E01[o], E23[o] (8 words)     96-byte entries
 *   lane phase, per register and received symbol:  X[e] (4 words; lane = destination o)             16-byte entries
 *
 * Exactness of 8-bit metrics: start metrics 0 / n S + 1 ("never wins": every state is reachable from state 0 within S
 * steps at cost <= n S; the reference's NUM_STATES + 1 in a 16-bit METRIC_TYPE plays the same role and gives the same
 * decisions, src/viterbiDecoder.c:236-258), spread <= n S afterwards, growth <= n per step, renormalisation every 24
 * steps: candidates <= n S + 25 n <= 87 < 128 for n <= 3, S <= 4.  tests/hostsim asserts the largest metric seen.
 */
#pragma once
#include "swar_generic.cuh"
#include <string.h>

namespace ced {

struct R4Code {
    int S, n;            /* 2-bit chunks in the state (K - 1), coded bits per segment */
    uint32_t tap[3];     /* generators over the 2K register bits, bit 0 on the newest input bit */
};

CED_HD uint32_t r4Label(const R4Code &c, uint32_t state, uint32_t in2)
{
    const uint32_t reg = (state << 2) | in2;
    uint32_t v = 0;
    for (int i = 0; i < c.n; i++)
        v |= parity32(reg & c.tap[i]) << i;
    return v;
}

CED_HD uint32_t r4Hd(uint32_t a, uint32_t b, int n)
{
    uint32_t x = (a ^ b) & ((1u << n) - 1u), d = 0;
    for (; x; x >>= 1)
        d += x & 1u;
    return d;
}

CED_HD uint32_t r4Rotl(uint32_t x, int r, int bits)
{
    const uint32_t m = (1u << bits) - 1u;
    r %= bits;
    return r == 0 ? (x & m) : (((x << r) | (x >> (bits - r))) & m);
}

template <int S>
struct R4Geom {
    static_assert(S >= 1 && S <= 4, "4 .. 256 states");
    static constexpr int kBits = 2 * S;
    static constexpr int kStates = 1 << kBits;
    static constexpr int kRegs = kStates / 4;
    static constexpr int kPhases = S;
    static constexpr int kQuads = kRegs >= 4 ? kRegs / 4 : 0;
    static constexpr int kHalfWords = kRegs >= 8 ? kRegs / 8 : 1;
    static constexpr int kWords = 2 * kHalfWords;
    static constexpr int kRenorm = 24;
    static constexpr int kChunk = 96;        /* trellis steps per staged tile (a multiple of S) */
    static constexpr int kTail = S;          /* tail segments */
    static constexpr int kStepBits = 2;      /* decoded bits per trellis step */
    CED_HD static constexpr int tableBytes(int V) { return ((S - 1) * kQuads * 96 + kRegs * 16) * V; }
    CED_HD static constexpr int phaseBase(int ph, int V) { return ph * kQuads * 96 * V; }
    /* byte offset of the received symbol's variant inside an entry row; `off` = rx * 32 as staged in the tile */
    CED_HD static constexpr uint32_t rxOffset(int ph, uint32_t off) { return ph < S - 1 ? off * 3u : off >> 1; }
    CED_HD static constexpr int entryStride(int ph, int V) { return (ph < S - 1 ? 96 : 16) * V; }
};

/* host: fill `table` (R4Geom<S>::tableBytes(1 << n) bytes) for code c */
template <int S>
inline void buildR4Table(const R4Code &c, uint8_t *table)
{
    using G = R4Geom<S>;
    const int V = 1 << c.n, B = G::kBits;
    const uint32_t top = 2u * (uint32_t)(S - 1);   /* the two oldest state bits */
    for (int ph = 0; ph < S; ph++) {
        const int q = 2 * (S - 1 - ph);
        uint8_t *base = table + G::phaseBase(ph, V);
        if (q >= 2) {
            const int rb = q - 2;                  /* register-index bits (rb, rb + 1) select the source e */
            for (int idx = 0; idx < G::kQuads; idx++) {
                const int low = (1 << rb) - 1;
                const int r0 = ((idx & ~low) << 2) | (idx & low);
                for (int rx = 0; rx < V; rx++) {
                    uint32_t w[24];
                    for (int i = 0; i < 24; i++)
                        w[i] = 0;
                    for (uint32_t l = 0; l < 4; l++) {
                        const uint32_t s0 = r4Rotl(4u * (uint32_t)r0 + l, 2 * ph, B);   /* source with e = 0 */
                        const uint32_t m = s0 & ((1u << top) - 1u);
                        for (uint32_t e = 0; e < 4; e++)
                            for (uint32_t o = 0; o < 4; o++)
                                w[4 * e + o] |= r4Hd(rx, r4Label(c, m | (e << top), o), c.n) << (8 * l);
                    }
                    for (int o = 0; o < 4; o++) {
                        w[16 + o] = w[4 + o] - w[o] + 0x80808080u;        /* E01[o] = X[1][o] - X[0][o] + guard */
                        w[20 + o] = w[12 + o] - w[8 + o] + 0x80808080u;   /* E23[o] */
                    }
                    memcpy(base + ((size_t)idx * V + rx) * 96, w, 96);
                }
            }
        } else {
            for (int r = 0; r < G::kRegs; r++)
                for (int rx = 0; rx < V; rx++) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    const uint32_t s0 = r4Rotl(4u * (uint32_t)r, 2 * ph, B);            /* lane 0 = source e = 0 */
                    const uint32_t m = s0 & ((1u << top) - 1u);
                    for (uint32_t e = 0; e < 4; e++)
                        for (uint32_t o = 0; o < 4; o++)                                /* lane o = destination o */
                            w[e] |= r4Hd(rx, r4Label(c, m | (e << top), o), c.n) << (8 * o);
                    memcpy(base + ((size_t)r * V + rx) * 16, w, 16);
                }
        }
    }
}

template <int S>
CED_HD void r4InitMetrics(uint32_t (&R)[R4Geom<S>::kRegs], int n)
{
    const uint32_t never = (uint32_t)(n * S + 1) * 0x01010101u;
#pragma unroll
    for (int r = 0; r < R4Geom<S>::kRegs; r++)
        R[r] = never;
    R[0] &= 0xFFFFFF00u; /* state 0 sits at position 0 in every phase */
}

/* winner of (c0, c1) given diff = c1 - c0 + guard: c0 unless c1 is strictly smaller; returns FF where c0 is kept */
CED_HD uint32_t r4Keep(uint32_t diff) { return signMask(diff); }

/* One trellis step of compile-time phase PH; `tab` = entries of (PH, rx), entry i at tab + i * stride. */
template <int S, int PH>
CED_HD void r4Step(uint32_t (&R)[R4Geom<S>::kRegs], const uint8_t *tab, int stride, uint32_t minusOne,
                   uint32_t (&T)[R4Geom<S>::kWords])
{
    using G = R4Geom<S>;
    constexpr int q = 2 * (S - 1 - PH);
    constexpr int H = G::kHalfWords;
#pragma unroll
    for (int w = 0; w < G::kWords; w++)
        T[w] = 0;
    if constexpr (q >= 2) {
        constexpr int rb = q - 2;
        constexpr int low = (1 << rb) - 1;
#pragma unroll
        for (int idx = 0; idx < G::kQuads; idx++) {
            const int r0 = ((idx & ~low) << 2) | (idx & low);
            const uint4 *e = reinterpret_cast<const uint4 *>(tab + (size_t)idx * stride);
            const uint4 x0 = e[0], x1 = e[1], x2 = e[2], x3 = e[3], e01 = e[4], e23 = e[5];
            const uint32_t s0 = R[r0], s1 = R[r0 + (1 << rb)], s2 = R[r0 + (2 << rb)], s3 = R[r0 + (3 << rb)];
            const uint32_t d10 = subOnFma(s1, s0, minusOne), d32 = subOnFma(s3, s2, minusOne);
            const uint32_t X0[4] = {x0.x, x0.y, x0.z, x0.w}, X1[4] = {x1.x, x1.y, x1.z, x1.w};
            const uint32_t X2[4] = {x2.x, x2.y, x2.z, x2.w}, X3[4] = {x3.x, x3.y, x3.z, x3.w};
            const uint32_t E01[4] = {e01.x, e01.y, e01.z, e01.w}, E23[4] = {e23.x, e23.y, e23.z, e23.w};
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const int r = r0 + (o << rb);
                const uint32_t c0 = s0 + X0[o], c1 = s1 + X1[o], c2 = s2 + X2[o], c3 = s3 + X3[o];
                const uint32_t k01 = r4Keep(d10 + E01[o]), k23 = r4Keep(d32 + E23[o]);
                const uint32_t w01 = sel(k01, c0, c1), w23 = sel(k23, c2, c3);
                const uint32_t k2 = r4Keep(subOnFma(w23, w01, minusOne) + kGuard);
                R[r] = sel(k2, w01, w23);
                const uint32_t bit = 0x01010101u << (r & 7);
                T[r >> 3] |= ~sel(k2, k01, k23) & bit;        /* bit 0: the odd member of the winning pair */
                T[H + (r >> 3)] |= ~k2 & bit;                 /* bit 1: the (2, 3) pair */
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < G::kRegs; r++) {
            const uint4 x = *reinterpret_cast<const uint4 *>(tab + (size_t)r * stride);
            const uint32_t v = R[r];
            const uint32_t c0 = prmt(v, 0u, 0x0000u) + x.x, c1 = prmt(v, 0u, 0x1111u) + x.y;
            const uint32_t c2 = prmt(v, 0u, 0x2222u) + x.z, c3 = prmt(v, 0u, 0x3333u) + x.w;
            const uint32_t k01 = r4Keep(subOnFma(c1, c0, minusOne) + kGuard);
            const uint32_t k23 = r4Keep(subOnFma(c3, c2, minusOne) + kGuard);
            const uint32_t w01 = sel(k01, c0, c1), w23 = sel(k23, c2, c3);
            const uint32_t k2 = r4Keep(subOnFma(w23, w01, minusOne) + kGuard);
            R[r] = sel(k2, w01, w23);
            const uint32_t bit = 0x01010101u << (r & 7);
            T[r >> 3] |= ~sel(k2, k01, k23) & bit;
            T[H + (r >> 3)] |= ~k2 & bit;
        }
    }
}

template <int S>
CED_HD void r4Renorm(uint32_t (&R)[R4Geom<S>::kRegs])
{
    constexpr int N = R4Geom<S>::kRegs;
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

/* one backward step through trellis step t: p = position after step t; returns the two decoded bits of step t (the
 * state's newest chunk, src/viterbiDecoderButterflyk1.c:244-249) and moves p to the source state (:252) */
template <int S>
CED_HD uint32_t r4TracebackStep(uint32_t &p, const uint32_t *words, int t)
{
    using G = R4Geom<S>;
    const int q = 2 * (S - 1 - (t % S));
    const uint32_t r = p >> 2, l = p & 3u, sh = 8u * l + (r & 7u);
    const uint32_t wi = G::kHalfWords > 1 ? (r >> 3) : 0u;
    const uint32_t dec = ((words[wi] >> sh) & 1u) | (((words[G::kHalfWords + wi] >> sh) & 1u) << 1);
    const uint32_t out = (p >> q) & 3u;
    p = (p & ~(3u << q)) | (dec << q);
    return out;
}

template <int S, int PH>
CED_HD uint32_t r4TracebackStepC(uint32_t &p, const uint32_t *words)
{
    using G = R4Geom<S>;
    constexpr int q = 2 * (S - 1 - PH);
    const uint32_t r = p >> 2, l = p & 3u, sh = 8u * l + (r & 7u);
    const uint32_t w0 = pickWord<G::kHalfWords>(words, r >> 3), w1 = pickWord<G::kHalfWords>(words + G::kHalfWords, r >> 3);
    const uint32_t dec = ((w0 >> sh) & 1u) | (((w1 >> sh) & 1u) << 1);
    const uint32_t out = (p >> q) & 3u;
    p = (p & ~(3u << q)) | (dec << q);
    return out;
}

/* what genForwardKernel / genTracebackKernel (swar_generic.cu) need to know about a trellis */
template <int S>
struct R4Policy : R4Geom<S> {
    using G = R4Geom<S>;
    template <int PH>
    CED_HD static void step(uint32_t (&R)[G::kRegs], const uint8_t *tab, int V, uint32_t minusOne, uint32_t (&T)[G::kWords])
    {
        r4Step<S, PH>(R, tab, G::entryStride(PH, V), minusOne, T);
    }
    CED_HD static void init(uint32_t (&R)[G::kRegs], int n) { r4InitMetrics<S>(R, n); }
    CED_HD static void renorm(uint32_t (&R)[G::kRegs]) { r4Renorm<S>(R); }
    CED_HD static uint32_t tbStep(uint32_t &p, const uint32_t *words, int t) { return r4TracebackStep<S>(p, words, t); }
    template <int PH>
    CED_HD static uint32_t tbStepC(uint32_t &p, const uint32_t *words) { return r4TracebackStepC<S, PH>(p, words); }
};

} // namespace ced
