/*
 * trellis_swar.cuh -- SIMD-in-word add-compare-select for the K=7 rate-1/2 trellis.
 *
 * One thread owns one frame.  The 64 path metrics are unsigned bytes packed four
 * to a 32-bit register (16 registers R[0..15]); all arithmetic is 8-bit and the
 * decisions are identical to src/viterbiDecoderButterflyk1.c:101-149 of the
 * reference (strict '>' compare, tie -> lower predecessor).  See DESIGN.md 4.
 *
 * In-place butterflies with a rotating state labelling
 * -----------------------------------------------------
 * A "position" p in [0,64) names byte lane (p & 3) of register (p >> 2).  Before
 * trellis step t, with phase ph = t mod 6, state s lives at position rotr6(s, ph)
 * (so position p holds state rotl6(p, ph)).  Butterfly j reads states j and j+32
 * -- two positions that differ only in position bit q = 5 - ph -- and writes the
 * successors 2j and 2j+1 back to the same two positions; that is exactly the
 * phase ph+1 labelling, so no data moves between steps and after 6 steps the
 * labelling is the identity again.
 *   ph 0..3: q is a register bit -> pairs are (R[r], R[r | 1<<(q-2)]), same lanes.
 *   ph 4,5 : q is a lane bit     -> partner = same register with lanes swapped
 *            (one PRMT); each lane computes min(self + d, partner + (2-d)).
 *
 * Branch metrics: d(j) = HD(sym[j], rx) with sym[] the reference's
 * edgeCodedBitsSymm (src/viterbiDecoderButterflyk1.c:24-29).  sym is GF(2)-linear
 * in j, so the class of lane l of register r is regCls(r) ^ laneCls(l): per phase
 * and received symbol only 4 distinct packed words X[0..3] exist (X[k] lane l =
 * HD(rx, k ^ laneCls(l, ph))), and the complement 2-d of X[k] is X[k^3].
 *
 * Compare/select without byte-min hardware: all candidates are < 128, so
 *   diff = cand1 + 0x80 - cand0   (per byte, never borrows across lanes)
 * has bit 7 set iff cand1 >= cand0; PRMT replicates that bit over the byte and
 * one LOP3 selects.  Decision bits (1 = predecessor j+32 won, as stored in the
 * reference's tracebackBufs, :145-149,185-187) are gathered with one LOP3 per
 * register into two 32-bit survivor words: position p -> word p>>5,
 * bit 8*(p&3) + ((p>>2)&7).
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CED_HD __host__ __device__ __forceinline__
#define CED_HDC __host__ __device__ constexpr
#else
#define CED_HD inline
#define CED_HDC constexpr
#endif

#if !defined(__CUDACC__)
struct uint4 { uint32_t x, y, z, w; }; /* host stand-ins for tests/hostsim (never compiled with the CUDA headers) */
struct uint2 { uint32_t x, y; };
#endif

namespace ced {

CED_HDC uint32_t rotl6(uint32_t x, int r)
{
    return r == 0 ? (x & 63u) : (((x << r) | (x >> (6 - r))) & 63u);
}
CED_HDC uint32_t rotr6(uint32_t x, int r)
{
    return r == 0 ? (x & 63u) : (((x >> r) | (x << (6 - r))) & 63u);
}
CED_HDC uint32_t parity32(uint32_t x)
{
    x ^= x >> 16;
    x ^= x >> 8;
    x ^= x >> 4;
    x ^= x >> 2;
    x ^= x >> 1;
    return x & 1u;
}
CED_HDC uint32_t hd2(uint32_t a, uint32_t b)
{
    return ((a ^ b) & 1u) + (((a ^ b) >> 1) & 1u);
}

/* A K=7, n=2 code given by its two Proakis-convention generators (octal 0113,
 * 0171 for src/defaultParams/convCodeParams.c:6). */
template <uint32_t G0, uint32_t G1>
struct K7Code {
    static constexpr int K = 7;
    static constexpr bool kRuntime = false;
    static constexpr int kCodedBits = 2;
    static constexpr int kRenormPeriod = 96;            /* DESIGN.md 4.3 */
    static constexpr uint32_t kSymMask = 0x03030303u;   /* calcHammingDist(..., n): only the low n bits count */
    static constexpr uint32_t g0 = G0, g1 = G1;
    /* src/convEncode.c:163-175: bit-reverse so bit 0 taps the newest input */
    static CED_HDC uint32_t rev7(uint32_t g)
    {
        uint32_t r = 0;
        for (int i = 0; i < 7; i++)
            r |= ((g >> i) & 1u) << (6 - i);
        return r;
    }
    static constexpr uint32_t tap0 = rev7(G0), tap1 = rev7(G1);
    /* both generators tap the newest and the oldest bit (src/viterbiDecoder.c:20-24) */
    static constexpr bool symmetric = (G0 & 1u) && ((G0 >> 6) & 1u) && (G1 & 1u) && ((G1 >> 6) & 1u);
    /* output segment for shift-register contents reg (src/convEncode.c:132-161) */
    static CED_HDC uint32_t segment(uint32_t reg)
    {
        return parity32(reg & tap0) | (parity32(reg & tap1) << 1);
    }
    /* edgeCodedBitsSymm[j], j < 32 (src/viterbiDecoderButterflyk1.c:24-29) */
    static CED_HDC uint32_t sym(uint32_t j) { return segment((j << 1) & 127u); }
    /* class of position p at phase ph: sym of the butterfly index (state bit 5 dropped) */
    static CED_HDC uint32_t cls(uint32_t p, int ph) { return sym(rotl6(p, ph) & 31u); }
    static CED_HDC uint32_t regCls(int r, int ph) { return cls((uint32_t)r << 2, ph); }
    static CED_HDC uint32_t laneCls(int l, int ph) { return cls((uint32_t)l, ph); }
    /* packed branch-metric word X[k] for phase ph and received symbol rx */
    static CED_HDC uint32_t bmWord(int ph, uint32_t rx, uint32_t k)
    {
        uint32_t w = 0;
        for (int l = 0; l < 4; l++)
            w |= hd2(rx & 3u, k ^ laneCls(l, ph)) << (8 * l);
        return w;
    }
};

using DefaultK7 = K7Code<0113, 0171>;

/* Guard word added to the "upper predecessor" candidate before the compare (see acsStep):
 * 0x80 per lane; in the lane phases the lanes holding the upper state use 0x7F so that a tie
 * still goes to the lower predecessor. */
CED_HDC uint32_t guardWord(int ph)
{
    return ph == 4 ? 0x7F7F8080u : ph == 5 ? 0x7F807F80u : 0x80808080u;
}

/* ---- byte-lane primitives ---- */
CED_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    /* PTX prmt (default mode): selector nibble bit 3 replicates the selected byte's sign bit.
     * The __byte_perm() intrinsic masks that bit away, so it cannot be used here. */
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    /* host model of PRMT (default mode), used only by tests/hostsim */
    uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t out = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t nib = (sel >> (4 * i)) & 0xFu;
        uint32_t byte = (uint32_t)(src >> (8 * (nib & 7u))) & 0xFFu;
        if (nib & 8u)
            byte = (byte & 0x80u) ? 0xFFu : 0x00u;
        out |= byte << (8 * i);
    }
    return out;
#endif
}

/* 0xFF in every byte whose bit 7 is set */
CED_HD uint32_t signMask(uint32_t x) { return prmt(x, 0u, 0xba98u); }
CED_HD uint32_t sel(uint32_t mask, uint32_t a, uint32_t b) { return (a & mask) | (b & ~mask); }

constexpr uint32_t kGuard = 0x80808080u;
constexpr uint32_t kInitMetricWord = 0x41414141u; /* NUM_STATES+1 = 65 in every lane (:59-67) */

CED_HD void initMetrics(uint32_t (&R)[16])
{
#pragma unroll
    for (int r = 0; r < 16; r++)
        R[r] = kInitMetricWord;
    R[0] = kInitMetricWord & 0xFFFFFF00u; /* STARTING_STATE 0 -> metric 0 */
}

/* a - b issued as a true multiply-add (b * minusOne + a) so that it runs on the FMA pipe;
 * `minusOne` is 0xFFFFFFFF passed at run time so the compiler cannot fold it back into an
 * ALU-pipe IADD3. */
CED_HD uint32_t subOnFma(uint32_t a, uint32_t b, uint32_t minusOne)
{
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(minusOne), "r"(a));
    return d;
#else
    return b * minusOne + a;
#endif
}

/*
 * One trellis step at compile-time phase PH.  X = the 4 packed branch-metric words for
 * (PH, rx); E[k] = X[k^3] - X[k] + guardWord(PH).  On return T0/T1 hold the 64 decision bits.
 *
 * Compare words.  For a butterfly (lo = metrics of states j, hi = of j+32, d = X[k], dc = X[k^3]):
 *   a1 + G - a0 = (hi - lo) + E[k]       b1 + G - b0 = (hi - lo) + E[k^3]
 * (byte-wise identical to the direct form because every true per-lane result lies in [0,255]).
 * Pipe balance: profiles/r1_v1_* showed the ALU pipe 82 % busy and the FMA pipe 19 % with the
 * direct 3-input IADD3 form; hi - lo is therefore issued once per butterfly as an IMAD and the
 * two + E as 2-input adds, which ptxas places on the FMA pipe (IMAD.IADD).
 */
template <class Code, int PH>
CED_HD void acsStep(uint32_t (&R)[16], const uint32_t (&X)[4], const uint32_t (&E)[4], uint32_t minusOne,
                    uint32_t &T0, uint32_t &T1)
{
    constexpr int q = 5 - PH;
    uint32_t t0 = 0, t1 = 0;
    if constexpr (q >= 2) {
        constexpr int rb = q - 2;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if ((r >> rb) & 1)
                continue;
            const int rh = r | (1 << rb);
            const uint32_t k = Code::regCls(r, PH);
            const uint32_t d = X[k], dc = X[k ^ 3u];
            const uint32_t lo = R[r], hi = R[rh];
            const uint32_t a0 = lo + d, a1 = hi + dc;   /* successors 2j   (:109-110) */
            const uint32_t b0 = lo + dc, b1 = hi + d;   /* successors 2j+1 (:113-114) */
            const uint32_t delta = subOnFma(hi, lo, minusOne);
            /* FF: candidate from j+32 >= candidate from j -> keep the lower predecessor */
            const uint32_t ma = signMask(delta + E[k]);
            const uint32_t mb = signMask(delta + E[k ^ 3u]);
            R[r] = sel(ma, a0, a1);
            R[rh] = sel(mb, b0, b1);
            const uint32_t ca = 0x01010101u << (r & 7), cb = 0x01010101u << (rh & 7);
            if (r < 8) t0 |= ~ma & ca; else t1 |= ~ma & ca;
            if (rh < 8) t0 |= ~mb & cb; else t1 |= ~mb & cb;
        }
    } else {
        /* lanes with position bit q set hold the upper state (j+32) of their pair */
        constexpr uint32_t swapSel = (q == 1) ? 0x1032u : 0x2301u;
        constexpr uint32_t upper = (q == 1) ? 0xFFFF0000u : 0xFF00FF00u;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const uint32_t k = Code::regCls(r, PH);
            const uint32_t self = R[r] + X[k];
            const uint32_t swapped = prmt(R[r], 0u, swapSel);
            const uint32_t cross = swapped + X[k ^ 3u];
            /* cross + guard - self; guardWord(PH) makes it
             *   lower lanes: FF iff cross >= self (keep self, decision 0)
             *   upper lanes: FF iff cross >  self (keep self, decision 1) */
            const uint32_t m = signMask(subOnFma(swapped, R[r], minusOne) + E[k]);
            R[r] = sel(m, self, cross);
            const uint32_t c = 0x01010101u << (r & 7);
            if (r < 8) t0 |= ~m & c; else t1 |= ~m & c;
        }
        t0 ^= upper;
        t1 ^= upper;
    }
    T0 = t0;
    T1 = t1;
}

/* ---- any symmetric K=7, n=2 code chosen at run time --------------------------------
 * K7Code fixes, at compile time, which of the four branch-metric words every register uses in every
 * phase.  For generators known only at run time the same information goes into a table that the host
 * builds once per code and the kernel keeps in shared memory: for every (phase, received symbol) sixteen
 * 8-byte entries, one per butterfly (register-pair phases: 8 used) or per register (lane phases), holding
 * { X[k], E[k] } for the class k acsStep would have picked; the complements follow arithmetically,
 *   X[k^full] = n * 0x01010101 - X[k]        E[k^full] = 2 * guardWord - E[k]     (mod 2^32, exact)
 * The lanes of a warp are different frames with different symbols, so one load touches up to four entries
 * (rx = 0..3): they are stored next to each other, [phase][entry][rx], 32 contiguous bytes = 8 banks, which
 * makes the load conflict-free.  (Measured on the way: [phase][rx][entry] puts the four on the same banks --
 * 4-way conflicts, shared-memory pipe 95 % busy, 87 Gbit/s; 16-byte entries holding all four operands:
 * 52.7 Gbit/s.) */
struct Word2 { uint32_t x, y; }; /* layout of a uint2, usable in host code without CUDA headers */

/* tag: "the code is whatever the step table in shared memory says"; N = coded bits per segment (2 or 3).
 * With N = 3 a branch costs up to 3, so the metrics are renormalised every 24 steps: spread <= 3 * 6 = 18 after
 * a renorm, + 3 * 24 growth + 3 for the candidate = 93 < 128 (same argument as DESIGN.md 4.3). */
template <int N>
struct RuntimeK7 {
    static_assert(N == 2 || N == 3, "step tables are built for 2 or 3 coded bits");
    static constexpr bool kRuntime = true;
    static constexpr int kCodedBits = N;
    static constexpr int kVariants = 1 << N;                 /* received symbols */
    static constexpr int kTableEntries = 6 * 16 * kVariants; /* Word2 each */
    static constexpr int kRenormPeriod = N == 2 ? 96 : 24;
    static constexpr uint32_t kSymMask = 0x01010101u * (uint32_t)(kVariants - 1);
};

struct K7Taps {
    uint32_t tap[3]; /* generators bit-reversed onto the shift register (src/convEncode.c:163-175) */
    int n;
    CED_HD uint32_t segment(uint32_t reg) const
    {
        uint32_t v = 0;
        for (int i = 0; i < n; i++)
            v |= parity32(reg & tap[i]) << i;
        return v;
    }
    CED_HD uint32_t cls(uint32_t p, int ph) const { return segment(((rotl6(p, ph) & 31u) << 1) & 127u); }
    CED_HD uint32_t bmWord(int ph, uint32_t rx, uint32_t k) const
    {
        uint32_t w = 0;
        for (uint32_t l = 0; l < 4; l++) {
            uint32_t diff = (rx ^ k ^ cls(l, ph)) & ((1u << n) - 1u), hd = 0;
            for (; diff; diff >>= 1)
                hd += diff & 1u;
            w |= hd << (8 * l);
        }
        return w;
    }
};

CED_HD K7Taps makeK7Taps(int n, const uint32_t *gens)
{
    K7Taps t = {{0u, 0u, 0u}, n};
    for (int j = 0; j < n; j++)
        for (int i = 0; i < 7; i++)
            t.tap[j] |= ((gens[j] >> i) & 1u) << (6 - i);
    return t;
}

/* out[6 * 16 * 2^n]: entry idx of (phase, rx) at out[(phase * 16 + idx) * 2^n + rx] */
inline void buildStepTable(const K7Taps &c, Word2 *out)
{
    const int variants = 1 << c.n;
    const uint32_t full = (uint32_t)variants - 1u;
    for (int ph = 0; ph < 6; ph++)
        for (uint32_t rx = 0; rx < (uint32_t)variants; rx++) {
            Word2 *e = out + ph * 16 * variants + (int)rx;
            const uint32_t gw = guardWord(ph);
            const int q = 5 - ph;
            int idx = 0;
            for (int r = 0; r < 16; r++) {
                if (q >= 2 && ((r >> (q - 2)) & 1))
                    continue;
                const uint32_t k = c.cls((uint32_t)r << 2, ph);
                const uint32_t d = c.bmWord(ph, rx, k), dc = c.bmWord(ph, rx, k ^ full);
                e[variants * idx].x = d;
                e[variants * idx].y = dc - d + gw;
                idx++;
            }
            for (; idx < 16; idx++)
                e[variants * idx].x = e[variants * idx].y = 0u;
        }
}

/* acsStep with the operands taken from the table entries of (PH, rx) -- tab points at entry 0 of that pair,
 * entry i sits at tab[2^N * i] -- same arithmetic, same decisions */
template <int PH, int N, class Entry>
CED_HD void acsStepTable(uint32_t (&R)[16], const Entry *tab, uint32_t minusOne, uint32_t &T0, uint32_t &T1)
{
    constexpr int q = 5 - PH;
    constexpr int stride = 1 << N;
    constexpr uint32_t kBitsPerLane = 0x01010101u * (uint32_t)N; /* X[k] + X[k ^ full] in every lane */
    uint32_t t0 = 0, t1 = 0;
    if constexpr (q >= 2) {
        constexpr int rb = q - 2;
        int idx = 0;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if ((r >> rb) & 1)
                continue;
            const int rh = r | (1 << rb);
            const Entry e = tab[stride * idx++];
            const uint32_t d = e.x, dc = subOnFma(kBitsPerLane, d, minusOne);
            const uint32_t lo = R[r], hi = R[rh];
            const uint32_t a0 = lo + d, a1 = hi + dc;
            const uint32_t b0 = lo + dc, b1 = hi + d;
            const uint32_t delta = subOnFma(hi, lo, minusOne);
            const uint32_t ma = signMask(delta + e.y);
            const uint32_t mb = signMask(delta + (2u * guardWord(PH) - e.y));
            R[r] = sel(ma, a0, a1);
            R[rh] = sel(mb, b0, b1);
            const uint32_t ca = 0x01010101u << (r & 7), cb = 0x01010101u << (rh & 7);
            if (r < 8) t0 |= ~ma & ca; else t1 |= ~ma & ca;
            if (rh < 8) t0 |= ~mb & cb; else t1 |= ~mb & cb;
        }
    } else {
        constexpr uint32_t swapSel = (q == 1) ? 0x1032u : 0x2301u;
        constexpr uint32_t upper = (q == 1) ? 0xFFFF0000u : 0xFF00FF00u;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const Entry e = tab[stride * r];
            const uint32_t self = R[r] + e.x;
            const uint32_t swapped = prmt(R[r], 0u, swapSel);
            const uint32_t cross = swapped + (kBitsPerLane - e.x);
            const uint32_t m = signMask(subOnFma(swapped, R[r], minusOne) + e.y);
            R[r] = sel(m, self, cross);
            const uint32_t c = 0x01010101u << (r & 7);
            if (r < 8) t0 |= ~m & c; else t1 |= ~m & c;
        }
        t0 ^= upper;
        t1 ^= upper;
    }
    T0 = t0;
    T1 = t1;
}

/* Subtract the minimum of the 64 metrics from all of them.  Decisions do not
 * depend on when this happens as long as no candidate reaches 128 (DESIGN.md
 * 4.3); the reference does it every 121 steps (:159-183), the batch kernel
 * every kRenormPeriod steps. */
CED_HD uint32_t byteMin(uint32_t a, uint32_t b)
{
    return sel(signMask(b + kGuard - a), a, b);
}
CED_HD void renorm(uint32_t (&R)[16])
{
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++)
        m[i] = byteMin(R[i], R[i + 8]);
#pragma unroll
    for (int i = 0; i < 4; i++)
        m[i] = byteMin(m[i], m[i + 4]);
    uint32_t v = byteMin(byteMin(m[0], m[2]), byteMin(m[1], m[3]));
    v = byteMin(v, prmt(v, 0u, 0x1032u));
    v = byteMin(v, prmt(v, 0u, 0x2301u));
#pragma unroll
    for (int r = 0; r < 16; r++)
        R[r] -= v;
}
constexpr int kRenormPeriod = 96;

/* ---- traceback bookkeeping (position in "survivor bit index" form) ----
 * b = 32*(p>>5) + 8*(p&3) + ((p>>2)&7).  Position bit q = 5-ph maps to b bit
 * kPairBitInB[ph]. */
CED_HDC int pairBitInB(int ph)
{
    /* q = 5-ph: 5,4,3,2,1,0 -> b bit 5,2,1,0,4,3 (one nibble per phase) */
    return (int)((0x340125u >> (4 * ph)) & 7u);
}

/* The two lane layouts of the packed metrics, as far as the traceback is concerned: where position bit
 * q = 5 - ph sits in the survivor bit index b.
 *   Lanes8  (this file):            b = 32*(p>>5) + 8*(p&3)  + ((p>>2)&7)
 *   Lanes16 (trellis_swar16.cuh):   b = 32*(p>>5) + 16*(p&1) + ((p>>1)&15)    q = 5..0 -> b bit 5,3,2,1,0,4 */
struct Lanes8 {
    static CED_HDC int pairBit(int ph) { return pairBitInB(ph); }
    /* survivor bit index -> position (== state when the next step has phase 0) */
    static CED_HD uint32_t toPosition(uint32_t b) { return (b & 32u) | ((b & 7u) << 2) | ((b >> 3) & 3u); }
};
struct Lanes16 {
    static CED_HDC int pairBit(int ph) { return (int)((0x401235u >> (4 * ph)) & 7u); }
    static CED_HD uint32_t toPosition(uint32_t b) { return (b & 32u) | ((b & 15u) << 1) | ((b >> 4) & 1u); }
};

CED_HD uint32_t rotr32(uint32_t x, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n); /* one SHF; only the low 5 bits of n count */
#else
    n &= 31u;
    return n ? (x >> n) | (x << (32u - n)) : x;
#endif
}

/* Windowed traceback start (continuous streams): after renorm() the smallest metric is 0; returns, in b form, the
 * lowest position holding it.  Called between slices, where the next phase is 0 and position == state, so ties go
 * to the lowest state.  The zero-byte finder is exact for bytes < 128 up to and including the lowest zero lane. */
CED_HD uint32_t bestPositionB(const uint32_t (&R)[16])
{
    uint32_t best = 0;
#pragma unroll
    for (int r = 15; r >= 0; r--) {
        const uint32_t z = (R[r] - 0x01010101u) & ~R[r] & 0x80808080u;
        if (z) {
#ifdef __CUDA_ARCH__
            const int first = __ffs((int)z) - 1;
#else
            const int first = __builtin_ffs((int)z) - 1;
#endif
            best = 4u * (uint32_t)r + (uint32_t)(first >> 3);
        }
    }
    return 32u * (best >> 5) + 8u * (best & 3u) + ((best >> 2) & 7u);
}

/* One backward step through trellis step t (phase ph).  On entry b locates the
 * survivor state after step t; returns that state's newest bit (the decoded bit
 * of step t, src/viterbiDecoderButterflyk1.c:244-249) and moves b to the
 * predecessor (:252). */
template <class L = Lanes8>
CED_HD uint32_t tracebackStep(uint32_t &b, uint32_t w0, uint32_t w1, int ph)
{
    const int qb = L::pairBit(ph);
    const uint32_t word = (b & 32u) ? w1 : w0;
    const uint32_t dec = (word >> (b & 31u)) & 1u;
    const uint32_t bit = (b >> qb) & 1u;
    b = (b & ~(1u << qb)) | (dec << qb);
    return bit;
}

/* backward step with the phase known at compile time */
template <class Lay, int PH>
CED_HD uint32_t tracebackStepC(uint32_t &b, uint32_t w0, uint32_t w1)
{
    constexpr int qb = Lay::pairBit(PH);
    const uint32_t word = (b & 32u) ? w1 : w0;
    const uint32_t dec = (word >> (b & 31u)) & 1u;
    const uint32_t bit = (b >> qb) & 1u;
    b = (b & ~(1u << qb)) | (dec << qb);
    return bit;
}

/* 8 steps = 4 pairs = one output byte; pair i of the group holds steps (base+6-2i, base+7-2i),
 * PH0 = phase of the group's first step (step base, a multiple of 8 inside a 24-step block). */
template <class Lay, int PH0>
CED_HD uint32_t tracebackByteC(uint32_t &b, const uint4 (&w)[4])
{
    uint32_t acc = 0;
    acc |= tracebackStepC<Lay, (PH0 + 7) % 6>(b, w[0].z, w[0].w) << 0;
    acc |= tracebackStepC<Lay, (PH0 + 6) % 6>(b, w[0].x, w[0].y) << 1;
    acc |= tracebackStepC<Lay, (PH0 + 5) % 6>(b, w[1].z, w[1].w) << 2;
    acc |= tracebackStepC<Lay, (PH0 + 4) % 6>(b, w[1].x, w[1].y) << 3;
    acc |= tracebackStepC<Lay, (PH0 + 3) % 6>(b, w[2].z, w[2].w) << 4;
    acc |= tracebackStepC<Lay, (PH0 + 2) % 6>(b, w[2].x, w[2].y) << 5;
    acc |= tracebackStepC<Lay, (PH0 + 1) % 6>(b, w[3].z, w[3].w) << 6;
    acc |= tracebackStepC<Lay, (PH0 + 0) % 6>(b, w[3].x, w[3].y) << 7;
    return acc;
}

} // namespace ced
