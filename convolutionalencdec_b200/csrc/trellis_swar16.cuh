/*
 * trellis_swar16.cuh -- SIMD-in-word add-compare-select with 16-bit path metrics, two per register, for
 * SOFT-decision decoding of the K=7 rate-1/2 code (north_star kernel (1) "soft or hard symbols").
 *
 * Same construction as trellis_swar.cuh (one thread = one frame, in-place butterflies, rotating state labels,
 * guard-bit compare, PRMT sign-replicate + LOP3 select), with 16-bit lanes because soft branch costs reach 512:
 *   position p in [0,64) = 16-bit lane (p & 1) of register (p >> 1), 32 registers R[0..31];
 *   before step t (phase ph = t mod 6) state s sits at position rotr6(s, ph); butterfly partners differ in
 *   position bit q = 5 - ph: a register bit for ph 0..4 (pure register renaming), the lane bit for ph 5
 *   (partner = same register, halves swapped, one PRMT).
 *
 * Branch costs.  A received segment is two int8 (s0 for generator 0, s1 for generator 1; BPSK bit 0 -> +).
 * The decoder's definition (include/ced_abi.h, ced_decode_soft_batch) weights every disagreeing coded bit by
 * its reliability, cost(c) = sum_i |s_i| [hard(s_i) != c_i] -- the reference's calcHammingDist
 * (src/viterbiDecoder.c:260-285) generalised.  The kernel uses the affinely equivalent correlation form
 *   x_i = 128 - s_i in [1,256]      cost'(c) = sum_i (c_i ? 256 - x_i : x_i) = 2 cost(c) + (256 - |s0| - |s1|)
 * Every path through a step picks up the same offset and all costs are doubled, so every comparison -- ties
 * included -- has the same outcome as with cost(c) (src/viterbiDecoderButterflyk1.c:129-130 tie rule kept).
 * What the form buys: the complement label costs 512 - cost'(c), a constant, exactly like 2 - d in the hard
 * kernel, so one staged word W = cost'(0) | cost'(1) << 16 per step is enough.
 *
 * Exactness of the 16-bit guard-bit compare: a branch costs <= 512, the smallest metric grows by <= 256 per
 * step (the two branches out of a state cost c and 512 - c), the spread is <= 6 * 512 once all states are
 * reachable and start metrics are 3584 (= 2 * n * 128 * K in correlation units) before that, so with a renormalisation every
 * 48 steps candidates stay below 3072 + 48 * 256 + 512 = 15872 < 32768: bit 15 is free for the guard.
 */
#pragma once
#include "trellis_swar.cuh"

namespace ced {

constexpr uint32_t kGuard16 = 0x80008000u;
constexpr uint32_t kSoftInit = 3584u;           /* "never wins": 2 * (n * 128 * K), above any reachable metric */
constexpr uint32_t kSoftFull = 0x02000200u;     /* cost'(c) + cost'(~c) in both lanes */
constexpr int kSoftRenormPeriod = 48;

CED_HDC uint32_t guardWord16(int ph)
{
    return ph == 5 ? 0x7FFF8000u : kGuard16; /* lane 1 holds the upper state in the lane phase: tie -> lower */
}

/* 0xFFFF in every 16-bit lane whose bit 15 is set */
CED_HD uint32_t signMask16(uint32_t x) { return prmt(x, 0u, 0xbb99u); }

CED_HD void initMetrics16(uint32_t (&R)[32])
{
#pragma unroll
    for (int r = 0; r < 32; r++)
        R[r] = kSoftInit * 0x10001u;
    R[0] = kSoftInit << 16; /* state 0 (position 0 in every phase) starts at 0 */
}

/* staged word of one received segment: cost'(0) | cost'(1) << 16 */
CED_HD uint32_t softWord(int s0, int s1)
{
    const uint32_t x0 = (uint32_t)(128 - s0), x1 = (uint32_t)(128 - s1);
    return (x0 + x1) | ((256u - x0 + x1) << 16);
}

/*
 * The four packed branch-cost words of a step and the compare offsets:
 *   X[k] = cost'(k ^ laneCls(0)) | cost'(k ^ laneCls(1)) << 16,  laneCls(0) = 0, laneCls(1) = c (per phase)
 *   E[k] = X[k^3] - X[k] + guardWord16(ph)                       (see acsStep in trellis_swar.cuh)
 * W = cost'(0) | cost'(1) << 16; 512 - W = cost'(3) | cost'(2) << 16.
 */
template <class Code, int PH>
CED_HD void softBranchWords(uint32_t W, uint32_t minusTwo, uint32_t (&X)[4], uint32_t (&E)[4])
{
    constexpr uint32_t c = Code::cls(1u, PH);
    const uint32_t Wc = kSoftFull - W;
    /* PRMT source bytes: 0-1 cost'(0), 2-3 cost'(1), 4-5 cost'(3), 6-7 cost'(2) */
    constexpr uint32_t half[4] = {0x10u, 0x32u, 0x76u, 0x54u}; /* selector byte pair of cost'(0..3) */
    constexpr uint32_t sel0 = half[0] | (half[0 ^ c] << 8), sel1 = half[1] | (half[1 ^ c] << 8);
    X[0] = prmt(W, Wc, sel0);
    X[1] = prmt(W, Wc, sel1);
    X[2] = kSoftFull - X[1];
    X[3] = kSoftFull - X[0];
    constexpr uint32_t g = guardWord16(PH);
    /* E[k] = (512 + g) - 2 X[k] as one multiply-add on the FMA pipe (minusTwo = -2 at run time) */
#if defined(__CUDA_ARCH__)
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(E[0]) : "r"(X[0]), "r"(minusTwo), "r"(kSoftFull + g));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(E[1]) : "r"(X[1]), "r"(minusTwo), "r"(kSoftFull + g));
#else
    E[0] = X[0] * minusTwo + (kSoftFull + g);
    E[1] = X[1] * minusTwo + (kSoftFull + g);
#endif
    E[2] = 2u * g - E[1];
    E[3] = 2u * g - E[0];
}

/* One trellis step at compile-time phase PH; T0/T1 receive the 64 decision bits (1 = predecessor j+32 won). */
template <class Code, int PH>
CED_HD void acsStep16(uint32_t (&R)[32], const uint32_t (&X)[4], const uint32_t (&E)[4], uint32_t minusOne,
                      uint32_t &T0, uint32_t &T1)
{
    constexpr int q = 5 - PH;
    uint32_t t0 = 0, t1 = 0;
    if constexpr (q >= 1) {
        constexpr int rb = q - 1;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if ((r >> rb) & 1)
                continue;
            const int rh = r | (1 << rb);
            const uint32_t k = Code::cls((uint32_t)r << 1, PH);
            const uint32_t d = X[k], dc = X[k ^ 3u];
            const uint32_t lo = R[r], hi = R[rh];
            const uint32_t a0 = lo + d, a1 = hi + dc;   /* successor 2j   (src/viterbiDecoderButterflyk1.c:109-110) */
            const uint32_t b0 = lo + dc, b1 = hi + d;   /* successor 2j+1 (:113-114) */
            const uint32_t delta = subOnFma(hi, lo, minusOne);
            const uint32_t ma = signMask16(delta + E[k]);       /* FFFF: keep the lower predecessor */
            const uint32_t mb = signMask16(delta + E[k ^ 3u]);
            R[r] = sel(ma, a0, a1);
            R[rh] = sel(mb, b0, b1);
            const uint32_t ca = 0x00010001u << (r & 15), cb = 0x00010001u << (rh & 15);
            if (r < 16) t0 |= ~ma & ca; else t1 |= ~ma & ca;
            if (rh < 16) t0 |= ~mb & cb; else t1 |= ~mb & cb;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const uint32_t k = Code::cls((uint32_t)r << 1, PH);
            const uint32_t self = R[r] + X[k];
            const uint32_t swapped = prmt(R[r], 0u, 0x1032u);
            const uint32_t cross = swapped + X[k ^ 3u];
            const uint32_t m = signMask16(subOnFma(swapped, R[r], minusOne) + E[k]);
            R[r] = sel(m, self, cross);
            const uint32_t c = 0x00010001u << (r & 15);
            if (r < 16) t0 |= ~m & c; else t1 |= ~m & c;
        }
        t0 ^= 0xFFFF0000u;
        t1 ^= 0xFFFF0000u;
    }
    T0 = t0;
    T1 = t1;
}

CED_HD uint32_t halfMin(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b); /* one VIMNMX.U16x2 */
#else
    const uint32_t lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
#endif
}

CED_HD void renorm16(uint32_t (&R)[32])
{
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        m[i] = halfMin(R[i], R[i + 16]);
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; i++)
            m[i] = halfMin(m[i], m[i + w]);
    uint32_t v = halfMin(m[0], prmt(m[0], 0u, 0x1032u));
#pragma unroll
    for (int r = 0; r < 32; r++)
        R[r] -= v;
}

} // namespace ced
