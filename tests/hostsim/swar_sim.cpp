// TEST INFRASTRUCTURE ONLY: runs the exact SWAR step / traceback functions of
// convolutionalencdec_b200/csrc/trellis_swar.cuh for ONE frame on the host so the
// rotating-label algebra can be checked against the oracle without a GPU.
// It is never linked into the product library.
#include "trellis_swar.cuh"
#include "trellis_swar16.cuh"
#include <vector>

using Code = ced::DefaultK7;

template <int PH>
static void stepPhase(uint32_t (&R)[16], uint32_t rx, uint32_t &t0, uint32_t &t1)
{
    uint32_t X[4], E[4];
    for (uint32_t k = 0; k < 4; k++)
        X[k] = Code::bmWord(PH, rx, k);
    for (uint32_t k = 0; k < 4; k++)
        E[k] = X[k ^ 3u] - X[k] + ced::guardWord(PH);
    ced::acsStep<Code, PH>(R, X, E, 0xFFFFFFFFu, t0, t1);
}

extern "C" int swar_sim_decode(const uint8_t *segs, int T, uint8_t *out, uint32_t *survOut, uint8_t *maxMetric,
                               int renormPeriod)
{
    uint32_t R[16];
    ced::initMetrics(R);
    std::vector<uint32_t> surv(2 * (size_t)T);
    uint8_t mx = 0;
    for (int t = 0; t < T; t++) {
        uint32_t t0 = 0, t1 = 0, rx = segs[t];
        switch (t % 6) {
        case 0: stepPhase<0>(R, rx, t0, t1); break;
        case 1: stepPhase<1>(R, rx, t0, t1); break;
        case 2: stepPhase<2>(R, rx, t0, t1); break;
        case 3: stepPhase<3>(R, rx, t0, t1); break;
        case 4: stepPhase<4>(R, rx, t0, t1); break;
        default: stepPhase<5>(R, rx, t0, t1); break;
        }
        surv[2 * t] = t0;
        surv[2 * t + 1] = t1;
        for (int r = 0; r < 16; r++)
            for (int l = 0; l < 4; l++) {
                uint8_t v = (R[r] >> (8 * l)) & 0xFF;
                if (v > mx) mx = v;
            }
        if ((t + 1) % renormPeriod == 0)
            ced::renorm(R);
    }
    if (maxMetric) *maxMetric = mx;
    if (survOut)
        for (size_t i = 0; i < surv.size(); i++) survOut[i] = surv[i];
    const int L = T - 6;
    uint32_t b = 0;
    for (int i = 0; i < (L + 7) / 8; i++) out[i] = 0;
    for (int t = T - 1; t >= 0; t--) {
        uint32_t bit = ced::tracebackStep(b, surv[2 * t], surv[2 * t + 1], t % 6);
        if (t < L)
            out[t / 8] |= (uint8_t)(bit << (7 - (t % 8)));
    }
    return (L + 7) / 8;
}

// The windowed procedure of ced_decode_window_batch for one stream (slices of callSegs, a multiple of 96;
// traceback from the best-metric position after every slice, from state 0 after the last one).
extern "C" int swar_sim_window(const uint8_t *segs, int T, int callSegs, int depth, uint8_t *out)
{
    uint32_t R[16];
    ced::initMetrics(R);
    std::vector<uint32_t> surv(2 * (size_t)T);
    const int L = T - 6;
    for (int i = 0; i < (L + 7) / 8; i++) out[i] = 0;
    int emitted = 0;
    for (int t = 0; t < T; t++) {
        uint32_t t0 = 0, t1 = 0, rx = segs[t];
        switch (t % 6) {
        case 0: stepPhase<0>(R, rx, t0, t1); break;
        case 1: stepPhase<1>(R, rx, t0, t1); break;
        case 2: stepPhase<2>(R, rx, t0, t1); break;
        case 3: stepPhase<3>(R, rx, t0, t1); break;
        case 4: stepPhase<4>(R, rx, t0, t1); break;
        default: stepPhase<5>(R, rx, t0, t1); break;
        }
        surv[2 * t] = t0;
        surv[2 * t + 1] = t1;
        const int P = t + 1;
        const bool lastCall = P == T;
        if (P % 96 == 0 && !lastCall)
            ced::renorm(R);
        if (!lastCall && !(P % callSegs == 0 && T - P > 0))
            continue;
        if (!lastCall && T - (P - callSegs) <= callSegs)
            continue; /* the remainder belongs to the last call */
        uint32_t b = lastCall ? 0u : ced::bestPositionB(R);
        const int hi = lastCall ? L : P - depth;
        for (int u = P - 1; u >= emitted; u--) {
            uint32_t bit = ced::tracebackStep(b, surv[2 * u], surv[2 * u + 1], u % 6);
            if (u < hi)
                out[u / 8] |= (uint8_t)(bit << (7 - (u % 8)));
        }
        if (hi > emitted) emitted = hi;
    }
    return emitted;
}

// Same frame decode through the run-time table path (acsStepTable + buildStepTable) for any symmetric K=7 code
// with n = 2 or 3 generators.
template <int PH, int N>
static void stepTable(uint32_t (&R)[16], const ced::Word2 *table, uint32_t rx, uint32_t &t0, uint32_t &t1)
{
    ced::acsStepTable<PH, N>(R, table + PH * 16 * (1 << N) + (int)(rx & ((1u << N) - 1u)), 0xFFFFFFFFu, t0, t1);
}

template <int N>
static int decodeRuntime(const uint32_t *gens, const uint8_t *segs, int T, uint8_t *out)
{
    std::vector<ced::Word2> table((size_t)ced::RuntimeK7<N>::kTableEntries);
    ced::buildStepTable(ced::makeK7Taps(N, gens), table.data());
    uint32_t R[16];
    ced::initMetrics(R);
    std::vector<uint32_t> surv(2 * (size_t)T);
    uint8_t mx = 0;
    for (int t = 0; t < T; t++) {
        uint32_t t0 = 0, t1 = 0, rx = segs[t];
        switch (t % 6) {
        case 0: stepTable<0, N>(R, table.data(), rx, t0, t1); break;
        case 1: stepTable<1, N>(R, table.data(), rx, t0, t1); break;
        case 2: stepTable<2, N>(R, table.data(), rx, t0, t1); break;
        case 3: stepTable<3, N>(R, table.data(), rx, t0, t1); break;
        case 4: stepTable<4, N>(R, table.data(), rx, t0, t1); break;
        default: stepTable<5, N>(R, table.data(), rx, t0, t1); break;
        }
        surv[2 * t] = t0;
        surv[2 * t + 1] = t1;
        for (int r = 0; r < 16; r++)
            for (int l = 0; l < 4; l++) {
                uint8_t v = (R[r] >> (8 * l)) & 0xFF;
                if (v > mx) mx = v;
            }
        if ((t + 1) % ced::RuntimeK7<N>::kRenormPeriod == 0)
            ced::renorm(R);
    }
    const int L = T - 6;
    uint32_t b = 0;
    for (int i = 0; i < (L + 7) / 8; i++) out[i] = 0;
    for (int t = T - 1; t >= 0; t--) {
        uint32_t bit = ced::tracebackStep(b, surv[2 * t], surv[2 * t + 1], t % 6);
        if (t < L)
            out[t / 8] |= (uint8_t)(bit << (7 - (t % 8)));
    }
    return mx; /* largest metric seen: must stay below 128 - n for the guard-bit compare */
}

extern "C" int swar_sim_decode_rt(int n, const uint32_t *gens, const uint8_t *segs, int T, uint8_t *out)
{
    return n == 3 ? decodeRuntime<3>(gens, segs, T, out) : decodeRuntime<2>(gens, segs, T, out);
}

// decision of state s after step t from the packed words (for comparing with the oracle's survivors)
extern "C" int swar_sim_decision(const uint32_t *surv, int t, int s)
{
    uint32_t p = ced::rotr6((uint32_t)s, (t + 1) % 6);
    uint32_t word = surv[2 * t + (p >> 5)];
    uint32_t bit = 8 * (p & 3) + ((p >> 2) & 7);
    return (word >> bit) & 1;
}

// Soft-decision decode of one frame with the 16-bit-lane step functions of trellis_swar16.cuh (what
// k7SoftForwardKernel runs per thread) and the generic traceback step for the Lanes16 survivor layout.
template <int PH>
static void softPhase(uint32_t (&R)[32], uint32_t W, uint32_t &t0, uint32_t &t1)
{
    uint32_t X[4], E[4];
    ced::softBranchWords<Code, PH>(W, 0xFFFFFFFEu, X, E);
    ced::acsStep16<Code, PH>(R, X, E, 0xFFFFFFFFu, t0, t1);
}

extern "C" int swar_sim_decode_soft(const int8_t *soft, int T, uint8_t *out, uint32_t *maxMetric)
{
    uint32_t R[32];
    ced::initMetrics16(R);
    std::vector<uint32_t> surv(2 * (size_t)T);
    uint32_t mx = 0;
    for (int t = 0; t < T; t++) {
        uint32_t t0 = 0, t1 = 0;
        const uint32_t W = ced::softWord(soft[2 * t], soft[2 * t + 1]);
        switch (t % 6) {
        case 0: softPhase<0>(R, W, t0, t1); break;
        case 1: softPhase<1>(R, W, t0, t1); break;
        case 2: softPhase<2>(R, W, t0, t1); break;
        case 3: softPhase<3>(R, W, t0, t1); break;
        case 4: softPhase<4>(R, W, t0, t1); break;
        default: softPhase<5>(R, W, t0, t1); break;
        }
        surv[2 * t] = t0;
        surv[2 * t + 1] = t1;
        for (int r = 0; r < 32; r++) {
            if ((R[r] & 0xFFFFu) > mx) mx = R[r] & 0xFFFFu;
            if ((R[r] >> 16) > mx) mx = R[r] >> 16;
        }
        if ((t + 1) % ced::kSoftRenormPeriod == 0)
            ced::renorm16(R);
    }
    if (maxMetric) *maxMetric = mx;
    const int L = T - 6;
    uint32_t b = 0;
    for (int i = 0; i < (L + 7) / 8; i++) out[i] = 0;
    for (int t = T - 1; t >= 0; t--) {
        uint32_t bit = ced::tracebackStep<ced::Lanes16>(b, surv[2 * t], surv[2 * t + 1], t % 6);
        if (t < L)
            out[t / 8] |= (uint8_t)(bit << (7 - (t % 8)));
    }
    return (L + 7) / 8;
}

// The in-kernel traceback of k7FusedKernel for one frame (trellis_fused.cuh): decisions live in a ring of
// kRingPairs step pairs only; after every 96 steps the chunk pass walks the window, the final pass ends the frame.
// *flagged = 1 if any pass failed its check (the kernel then hands the frame to the two-kernel path).
#include "trellis_fused.cuh"

template <class Geo>
static int simFused(const uint8_t *segs, int T, uint8_t *out, int *flagged, int bestStart)
{
    uint32_t R[16];
    ced::initMetrics(R);
    std::vector<uint4> ring(Geo::kRingPairs);
    const int L = T - 6;
    for (int i = 0; i < (L + 7) / 8; i++) out[i] = 0xEE; /* every byte must be written by a pass */
    uint32_t expect = 0;
    bool ok = true;
    auto loadBlock = [&](int blk, uint4 (&r)[12]) {
        for (int i = 0; i < 12; i++)
            r[i] = ring[(size_t)((12 * blk + 11 - i) % Geo::kRingPairs)];
    };
    auto storeBytes = [&](int blk, uint32_t o0, uint32_t o1, uint32_t o2) {
        out[3 * blk] = (uint8_t)o0;
        out[3 * blk + 1] = (uint8_t)o1;
        out[3 * blk + 2] = (uint8_t)o2;
    };
    for (int t = 0; t < T; t++) {
        uint32_t t0 = 0, t1 = 0, rx = segs[t];
        switch (t % 6) {
        case 0: stepPhase<0>(R, rx, t0, t1); break;
        case 1: stepPhase<1>(R, rx, t0, t1); break;
        case 2: stepPhase<2>(R, rx, t0, t1); break;
        case 3: stepPhase<3>(R, rx, t0, t1); break;
        case 4: stepPhase<4>(R, rx, t0, t1); break;
        default: stepPhase<5>(R, rx, t0, t1); break;
        }
        uint4 &row = ring[(size_t)((t / 2) % Geo::kRingPairs)];
        if (t & 1) { row.z = t0; row.w = t1; } else { row.x = t0; row.y = t1; }
        const int now = t + 1;
        if (now == T) {
            const int cc = (T - 1) / Geo::E;
            ok &= ced::fusedFinalPass<ced::Lanes8, Geo>(
                cc, T, 6, expect, [&](int m) { return ring[(size_t)(m % Geo::kRingPairs)]; }, loadBlock,
                [&](int i, uint32_t v) { out[i] = (uint8_t)v; }, storeBytes);
        } else if (now % 96 == 0) {
            ced::renorm(R);
            if (now % Geo::E == 0)
                ok &= ced::fusedChunkPass<ced::Lanes8, Geo>(now / Geo::E - 1, bestStart ? ced::bestPositionB(R) : 0u, expect,
                                                            loadBlock, storeBytes);
        }
    }
    *flagged = ok ? 0 : 1;
    return (L + 7) / 8;
}

extern "C" int swar_sim_fused(const uint8_t *segs, int T, uint8_t *out, int *flagged, int bestStart, int E, int D)
{
    if (E == 96 && D == 72) return simFused<ced::FusedGeom<96, 72>>(segs, T, out, flagged, bestStart);
    if (E == 96 && D == 48) return simFused<ced::FusedGeom<96, 48>>(segs, T, out, flagged, bestStart);
    if (E == 192 && D == 72) return simFused<ced::FusedGeom<192, 72>>(segs, T, out, flagged, bestStart);
    if (E == 192 && D == 96) return simFused<ced::FusedGeom<192, 96>>(segs, T, out, flagged, bestStart);
    if (E == 384 && D == 96) return simFused<ced::FusedGeom<384, 96>>(segs, T, out, flagged, bestStart);
    if (E == 384 && D == 120) return simFused<ced::FusedGeom<384, 120>>(segs, T, out, flagged, bestStart);
    return -1;
}

// Any k = 1 code with 4 .. 256 states and n = 2 / 3 through the table-driven step functions of swar_generic.cuh
// (what genForwardKernel / genTracebackKernel run per thread): returns the largest metric seen, -1 = not a code it takes.
#include "swar_generic.cuh"

template <int S>
static int decodeGen(const ced::GenCode &gc, const uint8_t *segs, int T, uint8_t *out)
{
    using G = ced::GenGeom<S>;
    const int V = 1 << gc.n;
    std::vector<uint8_t> table((size_t)G::tableBytes(V) + 16);
    ced::buildGenTable<S>(gc, table.data());
    uint32_t R[G::kRegs];
    ced::genInitMetrics<S>(R, gc.n);
    std::vector<uint32_t> surv((size_t)G::kWords * (size_t)T);
    int mx = 0;
    for (int t = 0; t < T; t++) {
        uint32_t Tw[G::kWords];
        const int ph = t % S;
        const uint32_t rx = segs[t] & (uint32_t)(V - 1);
        const int q = S - 1 - ph;
        const uint8_t *tab = table.data() + G::phaseBase(ph, V) + rx * (q >= 2 ? 32 : 16);
        const int stride = (q >= 2 ? 32 : 16) * V;
        switch (ph) {
        case 0: ced::genStep<S, 0>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        case 1: ced::genStep<S, 1>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        case 2: if constexpr (S > 2) ced::genStep<S, 2>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        case 3: if constexpr (S > 3) ced::genStep<S, 3>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        case 4: if constexpr (S > 4) ced::genStep<S, 4>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        case 5: if constexpr (S > 5) ced::genStep<S, 5>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        case 6: if constexpr (S > 6) ced::genStep<S, 6>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        default: if constexpr (S > 7) ced::genStep<S, 7>(R, tab, stride, 0xFFFFFFFFu, Tw); break;
        }
        for (int w = 0; w < G::kWords; w++)
            surv[(size_t)t * G::kWords + w] = Tw[w];
        for (int r = 0; r < G::kRegs; r++)
            for (int l = 0; l < 4; l++) {
                if (4 * r + l >= G::kStates) continue;
                const int v = (R[r] >> (8 * l)) & 0xFF;
                if (v > mx) mx = v;
            }
        if ((t + 1) % G::kRenorm == 0)
            ced::genRenorm<S>(R);
    }
    const int L = T - S;
    for (int i = 0; i < (L + 7) / 8; i++) out[i] = 0;
    uint32_t p = 0;
    for (int t = T - 1; t >= 0; t--) {
        const uint32_t bit = ced::genTracebackStep<S>(p, &surv[(size_t)t * G::kWords], t);
        if (t < L)
            out[t / 8] |= (uint8_t)(bit << (7 - (t % 8)));
    }
    return mx;
}

extern "C" int swar_sim_decode_gen(int K, int n, const uint32_t *gens, const uint8_t *segs, int T, uint8_t *out)
{
    ced::GenCode gc;
    gc.S = K - 1;
    gc.n = n;
    if (n < 2 || n > 3) return -1;
    for (int i = 0; i < 3; i++) {
        gc.tap[i] = 0;
        if (i < n)
            for (int b = 0; b < K; b++)   /* taps with bit 0 on the newest input bit (src/convEncode.c:163-175) */
                gc.tap[i] |= ((gens[i] >> b) & 1u) << (K - 1 - b);
    }
    switch (gc.S) {
    case 2: return decodeGen<2>(gc, segs, T, out);
    case 3: return decodeGen<3>(gc, segs, T, out);
    case 4: return decodeGen<4>(gc, segs, T, out);
    case 5: return decodeGen<5>(gc, segs, T, out);
    case 6: return decodeGen<6>(gc, segs, T, out);
    case 7: return decodeGen<7>(gc, segs, T, out);
    case 8: return decodeGen<8>(gc, segs, T, out);
    default: return -1;
    }
}

// k = 2 codes through the radix-4 step functions of swar_radix4.cuh (what genForwardKernel<R4Policy> runs per thread):
// returns the largest metric seen, -1 = not a code it takes.  out: (T - S) * 2 / 8 bytes.
#include "swar_radix4.cuh"

template <int S>
static int decodeR4(const ced::R4Code &c, const uint8_t *segs, int T, uint8_t *out)
{
    using G = ced::R4Geom<S>;
    const int V = 1 << c.n;
    std::vector<uint8_t> table((size_t)G::tableBytes(V) + 16);
    ced::buildR4Table<S>(c, table.data());
    uint32_t R[G::kRegs];
    ced::r4InitMetrics<S>(R, c.n);
    std::vector<uint32_t> surv((size_t)G::kWords * (size_t)T);
    int mx = 0;
    for (int t = 0; t < T; t++) {
        uint32_t Tw[G::kWords];
        const int ph = t % S;
        const uint32_t rx = segs[t] & (uint32_t)(V - 1);
        const uint8_t *tab = table.data() + G::phaseBase(ph, V) + G::rxOffset(ph, rx * 32u);
        switch (ph) {
        case 0: ced::R4Policy<S>::template step<0>(R, tab, V, 0xFFFFFFFFu, Tw); break;
        case 1: if constexpr (S > 1) ced::R4Policy<S>::template step<1>(R, tab, V, 0xFFFFFFFFu, Tw); break;
        case 2: if constexpr (S > 2) ced::R4Policy<S>::template step<2>(R, tab, V, 0xFFFFFFFFu, Tw); break;
        default: if constexpr (S > 3) ced::R4Policy<S>::template step<3>(R, tab, V, 0xFFFFFFFFu, Tw); break;
        }
        for (int w = 0; w < G::kWords; w++)
            surv[(size_t)t * G::kWords + w] = Tw[w];
        for (int r = 0; r < G::kRegs; r++)
            for (int l = 0; l < 4; l++) {
                const int v = (R[r] >> (8 * l)) & 0xFF;
                if (v > mx) mx = v;
            }
        if ((t + 1) % G::kRenorm == 0)
            ced::r4Renorm<S>(R);
    }
    const int L = T - S;                         /* information segments, 2 bits each */
    for (int i = 0; i < (2 * L + 7) / 8; i++) out[i] = 0;
    uint32_t p = 0;
    for (int t = T - 1; t >= 0; t--) {
        const uint32_t two = ced::r4TracebackStep<S>(p, &surv[(size_t)t * G::kWords], t);
        if (t < L)
            out[t / 4] |= (uint8_t)(two << (6 - 2 * (t % 4)));
    }
    return mx;
}

extern "C" int swar_sim_decode_r4(int K, int n, const uint32_t *gens, const uint8_t *segs, int T, uint8_t *out)
{
    ced::R4Code c;
    c.S = K - 1;
    c.n = n;
    if (n < 2 || n > 3 || K < 2 || K > 5) return -1;
    for (int i = 0; i < 3; i++) {
        c.tap[i] = 0;
        if (i < n)
            for (int b = 0; b < 2 * K; b++)   /* taps with bit 0 on the newest input bit (src/convEncode.c:163-175, k*K bits) */
                c.tap[i] |= ((gens[i] >> b) & 1u) << (2 * K - 1 - b);
    }
    switch (c.S) {
    case 1: return decodeR4<1>(c, segs, T, out);
    case 2: return decodeR4<2>(c, segs, T, out);
    case 3: return decodeR4<3>(c, segs, T, out);
    default: return decodeR4<4>(c, segs, T, out);
    }
}
