/*
 * trellis_fused.cuh -- the traceback that runs INSIDE the forward kernel (k7FusedKernel, decode_fused.cuh).
 *
 * The reference keeps every survivor decision of a packet and walks them back once, from state 0, after the last
 * step (src/viterbiDecoderButterflyk1.c:185-187 store, :200-256 walk).  On the GPU that is 8 bytes per frame-step
 * written to and read back from HBM (2.15 GB each way for 2^16 frames x 4096 bits) plus a second kernel.  Here a
 * frame keeps only its newest kRingSteps = E + D steps of decisions in an L2-resident ring, and after every E = 96
 * steps the thread that just produced them walks that window back:
 *
 *   chunk pass (segment cc, not the last one; now = 96 (cc + 1)):
 *      start in the best-metric state at time `now` (any survivor would do; this one merges soonest), walk D = 72
 *      steps without output                                                                      (acquisition)
 *      remember the state reached, s*(cc), at time now - D
 *      walk the E steps [96 cc - D, now - D) and emit their 96 bits                              (emission)
 *      the state reached at time 96 cc - D must be s*(cc - 1), the state the previous chunk pass started ITS
 *      emission from; if it is not, the frame is flagged
 *   final pass (last segment): start in state 0 at time T exactly like the reference (:205), drop the S tail steps,
 *      emit everything not emitted yet, and make the same check.
 *
 * Exactness.  The final pass walks the true path.  If its check passes, the previous pass started its emission on
 * the true path, so it emitted true bits and arrived at a true state for ITS check, and so on down to step 0: a
 * frame with no flag has exactly the reference's output.  A flagged frame (survivors from state 0 and from the true
 * path had not merged within D steps) is decoded again by the two-kernel path with the full survivor store.  Nothing is
 * approximated.  Flagged frames of 4096 bits on a BSC (tests/hostsim, 1500 frames per point; D = 48 / 72 / 96, best
 * start): p = 0.0377 (5 dB): 0.07 % / 0 / 0;  p = 0.06: 3.7 % / 0.13 % / 0;  p = 0.08 (3 dB, decoded BER 3e-2): 26 % /
 * 5 % / 0.9 %;  p = 0.12: 97 % / 70 % / 35 % (starting in state 0 instead: 4-8 times as many).
 *
 * Ring addressing: pairs of steps (one uint4 = 2 x 64 decision bits) in blocks of 12 pairs = 24 steps; the ring holds
 * kRingBlocks = 7 blocks; global block B lives in ring block B mod 7.  All windows are whole blocks because E and D
 * are multiples of 24 and segments start at multiples of 96 (phase 0).
 */
#pragma once
#include "trellis_swar.cuh"

namespace ced {

/* E = steps between passes = emission length (a multiple of the 96-step renormalisation period), D = acquisition
 * depth (a multiple of 24) */
template <int E_, int D_>
struct FusedGeom {
    static constexpr int E = E_, D = D_;
    static constexpr int kRingSteps = E + D;
    static constexpr int kRingPairs = kRingSteps / 2;     /* uint4 per frame */
    static constexpr int kRingBlocks = kRingSteps / 24;
    static_assert(E % 96 == 0 && D % 24 == 0, "fused geometry: passes at renormalisation boundaries, whole 24-step blocks");
};
using DefaultFusedGeom = FusedGeom<96, 72>;
constexpr int kFusedE = DefaultFusedGeom::E, kFusedD = DefaultFusedGeom::D;
constexpr int kRingPairs = DefaultFusedGeom::kRingPairs, kRingBlocks = DefaultFusedGeom::kRingBlocks;

/*
 * One backward step inside a 24-step block (TAU = step index within the block, phase TAU mod 6), written for the
 * instruction count -- the in-kernel traceback competes with the ACS for issue slots:
 *   w  = hi5 ? w1 : w0          the half of the 64 decisions b points into; hi5 = bit 5 of b, which only the one
 *                               step in six with pair bit 5 can change, so it lives in a predicate (SEL, off the chain)
 *   t  = rotr(w, b - qb)        the decision of position b lands on bit qb (one SHF; the count is taken mod 32)
 *   b  = bitselect(b, t, 1<<qb) the predecessor: pair bit replaced by the decision (one LOP3)
 * (tracebackStep in trellis_swar.cuh: the same result in about three times as many instructions.)
 */
template <class Lay, int TAU>
CED_HD void backStep(uint32_t &b, bool &hi5, uint32_t w0, uint32_t w1)
{
    constexpr uint32_t qb = (uint32_t)Lay::pairBit(TAU % 6);
    const uint32_t w = hi5 ? w1 : w0;
    const uint32_t t = rotr32(w, b - qb);
    b = (b & ~(1u << qb)) | (t & (1u << qb));
    if (qb == 5u)
        hi5 = (b & 32u) != 0u;
}

/* steps 2M+1 and 2M of the block from one ring row */
template <class Lay, int M>
CED_HD void backPair(uint32_t &b, bool &hi5, const uint4 &row)
{
    backStep<Lay, 2 * M + 1>(b, hi5, row.z, row.w);
    backStep<Lay, 2 * M>(b, hi5, row.x, row.y);
}

/*
 * Walk the 24 steps of one block backwards; r[i] = decisions of pair 12 blk + 11 - i (highest first).  The decoded
 * bits need no per-step work: the state after step t is the last six input bits (u_t in bit 0 ... u_{t-5} in bit 5,
 * src/viterbiDecoderButterflyk1.c:244-252), and at the top of every 6-step group the next phase is 0, where position
 * == state; so the six bits of steps 6k+5 .. 6k are Lay::toPosition(b) there, and in the MSb-first 24-bit word of the
 * block (step 0 = bit 23, :249) they sit at bits 18-6k .. 23-6k in exactly that order.
 */
template <class Lay, bool EMIT>
CED_HD uint32_t walkBlockBits(uint32_t &b, const uint4 (&r)[12])
{
    bool hi5 = (b & 32u) != 0u;
    uint32_t v = 0;
    if (EMIT) v = Lay::toPosition(b);                      /* steps 23..18 -> bits 0..5 */
    backPair<Lay, 11>(b, hi5, r[0]);
    backPair<Lay, 10>(b, hi5, r[1]);
    backPair<Lay, 9>(b, hi5, r[2]);
    if (EMIT) v |= Lay::toPosition(b) << 6;                /* steps 17..12 */
    backPair<Lay, 8>(b, hi5, r[3]);
    backPair<Lay, 7>(b, hi5, r[4]);
    backPair<Lay, 6>(b, hi5, r[5]);
    if (EMIT) v |= Lay::toPosition(b) << 12;               /* steps 11..6 */
    backPair<Lay, 5>(b, hi5, r[6]);
    backPair<Lay, 4>(b, hi5, r[7]);
    backPair<Lay, 3>(b, hi5, r[8]);
    if (EMIT) v |= Lay::toPosition(b) << 18;               /* steps 5..0 */
    backPair<Lay, 2>(b, hi5, r[9]);
    backPair<Lay, 1>(b, hi5, r[10]);
    backPair<Lay, 0>(b, hi5, r[11]);
    return v;
}

template <class Lay>
CED_HD void walkBlock(uint32_t &b, const uint4 (&r)[12], uint32_t &o2, uint32_t &o1, uint32_t &o0)
{
    const uint32_t v = walkBlockBits<Lay, true>(b, r);
    o0 = v >> 16;           /* steps 24 blk .. + 7 */
    o1 = (v >> 8) & 0xFFu;  /* steps 24 blk + 8 .. + 15 */
    o2 = v & 0xFFu;         /* steps 24 blk + 16 .. + 23 */
}

/*
 * Chunk pass after segment cc (cc >= 0, not the frame's last segment).  loadBlock(blk, r) fetches block blk of the
 * frame's decisions, storeBytes(blk, o0, o1, o2) writes output bytes 3 blk .. 3 blk + 2.  `expect` carries s* from
 * pass to pass.  Returns false if the frame must be flagged.
 */
template <class Lay, class Geo = DefaultFusedGeom, class LoadBlock, class StoreBytes>
CED_HD bool fusedChunkPass(int cc, uint32_t startB, uint32_t &expect, LoadBlock loadBlock, StoreBytes storeBytes)
{
    constexpr int kFusedE = Geo::E, kFusedD = Geo::D;
    uint32_t b = startB;                              /* any survivor will do; the best-metric one merges soonest */
    int blk = (kFusedE / 24) * (cc + 1) - 1;
    uint4 r[12];
    uint32_t o2, o1, o0;
#pragma unroll 1
    for (int i = 0; i < kFusedD / 24; i++, blk--) {
        loadBlock(blk, r);
        walkBlockBits<Lay, false>(b, r);
    }
    const uint32_t start = b;                         /* s*(cc): state at time 96 (cc + 1) - D */
    const int lo = blk - (kFusedE / 24 - 1) > 0 ? blk - (kFusedE / 24 - 1) : 0;
#pragma unroll 1
    for (; blk >= lo; blk--) {
        loadBlock(blk, r);
        walkBlock<Lay>(b, r, o2, o1, o0);
        storeBytes(blk, o0, o1, o2);
    }
    const bool ok = kFusedE * cc - kFusedD <= 0 || b == expect;   /* nothing was emitted below step 0 */
    expect = start;
    return ok;
}

/*
 * Final pass (segment cc is the frame's last; T steps in all, L = T - S information bits).  loadPair(m) fetches
 * one pair of steps for the ragged top, storeByte(i, v) writes output byte i.
 */
template <class Lay, class Geo = DefaultFusedGeom, class LoadPair, class LoadBlock, class StoreByte, class StoreBytes>
CED_HD bool fusedFinalPass(int cc, int T, int tailSteps, uint32_t expect, LoadPair loadPair, LoadBlock loadBlock,
                           StoreByte storeByte, StoreBytes storeBytes)
{
    constexpr int kFusedE = Geo::E, kFusedD = Geo::D;
    const int L = T - tailSteps;
    const int blocks24 = L / 24;
    uint32_t b = 0;                                   /* terminated frame: state 0 (src/viterbiDecoderButterflyk1.c:205) */
    int ph = (T - 1) % 6;
    uint32_t acc = 0;
#pragma unroll 1
    for (int m = T / 2 - 1; m >= blocks24 * 12; m--) {
        const uint4 w = loadPair(m);
        const int t = 2 * m;
        const uint32_t b1 = tracebackStep<Lay>(b, w.z, w.w, ph);
        ph = ph ? ph - 1 : 5;
        const uint32_t b0 = tracebackStep<Lay>(b, w.x, w.y, ph);
        ph = ph ? ph - 1 : 5;
        if (t < L) {                                  /* the S tail steps carry no output (:208-223) */
            acc = (acc >> 2) | (b1 << 6) | (b0 << 7); /* first visited (t % 8 == 7) ends as the LSb (:249) */
            if ((t & 7) == 0) {
                storeByte(t >> 3, acc);
                acc = 0;
            }
        }
    }
    const int lo = (kFusedE / 24) * cc - kFusedD / 24 > 0 ? (kFusedE / 24) * cc - kFusedD / 24 : 0;
    uint4 r[12];
    uint32_t o2, o1, o0;
#pragma unroll 1
    for (int blk = blocks24 - 1; blk >= lo; blk--) {
        loadBlock(blk, r);
        walkBlock<Lay>(b, r, o2, o1, o0);
        storeBytes(blk, o0, o1, o2);
    }
    return kFusedE * cc - kFusedD <= 0 || b == expect;
}

} // namespace ced
