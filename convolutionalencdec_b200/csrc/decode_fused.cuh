/*
 * decode_fused.cuh -- ONE kernel for the whole decode of a batch: forward ACS (exactly k7ForwardKernel's
 * arithmetic, decode_batch.cuh / trellis_swar.cuh) with the traceback running inside it on an L2-resident ring of
 * the newest 168 steps of decisions per frame (trellis_fused.cuh), instead of 8 bytes per frame-step spilled to
 * HBM and read back by a second kernel.
 *
 *   reference                                       here
 *   src/viterbiDecoderButterflyk1.c:85-196          forward steps, 96 per segment, decisions -> ring (st.global, L2)
 *   :200-256 (one walk from state 0 at the end)     chunk pass after every segment, final pass at the end
 *
 * What it buys (north_star (3) "survivor-decision store ... in shared memory or L2"): per decode of 2^16 frames
 * x 4096 bits the two-kernel path moves 4.7 GB through HBM (15.6 x the algorithmic bytes) and needs several calls
 * in flight to hide its HBM-bound second kernel behind the issue-bound first one; the fused kernel issues the
 * traceback's instructions (ALU pipe, ~1.75 backward steps per forward step) in the same instruction stream, so a
 * single ced_decode_batch call runs at the rate the two-kernel path only reaches with three calls in flight.
 *
 * Work distribution: the persistent unit scheduler of k7ForwardKernel, cohort-major so that the rings that are hot
 * at any time belong to at most two cohorts of kCohortGroups groups (ring slot = group mod 2 * kCohortGroups).
 * Frames that fail a pass check are appended to `list` (atomic counter) and decoded again by the two-kernel path.
 */
#pragma once
#include "decode_batch.cuh"
#include "trellis_fused.cuh"

namespace ced {

constexpr int kCohortGroups = 2048;   /* 2^16 frames; two cohorts of rings = 176 MB allocated, ~88 MB hot */

struct FusedArgs {
    uint4 *ring;         /* [ringSlots][kRingPairs][32]  newest decisions of the groups in flight           */
    uint32_t *expect;    /* [groups][32]  s* handed from pass to pass (and from unit to unit)               */
    unsigned int *flag;  /* [nFrames] 0 / 1, zeroed before the launch                                      */
    int *list;           /* [nFrames] frames to decode again                                               */
    int *count;          /* number of entries in list, zeroed before the launch                            */
    uint8_t *out;
    size_t outStride;
    int ringSlots;
};

template <class Code, class Fmt, bool ALIGNED, class Geo = DefaultFusedGeom>
__global__ void __launch_bounds__(kFwdThreads, 4)
k7FusedKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, BmTable table, FwdSched sched,
              int chunksPerUnit, FusedArgs fa)
{
    using G = TileGeom<Fmt, ALIGNED>;
    constexpr int kChunk = G::kChunk, kPitch = G::kPitch;
    constexpr int kFusedE = Geo::E, kRingPairs = Geo::kRingPairs, kRingBlocks = Geo::kRingBlocks;
    constexpr int kSeg = 96;                                 /* renormalisation period = forward segment */
    constexpr int kSegsPerTile = kChunk / kSeg;              /* 1 (byte format) or 2 (packed) */
    static_assert(!Code::kRuntime && Code::kRenormPeriod == kSeg, "fused kernel: compile-time codes, 96-step renorm");
    __shared__ uint4 sBm[6 * 4 * 2];
    __shared__ __align__(16) uint8_t sTile[kFwdThreads / 32][32 * kPitch];

    if (threadIdx.x < 48)
        sBm[threadIdx.x] = table.x[threadIdx.x];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile = sTile[warp];
    const uint8_t *bmBase = reinterpret_cast<const uint8_t *>(sBm);
    const uint32_t minusOne = table.minusOne;
    const unsigned groups = (unsigned)((nFrames + 31) / 32);
    const unsigned chunks = (unsigned)((T + kChunk - 1) / kChunk);
    const unsigned unitsPerGroup = (chunks + chunksPerUnit - 1) / chunksPerUnit;
    const unsigned total = groups * unitsPerGroup;
    const unsigned unitsPerCohort = (unsigned)kCohortGroups * unitsPerGroup;

    /* unit number -> (group, unit within the group): cohort-major, chunk-major inside a cohort */
    auto locate = [&](unsigned u, unsigned &g, unsigned &su) {
        const unsigned k = u / unitsPerCohort, r = u - k * unitsPerCohort;
        const unsigned g0 = k * (unsigned)kCohortGroups;
        const unsigned cg = min((unsigned)kCohortGroups, groups - g0);   /* groups in this cohort */
        su = r / cg;
        g = g0 + (r - su * cg);
    };
    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0)
            v = atomicAdd(sched.counter, 1u);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };

    unsigned u = grab();
    uint4 pre[G::kPiecesPerRow];
    if (u < total) {
        unsigned g, su;
        locate(u, g, su);
        loadTile<Fmt, ALIGNED>(pre, segs, stride, 32LL * g, nFrames, (int)(su * chunksPerUnit) * kChunk, T, lane);
    }

    while (u < total) {
        unsigned g, su;
        locate(u, g, su);
        const unsigned cFirst = su * chunksPerUnit, cEnd = min(chunks, cFirst + chunksPerUnit);
        const long long frame0 = 32LL * g;
        const bool live = frame0 + lane < nFrames;
        uint4 *stateSlot = sched.state + ((size_t)g * 4) * 32 + lane;
        uint4 *ringG = fa.ring + ((size_t)(g % (unsigned)fa.ringSlots) * kRingPairs) * 32 + lane;
        uint8_t *dst = fa.out + (size_t)(frame0 + lane) * fa.outStride;

        uint32_t R[16];
        uint32_t expect = 0;
        if (su == 0) {
            initMetrics(R);
        } else {
            if (lane == 0)
                while (ldAcquire(sched.done + g) < (int)su)
                    __nanosleep(200);
            __syncwarp();
            __threadfence();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 v = __ldcg(stateSlot + i * 32);
                R[4 * i] = v.x;
                R[4 * i + 1] = v.y;
                R[4 * i + 2] = v.z;
                R[4 * i + 3] = v.w;
            }
            expect = __ldcg(fa.expect + (size_t)g * 32 + lane);
        }
        bool ok = true;
        /* ring accessors of this lane's frame; rows are read with ld.global.cg: the previous owner of the group
         * (another SM) wrote some of them, and its stores were released through sched.done */
        auto loadBlock = [&](int blk, uint4 (&r)[12]) {
            const uint4 *p = ringG + (size_t)((blk % kRingBlocks) * 12) * 32;
#pragma unroll
            for (int i = 0; i < 12; i++)
                r[i] = __ldcg(p + (size_t)(11 - i) * 32);
        };
        auto storeBytes = [&](int blk, uint32_t o0, uint32_t o1, uint32_t o2) {
            if (live) {
                dst[3 * blk] = (uint8_t)o0;
                dst[3 * blk + 1] = (uint8_t)o1;
                dst[3 * blk + 2] = (uint8_t)o2;
            }
        };

        unsigned un = total;
        for (unsigned c = cFirst; c < cEnd; c++) {
            const int t0 = (int)c * kChunk;
            __syncwarp();
            storeTile<Fmt, ALIGNED>(tile, pre, lane, Code::kSymMask);
            if (c + 1 < cEnd) {
                loadTile<Fmt, ALIGNED>(pre, segs, stride, frame0, nFrames, t0 + kChunk, T, lane);
            } else {
                un = grab();
                if (un < total) {
                    unsigned gn, sn;
                    locate(un, gn, sn);
                    loadTile<Fmt, ALIGNED>(pre, segs, stride, 32LL * gn, nFrames, (int)(sn * chunksPerUnit) * kChunk, T, lane);
                }
            }
            __syncwarp();
            const uintptr_t rowAddr = reinterpret_cast<uintptr_t>(segs) + (size_t)(frame0 + lane) * stride +
                                      (size_t)(t0 / Fmt::kSegsPerByte);
            const uint8_t *p = tile + lane * kPitch + (ALIGNED ? 0u : (rowAddr & 15u) * Fmt::kSegsPerByte);

#pragma unroll 1
            for (int sg = 0; sg < kSegsPerTile; sg++) {
                const int ts = t0 + sg * kSeg;             /* first step of this segment (a multiple of 96: phase 0) */
                if (ts >= T)
                    break;
                const int steps = min(kSeg, T - ts);
                /* ---- forward: `steps` trellis steps, decisions of steps (2m, 2m+1) -> ring row m mod kRingPairs ---- */
                int pr = (ts / 2) % kRingPairs;
                const int full = steps / 6;
#pragma unroll 1
                for (int it = 0; it < full; it++) {
                    uint4 *o = ringG + (size_t)pr * 32;
                    uint4 s;
                    fwdStep<Code, 0>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 1>(R, bmBase, p, minusOne, s.z, s.w);
                    o[0] = s;
                    fwdStep<Code, 2>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 3>(R, bmBase, p, minusOne, s.z, s.w);
                    o[32] = s;
                    fwdStep<Code, 4>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 5>(R, bmBase, p, minusOne, s.z, s.w);
                    o[64] = s;
                    p += 6;
                    pr += 3;                              /* rows wrap only between iterations: kRingPairs and */
                    if (pr >= kRingPairs)                 /* every segment start are multiples of 12 pairs      */
                        pr -= kRingPairs;
                }
                const int rem = steps - 6 * full;          /* T is even: 0, 2 or 4 steps, end of the frame only */
                if (rem >= 2) {
                    uint4 *o = ringG + (size_t)pr * 32;
                    uint4 s;
                    fwdStep<Code, 0>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 1>(R, bmBase, p, minusOne, s.z, s.w);
                    o[0] = s;
                    if (rem >= 4) {
                        fwdStep<Code, 2>(R, bmBase, p, minusOne, s.x, s.y);
                        fwdStep<Code, 3>(R, bmBase, p, minusOne, s.z, s.w);
                        o[32] = s;
                    }
                }
                /* ---- traceback over the window that just became complete ---- */
                if (ts + steps >= T) {
                    ok &= fusedFinalPass<Lanes8, Geo>(
                        (T - 1) / kFusedE, T, kTailSteps, expect,
                        [&](int m) { return __ldcg(ringG + (size_t)(m % kRingPairs) * 32); }, loadBlock,
                        [&](int i, uint32_t v) { if (live) dst[i] = (uint8_t)v; }, storeBytes);
                } else {
                    renorm(R);                             /* every 96 steps, see DESIGN.md 4.3 */
                    if ((ts + kSeg) % kFusedE == 0)        /* a pass every E steps */
                        ok &= fusedChunkPass<Lanes8, Geo>((ts + kSeg) / kFusedE - 1, bestPositionB(R), expect, loadBlock,
                                                          storeBytes);
                }
            }
        } /* chunks of this unit */
        if (!ok && live) {
            const long long f = frame0 + lane;
            if (atomicExch(fa.flag + f, 1u) == 0u)
                fa.list[atomicAdd(fa.count, 1)] = (int)f;
        }
        if (cEnd < chunks) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                __stcg(stateSlot + i * 32, make_uint4(R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]));
            __stcg(fa.expect + (size_t)g * 32 + lane, expect);
            __threadfence();
            __syncwarp();
            if (lane == 0)
                stRelease(sched.done + g, (int)su + 1);
        }
        u = un;
    }
}

/*
 * ---------------------------------------------------------------------------------------------------------------
 * k7FusedWsKernel -- the same decode with the traceback on its OWN warp (warp specialisation).
 *
 * ncu on k7FusedKernel (profiles/r2_fused_inline_ncu.txt): the ring is larger than the ~55 MB of L2 a write-allocate
 * working set gets on B200 (DRAM bytes jump from 0.17 GB at 2^15 frames to 3.6 GB at 2^16), so every 24-step block of a
 * pass waits ~1 us for HBM with nothing else for that warp to issue: 15 % of all warp samples sit on the first use of a
 * ring row and the issue rate drops from 0.72 to 0.66.  A forward warp must never wait for the ring.  So:
 *
 *   warps 0..3  forward ACS exactly as in k7ForwardKernel; after the segment that completes a window they post a task
 *               (group, pass index, per-frame start positions) in shared memory and go on with the next segment
 *   warp 4      takes the tasks and walks the windows.  It streams the ring with 1-D bulk copies -- one
 *               cp.async.bulk of 6 KB per 24-step block (the 12 x 512-byte rows of a block are contiguous), two in
 *               flight, completion on an mbarrier -- so its own waiting costs the SM nothing but one idle warp slot.
 *
 * Passes are independent tasks: each writes the state it started its emission from, s*(cc), and the state it arrived
 * at, a(cc), to per-pass arrays; fusedVerifyKernel afterwards flags every frame with a(cc) != s*(cc-1).  The only
 * waiting is a forward warp about to overwrite ring rows that a posted pass has not read yet (one slack segment of
 * ring makes that rare): no warp ever waits for a warp that can wait for it.
 */
constexpr int kWsThreads = 160;
constexpr int kWsFwdWarps = 4;
constexpr int kStageBlocks = 5;                           /* bulk copies in flight per traceback warp */
constexpr int kBlockBytes = 12 * 32 * (int)sizeof(uint4); /* one 24-step block of a group's ring: 6144 contiguous bytes */

template <int E_, int D_>
struct WsGeom : FusedGeom<E_, D_> {
    static constexpr int kSlack = E_;                     /* the writer may run this far past a window before its pass is done */
    static constexpr int kRingSteps = E_ + D_ + kSlack;
    static constexpr int kRingPairs = kRingSteps / 2;
    static constexpr int kRingBlocks = kRingSteps / 24;
};

struct WsArgs {
    uint4 *ring;            /* [ringSlots][kRingPairs][32] */
    uint8_t *startState;    /* [groups][passes][32]  s*(cc) */
    uint8_t *arriveState;   /* [groups][passes][32]  a(cc)  */
    int *passDone;          /* [groups][passes] 0 / 1, zeroed before the launch */
    uint8_t *out;
    size_t outStride;
    int ringSlots;
    int passes;             /* ceil(T / E) */
};

struct PassTask {
    volatile int state;     /* 0 free, 1 posted */
    int g, cc, last;
    uint32_t startB[32];
};

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void *smemDst, const void *gmemSrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smemAddr(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

template <class Code, class Fmt, bool ALIGNED, class Geo>
__global__ void __launch_bounds__(kWsThreads, 3)
k7FusedWsKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, BmTable table, FwdSched sched,
                int chunksPerUnit, WsArgs wa)
{
    using G = TileGeom<Fmt, ALIGNED>;
    constexpr int kChunk = G::kChunk, kPitch = G::kPitch;
    constexpr int E = Geo::E, D = Geo::D, kRingPairs = Geo::kRingPairs, kRingBlocks = Geo::kRingBlocks;
    constexpr int kSeg = 96;
    constexpr int kSegsPerTile = kChunk / kSeg;
    static_assert(!Code::kRuntime && Code::kRenormPeriod == kSeg, "fused kernel: compile-time codes, 96-step renorm");
    __shared__ uint4 sBm[6 * 4 * 2];
    __shared__ __align__(16) uint8_t sTile[kWsFwdWarps][32 * kPitch];
    __shared__ __align__(128) uint4 sStage[kStageBlocks][12 * 32];
    __shared__ __align__(8) uint64_t sBar[kStageBlocks];
    __shared__ PassTask sTask[kWsFwdWarps];
    __shared__ volatile int sFwdDone[kWsFwdWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 48)
        sBm[threadIdx.x] = table.x[threadIdx.x];
    if (threadIdx.x < kWsFwdWarps) {
        sTask[threadIdx.x].state = 0;
        sFwdDone[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStageBlocks; i++)
            mbarInit(&sBar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned groups = (unsigned)((nFrames + 31) / 32);

    if (warp == kWsFwdWarps) {
        /* ================= traceback warp ================= */
        uint32_t parity[kStageBlocks] = {};
        int stage = 0;          /* stage of the next block to consume */
        for (;;) {
            /* look for a posted task; leave when every forward warp is done and nothing is posted */
            int w = -1;
            bool allDone = true;
            for (int i = 0; i < kWsFwdWarps; i++) {
                if (sTask[i].state == 1 && w < 0)
                    w = i;
                allDone &= sFwdDone[i] != 0;
            }
            if (w < 0) {
                if (allDone) {
                    bool none = true;
                    for (int i = 0; i < kWsFwdWarps; i++)
                        none &= sTask[i].state == 0;
                    if (none)
                        break;
                }
                __nanosleep(100);
                continue;
            }
            __threadfence_block();
            const int g = sTask[w].g, cc = sTask[w].cc, last = sTask[w].last;
            uint32_t b = sTask[w].startB[lane];
            __syncwarp();
            asm volatile("fence.proxy.async;" ::: "memory");   /* the forward warps' ring stores -> visible to bulk copies */
            const long long frame0 = 32LL * g;
            const bool live = frame0 + lane < nFrames;
            const uint4 *ringG = wa.ring + ((size_t)((unsigned)g % (unsigned)wa.ringSlots) * kRingPairs) * 32;
            uint8_t *dst = wa.out + (size_t)(frame0 + lane) * wa.outStride;

            /* the blocks this pass walks, top down: [hiBlk .. loBlk]; the first nAcq of them without output */
            const int L = T - kTailSteps;
            int hiBlk, nAcq;
            if (last) {
                hiBlk = L / 24 - 1;
                nAcq = 0;
            } else {
                hiBlk = (E / 24) * (cc + 1) - 1;
                nAcq = D / 24;
            }
            const int loBlk = max(0, (E / 24) * cc - D / 24);
            const int nBlk = hiBlk - loBlk + 1;
            const int stage0 = stage;
            auto issue = [&](int j) {        /* j-th block of the pass -> stage (stage0 + j) mod kStageBlocks (lane 0 only) */
                const int blk = hiBlk - j, st = (stage0 + j) % kStageBlocks;
                mbarExpectTx(&sBar[st], kBlockBytes);
                bulkLoad(sStage[st], ringG + (size_t)((blk % kRingBlocks) * 12) * 32, kBlockBytes, &sBar[st]);
            };
            if (lane == 0)
                for (int j = 0; j < min(kStageBlocks, nBlk); j++)
                    issue(j);
            if (last) {
                /* ragged top of the frame: the S tail steps and what is left above the last whole block */
                int ph = (T - 1) % 6;
                uint32_t acc = 0;
                for (int m = T / 2 - 1; m >= (L / 24) * 12; m--) {
                    const uint4 wv = __ldcg(ringG + (size_t)(m % kRingPairs) * 32 + lane);
                    const int t = 2 * m;
                    const uint32_t b1 = tracebackStep<Lanes8>(b, wv.z, wv.w, ph);
                    ph = ph ? ph - 1 : 5;
                    const uint32_t b0 = tracebackStep<Lanes8>(b, wv.x, wv.y, ph);
                    ph = ph ? ph - 1 : 5;
                    if (t < L) {
                        acc = (acc >> 2) | (b1 << 6) | (b0 << 7);
                        if ((t & 7) == 0) {
                            if (live)
                                dst[t >> 3] = (uint8_t)acc;
                            acc = 0;
                        }
                    }
                }
            }
            uint32_t startState = b;
            for (int j = 0; j < nBlk; j++) {
                const int st = stage;
                mbarWait(&sBar[st], parity[st]);
                parity[st] ^= 1u;
                uint4 r[12];
#pragma unroll
                for (int i = 0; i < 12; i++)
                    r[i] = sStage[st][(11 - i) * 32 + lane];
                __syncwarp();                                 /* every lane has its rows: the stage may be refilled */
                if (lane == 0 && j + kStageBlocks < nBlk)
                    issue(j + kStageBlocks);                  /* lands in the stage just freed */
                stage = (stage + 1) % kStageBlocks;
                if (j < nAcq) {
                    walkBlockBits<Lanes8, false>(b, r);
                    if (j == nAcq - 1)
                        startState = b;
                } else {
                    const uint32_t v = walkBlockBits<Lanes8, true>(b, r);
                    const int blk = hiBlk - j;
                    if (live) {
                        dst[3 * blk] = (uint8_t)(v >> 16);
                        dst[3 * blk + 1] = (uint8_t)(v >> 8);
                        dst[3 * blk + 2] = (uint8_t)v;
                    }
                }
            }
            const size_t slot = ((size_t)g * wa.passes + cc) * 32 + lane;
            wa.startState[slot] = (uint8_t)startState;
            wa.arriveState[slot] = (uint8_t)b;
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                stRelease(wa.passDone + (size_t)g * wa.passes + cc, 1);
                sTask[w].state = 0;
            }
        }
        return;
    }

    /* ================= forward warps ================= */
    uint8_t *tile = sTile[warp];
    const uint8_t *bmBase = reinterpret_cast<const uint8_t *>(sBm);
    const uint32_t minusOne = table.minusOne;
    const unsigned chunks = (unsigned)((T + kChunk - 1) / kChunk);
    const unsigned unitsPerGroup = (chunks + chunksPerUnit - 1) / chunksPerUnit;
    const unsigned total = groups * unitsPerGroup;
    const unsigned unitsPerCohort = (unsigned)kCohortGroups * unitsPerGroup;
    auto locate = [&](unsigned u, unsigned &g, unsigned &su) {
        const unsigned k = u / unitsPerCohort, r = u - k * unitsPerCohort;
        const unsigned g0 = k * (unsigned)kCohortGroups;
        const unsigned cg = min((unsigned)kCohortGroups, groups - g0);
        su = r / cg;
        g = g0 + (r - su * cg);
    };
    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0)
            v = atomicAdd(sched.counter, 1u);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };
    PassTask &task = sTask[warp];

    unsigned u = grab();
    uint4 pre[G::kPiecesPerRow];
    if (u < total) {
        unsigned g, su;
        locate(u, g, su);
        loadTile<Fmt, ALIGNED>(pre, segs, stride, 32LL * g, nFrames, (int)(su * chunksPerUnit) * kChunk, T, lane);
    }
    while (u < total) {
        unsigned g, su;
        locate(u, g, su);
        const unsigned cFirst = su * chunksPerUnit, cEnd = min(chunks, cFirst + chunksPerUnit);
        const long long frame0 = 32LL * g;
        uint4 *stateSlot = sched.state + ((size_t)g * 4) * 32 + lane;
        uint4 *ringG = wa.ring + ((size_t)(g % (unsigned)wa.ringSlots) * kRingPairs) * 32 + lane;

        uint32_t R[16];
        if (su == 0) {
            initMetrics(R);
        } else {
            if (lane == 0)
                while (ldAcquire(sched.done + g) < (int)su)
                    __nanosleep(200);
            __syncwarp();
            __threadfence();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 v = __ldcg(stateSlot + i * 32);
                R[4 * i] = v.x;
                R[4 * i + 1] = v.y;
                R[4 * i + 2] = v.z;
                R[4 * i + 3] = v.w;
            }
        }
        unsigned un = total;
        for (unsigned c = cFirst; c < cEnd; c++) {
            const int t0 = (int)c * kChunk;
            __syncwarp();
            storeTile<Fmt, ALIGNED>(tile, pre, lane, Code::kSymMask);
            if (c + 1 < cEnd) {
                loadTile<Fmt, ALIGNED>(pre, segs, stride, frame0, nFrames, t0 + kChunk, T, lane);
            } else {
                un = grab();
                if (un < total) {
                    unsigned gn, sn;
                    locate(un, gn, sn);
                    loadTile<Fmt, ALIGNED>(pre, segs, stride, 32LL * gn, nFrames, (int)(sn * chunksPerUnit) * kChunk, T, lane);
                }
            }
            __syncwarp();
            const uintptr_t rowAddr = reinterpret_cast<uintptr_t>(segs) + (size_t)(frame0 + lane) * stride +
                                      (size_t)(t0 / Fmt::kSegsPerByte);
            const uint8_t *p = tile + lane * kPitch + (ALIGNED ? 0u : (rowAddr & 15u) * Fmt::kSegsPerByte);
#pragma unroll 1
            for (int sg = 0; sg < kSegsPerTile; sg++) {
                const int ts = t0 + sg * kSeg;
                if (ts >= T)
                    break;
                const int steps = min(kSeg, T - ts);
                /* the rows this segment overwrites belong to windows whose passes must have been walked: every pass
                 * with its top at or below ts - slack (the newest of them: the others were waited for earlier) */
                const int need = (ts - Geo::kSlack) / E;      /* passes with top <= ts - slack */
                if (ts >= Geo::kSlack + E && lane == 0)
                    while (ldAcquire(wa.passDone + (size_t)g * wa.passes + (need - 1)) == 0)
                        __nanosleep(100);
                __syncwarp();
                int pr = (ts / 2) % kRingPairs;
                const int full = steps / 6;
#pragma unroll 1
                for (int it = 0; it < full; it++) {
                    uint4 *o = ringG + (size_t)pr * 32;
                    uint4 s;
                    fwdStep<Code, 0>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 1>(R, bmBase, p, minusOne, s.z, s.w);
                    o[0] = s;
                    fwdStep<Code, 2>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 3>(R, bmBase, p, minusOne, s.z, s.w);
                    o[32] = s;
                    fwdStep<Code, 4>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 5>(R, bmBase, p, minusOne, s.z, s.w);
                    o[64] = s;
                    p += 6;
                    pr += 3;
                    if (pr >= kRingPairs)
                        pr -= kRingPairs;
                }
                const int rem = steps - 6 * full;
                if (rem >= 2) {
                    uint4 *o = ringG + (size_t)pr * 32;
                    uint4 s;
                    fwdStep<Code, 0>(R, bmBase, p, minusOne, s.x, s.y);
                    fwdStep<Code, 1>(R, bmBase, p, minusOne, s.z, s.w);
                    o[0] = s;
                    if (rem >= 4) {
                        fwdStep<Code, 2>(R, bmBase, p, minusOne, s.x, s.y);
                        fwdStep<Code, 3>(R, bmBase, p, minusOne, s.z, s.w);
                        o[32] = s;
                    }
                }
                const bool last = ts + steps >= T;
                if (!last)
                    renorm(R);
                if (last || (ts + kSeg) % E == 0) {
                    /* post the pass over the window that just became complete */
                    const uint32_t start = last ? 0u : bestPositionB(R);
                    __threadfence();                           /* the ring rows are in L2 before the task is visible */
                    if (lane == 0)
                        while (task.state != 0)
                            __nanosleep(100);
                    __syncwarp();
                    task.startB[lane] = start;
                    if (lane == 0) {
                        task.g = (int)g;
                        task.cc = last ? (T - 1) / E : (ts + kSeg) / E - 1;
                        task.last = last ? 1 : 0;
                    }
                    __syncwarp();
                    __threadfence_block();
                    if (lane == 0)
                        task.state = 1;
                }
            }
        }
        if (cEnd < chunks) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                __stcg(stateSlot + i * 32, make_uint4(R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]));
            __threadfence();
            __syncwarp();
            if (lane == 0)
                stRelease(sched.done + g, (int)su + 1);
        }
        u = un;
    }
    __syncwarp();
    if (lane == 0)
        sFwdDone[warp] = 1;
}

/* a(cc) must be s*(cc-1) for every pass that had a predecessor emitting below it (trellis_fused.cuh): frames that fail
 * are appended to the list for the two-kernel path */
static __global__ void fusedVerifyKernel(const uint8_t *__restrict__ startState, const uint8_t *__restrict__ arriveState,
                                         int nFrames, int passes, int E, int D, int *__restrict__ list, int *__restrict__ count)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nFrames)
        return;
    const size_t base = (size_t)(f >> 5) * passes * 32 + (f & 31);
    bool ok = true;
    for (int cc = 1; cc < passes; cc++)
        if (E * cc - D > 0)
            ok &= arriveState[base + (size_t)cc * 32] == startState[base + (size_t)(cc - 1) * 32];
    if (!ok)
        list[atomicAdd(count, 1)] = f;
}

/* ---- the flagged frames go through the two-kernel path: their symbol rows are gathered into a dense buffer,
 * decoded there (k7ForwardKernel / k7TracebackKernel with the frame count read from device memory) and the
 * decoded rows scattered back.  All three kernels do nothing when *count == 0. ---- */
static __global__ void gatherRowsKernel(const uint8_t *__restrict__ src, size_t srcStride, const int *__restrict__ list,
                                 const int *__restrict__ count, uint8_t *__restrict__ dst, size_t dstStride, int rowBytes)
{
    const int n = *count;
    const int per = (rowBytes + 15) / 16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * per;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / per), pc = (int)(i - (long long)row * per);
        const uint8_t *s = src + (size_t)list[row] * srcStride + (size_t)pc * 16;
        uint8_t *d = dst + (size_t)row * dstStride + (size_t)pc * 16;      /* dstStride is a multiple of 16 */
        const int nb = min(16, rowBytes - pc * 16);
        if (nb == 16 && (reinterpret_cast<uintptr_t>(s) & 15u) == 0) {
            *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(s);
        } else {
            for (int b = 0; b < nb; b++)
                d[b] = s[b];
        }
    }
}

static __global__ void scatterRowsKernel(const uint8_t *__restrict__ src, size_t srcStride, const int *__restrict__ list,
                                  const int *__restrict__ count, uint8_t *__restrict__ dst, size_t dstStride, int rowBytes)
{
    const int n = *count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * rowBytes;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / rowBytes), b = (int)(i - (long long)row * rowBytes);
        dst[(size_t)list[row] * dstStride + b] = src[(size_t)row * srcStride + b];
    }
}

} // namespace ced
