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

template <class Code, class Fmt, bool ALIGNED>
__global__ void __launch_bounds__(kFwdThreads, 4)
k7FusedKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, BmTable table, FwdSched sched,
              int chunksPerUnit, FusedArgs fa)
{
    using G = TileGeom<Fmt, ALIGNED>;
    constexpr int kChunk = G::kChunk, kPitch = G::kPitch;
    constexpr int kSegsPerTile = kChunk / kFusedE;          /* 96-step segments per staged tile: 1 (byte) / 2 (packed) */
    static_assert(!Code::kRuntime && Code::kRenormPeriod == kFusedE, "fused kernel: compile-time codes, 96-step renorm");
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
                const int ts = t0 + sg * kFusedE;          /* first step of this segment (a multiple of 96: phase 0) */
                if (ts >= T)
                    break;
                const int cc = ts / kFusedE;
                const int steps = min(kFusedE, T - ts);
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
                    ok &= fusedFinalPass<Lanes8>(
                        cc, T, kTailSteps, expect, [&](int m) { return __ldcg(ringG + (size_t)(m % kRingPairs) * 32); },
                        loadBlock, [&](int i, uint32_t v) { if (live) dst[i] = (uint8_t)v; }, storeBytes);
                } else {
                    renorm(R);                             /* every 96 steps, see DESIGN.md 4.3 */
                    ok &= fusedChunkPass<Lanes8>(cc, bestPositionB(R), expect, loadBlock, storeBytes);
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
