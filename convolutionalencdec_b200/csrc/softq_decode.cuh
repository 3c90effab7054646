/*
 * softq_decode.cuh -- 3-bit SOFT-decision Viterbi decode of the K=7 rate-1/2 code at (nearly) the hard decoder's speed.
 *
 * ced_decode_batch_soft (soft_decode.cuh) takes int8 reliabilities and needs 16-bit path metrics: twice the instructions
 * of the hard kernel (96 vs 164 Gbit/s) and twice the input bytes.  Receivers usually quantise soft decisions to three
 * bits; with eight levels the whole hard kernel carries over -- byte metrics, the same in-place butterflies, the same
 * survivor stream and traceback -- and only the branch costs change:
 *
 *   a coded bit arrives as a level x in 0..7 (0 = surely bit 0 ... 7 = surely bit 1), i.e. the reliability
 *   s = 7 - 2x in {+7, +5, ..., -7} of ced_decode_batch_soft's definition; a segment is ONE byte x0 | x1 << 3
 *   (generator 0 in the low field) -- the wire size of the hard format
 *   cost'(label c) = sum_i (c_i ? 7 - x_i : x_i) = cost(c) + (14 - |s0| - |s1|) / 2
 * where cost(c) = sum_i |s_i| [hard(s_i) != c_i] is the reliability-weighted calcHammingDist (src/viterbiDecoder.c:
 * 260-285) that ced_decode_batch_soft uses: every path through a step picks up the same offset, so every comparison,
 * ties included, is the int8 soft decoder's on the inputs s -- which is what the tests hold this kernel to.  The
 * complement label costs 14 - cost'(c), a constant, exactly like 2 - d in the hard kernel, so the step function
 * (acsStep, trellis_swar.cuh) is used unchanged with a 64-entry-per-phase table instead of a 4-entry one.
 *
 * Exactness of 8-bit metrics: a branch costs <= 14, the smallest metric grows by <= 7 per step (the two branches out of a
 * state cost c and 14 - c), and every state is reachable from the best state of 6 steps ago at <= 6 * 14 = 84: with a
 * renormalisation every 6 steps metrics stay <= 84 and candidates <= 98 < 128, so the guard-bit compare never borrows.
 * There is no room for a "never wins" start value next to that (unreachable states would climb to 84 + 84), so the
 * first six steps are run with FORCED decisions instead: before step 6 the upper predecessor j + 32 of a butterfly is
 * never reachable from state 0, so the reference's decoder keeps the lower one there (decision 0) whatever the
 * metrics; starting all metrics at 0 and taking the lower predecessor unconditionally for six steps gives every state
 * exactly the cost of its unique path from state 0 -- the metrics the reference form has after six steps.
 */
#pragma once
#include "decode_batch.cuh"

namespace ced {

constexpr int kSoftQLevels = 8;
constexpr int kSoftQTableUint4 = 6 * 64 * 2;   /* X[phase][symbol][0..3] then E[phase][symbol][0..3]: 16-byte entries, so the
                                                  64 symbols of a phase spread over all 32 banks (32-byte {X, E} entries would
                                                  use 16 of them) */

/* cost'(c) of a segment byte for the 2-bit label c */
CED_HD uint32_t softqCost(uint32_t sym, uint32_t c)
{
    const uint32_t x0 = sym & 7u, x1 = (sym >> 3) & 7u;
    return ((c & 1u) ? 7u - x0 : x0) + ((c & 2u) ? 7u - x1 : x1);
}

/* host: the branch-cost table of Code (6 * 64 * 32 bytes) */
template <class Code>
inline void buildSoftQTable(uint32_t *t)
{
    for (int ph = 0; ph < 6; ph++)
        for (uint32_t sym = 0; sym < 64; sym++) {
            uint32_t X[4] = {0, 0, 0, 0};
            for (uint32_t k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++)
                    X[k] |= softqCost(sym, k ^ Code::laneCls(l, ph)) << (8 * l);
            uint32_t *x = t + ((size_t)ph * 64 + sym) * 4, *e = t + (size_t)6 * 64 * 4 + ((size_t)ph * 64 + sym) * 4;
            for (uint32_t k = 0; k < 4; k++) {
                x[k] = X[k];
                e[k] = X[k ^ 3u] - X[k] + guardWord(ph);
            }
        }
}

/* the first six steps: the lower predecessor unconditionally, all decisions 0 (see the header) */
template <class Code, int PH>
CED_HD void acsStepForced(uint32_t (&R)[16], const uint32_t (&X)[4])
{
    constexpr int q = 5 - PH;
    if constexpr (q >= 2) {
        constexpr int rb = q - 2;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if ((r >> rb) & 1)
                continue;
            const int rh = r | (1 << rb);
            const uint32_t k = Code::regCls(r, PH);
            const uint32_t lo = R[r];
            R[r] = lo + X[k];          /* successor 2j   from j */
            R[rh] = lo + X[k ^ 3u];    /* successor 2j+1 from j */
        }
    } else {
        constexpr uint32_t swapSel = (q == 1) ? 0x1032u : 0x2301u;
        constexpr uint32_t upper = (q == 1) ? 0xFFFF0000u : 0xFF00FF00u;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const uint32_t k = Code::regCls(r, PH);
            const uint32_t self = R[r] + X[k];                          /* lower lanes: own state j -> 2j */
            const uint32_t cross = prmt(R[r], 0u, swapSel) + X[k ^ 3u]; /* upper lanes: partner j -> 2j+1 */
            R[r] = sel(~upper, self, cross);
        }
    }
}

/*
 * Forward pass: k7ForwardKernel's persistent unit scheduler and tile staging (decode_batch.cuh) with the 12 KB cost
 * table in shared memory, the symbol byte scaled to a table offset at its use, a renormalisation after every 6-step
 * iteration and the forced first iteration.
 */
template <class Code, bool ALIGNED, bool CARRY = false>
__global__ void __launch_bounds__(kFwdThreads)
k7SoftQForwardKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, uint4 *__restrict__ surv,
                     const uint4 *__restrict__ table, uint32_t minusOne, FwdSched sched, int chunksPerUnit,
                     FwdWindow win = FwdWindow())
{
    using G = TileGeom<ByteSymbols, ALIGNED>;
    constexpr int kChunk = G::kChunk, kPitch = G::kPitch;
    __shared__ uint4 sBm[kSoftQTableUint4];
    __shared__ __align__(16) uint8_t sTile[kFwdThreads / 32][32 * kPitch];
    for (int i = threadIdx.x; i < kSoftQTableUint4; i += kFwdThreads)
        sBm[i] = table[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile = sTile[warp];
    const uint8_t *bmBase = reinterpret_cast<const uint8_t *>(sBm);
    /* continuous streams (ced_decode_window_batch_softq): see FwdWindow in decode_batch.cuh */
    const size_t pairs = CARRY ? (size_t)win.survPairs : (size_t)(T / 2);
    const size_t pairOffset = CARRY ? (size_t)win.pairOffset : 0;
    const bool fresh = !CARRY || win.metricsIn == nullptr;     /* the stream starts here: forced first six steps */
    const unsigned groups = (unsigned)((nFrames + 31) / 32);
    const unsigned chunks = (unsigned)((T + kChunk - 1) / kChunk);
    const unsigned unitsPerGroup = (chunks + chunksPerUnit - 1) / chunksPerUnit;
    const unsigned total = groups * unitsPerGroup;

    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0)
            v = atomicAdd(sched.counter, 1u);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };
    /* stage the tile: the symbol bytes masked to their six bits (the table offset is formed at the use) */
    auto stage = [&](const uint4 (&v)[G::kPiecesPerRow]) {
#pragma unroll
        for (int i = 0; i < G::kPiecesPerRow; i++) {
            const int piece = i * 32 + lane;
            const int row = piece / G::kPiecesPerRow, pc = piece % G::kPiecesPerRow;
            *reinterpret_cast<uint4 *>(tile + row * kPitch + pc * G::kSegsPerPiece) =
                make_uint4(v[i].x & 0x3F3F3F3Fu, v[i].y & 0x3F3F3F3Fu, v[i].z & 0x3F3F3F3Fu, v[i].w & 0x3F3F3F3Fu);
        }
    };
    auto tableWords = [&](int ph, uint32_t sym, uint32_t (&X)[4], uint32_t (&E)[4]) {
        const uint8_t *e = bmBase + ph * 1024 + sym * 16u;
        const uint4 x = *reinterpret_cast<const uint4 *>(e);
        const uint4 g = *reinterpret_cast<const uint4 *>(e + 6 * 1024);
        X[0] = x.x; X[1] = x.y; X[2] = x.z; X[3] = x.w;
        E[0] = g.x; E[1] = g.y; E[2] = g.z; E[3] = g.w;
    };

    unsigned u = grab();
    uint4 pre[G::kPiecesPerRow];
    if (u < total)
        loadTile<ByteSymbols, ALIGNED>(pre, segs, stride, 32LL * (u % groups), nFrames, (int)((u / groups) * chunksPerUnit) * kChunk,
                                      T, lane);
    while (u < total) {
        const unsigned g = u % groups, su = u / groups;
        const unsigned cFirst = su * chunksPerUnit, cEnd = min(chunks, cFirst + chunksPerUnit);
        const long long frame0 = 32LL * g;
        const bool live = frame0 + lane < nFrames;
        uint4 *stateSlot = sched.state + ((size_t)g * 4) * 32 + lane;
        uint32_t R[16];
        if (su == 0) {
#pragma unroll
            for (int r = 0; r < 16; r++)
                R[r] = 0u;
            if constexpr (CARRY) {
                if (win.metricsIn) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 v = __ldcg(win.metricsIn + ((size_t)g * 4 + i) * 32 + lane);
                        R[4 * i] = v.x;
                        R[4 * i + 1] = v.y;
                        R[4 * i + 2] = v.z;
                        R[4 * i + 3] = v.w;
                    }
                }
            }
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
            stage(pre);
            if (c + 1 < cEnd) {
                loadTile<ByteSymbols, ALIGNED>(pre, segs, stride, frame0, nFrames, t0 + kChunk, T, lane);
            } else {
                un = grab();
                if (un < total)
                    loadTile<ByteSymbols, ALIGNED>(pre, segs, stride, 32LL * (un % groups), nFrames,
                                                  (int)((un / groups) * chunksPerUnit) * kChunk, T, lane);
            }
            __syncwarp();
            const uintptr_t rowAddr = reinterpret_cast<uintptr_t>(segs) + (size_t)(frame0 + lane) * stride + (size_t)t0;
            const uint8_t *p = tile + lane * kPitch + (ALIGNED ? 0u : (rowAddr & 15u));
            uint4 *o = surv + ((size_t)g * pairs + pairOffset + (size_t)(t0 / 2)) * 32 + lane;
            const int steps = min(kChunk, T - t0);
            int done = 0;
            uint32_t X[4], E[4];
            if (c == 0 && fresh) {
                /* steps 0..5: forced (T > 6 always: a frame has at least 8 information bits) */
                tableWords(0, p[0], X, E); acsStepForced<Code, 0>(R, X);
                tableWords(1, p[1], X, E); acsStepForced<Code, 1>(R, X);
                tableWords(2, p[2], X, E); acsStepForced<Code, 2>(R, X);
                tableWords(3, p[3], X, E); acsStepForced<Code, 3>(R, X);
                tableWords(4, p[4], X, E); acsStepForced<Code, 4>(R, X);
                tableWords(5, p[5], X, E); acsStepForced<Code, 5>(R, X);
                if (live) {
                    const uint4 z = make_uint4(0, 0, 0, 0);
                    o[0] = z;
                    o[32] = z;
                    o[64] = z;
                }
                p += 6;
                o += 96;
                done = 6;
            }
            for (; done + 6 <= steps; done += 6) {
                uint4 s;
                tableWords(0, p[0], X, E); acsStep<Code, 0>(R, X, E, minusOne, s.x, s.y);
                tableWords(1, p[1], X, E); acsStep<Code, 1>(R, X, E, minusOne, s.z, s.w);
                if (live) o[0] = s;
                tableWords(2, p[2], X, E); acsStep<Code, 2>(R, X, E, minusOne, s.x, s.y);
                tableWords(3, p[3], X, E); acsStep<Code, 3>(R, X, E, minusOne, s.z, s.w);
                if (live) o[32] = s;
                tableWords(4, p[4], X, E); acsStep<Code, 4>(R, X, E, minusOne, s.x, s.y);
                tableWords(5, p[5], X, E); acsStep<Code, 5>(R, X, E, minusOne, s.z, s.w);
                if (live) o[64] = s;
                p += 6;
                o += 96;
                renorm(R);   /* every 6 steps: metrics <= 84 at all times (see the header) */
            }
            const int rem = steps - done;   /* T is even: 0, 2 or 4 steps, end of the frame only */
            if (rem >= 2) {
                uint4 s;
                tableWords(0, p[0], X, E); acsStep<Code, 0>(R, X, E, minusOne, s.x, s.y);
                tableWords(1, p[1], X, E); acsStep<Code, 1>(R, X, E, minusOne, s.z, s.w);
                if (live) o[0] = s;
            }
            if (rem >= 4) {
                uint4 s;
                tableWords(2, p[2], X, E); acsStep<Code, 2>(R, X, E, minusOne, s.x, s.y);
                tableWords(3, p[3], X, E); acsStep<Code, 3>(R, X, E, minusOne, s.z, s.w);
                if (live) o[32] = s;
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
        } else if constexpr (CARRY) {
            if (win.metricsOut) {
                /* a slice is a multiple of 96 steps, so the metrics were renormalised by the last iteration (smallest =
                 * 0) and the next step has phase 0: position == state */
                const uint32_t best = bestPositionB(R);
                if (live) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        __stcg(win.metricsOut + ((size_t)g * 4 + i) * 32 + lane,
                               make_uint4(R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]));
                    win.startPos[(size_t)g * 32 + lane] = best;
                }
            }
        }
        u = un;
    }
}

/* int8 reliabilities (two per segment, + = bit 0) -> one byte per segment x0 | x1 << 3: uniform 8-level quantiser with
 * step `delta` (decision thresholds 0, +-delta, +-2 delta, +-3 delta) */
__global__ void quantizeSoftKernel(const int8_t *__restrict__ soft, size_t softStride, int nFrames, int segsPerFrame,
                                   uint8_t *__restrict__ syms, size_t symStride, float invDelta)
{
    const long long total = (long long)nFrames * segsPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / segsPerFrame;
        const int t = (int)(i - f * segsPerFrame);
        const int8_t *s = soft + (size_t)f * softStride + 2 * (size_t)t;
        auto level = [invDelta](int v) {
            const int k = (int)floorf((float)v * invDelta);   /* ... -2 -1 | 0 1 2 ... in units of delta */
            return (uint32_t)min(7, max(0, 3 - k));          /* k >= 3 -> 0 (surely bit 0) ... k <= -4 -> 7 */
        };
        syms[(size_t)f * symStride + t] = (uint8_t)(level(s[0]) | (level(s[1]) << 3));
    }
}

} // namespace ced
