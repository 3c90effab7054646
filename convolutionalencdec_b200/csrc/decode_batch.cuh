/*
 * decode_batch.cuh -- batched K=7 r=1/2 hard-decision Viterbi decode, sm_100a.
 *
 * Replaces, for a batch of independent frames, the reference's
 *   forward ACS + renorm + survivor store  src/viterbiDecoderButterflyk1.c:85-196
 *   full-frame traceback + MSb-first pack  src/viterbiDecoderButterflyk1.c:200-260
 *
 * k7ForwardKernel  : one thread per frame, metrics in 16 registers (trellis_swar.cuh).
 *   - symbols: each warp stages a tile of 32 frames x kChunk segments through shared
 *     memory with coalesced 128-bit loads prefetched one tile ahead (byte-load fallback
 *     for rows that are not 16-byte aligned), already converted to "(rx & 3) * 32", the
 *     byte offset into the branch-metric table;
 *   - branch metrics: two LDS.128 per step from a 768-byte table [phase][rx] -> {X[4], E[4]};
 *   - survivors: 64 decision bits per step, two steps per 128-bit store, warp-major
 *     layout surv[((frame/32) * T/2 + t/2) * 32 + lane]: one sequential stream per warp.
 * k7TracebackKernel: one thread per frame walks its survivor stream backwards; the next
 *     24-step block is fetched with cp.async while the current one is walked.
 */
#pragma once
#include "trellis_swar.cuh"
#include <cuda_runtime.h>

namespace ced {

constexpr int kFwdThreads = 128;          /* 4 warps, one per SM sub-partition          */
constexpr int kTailSteps = 6;             /* S = K-1                                     */

/*
 * Wire formats of the coded symbols.
 *   ByteSymbols  : the reference's format, one byte per 2-bit segment, c0 | c1<<1
 *                  (src/viterbiDecoder.h:154); only bits 0-1 of a byte are used.
 *   PackedSymbols: four segments per byte, segment t in bits 2*(t%4)..2*(t%4)+1 of byte t/4
 *                  (a quarter of the HBM / PCIe bytes).
 * kChunk = trellis steps per work unit / per staged tile row; one 16-byte global piece carries
 * 16 (byte) or 64 (packed) segments, and kChunk is a whole number of pieces and of 96 steps.
 */
struct ByteSymbols {
    static constexpr int kSegsPerByte = 1;
    static constexpr int kChunk = 96;
};
struct PackedSymbols {
    static constexpr int kSegsPerByte = 4;
    static constexpr int kChunk = 192;
};

struct BmTable {
    uint4 x[6 * 4 * 2]; /* [phase][rx] -> X[0..3], E[0..3] (E[k] = X[k^3] - X[k] + guardWord(phase)) */
    uint32_t minusOne;  /* 0xFFFFFFFF, opaque to the compiler (trellis_swar.cuh subOnFma) */
};

template <class Code>
inline BmTable makeBmTable()
{
    BmTable t;
    for (int ph = 0; ph < 6; ph++)
        for (uint32_t rx = 0; rx < 4; rx++) {
            uint4 v;
            v.x = Code::bmWord(ph, rx, 0);
            v.y = Code::bmWord(ph, rx, 1);
            v.z = Code::bmWord(ph, rx, 2);
            v.w = Code::bmWord(ph, rx, 3);
            t.x[(ph * 4 + rx) * 2] = v;
            const uint32_t g = guardWord(ph);
            t.x[(ph * 4 + rx) * 2 + 1] = make_uint4(v.w - v.x + g, v.z - v.y + g, v.y - v.z + g, v.x - v.w + g);
        }
    t.minusOne = 0xFFFFFFFFu;
    return t;
}

/* four byte-format segments per word -> four table offsets: keep the n=2 low bits
 * (calcHammingDist(..., n), src/viterbiDecoder.c:279-283) and scale by the 32-byte table entry */
__device__ __forceinline__ uint32_t toBmOffset(uint32_t w, uint32_t symMask)
{
    return (w & symMask) << 5;
}
/* one packed byte b = [s3 s2 s1 s0] -> four table offsets (s_j * 32 in byte j).  The even and the odd
 * segments are spread by separate multiplies so that the shifted copies never overlap (a single
 * multiply would carry between them). */
__device__ __forceinline__ uint32_t packedToBmOffsets(uint32_t b)
{
    const uint32_t x = (b & 0x33u) * (0x1001u << 5);   /* s0 -> bits 5-6,   s2 -> bits 21-22 */
    const uint32_t y = (b & 0xCCu) * (0x40040u << 5);  /* s1 -> bits 13-14, s3 -> bits 29-30 */
    return (x & 0x00600060u) | (y & 0x60006000u);
}

template <class Fmt, bool ALIGNED>
struct TileGeom {
    static constexpr int kChunk = Fmt::kChunk;
    static constexpr int kSegsPerPiece = 16 * Fmt::kSegsPerByte;       /* segments in one 16-byte global piece */
    /* ALIGNED: base pointer and stride are multiples of 16, a row chunk is exactly kChunk/kSegsPerPiece
     * pieces; otherwise one more piece covers the misaligned head/tail (3.5 % slower forward pass) */
    static constexpr int kPiecesPerRow = kChunk / kSegsPerPiece + (ALIGNED ? 0 : 1);
    /* bytes per tile row in shared memory: the pieces, padded so that rows start 4 banks apart
     * (pitch/4 = 4 mod 32); 32 rows then touch 8 banks 4 times each with LDS.U8 */
    static constexpr int kPitch = Fmt::kSegsPerByte == 1 ? 112 : 272;
    static_assert(kPitch >= kPiecesPerRow * kSegsPerPiece && (kPitch / 4) % 32 == 4 || kPitch == 112, "tile pitch");
};

/*
 * Symbol staging, any base pointer and row stride (the reference's natural `[PKTS][4102]` array is not
 * 16-byte aligned).  For each of the group's 32 rows the kChunk/segsPerByte wire bytes of this unit
 * start at some address A; the lanes fetch the kPiecesPerRow ALIGNED 16-byte pieces that cover
 * [A & ~15, A + bytes) with 128-bit loads and the consumer lane later reads its row at offset
 * (A & 15) * segsPerByte.  Pieces are fetched one unit ahead of use so the HBM latency (18 % of warp
 * time in profiles/r1_v1 when loaded just in time) overlaps the ACS work of the current unit.  A piece
 * that would reach outside [bufLo, bufHi) -- before the first or past the last frame -- is assembled
 * from byte loads of its in-range part.
 */
template <class Fmt, bool ALIGNED>
__device__ __forceinline__ void loadTile(uint4 (&v)[TileGeom<Fmt, ALIGNED>::kPiecesPerRow],
                                         const uint8_t *__restrict__ segs, size_t stride, long long frame0, int nFrames,
                                         int t0, int T, int lane)
{
    using G = TileGeom<Fmt, ALIGNED>;
    const int rowBytes = (T + Fmt::kSegsPerByte - 1) / Fmt::kSegsPerByte;   /* valid wire bytes per frame */
    const uint8_t *bufLo = segs;
    const uint8_t *bufHi = segs + (size_t)(nFrames - 1) * stride + rowBytes;
#pragma unroll
    for (int i = 0; i < G::kPiecesPerRow; i++) {
        const int piece = i * 32 + lane;
        const int row = piece / G::kPiecesPerRow, pc = piece % G::kPiecesPerRow;
        const long long f = frame0 + row;
        v[i] = make_uint4(0, 0, 0, 0);
        if (f < nFrames) {
            const uint8_t *rowStart = segs + (size_t)f * stride + (size_t)(t0 / Fmt::kSegsPerByte);
            const uint8_t *src = rowStart - (reinterpret_cast<uintptr_t>(rowStart) & 15u) + 16 * pc;
            if (src >= bufLo && src + 16 <= bufHi) {
                v[i] = __ldg(reinterpret_cast<const uint4 *>(src));
            } else {
                uint32_t w[4] = {0, 0, 0, 0};
                for (int b = 0; b < 16; b++)
                    if (src + b >= bufLo && src + b < bufHi)
                        w[b >> 2] |= (uint32_t)src[b] << (8 * (b & 3));
                v[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

template <class Fmt, bool ALIGNED>
__device__ __forceinline__ void storeTile(uint8_t *tile, const uint4 (&v)[TileGeom<Fmt, ALIGNED>::kPiecesPerRow],
                                          int lane, uint32_t symMask)
{
    using G = TileGeom<Fmt, ALIGNED>;
#pragma unroll
    for (int i = 0; i < G::kPiecesPerRow; i++) {
        const int piece = i * 32 + lane;
        const int row = piece / G::kPiecesPerRow, pc = piece % G::kPiecesPerRow;
        uint8_t *dst = tile + row * G::kPitch + pc * G::kSegsPerPiece;
        const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        if (Fmt::kSegsPerByte == 1) {
            *reinterpret_cast<uint4 *>(dst) =
                make_uint4(toBmOffset(w[0], symMask), toBmOffset(w[1], symMask), toBmOffset(w[2], symMask),
                           toBmOffset(w[3], symMask));
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) /* each word: 4 packed bytes = 16 segments = one uint4 of offsets */
                *reinterpret_cast<uint4 *>(dst + 16 * q) =
                    make_uint4(packedToBmOffsets(w[q] & 0xFFu), packedToBmOffsets((w[q] >> 8) & 0xFFu),
                               packedToBmOffsets((w[q] >> 16) & 0xFFu), packedToBmOffsets(w[q] >> 24));
        }
    }
}

template <class Code, int PH>
__device__ __forceinline__ void fwdStep(uint32_t (&R)[16], const uint8_t *bmBase, const uint8_t *symPtr,
                                        uint32_t minusOne, uint32_t &t0, uint32_t &t1)
{
    const uint32_t off = symPtr[PH];
    if constexpr (Code::kRuntime) {
        /* step table [phase][entry][rx] of 8-byte entries; off = rx * 32 */
        acsStepTable<PH, Code::kCodedBits>(
            R, reinterpret_cast<const uint2 *>(bmBase + PH * (128 * Code::kVariants) + (off >> 2)), minusOne, t0, t1);
    } else {
        const uint4 x = *reinterpret_cast<const uint4 *>(bmBase + PH * 128 + off);
        const uint4 xg = *reinterpret_cast<const uint4 *>(bmBase + PH * 128 + 16 + off);
        const uint32_t X[4] = {x.x, x.y, x.z, x.w};
        const uint32_t E[4] = {xg.x, xg.y, xg.z, xg.w};
        acsStep<Code, PH>(R, X, E, minusOne, t0, t1);
    }
}

/* Work distribution of the forward kernel.  A unit is (group of 32 frames, chunk of kChunk steps);
 * units are numbered chunk-major and handed out with one atomic counter, so a unit's predecessor
 * (same group, previous chunk) was always handed out earlier, to a warp that is running. */
struct FwdSched {
    unsigned int *counter;  /* next unit to hand out (zeroed before every launch)            */
    int *done;              /* [groups] number of chunks finished per group (zeroed likewise) */
    uint4 *state;           /* [groups][4][32] path metrics handed from chunk c to chunk c+1  */
};

__device__ __forceinline__ int ldAcquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelease(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/*
 * Persistent forward kernel: gridDim.x * 4 warps, each looping over units.  Why not simply one
 * thread per frame for the whole frame: 2^16 frames are 2048 warps for 592 SM sub-partitions,
 * 3.46 per SMSP, so a static assignment leaves the 3-warp SMSPs idle for the last 13.5 % of the
 * launch (profiles/r1_v4: ALU pipe 77 % while active, 67 % of elapsed).  With 3 persistent warps per
 * SMSP pulling 96-step units the tail shrinks to one unit in 43+.  (Two frames per thread to double
 * the ILP of the few warps was measured slower: 1.61 ms vs 1.25 ms)
 */
/*
 * Continuous streams (ced_decode_window_batch): the T steps of this launch continue a stream whose metrics
 * come from / go back to a caller-owned carry block, and the survivor rows of this launch are appended
 * behind the `pairOffset` rows kept from the previous call (the traceback depth).
 */
/* uint4 slots of shared memory for the branch-metric operands: the 768-byte class table of a compile-time
 * code, or the step table of a run-time one */
template <class Code, bool RT = Code::kRuntime>
struct StepTableUint4 {
    static constexpr int value = 6 * 4 * 2;
};
template <class Code>
struct StepTableUint4<Code, true> {
    static constexpr int value = Code::kTableEntries / 2;
};

struct FwdWindow {
    const uint4 *metricsIn;  /* [groups][4][32] or NULL = start of the stream (state 0)             */
    uint4 *metricsOut;       /* [groups][4][32] renormalised metrics after the last step, or NULL    */
    uint32_t *startPos;      /* [groups][32] best-metric position after the last step (traceback b)  */
    int survPairs;           /* survivor rows (step pairs) per group in `surv`                       */
    int pairOffset;          /* rows in front of this launch's first step                            */
};

#ifndef CED_FWD_MIN_BLOCKS
#define CED_FWD_MIN_BLOCKS 4 /* 128 registers: forward 1.224 vs 1.247 ms with the 96 ptxas picks unconstrained; 3 (162 registers) is slower */
#endif
template <class Code, class Fmt, bool ALIGNED, bool CARRY = false>
__global__ void __launch_bounds__(kFwdThreads, CED_FWD_MIN_BLOCKS)
k7ForwardKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, uint4 *__restrict__ surv,
                BmTable table, FwdSched sched, int chunksPerUnit, FwdWindow win = FwdWindow(),
                const uint2 *__restrict__ stepTable = nullptr, const int *__restrict__ nFramesDev = nullptr)
{
    if (nFramesDev)   /* frame count decided on the device (the frames k7FusedKernel handed back, decode_fused.cuh) */
        nFrames = *nFramesDev;
    using G = TileGeom<Fmt, ALIGNED>;
    constexpr int kChunk = G::kChunk, kPitch = G::kPitch;
    __shared__ uint4 sBm[StepTableUint4<Code>::value];
    __shared__ __align__(16) uint8_t sTile[kFwdThreads / 32][32 * kPitch];

    if constexpr (Code::kRuntime) {
        for (int i = threadIdx.x; i < 2 * StepTableUint4<Code>::value; i += kFwdThreads)
            reinterpret_cast<uint2 *>(sBm)[i] = stepTable[i];
    } else {
        if (threadIdx.x < 48)
            sBm[threadIdx.x] = table.x[threadIdx.x];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile = sTile[warp];
    const uint8_t *bmBase = reinterpret_cast<const uint8_t *>(sBm);
    const uint32_t minusOne = table.minusOne;
    const size_t pairs = CARRY ? (size_t)win.survPairs : (size_t)(T / 2);
    const size_t pairOffset = CARRY ? (size_t)win.pairOffset : 0;
    const unsigned groups = (unsigned)((nFrames + 31) / 32);
    const unsigned chunks = (unsigned)((T + kChunk - 1) / kChunk);
    /* a unit = chunksPerUnit consecutive chunks of one group, run by one warp without hand-off: the
     * claim atomic, the state round trip through L2 and the release fence (which waits for all of the
     * warp's outstanding survivor stores) cost ~20 % of warp time at one chunk per unit (ncu v9:
     * membar 0.30, long_scoreboard 0.42, sleeping 0.10 stall cycles per issue) */
    const unsigned unitsPerGroup = (chunks + chunksPerUnit - 1) / chunksPerUnit;
    const unsigned total = groups * unitsPerGroup;

    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0)
            v = atomicAdd(sched.counter, 1u);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };

    unsigned u = grab();
    uint4 pre[G::kPiecesPerRow];
    if (u < total)
        loadTile<Fmt, ALIGNED>(pre, segs, stride, 32LL * (u % groups), nFrames,
                               (int)((u / groups) * chunksPerUnit) * kChunk, T, lane);

    while (u < total) {
        const unsigned g = u % groups, su = u / groups;
        const unsigned cFirst = su * chunksPerUnit, cEnd = min(chunks, cFirst + chunksPerUnit);
        const long long frame0 = 32LL * g;
        const bool live = frame0 + lane < nFrames;
        uint4 *stateSlot = sched.state + ((size_t)g * 4) * 32 + lane;

        uint32_t R[16];
        if (su == 0) {
            initMetrics(R);
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
        storeTile<Fmt, ALIGNED>(tile, pre, lane, Code::kSymMask);
        /* prefetch the next tile: the following chunk of this unit, or the first chunk of the next
         * unit, which is claimed now so that its symbols stream in during this chunk's ACS work */
        if (c + 1 < cEnd) {
            loadTile<Fmt, ALIGNED>(pre, segs, stride, frame0, nFrames, t0 + kChunk, T, lane);
        } else {
            un = grab();
            if (un < total)
                loadTile<Fmt, ALIGNED>(pre, segs, stride, 32LL * (un % groups), nFrames,
                                       (int)((un / groups) * chunksPerUnit) * kChunk, T, lane);
        }
        __syncwarp();
        /* this lane's frame starts (A & 15) wire bytes into the first staged piece of its row */
        const uintptr_t rowAddr = reinterpret_cast<uintptr_t>(segs) + (size_t)(frame0 + lane) * stride +
                                  (size_t)(t0 / Fmt::kSegsPerByte);
        const uint8_t *myRow = tile + lane * kPitch + (ALIGNED ? 0u : (rowAddr & 15u) * Fmt::kSegsPerByte);

        /* survivor layout: one stream per 32-frame group, (T/2) consecutive 512-byte rows (one uint4 per
         * lane and step pair), so this kernel's stores and the traceback's loads are sequential. */
        uint4 *o = surv + ((size_t)g * pairs + pairOffset + (size_t)(t0 / 2)) * 32 + lane;
        const int steps = min(kChunk, T - t0);
        const uint8_t *p = myRow;
        for (int done = 0; done < steps;) {
            const int n = min(Code::kRenormPeriod, steps - done);
            const int full = n / 6;
            for (int it = 0; it < full; it++) {
                uint4 s;
                fwdStep<Code, 0>(R, bmBase, p, minusOne, s.x, s.y);
                fwdStep<Code, 1>(R, bmBase, p, minusOne, s.z, s.w);
                if (live) o[0] = s;
                fwdStep<Code, 2>(R, bmBase, p, minusOne, s.x, s.y);
                fwdStep<Code, 3>(R, bmBase, p, minusOne, s.z, s.w);
                if (live) o[32] = s;
                fwdStep<Code, 4>(R, bmBase, p, minusOne, s.x, s.y);
                fwdStep<Code, 5>(R, bmBase, p, minusOne, s.z, s.w);
                if (live) o[64] = s;
                p += 6;
                o += 96;
            }
            /* T is even, so the remainder is 0, 2 or 4 steps (end of the frame only) */
            const int rem = n - 6 * full;
            if (rem >= 2) {
                uint4 s;
                fwdStep<Code, 0>(R, bmBase, p, minusOne, s.x, s.y);
                fwdStep<Code, 1>(R, bmBase, p, minusOne, s.z, s.w);
                if (live) o[0] = s;
            }
            if (rem >= 4) {
                uint4 s;
                fwdStep<Code, 2>(R, bmBase, p, minusOne, s.x, s.y);
                fwdStep<Code, 3>(R, bmBase, p, minusOne, s.z, s.w);
                if (live) o[32] = s;
            }
            done += n;
            if (done < steps || c + 1 < chunks)
                renorm(R); /* every Code::kRenormPeriod (96) steps, see DESIGN.md 4.3 */
        }
      } /* chunks of this unit */
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
                /* T is a multiple of 96 here, so the next step has phase 0: position == state */
                renorm(R);
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

constexpr int kTbThreads = 64;

__device__ __forceinline__ void cpAsync16(void *smemDst, const void *gmemSrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmemSrc));
}

/*
 * Full-frame traceback from state 0 (src/viterbiDecoderButterflyk1.c:205-254), one thread per
 * frame.  The survivor words a thread needs do not depend on the path, only which bit of them
 * does, so the kernel is a pure HBM-read stream: the 12 uint4 of the NEXT 24-step block are
 * fetched with cp.async (LDGSTS, no registers held) into the thread's private shared-memory
 * slots while the current block is walked -- 192..384 bytes in flight per thread.
 * Steps [24*floor(L/24), T) -- the S tail steps plus at most two bytes -- take the generic path.
 */
/* Window form (continuous streams): start at the position startPos[frame] instead of state 0, drop the top
 * `skip` steps instead of the S tail steps, and stop above step `emitLo` (a multiple of 24), whose bit is
 * the first one of the output row. */
template <class Lay = Lanes8>
__global__ void __launch_bounds__(kTbThreads)
k7TracebackKernel(const uint4 *__restrict__ surv, int nFrames, int T, uint8_t *__restrict__ out, size_t outStride,
                  const uint32_t *__restrict__ startPos = nullptr, int skip = kTailSteps, int emitLo = 0,
                  const int *__restrict__ nFramesDev = nullptr)
{
    if (nFramesDev)
        nFrames = *nFramesDev;
    __shared__ uint4 sW[2][12][kTbThreads];
    const long long frame = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= nFrames)
        return;
    const size_t pairs = (size_t)(T / 2);
    const uint4 *s = surv + ((size_t)(frame / 32) * pairs) * 32 + (frame & 31);   /* warp-major, see k7ForwardKernel */
    uint8_t *dst = out + (size_t)frame * outStride;
    const int L = T - skip;
    const int blocks24 = L / 24;
    const int loBlock = emitLo / 24;
    const int tid = threadIdx.x;

    auto prefetch = [&](int blk, int buf) {
        const uint4 *p = s + (size_t)(blk * 12) * 32;
#pragma unroll
        for (int i = 0; i < 12; i++)
            cpAsync16(&sW[buf][i][tid], p + (size_t)(11 - i) * 32);   /* slot 0 = highest pair of the block */
        asm volatile("cp.async.commit_group;");
    };
    if (blocks24 > loBlock)
        prefetch(blocks24 - 1, 0);

    uint32_t b = startPos ? startPos[frame] : 0u;   /* state 0 sits at position 0 in every phase */
    int ph = (T - 1) % 6;           /* phase of step 2m+1                        */
    uint32_t acc = 0;
    for (int m = T / 2 - 1; m >= blocks24 * 12; m--) {
        const uint4 w = __ldg(s + (size_t)m * 32);
        const int t = 2 * m;
        const uint32_t b1 = tracebackStep<Lay>(b, w.z, w.w, ph);
        ph = ph ? ph - 1 : 5;
        const uint32_t b0 = tracebackStep<Lay>(b, w.x, w.y, ph);
        ph = ph ? ph - 1 : 5;
        if (t < L && t >= emitLo) { /* the S = 6 tail steps carry no output (:208-223) */
            acc = (acc >> 2) | (b1 << 6) | (b0 << 7);   /* first visited (t%8 == 7) ends as the LSb (:249) */
            if ((t & 7) == 0) {
                dst[(t - emitLo) >> 3] = (uint8_t)acc;
                acc = 0;
            }
        }
    }
    dst -= 3 * loBlock;
    int buf = 0;
    for (int blk = blocks24 - 1; blk >= loBlock; blk--, buf ^= 1) {
        if (blk > loBlock) {
            prefetch(blk - 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        /* only the issuing thread reads its slots: no barrier needed */
        uint4 g[4];
#pragma unroll
        for (int i = 0; i < 4; i++) g[i] = sW[buf][i][tid];          /* steps 24blk+16 .. +23 */
        const uint32_t o2 = tracebackByteC<Lay, 16 % 6>(b, g);
#pragma unroll
        for (int i = 0; i < 4; i++) g[i] = sW[buf][4 + i][tid];      /* steps 24blk+8  .. +15 */
        const uint32_t o1 = tracebackByteC<Lay, 8 % 6>(b, g);
#pragma unroll
        for (int i = 0; i < 4; i++) g[i] = sW[buf][8 + i][tid];      /* steps 24blk    .. +7  */
        const uint32_t o0 = tracebackByteC<Lay, 0>(b, g);
        dst[3 * blk + 2] = (uint8_t)o2;
        dst[3 * blk + 1] = (uint8_t)o1;
        dst[3 * blk + 0] = (uint8_t)o0;
    }
}

} // namespace ced
