/*
 * decode_batch.cuh -- batched K=7 r=1/2 hard-decision Viterbi decode, sm_100a.
 *
 * Replaces, for a batch of independent frames, the reference's
 *   forward ACS + renorm + survivor store  src/viterbiDecoderButterflyk1.c:85-196
 *   full-frame traceback + MSb-first pack  src/viterbiDecoderButterflyk1.c:200-260
 *
 * k7ForwardKernel  : one thread per frame, metrics in 16 registers (trellis_swar.cuh).
 *   - symbols: each warp stages a tile of 32 frames x kChunk segments through shared
 *     memory with coalesced 128-bit loads (8-byte-per-lane fallback path for rows
 *     that are not 16-byte aligned), already converted to "(rx & 3) * 16", the
 *     byte offset into the branch-metric table;
 *   - branch metrics: one LDS.128 per step from a 384-byte table [phase][rx][4 words];
 *   - survivors: 64 decision bits per step, two steps per 128-bit store, layout
 *     surv[(t/2) * framesPad + frame] so a warp writes 512 contiguous bytes.
 * k7TracebackKernel: one thread per frame walks the survivor words backwards (loads
 *     are independent of the state, so 4 are kept in flight), emits one byte per 8 steps.
 */
#pragma once
#include "trellis_swar.cuh"
#include <cuda_runtime.h>

namespace ced {

constexpr int kFwdThreads = 128;          /* 4 warps, one per SM sub-partition          */
constexpr int kChunk = 96;                /* segments staged per tile row (16 x 6 steps) */
constexpr int kPitch = kChunk + 16;       /* bytes per tile row in shared memory         */
constexpr int kTailSteps = 6;             /* S = K-1                                     */

struct BmTable {
    uint4 x[6 * 4]; /* [phase][rx] -> X[0..3] */
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
            t.x[ph * 4 + rx] = v;
        }
    return t;
}

__device__ __forceinline__ uint32_t toBmOffset(uint32_t w)
{
    /* four segments per word: keep the n=2 low bits (calcHammingDist(..., n),
     * src/viterbiDecoder.c:279-283) and scale by sizeof(uint4) */
    return (w & 0x03030303u) << 4;
}

/* Stage segments [t0, t0+kChunk) of the warp's 32 frames into `tile`. */
__device__ __forceinline__ void stageTile(uint8_t *tile, const uint8_t *__restrict__ segs, size_t stride,
                                          long long frame0, int nFrames, int t0, int T, int lane, bool aligned16)
{
    if (aligned16) {
        /* 6 x 16 bytes per row; 32 rows -> 192 pieces over 32 lanes */
#pragma unroll
        for (int i = 0; i < (32 * (kChunk / 16)) / 32; i++) {
            const int piece = i * 32 + lane;
            const int row = piece / (kChunk / 16), col = (piece % (kChunk / 16)) * 16;
            const long long f = frame0 + row;
            uint4 v = make_uint4(0, 0, 0, 0);
            /* rows are padded to a multiple of 16 by the aligned16 contract only up
             * to `stride`; never read beyond the row */
            if (f < nFrames && t0 + col < T) {
                const uint8_t *src = segs + (size_t)f * stride + (size_t)(t0 + col);
                if ((size_t)(t0 + col + 16) <= stride)
                    v = __ldg(reinterpret_cast<const uint4 *>(src));
                else {
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (int b = 0; b < 16 && t0 + col + b < T; b++)
                        w[b >> 2] |= (uint32_t)src[b] << (8 * (b & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            v.x = toBmOffset(v.x);
            v.y = toBmOffset(v.y);
            v.z = toBmOffset(v.z);
            v.w = toBmOffset(v.w);
            *reinterpret_cast<uint4 *>(tile + row * kPitch + col) = v;
        }
    } else {
        /* any alignment: lane l fetches bytes col = 4*(i*32+l) .. +3 of each row pass */
        for (int row = 0; row < 32; row++) {
            const long long f = frame0 + row;
            if (lane < kChunk / 4) {
                uint32_t w = 0;
                if (f < nFrames) {
                    const uint8_t *src = segs + (size_t)f * stride + (size_t)t0 + 4 * lane;
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        if (t0 + 4 * lane + b < T)
                            w |= (uint32_t)__ldg(src + b) << (8 * b);
                }
                *reinterpret_cast<uint32_t *>(tile + row * kPitch + 4 * lane) = toBmOffset(w);
            }
        }
    }
}

template <class Code, int PH>
__device__ __forceinline__ void fwdStep(uint32_t (&R)[16], const uint8_t *bmBase, const uint8_t *symPtr,
                                        uint32_t &t0, uint32_t &t1)
{
    const uint32_t off = symPtr[PH];
    const uint4 x = *reinterpret_cast<const uint4 *>(bmBase + PH * 64 + off);
    const uint32_t X[4] = {x.x, x.y, x.z, x.w};
    acsStep<Code, PH>(R, X, t0, t1);
}

template <class Code>
__global__ void __launch_bounds__(kFwdThreads)
k7ForwardKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, uint4 *__restrict__ surv,
                int framesPad, int aligned16, BmTable table)
{
    __shared__ uint4 sBm[6 * 4];
    __shared__ __align__(16) uint8_t sTile[kFwdThreads / 32][32 * kPitch];

    if (threadIdx.x < 24)
        sBm[threadIdx.x] = table.x[threadIdx.x];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long frame = (long long)blockIdx.x * kFwdThreads + threadIdx.x;
    const long long frame0 = frame - lane;
    const bool live = frame < nFrames;
    uint8_t *tile = sTile[warp];
    const uint8_t *myRow = tile + lane * kPitch;
    const uint8_t *bmBase = reinterpret_cast<const uint8_t *>(sBm);
    uint4 *out = surv + frame;

    uint32_t R[16];
    initMetrics(R);
    int sinceRenorm = 0;

    for (int t0 = 0; t0 < T; t0 += kChunk) {
        __syncwarp();
        stageTile(tile, segs, stride, frame0, nFrames, t0, T, lane, aligned16 != 0);
        __syncwarp();
        const int steps = min(kChunk, T - t0);
        const int full = steps / 6;
        const uint8_t *p = myRow;
        uint4 *o = out + (size_t)(t0 / 2) * framesPad;
        for (int it = 0; it < full; it++) {
            uint4 s;
            fwdStep<Code, 0>(R, bmBase, p, s.x, s.y);
            fwdStep<Code, 1>(R, bmBase, p, s.z, s.w);
            if (live) o[0] = s;
            fwdStep<Code, 2>(R, bmBase, p, s.x, s.y);
            fwdStep<Code, 3>(R, bmBase, p, s.z, s.w);
            if (live) o[framesPad] = s;
            fwdStep<Code, 4>(R, bmBase, p, s.x, s.y);
            fwdStep<Code, 5>(R, bmBase, p, s.z, s.w);
            if (live) o[2 * (size_t)framesPad] = s;
            p += 6;
            o += 3 * (size_t)framesPad;
        }
        /* T is even, so the remainder is 0, 2 or 4 steps (last chunk only) */
        const int rem = steps - 6 * full;
        if (rem >= 2) {
            uint4 s;
            fwdStep<Code, 0>(R, bmBase, p, s.x, s.y);
            fwdStep<Code, 1>(R, bmBase, p, s.z, s.w);
            if (live) o[0] = s;
        }
        if (rem >= 4) {
            uint4 s;
            fwdStep<Code, 2>(R, bmBase, p, s.x, s.y);
            fwdStep<Code, 3>(R, bmBase, p, s.z, s.w);
            if (live) o[framesPad] = s;
        }
        sinceRenorm += kChunk;
        if (sinceRenorm >= kRenormPeriod) {
            renorm(R);
            sinceRenorm = 0;
        }
    }
}

/* Full-frame traceback from state 0 (src/viterbiDecoderButterflyk1.c:205-254). */
__global__ void __launch_bounds__(128)
k7TracebackKernel(const uint4 *__restrict__ surv, int framesPad, int nFrames, int T, uint8_t *__restrict__ out,
                  size_t outStride)
{
    const long long frame = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= nFrames)
        return;
    const uint4 *s = surv + frame;
    uint8_t *dst = out + (size_t)frame * outStride;
    uint32_t b = 0;                 /* state 0 sits at position 0 in every phase */
    int m = T / 2 - 1;              /* index of the step pair (2m, 2m+1)         */
    int ph = (T - 1) % 6;           /* phase of step 2m+1                        */
    /* the S = 6 tail steps carry no output (:208-223) */
#pragma unroll
    for (int i = 0; i < kTailSteps / 2; i++, m--) {
        const uint4 w = __ldg(s + (size_t)m * framesPad);
        tracebackStep(b, w.z, w.w, ph);
        ph = ph ? ph - 1 : 5;
        tracebackStep(b, w.x, w.y, ph);
        ph = ph ? ph - 1 : 5;
    }
    /* L = T-6 is a multiple of 8: one output byte per 4 pairs */
    for (int byteIdx = (T - kTailSteps) / 8 - 1; byteIdx >= 0; byteIdx--, m -= 4) {
        uint4 w[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
            w[i] = __ldg(s + (size_t)(m - i) * framesPad);
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            /* step 2(m-i)+1 = bit (6 - 2i ... ) of the byte: steps are visited in
             * descending order, the first visited (t%8 == 7) is the LSb (:249) */
            acc |= tracebackStep(b, w[i].z, w[i].w, ph) << (2 * i);
            ph = ph ? ph - 1 : 5;
            acc |= tracebackStep(b, w[i].x, w[i].y, ph) << (2 * i + 1);
            ph = ph ? ph - 1 : 5;
        }
        dst[byteIdx] = (uint8_t)acc;
    }
}

} // namespace ced
