/*
 * stream_kernels.cuh -- the per-frame, chunk-by-chunk decoder behind
 * VITERBI_DECODER_HARD (src/viterbiDecoderButterflyk1.c:82-263).
 *
 * One CTA per call, one thread per butterfly (N/2 threads, at least a warp).
 * Unlike the batch kernel this one keeps the reference's *absolute* metrics:
 * uint8 arithmetic that wraps on store (:109-115) and renormalisation when the
 * carried renormCounter reaches 120 (:159-183), because the host reads
 * nodeMetricsCur between calls (handTracedTest/handTraced.c:72-111).  General
 * (non-symmetric) butterflies are used so the K=3 test code g={7,6} works:
 *   a0 = m[j]     + HD(edge[0][j],     rx)    a1 = m[j+N/2] + HD(edge[0][j+N/2], rx)
 *   b0 = m[j]     + HD(edge[1][j],     rx)    b1 = m[j+N/2] + HD(edge[1][j+N/2], rx)
 * which for symmetric generators equals the reference's e / n-e form (:104-115).
 * Decisions are ballot-packed: step t, warp w -> surv[t*W + 2w] (successors 2j)
 * and surv[t*W + 2w + 1] (successors 2j+1), bit = j & 31.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ced {

constexpr int kStreamSegChunk = 4096;  /* segments staged in shared memory per pass */
constexpr int kStreamTbChunk = 512;    /* trellis steps of survivors staged per pass */

struct StreamArgs {
    int K, n, N, W;              /* W = uint32 survivor words per step            */
    uint32_t iteration;          /* steps already taken for this frame            */
    uint32_t renormCounter;
    int segmentsIn;
    int last;
    const uint8_t *edge;         /* [2][N]                                        */
    const uint8_t *metricsIn;    /* [N] path metrics before the call              */
    uint8_t *metrics;            /* [N] path metrics after the call               */
    const uint8_t *segs;         /* [segmentsIn]                                  */
    uint32_t *surv;              /* [capacity][W]                                 */
    uint32_t *stateOut;          /* [0] = renormCounter after the call            */
    uint8_t *out;                /* decoded bytes (last only)                     */
};

__device__ __forceinline__ void streamDecodeBlockBody(const StreamArgs &a)
{
    __shared__ uint8_t sMetric[2][256];
    __shared__ uint8_t sSeg[kStreamSegChunk];
    __shared__ uint32_t sSurv[kStreamTbChunk * 8];
    __shared__ uint8_t sWarpMin[4];

    const int N = a.N, H = N / 2, j = threadIdx.x;
    const int lane = j & 31, warp = j >> 5;
    const bool active = j < H;
    const uint32_t nmask = (1u << a.n) - 1u;

    uint32_t e00 = 0, e0h = 0, e10 = 0, e1h = 0;
    if (active) {
        e00 = a.edge[j];
        e0h = a.edge[j + H];
        e10 = a.edge[N + j];
        e1h = a.edge[N + j + H];
        sMetric[0][j] = a.metricsIn[j];
        sMetric[0][j + H] = a.metricsIn[j + H];
    }
    uint32_t renormCounter = a.renormCounter;
    int cur = 0;
    __syncthreads();

    for (int base = 0; base < a.segmentsIn; base += kStreamSegChunk) {
        const int cnt = min(kStreamSegChunk, a.segmentsIn - base);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x)
            sSeg[i] = a.segs[base + i];
        __syncthreads();
        for (int i = 0; i < cnt; i++) {
            const uint32_t rx = sSeg[i];
            uint32_t da = 0, db = 0;
            uint8_t na = 0xFF, nb = 0xFF;
            if (active) {
                const uint8_t lo = sMetric[cur][j], hi = sMetric[cur][j + H];
                const uint8_t a0 = (uint8_t)(lo + __popc((e00 ^ rx) & nmask));
                const uint8_t a1 = (uint8_t)(hi + __popc((e0h ^ rx) & nmask));
                const uint8_t b0 = (uint8_t)(lo + __popc((e10 ^ rx) & nmask));
                const uint8_t b1 = (uint8_t)(hi + __popc((e1h ^ rx) & nmask));
                da = a0 > a1;
                db = b0 > b1;
                na = da ? a1 : a0;
                nb = db ? b1 : b0;
            }
            const uint32_t wa = __ballot_sync(0xFFFFFFFFu, da);
            const uint32_t wb = __ballot_sync(0xFFFFFFFFu, db);
            if (lane == 0 && warp * 32 < H) {
                uint32_t *row = a.surv + (size_t)(a.iteration + base + i) * a.W;
                row[2 * warp] = wa;
                row[2 * warp + 1] = wb;
            }
            if (renormCounter >= 120) { /* uniform across the CTA */
                uint32_t mn = na < nb ? na : nb;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, mn, o);
                    mn = other < mn ? other : mn;
                }
                if (lane == 0)
                    sWarpMin[warp] = (uint8_t)mn;
                __syncthreads();
                for (int w2 = 0; w2 * 32 < (int)blockDim.x; w2++)
                    mn = sWarpMin[w2] < mn ? sWarpMin[w2] : mn;
                na = (uint8_t)(na - mn);
                nb = (uint8_t)(nb - mn);
                renormCounter = 0;
            } else {
                renormCounter++;
            }
            if (active) {
                sMetric[cur ^ 1][2 * j] = na;
                sMetric[cur ^ 1][2 * j + 1] = nb;
            }
            cur ^= 1;
            __syncthreads();
        }
    }
    if (active) {
        a.metrics[j] = sMetric[cur][j];
        a.metrics[j + H] = sMetric[cur][j + H];
    }
    if (threadIdx.x == 0)
        a.stateOut[0] = renormCounter;
    if (!a.last)
        return;

    /* traceback (:200-256): state 0, S unrecorded tail steps, then MSb-first bytes */
    __threadfence_block();
    __syncthreads();
    const int S = a.K - 1;
    const int T = (int)a.iteration + a.segmentsIn;
    uint32_t state = 0, acc = 0;
    for (int hiStep = T; hiStep > 0; hiStep -= kStreamTbChunk) {
        const int loStep = max(0, hiStep - kStreamTbChunk);
        const int words = (hiStep - loStep) * a.W;
        for (int i = threadIdx.x; i < words; i += blockDim.x)
            sSurv[i] = a.surv[(size_t)loStep * a.W + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int t = hiStep - 1; t >= loStep; t--) {
                const uint32_t jj = state >> 1;
                const uint32_t word = sSurv[(t - loStep) * a.W + 2 * (jj >> 5) + (state & 1u)];
                const uint32_t dec = (word >> (jj & 31u)) & 1u;
                if (t < T - S) {
                    acc = (acc >> 1) | ((state & 1u) << 7);
                    if ((t & 7) == 0) {
                        a.out[t >> 3] = (uint8_t)acc;
                        acc = 0;
                    }
                }
                state = (state >> 1) | (dec << (S - 1));
            }
        }
        __syncthreads();
    }
}

/*
 * Latency-optimised variant for trellises of up to 64 states (K <= 7): ONE warp, lane j = butterfly j,
 * metrics m[j], m[j+H] in registers, successors exchanged with two __shfl_sync per step (no shared
 * memory, no block barrier).  Same arithmetic, renormalisation schedule and survivor packing as
 * streamDecodeKernel.  The traceback is done by lane 0 over survivor rows staged in shared memory;
 * the loads do not depend on the path, so they are issued 8 steps ahead of the bit extraction.
 */
constexpr int kStreamWarpSegChunk = 2048;

__device__ __forceinline__ void streamDecodeWarpBody(const StreamArgs &a)
{
    __shared__ __align__(16) uint8_t sSeg[kStreamWarpSegChunk + 16];
    __shared__ uint2 sSurv[kStreamTbChunk];

    const int N = a.N, H = N / 2, j = threadIdx.x;
    const bool active = j < H;
    const uint32_t nmask = (1u << a.n) - 1u;
    uint32_t e00 = 0, e0h = 0, e10 = 0, e1h = 0, lo = 0xFF, hi = 0xFF;
    if (active) {
        e00 = a.edge[j];
        e0h = a.edge[j + H];
        e10 = a.edge[N + j];
        e1h = a.edge[N + j + H];
        lo = a.metricsIn[j];
        hi = a.metricsIn[j + H];
    }
    /* successor 2j' / 2j'+1 of butterfly j' becomes state j (lo) resp. j+H (hi) of the next step */
    const int srcLo = j >> 1, srcHi = (j + H) >> 1;
    const bool oddLo = j & 1, oddHi = (j + H) & 1;
    uint32_t renormCounter = a.renormCounter;
    uint2 *survOut = reinterpret_cast<uint2 *>(a.surv) + a.iteration;

    for (int base = 0; base < a.segmentsIn; base += kStreamWarpSegChunk) {
        const int cnt = min(kStreamWarpSegChunk, a.segmentsIn - base);
        for (int i = j; i < cnt; i += 32)
            sSeg[i] = a.segs[base + i];
        if (j == 0)
            sSeg[cnt] = 0; /* read one step ahead below */
        __syncwarp();
        /* branch metrics of step i+1 are computed during step i: the only loop-carried chain is
         * add -> compare -> select -> shuffle */
        uint32_t rx = sSeg[0];
        uint32_t h00 = __popc((e00 ^ rx) & nmask), h0h = __popc((e0h ^ rx) & nmask);
        uint32_t h10 = __popc((e10 ^ rx) & nmask), h1h = __popc((e1h ^ rx) & nmask);
        for (int i = 0; i < cnt; i++) {
            const uint32_t rxn = sSeg[i + 1];
            const uint32_t n00 = __popc((e00 ^ rxn) & nmask), n0h = __popc((e0h ^ rxn) & nmask);
            const uint32_t n10 = __popc((e10 ^ rxn) & nmask), n1h = __popc((e1h ^ rxn) & nmask);
            const uint32_t a0 = (lo + h00) & 0xFFu; /* uint8 arithmetic (:109-115) */
            const uint32_t a1 = (hi + h0h) & 0xFFu;
            const uint32_t b0 = (lo + h10) & 0xFFu;
            const uint32_t b1 = (hi + h1h) & 0xFFu;
            const bool da = active && a0 > a1, db = active && b0 > b1;
            uint32_t na = da ? a1 : a0, nb = db ? b1 : b0;
            const uint32_t wa = __ballot_sync(0xFFFFFFFFu, da);
            const uint32_t wb = __ballot_sync(0xFFFFFFFFu, db);
            if (j == 0)
                survOut[base + i] = make_uint2(wa, wb);
            if (renormCounter >= 120) { /* uniform across the warp */
                uint32_t mn = active ? min(na, nb) : 0xFFu;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
                na = (na - mn) & 0xFFu;
                nb = (nb - mn) & 0xFFu;
                renormCounter = 0;
            } else {
                renormCounter++;
            }
            const uint32_t ab = na | (nb << 8);
            const uint32_t v = __shfl_sync(0xFFFFFFFFu, ab, srcLo);
            const uint32_t w = __shfl_sync(0xFFFFFFFFu, ab, srcHi);
            lo = oddLo ? (v >> 8) : (v & 0xFFu);
            hi = oddHi ? (w >> 8) : (w & 0xFFu);
            h00 = n00;
            h0h = n0h;
            h10 = n10;
            h1h = n1h;
        }
        __syncwarp();
    }
    if (active) {
        a.metrics[j] = (uint8_t)lo;
        a.metrics[j + H] = (uint8_t)hi;
    }
    if (j == 0)
        a.stateOut[0] = renormCounter;
    if (!a.last)
        return;

    __threadfence_block();
    __syncwarp();
    const int S = a.K - 1;
    const int T = (int)a.iteration + a.segmentsIn;
    const int L = T - S;
    const uint2 *surv = reinterpret_cast<const uint2 *>(a.surv);
    uint32_t state = 0, acc = 0;
    auto back = [&](const uint2 w) -> uint32_t { /* one step back; returns the decoded bit of that step */
        const uint32_t word = (state & 1u) ? w.y : w.x;
        const uint32_t dec = (word >> (state >> 1)) & 1u;
        const uint32_t bit = state & 1u;
        state = (state >> 1) | (dec << (S - 1));
        return bit;
    };
    /* generic part: the S tail steps and the (at most 7) steps above the last byte boundary */
    int t = T - 1;
    if (j == 0) {
        for (; t >= 0 && (t >= L || (t & 7) != 7); t--) {
            const uint32_t bit = back(surv[t]);
            if (t < L) {
                acc = (acc >> 1) | (bit << 7);
                if ((t & 7) == 0) {
                    a.out[t >> 3] = (uint8_t)acc;
                    acc = 0;
                }
            }
        }
    }
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    /* now t + 1 is a multiple of 8: whole bytes, survivor rows staged through shared memory */
    for (int hiStep = t + 1; hiStep > 0; hiStep -= kStreamTbChunk) {
        const int loStep = max(0, hiStep - kStreamTbChunk);
        for (int i = j; i < hiStep - loStep; i += 32)
            sSurv[i] = surv[loStep + i];
        __syncwarp();
        if (j == 0) {
            for (int tb = hiStep - 8; tb >= loStep; tb -= 8) {
                uint2 w[8];
#pragma unroll
                for (int q = 0; q < 8; q++)
                    w[q] = sSurv[tb - loStep + q];
                uint32_t byte = 0;
#pragma unroll
                for (int q = 7; q >= 0; q--)
                    byte |= back(w[q]) << (7 - q); /* step tb+q is bit 7-q of byte tb/8 (:249) */
                a.out[tb >> 3] = (uint8_t)byte;
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) streamDecodeKernel(StreamArgs a)
{
    streamDecodeBlockBody(a);
}
__global__ void __launch_bounds__(32) streamDecodeWarpKernel(StreamArgs a)
{
    streamDecodeWarpBody(a);
}

/*
 * Batch decoding for code parameters the SWAR kernel is not built for (K = 3, 5, 9, other
 * generators, n up to 8): every CTA takes whole frames (grid-stride) through the same bodies as the
 * per-frame API -- one warp per frame for <= 64 states, N/2 threads above.  Decisions do not depend on
 * the renormalisation schedule, so the reference's schedule is simply kept.
 */
struct GenericBatchArgs {
    int K, n, N, W;
    int nFrames, T;
    const uint8_t *edge;         /* [2][N] */
    const uint8_t *initMetrics;  /* [N]: 0 for state 0, N+1 elsewhere (:59-67) */
    const uint8_t *segs;
    size_t segStride;
    uint32_t *surv;              /* [nFrames][T][W] */
    uint8_t *out;
    size_t outStride;
    uint8_t *metricsSink;        /* [gridDim.x][256] final metrics, unused */
    uint32_t *stateSink;         /* [gridDim.x] */
};

template <bool WARP>
__global__ void __launch_bounds__(128) genericBatchDecodeKernel(GenericBatchArgs g)
{
    for (long long f = blockIdx.x; f < g.nFrames; f += gridDim.x) {
        StreamArgs a;
        a.K = g.K;
        a.n = g.n;
        a.N = g.N;
        a.W = g.W;
        a.iteration = 0;
        a.renormCounter = 0;
        a.segmentsIn = g.T;
        a.last = 1;
        a.edge = g.edge;
        a.metricsIn = g.initMetrics;
        a.metrics = g.metricsSink + (size_t)blockIdx.x * 256;
        a.segs = g.segs + (size_t)f * g.segStride;
        a.surv = g.surv + (size_t)f * (size_t)g.T * g.W;
        a.stateOut = g.stateSink + blockIdx.x;
        a.out = g.out + (size_t)f * g.outStride;
        if constexpr (WARP)
            streamDecodeWarpBody(a);
        else
            streamDecodeBlockBody(a);
        __syncthreads(); /* shared staging buffers are reused by the next frame */
    }
}

} // namespace ced
