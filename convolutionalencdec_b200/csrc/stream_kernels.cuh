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
    uint8_t *metrics;            /* [N] in/out                                    */
    const uint8_t *segs;         /* [segmentsIn]                                  */
    uint32_t *surv;              /* [capacity][W]                                 */
    uint32_t *stateOut;          /* [0] = renormCounter after the call            */
    uint8_t *out;                /* decoded bytes (last only)                     */
};

__global__ void __launch_bounds__(128) streamDecodeKernel(StreamArgs a)
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
        sMetric[0][j] = a.metrics[j];
        sMetric[0][j + H] = a.metrics[j + H];
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

} // namespace ced
