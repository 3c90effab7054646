/*
 * radix_k.cu -- rate-k/n codes with k > 1 (VERDICT r1 missing 3): batched encoder and Viterbi decoder
 * behind ced_encode_batch_k / ced_decode_batch_k.
 *
 * The reference describes such a code as ONE shift register of k*K bits that takes k message bits per coded segment
 * (src/convEncode.h:8-18, src/convEncode.c:46-130; generators are k*K-bit masks, src/convEncode.c:163-175) and decodes
 * it on a trellis of 2^(k*S) states with 2^k branches into every state (src/viterbiDecoder.c:95-128):
 *
 *   destination d:  edgeOut = d mod 2^k,  sources  d / 2^k + edgeIn * 2^((S-1)k),  edgeIn = 0 .. 2^k - 1
 *   path metric   = source metric + calcHammingDist(label(source, edgeOut), rx, n)
 *   survivor      = the smallest, the LOWEST edgeIn on equal metrics (argminPathMetrics: '<=' trees keep the left entry)
 *   start metrics = 0 / NUM_STATES + 1, no renormalisation (METRIC_TYPE is sized for the packet, src/viterbiDecoder.h:52-61)
 *
 * The reference's own k > 1 decoder does not produce output at HEAD (its register-exchange traceback shifts a uint8_t by
 * (5K-1)k bits); the traceback here is the full-frame walk from state 0 that its butterfly decoder performs at `last`
 * (src/viterbiDecoderButterflyk1.c:200-256), whose formulas are written for general k.
 *
 * Mapping: one warp per frame, lane l owns destination states l, l + 32, ...; path metrics are ints in shared memory,
 * double-buffered; the k decision bits of 32 states become k __ballot_sync words, so a step stores W = k * max(1, N/32)
 * words which the lanes write as one row.  A second kernel walks the rows back, one thread per frame.  This is the
 * general kernel of the family (the hand-scheduled SIMD-in-word kernels are k = 1 only): 2^k candidates per state and
 * step instead of 2.
 */
#include "ced_internal.cuh"

namespace ced {

struct KCode {
    int K, k, n, S, N, P, top, W;   /* N = 2^(kS) states, P = 2^k branches, top = (S-1)k, W = survivor words per step */
    uint32_t taps[CED_MAX_N];       /* generators with bit 0 on the newest input bit */
};

__device__ __forceinline__ uint32_t kSegment(const KCode &c, uint32_t reg)
{
    uint32_t v = 0;
    for (int i = 0; i < c.n; i++)
        v |= (uint32_t)(__popc(reg & c.taps[i]) & 1) << i;
    return v;
}

/* One thread per coded segment: the register after segment t holds message bits k(t+1) - kK .. k(t+1) - 1, newest in
 * bit 0 (src/convEncode.c:56-97); bits before the message and in the S tail segments are 0 (:100-122). */
__global__ void kEncodeBatchKernel(KCode c, const uint8_t *__restrict__ msg, size_t msgStride, int nFrames, int frameBytes,
                                   uint8_t *__restrict__ segs, size_t segStride)
{
    const int T = frameBytes * 8 / c.k + c.S;
    const long long total = (long long)nFrames * T;
    const int regBits = c.k * c.K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / T;
        const int t = (int)(i - f * T);
        const uint8_t *m = msg + (size_t)f * msgStride;
        const int end = c.k * (t + 1); /* one past the newest message bit in the register */
        uint32_t reg = 0;
        for (int j = 0; j < regBits; j++) {
            const int bit = end - 1 - j;
            if (bit >= 0 && bit < frameBytes * 8)
                reg |= (uint32_t)((m[bit >> 3] >> (7 - (bit & 7))) & 1u) << j;
        }
        segs[(size_t)f * segStride + t] = (uint8_t)kSegment(c, reg);
    }
}

constexpr int kRkWarps = 4;

/* Forward recursion, one warp per frame.  Shared memory: edge labels [P][N] bytes, then int metrics [warp][2][N]. */
__global__ void __launch_bounds__(kRkWarps * 32)
kForwardKernel(KCode c, const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, uint32_t *__restrict__ surv)
{
    extern __shared__ __align__(16) uint8_t sMem[];
    uint8_t *sEdge = sMem;
    int *sMetric = reinterpret_cast<int *>(sMem + ((c.P * c.N + 15) / 16) * 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < c.P * c.N; i += blockDim.x) {
        const int e = i / c.N, s = i - e * c.N;
        sEdge[i] = (uint8_t)kSegment(c, ((uint32_t)s << c.k) | (uint32_t)e); /* viterbiInit, src/viterbiDecoder.c:32-50 */
    }
    __syncthreads();
    int *m0 = sMetric + warp * 2 * c.N, *m1 = m0 + c.N;
    const uint32_t nmask = (1u << c.n) - 1u;
    const int slots = c.N >= 32 ? c.N / 32 : 1;
    for (long long f = (long long)blockIdx.x * kRkWarps + warp; f < nFrames; f += (long long)gridDim.x * kRkWarps) {
        for (int s = lane; s < c.N; s += 32)
            m0[s] = s == 0 ? 0 : c.N + 1; /* resetViterbiDecoderHard, src/viterbiDecoder.c:236-258 */
        __syncwarp();
        const uint8_t *row = segs + (size_t)f * stride;
        uint32_t *out = surv + (size_t)f * (size_t)T * c.W;
        int *cur = m0, *nxt = m1;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const uint32_t mine = (t0 + lane < T) ? row[t0 + lane] : 0u; /* 32 segments per load, one per lane */
            const int steps = min(32, T - t0);
            for (int i = 0; i < steps; i++) {
                const uint32_t rx = __shfl_sync(0xFFFFFFFFu, mine, i);
                uint32_t word = 0; /* lane j keeps survivor word j of this step */
                for (int sl = 0; sl < slots; sl++) {
                    const int d = sl * 32 + lane;
                    uint32_t best = 0, bestIn = 0;
                    if (d < c.N) {
                        const int edgeOut = d & (c.P - 1), base = d >> c.k;
                        const uint8_t *lab = sEdge + edgeOut * c.N;
                        for (int in = 0; in < c.P; in++) {
                            const int src = base + (in << c.top);
                            const uint32_t pm = (uint32_t)cur[src] + (uint32_t)__popc((lab[src] ^ rx) & nmask);
                            if (in == 0 || pm < best) { /* strict '<': the lowest edgeIn keeps a tie */
                                best = pm;
                                bestIn = (uint32_t)in;
                            }
                        }
                        nxt[d] = (int)best;
                    }
                    for (int b = 0; b < c.k; b++) {
                        const uint32_t w = __ballot_sync(0xFFFFFFFFu, (bestIn >> b) & 1u);
                        if (lane == sl * c.k + b)
                            word = w;
                    }
                }
                if (lane < c.W)
                    out[(size_t)(t0 + i) * c.W + lane] = word;
                __syncwarp();
                int *tmp = cur;
                cur = nxt;
                nxt = tmp;
            }
        }
        __syncwarp();
    }
}

/* Full traceback from state 0 (src/viterbiDecoderButterflyk1.c:200-256, k-generic formulas), one thread per frame. */
__global__ void kTracebackKernel(KCode c, const uint32_t *__restrict__ surv, int nFrames, int T, uint8_t *__restrict__ out,
                                 size_t outStride)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nFrames)
        return;
    const uint32_t *rows = surv + (size_t)f * (size_t)T * c.W;
    uint8_t *dst = out + (size_t)f * outStride;
    auto decision = [&](int t, uint32_t state) {
        const uint32_t *r = rows + (size_t)t * c.W + (state >> 5) * c.k;
        uint32_t dec = 0;
        for (int b = 0; b < c.k; b++)
            dec |= ((r[b] >> (state & 31u)) & 1u) << b;
        return dec;
    };
    uint32_t state = 0;
    for (int t = T - 1; t >= T - c.S; t--) /* tail: no output (:208-223) */
        state = (state >> c.k) | (decision(t, state) << c.top);
    uint32_t acc = 0;
    for (int t = T - c.S - 1; t >= 0; t--) {
        acc = (acc >> c.k) | ((state & (uint32_t)(c.P - 1)) << (8 - c.k)); /* :249 */
        if ((t * c.k) % 8 == 0) {
            dst[t * c.k / 8] = (uint8_t)acc;
            acc = 0;
        }
        state = (state >> c.k) | (decision(t, state) << c.top); /* :252 */
    }
}

} // namespace ced

static int makeKCode(const ced_code_t *code, int inputBits, ced::KCode *kc, const char *who)
{
    if (!code || (inputBits != 2 && inputBits != 4) || code->constraintLen < 2 || code->codedBits < 1 ||
        code->codedBits > CED_MAX_N || inputBits * code->constraintLen > 32 ||
        inputBits * (code->constraintLen - 1) > 8) {
        setError("%s: k in {1, 2, 4}, n <= %d, k*K <= 32 and at most %d states (k*(K-1) <= 8)", who, CED_MAX_N, CED_MAX_STATES);
        return CED_ERR_UNSUPPORTED;
    }
    kc->K = code->constraintLen;
    kc->k = inputBits;
    kc->n = code->codedBits;
    kc->S = kc->K - 1;
    kc->N = 1 << (kc->k * kc->S);
    kc->P = 1 << kc->k;
    kc->top = (kc->S - 1) * kc->k;
    kc->W = kc->k * (kc->N >= 32 ? kc->N / 32 : 1);
    for (int i = 0; i < CED_MAX_N; i++)
        kc->taps[i] = i < kc->n ? reverseBits(code->gen[i], kc->k * kc->K) : 0u;
    return CED_OK;
}

int ced_encode_batch_k(ced_ctx *c, const ced_code_t *code, int inputBits, const uint8_t *dMsg, size_t msgStride, int nFrames,
                       int frameBytes, uint8_t *dSegs, size_t segStride, void *stream)
{
    if (inputBits == 1)
        return ced_encode_batch(c, code, dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, stream);
    ced::KCode kc;
    const int rc = makeKCode(code, inputBits, &kc, "ced_encode_batch_k");
    if (rc != CED_OK)
        return rc;
    if (!c || nFrames < 0 || frameBytes <= 0 || (nFrames > 0 && (!dMsg || !dSegs)) || msgStride < (size_t)frameBytes ||
        segStride < (size_t)(frameBytes * 8 / kc.k + kc.S)) {
        setError("ced_encode_batch_k: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const long long total = (long long)nFrames * (frameBytes * 8 / kc.k + kc.S);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)c->sms * 32);
    ced::kEncodeBatchKernel<<<blocks, 256, 0, s>>>(kc, dMsg, msgStride, nFrames, frameBytes, dSegs, segStride);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int ced_decode_batch_k(ced_ctx *c, const ced_code_t *code, int inputBits, const uint8_t *dSegs, size_t segStride, int nFrames,
                       int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    if (inputBits == 1)
        return ced_decode_batch(c, code, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream);
    ced::KCode kc;
    const int rc = makeKCode(code, inputBits, &kc, "ced_decode_batch_k");
    if (rc != CED_OK)
        return rc;
    if (!c || nFrames < 0 || frameBits <= 0 || (frameBits & 7) || frameBits % kc.k || (nFrames > 0 && (!dSegs || !dOut))) {
        setError("ced_decode_batch_k: bad argument (frameBits must be a multiple of 8 and of k)");
        return CED_ERR_ARG;
    }
    const int T = frameBits / kc.k + kc.S;
    if (segStride < (size_t)T || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch_k: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    if (kc.k == 2) { /* thread-per-frame SIMD-in-word kernels with radix-4 butterflies (swar_radix4.cuh) take n = 2, 3 */
        const int rr = cedDecodeBatchSwarRadix4(c, code, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream);
        if (rr != CED_ERR_UNSUPPORTED)
            return rr;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    ced_ctx::Work &wk = c->work[0];
    const size_t perFrame = (size_t)T * kc.W * sizeof(uint32_t);
    const size_t waveMax = std::max<size_t>(64, std::min<size_t>(c->maxWaveFrames, kMaxScratchBytes / perFrame) / 64 * 64);
    const size_t need = std::min<size_t>((size_t)nFrames, waveMax) * perFrame;
    if (wk.scratch.bytes < need) {
        CED_CUDA(cudaDeviceSynchronize());
        const int rc2 = wk.scratch.ensure(need);
        if (rc2 != CED_OK)
            return rc2;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    const size_t smem = (size_t)((kc.P * kc.N + 15) / 16) * 16 + (size_t)ced::kRkWarps * 2 * kc.N * sizeof(int);
    for (size_t f0 = 0; f0 < (size_t)nFrames; f0 += waveMax) {
        const int wave = (int)std::min<size_t>(waveMax, (size_t)nFrames - f0);
        const int blocks = std::max(1, std::min(c->sms * 8, (wave + ced::kRkWarps - 1) / ced::kRkWarps));
        ced::kForwardKernel<<<blocks, ced::kRkWarps * 32, smem, s>>>(kc, dSegs + f0 * segStride, segStride, wave, T,
                                                                    reinterpret_cast<uint32_t *>(wk.scratch.p));
        ced::kTracebackKernel<<<(wave + 127) / 128, 128, 0, s>>>(kc, reinterpret_cast<const uint32_t *>(wk.scratch.p), wave, T,
                                                               dOut + f0 * outStride, outStride);
        c->launches += 2;
    }
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}
