/*
 * warp_frame.cu -- ced_decode_batch for SMALL batches: one warp per frame, the trellis states spread over the lanes.
 *
 * The thread-per-frame kernels (decode_batch.cuh, swar_generic.cu) need >= 2^11 frames before every sub-partition of the
 * GPU has a warp, and a frame is a chain of T dependent steps of ~121 instructions: 73 ns per step, however few frames
 * there are (the reference's own shapes are 16 packets, speedDecode/speedDecode.c:18-19, and 10,000 packets one by one,
 * berTestK7/berTestK7.c:109-165).  Here a frame belongs to ONE warp and a step is a handful of instructions deep:
 *
 *  forward (src/viterbiDecoderButterflyk1.c:85-196).  Metrics are exact 16-bit integers, two per register.
 *    radix 2 (any k = 1 code with <= 64 states): lane j is butterfly j -- it holds m[j], m[j+H] and produces m'[2j],
 *    m'[2j+1]: A = (m[j], m[j]) + (c(j->2j), c(j->2j+1)), B = (m[j+H], m[j+H]) + (...), m' = min(A, B) per half
 *    (VIMNMX.U16x2), decision = A != min (the lower predecessor wins a tie, :129-130); two independent shuffles and
 *    two PRMT re-gather the operands of the next step.
 *    radix 4 (64 states): two trellis steps per gather.  Lane l = 2q + h holds (m[q], m[q+16]) and (m[q+32], m[q+48]);
 *    step t gives the two intermediate states l and l + 32 (each of the 64 is computed by exactly one lane), step t+1 is
 *    the radix-2 butterfly of those two inside the lane: m''[2l], m''[2l+1].  Four independent shuffles + two PRMT per
 *    TWO steps, so the dependent chain per step is ~26 cycles instead of ~40.
 *    Branch costs come from a per-lane table in shared memory indexed by the received symbol (one LDS.64 per step,
 *    fetched a block of four steps ahead); the frame's symbols are staged in shared memory with 16-byte loads first.
 *  decisions: __ballot_sync-packed, one 8-byte row per step, kept in SHARED memory (never written to HBM).
 *  traceback (:200-256): warp-parallel and exact.  The T steps are cut into 32 segments; lane i walks segment i after a
 *    warm-up of one segment length from state 0 above it.  Paths merge, so nearly always every lane enters its segment
 *    in the state the lane above leaves in; where it does not, that lane walks its segment again from the right state,
 *    until every hand-over agrees (the top lane starts in state 0 at the last step and is exact by construction, so the
 *    loop ends after at most 32 rounds and the result is the sequential walk's, bit for bit).  Bytes are packed MSb
 *    first (:244-249) in shared memory and leave as one row.
 *
 * No global scratch: a CTA is one warp with (9 T + 4.5 K) bytes of shared memory.
 */
#include "warp_frame.cuh"

namespace ced {

template <bool R4>
__global__ void __launch_bounds__(32) wfDecodeKernel(const __grid_constant__ WfArgs a)
{
    extern __shared__ __align__(16) uint8_t wfSmem[];
    uint2 *sSurv = reinterpret_cast<uint2 *>(wfSmem + kWfCostBytes);        /* row of step t at t + t / seg */
    uint8_t *sOut = reinterpret_cast<uint8_t *>(sSurv + a.survRows);
    uint32_t *sOffs = reinterpret_cast<uint32_t *>(sOut + a.outPad);        /* table offset of every unit (+ 12 of padding) */

    const int lane = threadIdx.x, T = a.T, S = a.S, L = T - S, seg = a.seg;
    const int H = 1 << (S - 1);
    const uint32_t rxMask = (1u << a.n) - 1u;
    wfBuildCostTable<R4>(a.cost, wfSmem, lane);

    /* operands of the next step: which lane holds them, which half */
    uint32_t src0, src1, src2 = 0, src3 = 0, sel0, sel1;
    if (R4) {
        const uint32_t q = (uint32_t)lane >> 1;
        src0 = q >> 1;
        src1 = src0 + 8;
        src2 = src0 + 16;
        src3 = src0 + 24;
        sel0 = sel1 = (q & 1u) ? 0x7632u : 0x5410u;
    } else {
        const uint32_t j = (uint32_t)lane & (uint32_t)(H - 1);   /* lanes >= H (fewer than 64 states) mirror a butterfly */
        src0 = j >> 1;
        src1 = (j + (uint32_t)H) >> 1;
        sel0 = (j & 1u) ? 0x3232u : 0x1010u;
        sel1 = ((j + (uint32_t)H) & 1u) ? 0x3232u : 0x1010u;
    }

    for (int f = blockIdx.x; f < a.nFrames; f += gridDim.x) {
        /* stage the symbols (whole 16-byte pieces around the row) where the decisions will go, then turn them into the
         * table offset of every unit -- a unit is a pair of steps (radix 4: 16 B of costs per lane and symbol pair) or a
         * step (radix 2: 8 B per lane and symbol) -- so that the dependent loop below spends one add and one load on it */
        const uintptr_t rowAddr = reinterpret_cast<uintptr_t>(a.segs) + (size_t)f * a.segStride;
        const uint32_t off = (uint32_t)(rowAddr & 15u);
        const uint4 *src = reinterpret_cast<const uint4 *>(rowAddr - off);
        const int nq = (int)(off + (uint32_t)(R4 && a.packed ? (T + 3) >> 2 : T) + 15u) >> 4;
        __syncwarp();
        for (int q = lane; q < nq; q += 32)
            reinterpret_cast<uint4 *>(sSurv)[q] = __ldg(src + q);
        __syncwarp();
        const uint32_t *symW = reinterpret_cast<const uint32_t *>(sSurv) + (off >> 2);
        const uint32_t sh = (off & 3u) * 8u;
        const int units = R4 ? T >> 1 : T;
        for (int g = lane; 4 * g < units + 12; g += 32) {   /* four units per lane and round; padding units get offset 0 */
            uint32_t o[4];
            if (R4 && a.packed) {   /* four symbols to a byte: a nibble is a pair of symbols */
                const uint8_t *symB = reinterpret_cast<const uint8_t *>(sSurv) + off;
                const uint32_t b0 = symB[2 * g], b1 = symB[2 * g + 1];
                o[0] = (b0 & 15u) * 512u;
                o[1] = (b0 >> 4) * 512u;
                o[2] = (b1 & 15u) * 512u;
                o[3] = (b1 >> 4) * 512u;
            } else if (R4) {
                const uint32_t w0 = __funnelshift_r(symW[2 * g], symW[2 * g + 1], sh);
                const uint32_t w1 = __funnelshift_r(symW[2 * g + 1], symW[2 * g + 2], sh);
                o[0] = ((w0 & 3u) | ((w0 >> 6) & 12u)) * 512u;
                o[1] = (((w0 >> 16) & 3u) | ((w0 >> 22) & 12u)) * 512u;
                o[2] = ((w1 & 3u) | ((w1 >> 6) & 12u)) * 512u;
                o[3] = (((w1 >> 16) & 3u) | ((w1 >> 22) & 12u)) * 512u;
            } else {
                const uint32_t w0 = __funnelshift_r(symW[g], symW[g + 1], sh);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    o[k] = ((w0 >> (8 * k)) & rxMask) * 256u;
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (4 * g + k >= units)
                    o[k] = 0;
            reinterpret_cast<uint4 *>(sOffs)[g] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();

        uint32_t X, Y;
        const uint32_t m0 = a.initMetric;
        if (R4) {
            const uint32_t q = (uint32_t)lane >> 1;
            X = (q == 0 ? 0u : m0) | (m0 << 16);
            Y = m0 | (m0 << 16);
        } else {
            const uint32_t lo = (lane & (H - 1)) == 0 ? 0u : m0;
            X = lo | (lo << 16);
            Y = m0 | (m0 << 16);
        }
        using Cost = typename std::conditional<R4, uint4, uint2>::type;
        const uint8_t *costBase = wfSmem + lane * sizeof(Cost);
        auto costAt = [&](uint32_t o) -> Cost { return *reinterpret_cast<const Cost *>(costBase + o); };
        uint2 *rowPtr = sSurv;
        auto unit = [&](const Cost cc) {
            if constexpr (R4) {   /* two steps */
                bool h1, l1, h2, l2;
                const uint32_t A = X + cc.x, B = Y + cc.y;
                const uint32_t I = __vibmin_u16x2(A, B, &h1, &l1);   /* predicate: the first operand is the minimum */
                const uint32_t w0 = __ballot_sync(0xFFFFFFFFu, !l1);
                const uint32_t w1 = __ballot_sync(0xFFFFFFFFu, !h1);
                const uint32_t A2 = __byte_perm(I, 0, 0x1010) + cc.z, B2 = __byte_perm(I, 0, 0x3232) + cc.w;
                const uint32_t O = __vibmin_u16x2(A2, B2, &h2, &l2);
                const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, O, src0);
                const uint32_t v1 = __shfl_sync(0xFFFFFFFFu, O, src1);
                const uint32_t v2 = __shfl_sync(0xFFFFFFFFu, O, src2);
                const uint32_t v3 = __shfl_sync(0xFFFFFFFFu, O, src3);
                const uint32_t wa = __ballot_sync(0xFFFFFFFFu, !l2);
                const uint32_t wb = __ballot_sync(0xFFFFFFFFu, !h2);
                if (lane == 0) {
                    rowPtr[0] = make_uint2(w0, w1);
                    rowPtr[1] = make_uint2(wa, wb);
                }
                rowPtr += 2;
                X = __byte_perm(v0, v1, sel0);
                Y = __byte_perm(v2, v3, sel0);
            } else {
                bool h1, l1;
                const uint32_t A = X + cc.x, B = Y + cc.y;
                const uint32_t N = __vibmin_u16x2(A, B, &h1, &l1);
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, N, src0);
                const uint32_t w = __shfl_sync(0xFFFFFFFFu, N, src1);
                const uint32_t wa = __ballot_sync(0xFFFFFFFFu, !l1);
                const uint32_t wb = __ballot_sync(0xFFFFFFFFu, !h1);
                if (lane == 0)
                    rowPtr[0] = make_uint2(wa, wb);
                rowPtr += 1;
                X = __byte_perm(v, 0, sel0);
                Y = __byte_perm(w, 0, sel1);
            }
        };
        /* blocks of four units; the offsets are fetched two blocks ahead, the costs one */
        const int blocks = units >> 2, blocksPerSeg = seg / (R4 ? 8 : 4);
        int untilBoundary = blocksPerSeg;
        const uint4 *offs4 = reinterpret_cast<const uint4 *>(sOffs);
        uint4 oNext = offs4[1];
        Cost cA[4], cB[4];
        {
            const uint4 o0 = offs4[0];
            cA[0] = costAt(o0.x);
            cA[1] = costAt(o0.y);
            cA[2] = costAt(o0.z);
            cA[3] = costAt(o0.w);
        }
        auto block = [&](int b, const Cost (&cur)[4], Cost (&nxt)[4]) {
            const uint4 o2 = offs4[b + 2];   /* padded: stays inside the array */
            nxt[0] = costAt(oNext.x);
            nxt[1] = costAt(oNext.y);
            nxt[2] = costAt(oNext.z);
            nxt[3] = costAt(oNext.w);
            unit(cur[0]);
            unit(cur[1]);
            unit(cur[2]);
            unit(cur[3]);
            oNext = o2;
            if (--untilBoundary == 0) {   /* the next segment's rows start one slot further on */
                rowPtr++;
                untilBoundary = blocksPerSeg;
            }
        };
        int b = 0;
        for (; b + 2 <= blocks; b += 2) {
            block(b, cA, cB);
            block(b + 1, cB, cA);
        }
        if (b < blocks) {
            block(b, cA, cB);
#pragma unroll
            for (int k = 0; k < 4; k++)
                cA[k] = cB[k];
            b++;
        }
        for (int u = 4 * b, k = 0; u < units; u++, k++)   /* at most three units; cA holds their costs */
            unit(k == 0 ? cA[0] : k == 1 ? cA[1] : cA[2]);
        __syncwarp();

        wfTraceback<R4>(sSurv, T, S, seg, sOut, lane);   /* warp-parallel, exact */
        __syncwarp();
        uint8_t *dst = a.out + (size_t)f * a.outStride;
        for (int i = lane; i < (L >> 3); i += 32)
            dst[i] = sOut[i];
    }
}

} // namespace ced

/* bytes of shared memory a frame of T steps needs */
static size_t wfSmemBytes(int T, int seg, bool r4, int *survRows, int *outPad)
{
    *survRows = (T + T / seg + 9) & ~1;   /* even: what follows stays 16-byte aligned; also holds the staged symbols (T + 46 bytes) */
    *outPad = ((T >> 3) + 15) / 16 * 16 + 16;
    const size_t units = r4 ? (size_t)T / 2 : (size_t)T;
    return (size_t)ced::kWfCostBytes + (size_t)*survRows * sizeof(uint2) + (size_t)*outPad + (units + 12 + 3) / 4 * 16;
}

/* what a call would run as: radix, traceback segment, shared memory, resident CTAs per SM */
struct WfPlan {
    bool ok, r4;
    int seg, survRows, outPad, perSm;
    size_t smem;
};

static WfPlan wfPlan(const ced_code_t *code, int frameBits)
{
    WfPlan p = {};
    if (!code || code->constraintLen < 3 || code->constraintLen > 7 || code->codedBits < 1 || code->codedBits > 3 || frameBits <= 0)
        return p;
    const int K = code->constraintLen, n = code->codedBits, T = frameBits + K - 1;
    const char *envAnyK = getenv("CED_WARP_FRAME_ANY_K");   /* read per call: tests flip them */
    /* fewer than 32 states leave most lanes idle, and the thread-per-frame kernels of such codes are short chains
     * themselves (K = 5: ~24 ns per step against ~30 here): only K = 6 and 7 by default */
    if (K < 6 && !(envAnyK && atoi(envAnyK) != 0))
        return p;
    if ((size_t)(uint8_t)((1 << (K - 1)) + 1) + (size_t)n * (size_t)T > 65535u)   /* exact 16-bit metrics without renormalisation */
        return p;
    const char *envRadix = getenv("CED_WARP_FRAME_RADIX");
    p.r4 = K == 7 && n == 2 && !(envRadix && atoi(envRadix) == 2);
    p.seg = (((T + 31) / 32) + 7) & ~7;
    p.smem = wfSmemBytes(T, p.seg, p.r4, &p.survRows, &p.outPad);
    if (p.smem > 200 * 1024)
        return p;
    p.perSm = (int)std::max<size_t>(1, std::min<size_t>(32, (size_t)(220 * 1024) / (p.smem + 1024)));
    p.ok = true;
    return p;
}

/* branch costs per lane and received symbol, four bytes per entry (WfArgs::cost) */
void cedWarpFrameCosts(const ced_code_t *code, bool r4, uint32_t (&cost)[2][ced::kWfMaxV][32])
{
    const int K = code->constraintLen, n = code->codedBits, H = (1 << (K - 1)) / 2;
    /* trellis labels as viterbiInit builds them (src/viterbiDecoder.c:32-50): edge[b][st] = coded segment of the branch
     * that leaves state st with input bit b */
    uint32_t taps[3] = {0, 0, 0};
    for (int i = 0; i < n; i++)
        taps[i] = reverseBits(code->gen[i], K);
    auto edge = [&](int b, int st) -> uint32_t {
        const uint32_t reg = (((uint32_t)st << 1) | (uint32_t)b) & ((1u << K) - 1u);
        uint32_t segv = 0;
        for (int i = 0; i < n; i++)
            segv |= (uint32_t)(__builtin_popcount(reg & taps[i]) & 1) << i;
        return segv;
    };
    auto hd = [&](uint32_t label, uint32_t rx) -> uint32_t { return (uint32_t)__builtin_popcount((label ^ rx) & ((1u << n) - 1u)); };
    memset(cost, 0, sizeof(cost));
    for (int rx = 0; rx < (1 << n); rx++)
        for (int l = 0; l < 32; l++) {
            /* second step of a pair / the radix-2 step: butterfly j = l: (j -> 2j, j -> 2j+1), (j+H -> 2j, j+H -> 2j+1) */
            const int j = l & (H - 1);
            cost[1][rx][l] = hd(edge(0, j), rx) | hd(edge(1, j), rx) << 8 | hd(edge(0, j + H), rx) << 16 | hd(edge(1, j + H), rx) << 24;
            if (r4) {
                /* first step: lane l = 2q + h makes states l (from q, q+32) and l + 32 (from q+16, q+48); input bit h */
                const int q = l >> 1, h = l & 1;
                cost[0][rx][l] = hd(edge(h, q), rx) | hd(edge(h, q + 16), rx) << 8 | hd(edge(h, q + 32), rx) << 16 |
                                   hd(edge(h, q + 48), rx) << 24;
            }
        }
}

/*
 * Whether ced_decode_batch hands a batch to this kernel: k = 1 codes with 32 or 64 states and n <= 3, byte format, and
 * few enough frames -- at most three (radix 4; radix 2: two) rounds of resident one-warp CTAs.  Measured on B200
 * (tools/small_batch_throughput.py, profiles/small_batch_r2.txt): 2048-bit frames 58 us up to 592 frames, 75 us at
 * 1024, 135 us at 2048 against 285 .. 320 us on the thread-per-frame kernels; 4096 frames: 258 against 325 us.
 * CED_WARP_FRAME_MAX = n overrides the frame limit (0 = never).
 */
bool cedWarpFrameTakes(const ced_ctx *c, const ced_code_t *code, int nFrames, int frameBits, bool packed)
{
    if (!c || nFrames <= 0)
        return false;
    const WfPlan p = wfPlan(code, frameBits);
    if (!p.ok || (packed && !p.r4))
        return false;
    if (const char *e = getenv("CED_WARP_FRAME_MAX"))
        return nFrames <= atoi(e);
    return nFrames <= (p.r4 ? 3 : 2) * c->sms * p.perSm;
}

/* ced_decode_batch for a batch cedWarpFrameTakes() said yes to (CED_ERR_UNSUPPORTED otherwise) */
int cedDecodeBatchWarpFrame(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                            int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot, bool packed)
{
    const WfPlan plan = wfPlan(code, frameBits);
    if (!c || !plan.ok || (packed && !plan.r4))
        return CED_ERR_UNSUPPORTED;
    if (plan.r4) {   /* so few frames that cutting them in time as well pays (warp_split.cu) */
        const int rc = cedDecodeBatchWarpSplit(c, code, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream, slot, packed);
        if (rc != CED_ERR_UNSUPPORTED)
            return rc;
    }
    const int K = code->constraintLen, S = K - 1, n = code->codedBits, N = 1 << S;
    const int T = frameBits + S;
    const uint32_t init = (uint32_t)(uint8_t)(N + 1);
    if (segStride < (packed ? (size_t)(T + 3) / 4 : (size_t)T) || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    ced::WfArgs a;
    a.packed = packed ? 1 : 0;
    a.seg = plan.seg;
    a.survRows = plan.survRows;
    a.outPad = plan.outPad;
    const bool r4 = plan.r4;
    const size_t smem = plan.smem;
    if (nFrames == 0)
        return CED_OK;
    cedWarpFrameCosts(code, r4, a.cost);
    a.segs = dSegs;
    a.segStride = segStride;
    a.out = dOut;
    a.outStride = outStride;
    a.nFrames = nFrames;
    a.T = T;
    a.S = S;
    a.n = n;
    a.initMetric = init;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    auto kernel = r4 ? ced::wfDecodeKernel<true> : ced::wfDecodeKernel<false>;
    CED_CUDA(cedWarpEnsureSmem(c->device, r4 ? 0 : 1, kernel, smem));
    const int blocks = std::min(nFrames, c->sms * plan.perSm);
    kernel<<<blocks, 32, smem, s>>>(a);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

