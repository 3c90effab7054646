/*
 * warp_split.cu -- few frames, many SMs: a frame is cut into blocks IN TIME and every block gets its own warp.
 *
 * warp_frame.cu gives a frame to one warp: 21 cycles per trellis step, 53 us for a 2048-bit frame however idle the rest
 * of the GPU is.  The forward recursion (src/viterbiDecoderButterflyk1.c:85-196) is sequential only through its 64 path
 * metrics, and only their DIFFERENCES matter: add-compare-select compares sums of metrics and branch costs, so two runs
 * whose metric vectors differ by a constant take the same decisions from there on, ties included (:129-130).  After a
 * few constraint lengths the differences no longer depend on where a run started -- the survivors have merged -- which
 * is what every sliding-window decoder relies on.  Here it is used speculatively and CHECKED, so the result stays exact:
 *
 *  wsBlockKernel, one warp per (frame, block c): starts D steps before its block from all-equal metrics (blocks that
 *    reach back to step 0 start from the reference's initial metrics and are exact by construction), runs the warm-up
 *    without recording anything, stores its metric vector minus its minimum at the block start, runs the block with the
 *    radix-4 step of warp_frame.cu -- decisions as ballot rows to global memory --, stores the vector at the block end.
 *  wsJoinKernel, one warp per frame: block c is exact if block c-1 is and the vector c started its block with equals the
 *    one c-1 ended with -- then c took the very decisions the sequential decoder takes.  All hand-overs are compared at
 *    once; where one differs (rare at useful noise levels, common in pure noise) that block is run again from the true
 *    vector, its end vector is replaced and the next hand-over compared again.  Then the frame's rows are staged in
 *    shared memory and walked back by the warp-parallel traceback of warp_frame.cuh (:200-256).
 *
 * 64-state rate-1/2 codes (the reference's default parameters), byte format.
 */
#include "warp_frame.cuh"

namespace ced {

struct WsArgs {
    const uint8_t *segs;
    size_t segStride;
    uint8_t *out;
    size_t outStride;
    int nFrames, T;
    int packed;             /* symbols four to a byte (ced_decode_batch_packed) */
    int len, B, D;          /* steps per block (a multiple of 8), blocks per frame, warm-up steps (a multiple of 8) */
    int seg, survRows, outPad;   /* traceback geometry of the join kernel (as WfArgs) */
    int rowPitch;           /* decision rows per frame in `rows` */
    uint2 *rows;            /* [nFrames][rowPitch] */
    uint2 *vecs;            /* [nFrames][B][2][32]: (X, Y) of every lane at the block start / at the block end, minimum 0 */
    unsigned int *doneCounter;            /* one-packet calls: device counter and the host-visible word that receives its */
    volatile unsigned int *doneFlag;      /* new value when the packet's bytes are out (NULL: no doorbell) */
    uint2 start[32];        /* (X, Y) of every lane at step 0: the frame's initial metrics (:59-67, or the caller's) */
    uint32_t cost[2][kWfMaxV][32];
};

/* the radix-4 forward recursion of one warp (warp_frame.cu), as an object so that both kernels can run pieces of it */
struct WsForward {
    uint32_t X, Y, src0, src1, src2, src3, sel;
    const uint8_t *costBase;
    int lane;

    __device__ __forceinline__ void setup(int l, const uint8_t *costTable)
    {
        lane = l;
        const uint32_t q = (uint32_t)l >> 1;
        src0 = q >> 1;
        src1 = src0 + 8;
        src2 = src0 + 16;
        src3 = src0 + 24;
        sel = (q & 1u) ? 0x7632u : 0x5410u;
        costBase = costTable + l * sizeof(uint4);
    }
    __device__ __forceinline__ uint4 costAt(uint32_t o) const { return *reinterpret_cast<const uint4 *>(costBase + o); }

    template <bool DEC>
    __device__ __forceinline__ void unit(const uint4 cc, uint2 *&rowPtr)
    {
        bool h1, l1, h2, l2;
        const uint32_t A = X + cc.x, B = Y + cc.y;
        const uint32_t I = __vibmin_u16x2(A, B, &h1, &l1);
        const uint32_t A2 = __byte_perm(I, 0, 0x1010) + cc.z, B2 = __byte_perm(I, 0, 0x3232) + cc.w;
        const uint32_t O = __vibmin_u16x2(A2, B2, &h2, &l2);
        const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, O, src0);
        const uint32_t v1 = __shfl_sync(0xFFFFFFFFu, O, src1);
        const uint32_t v2 = __shfl_sync(0xFFFFFFFFu, O, src2);
        const uint32_t v3 = __shfl_sync(0xFFFFFFFFu, O, src3);
        if (DEC) {
            const uint32_t w0 = __ballot_sync(0xFFFFFFFFu, !l1);
            const uint32_t w1 = __ballot_sync(0xFFFFFFFFu, !h1);
            const uint32_t wa = __ballot_sync(0xFFFFFFFFu, !l2);
            const uint32_t wb = __ballot_sync(0xFFFFFFFFu, !h2);
            if (lane == 0)
                *reinterpret_cast<uint4 *>(rowPtr) = make_uint4(w0, w1, wa, wb);   /* rows of an even step are 16-byte aligned */
            rowPtr += 2;
        }
        X = __byte_perm(v0, v1, sel);
        Y = __byte_perm(v2, v3, sel);
    }

    /* `units` pairs of steps whose table offsets start at offs (16-byte aligned, >= 12 entries of padding behind them) */
    template <bool DEC>
    __device__ __forceinline__ void run(const uint32_t *offs, int units, uint2 *&rowPtr)
    {
        const uint4 *offs4 = reinterpret_cast<const uint4 *>(offs);
        const int blocks = units >> 2;
        uint4 oNext = offs4[1];
        uint4 cA[4], cB[4];
        {
            const uint4 o0 = offs4[0];
            cA[0] = costAt(o0.x);
            cA[1] = costAt(o0.y);
            cA[2] = costAt(o0.z);
            cA[3] = costAt(o0.w);
        }
        auto block = [&](int b, const uint4 (&cur)[4], uint4 (&nxt)[4]) {
            const uint4 o2 = offs4[b + 2];
            nxt[0] = costAt(oNext.x);
            nxt[1] = costAt(oNext.y);
            nxt[2] = costAt(oNext.z);
            nxt[3] = costAt(oNext.w);
            unit<DEC>(cur[0], rowPtr);
            unit<DEC>(cur[1], rowPtr);
            unit<DEC>(cur[2], rowPtr);
            unit<DEC>(cur[3], rowPtr);
            oNext = o2;
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
        for (int u = 4 * b, k = 0; u < units; u++, k++)
            unit<DEC>(k == 0 ? cA[0] : k == 1 ? cA[1] : cA[2], rowPtr);
    }

    /* subtract the smallest of the 64 metrics: differences are all that matters, and this makes vectors comparable */
    __device__ __forceinline__ void normalise()
    {
        uint32_t m = __vminu2(X, Y);
        m = min(m & 0xFFFFu, m >> 16);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        m |= m << 16;
        X -= m;   /* every half is >= m: no borrow between the halves */
        Y -= m;
    }
};

/* stage the symbols of steps [lo, lo + nSteps) of a row in `stage` (16-byte aligned shared memory of nSteps + 64 bytes)
 * and write the table offset of every pair of steps to offs (nSteps / 2 entries + 12 zeros) */
/* first: the lane's first 16-byte piece of those symbols, fetched by the caller ahead of time (wsFirstPiece), or NULL */
__device__ __forceinline__ uint4 wsFirstPiece(const uint8_t *row, int lo, int nSteps, int lane, int packed)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(row) + (size_t)(packed ? lo >> 2 : lo);   /* lo is a multiple of 8 */
    const uint32_t off = (uint32_t)(addr & 15u);
    const int nq = (int)(off + (uint32_t)(packed ? (nSteps + 3) >> 2 : nSteps) + 15u) >> 4;
    return lane < nq ? __ldg(reinterpret_cast<const uint4 *>(addr - off) + lane) : make_uint4(0, 0, 0, 0);
}

__device__ __forceinline__ void wsOffsets(const uint8_t *row, int lo, int nSteps, uint32_t *stage, uint32_t *offs, int lane,
                                          int packed, const uint4 *first = nullptr)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(row) + (size_t)(packed ? lo >> 2 : lo);
    const uint32_t off = (uint32_t)(addr & 15u);
    const uint4 *src = reinterpret_cast<const uint4 *>(addr - off);
    const int nq = (int)(off + (uint32_t)(packed ? (nSteps + 3) >> 2 : nSteps) + 15u) >> 4;
    __syncwarp();
    if (first && lane < nq)
        reinterpret_cast<uint4 *>(stage)[lane] = *first;
    for (int q = lane + (first ? 32 : 0); q < nq; q += 32)
        reinterpret_cast<uint4 *>(stage)[q] = __ldg(src + q);
    __syncwarp();
    const uint32_t *symW = stage + (off >> 2);
    const uint32_t sh = (off & 3u) * 8u;
    const int units = nSteps >> 1;
    for (int g = lane; 4 * g < units + 12; g += 32) {
        uint32_t o[4];
        if (packed) {   /* a nibble is a pair of symbols */
            const uint8_t *symB = reinterpret_cast<const uint8_t *>(stage) + off;
            const uint32_t b0 = symB[2 * g], b1 = symB[2 * g + 1];
            o[0] = (b0 & 15u) * 512u;
            o[1] = (b0 >> 4) * 512u;
            o[2] = (b1 & 15u) * 512u;
            o[3] = (b1 >> 4) * 512u;
        } else {
            const uint32_t w0 = __funnelshift_r(symW[2 * g], symW[2 * g + 1], sh);
            const uint32_t w1 = __funnelshift_r(symW[2 * g + 1], symW[2 * g + 2], sh);
            o[0] = ((w0 & 3u) | ((w0 >> 6) & 12u)) * 512u;
            o[1] = (((w0 >> 16) & 3u) | ((w0 >> 22) & 12u)) * 512u;
            o[2] = ((w1 & 3u) | ((w1 >> 6) & 12u)) * 512u;
            o[3] = (((w1 >> 16) & 3u) | ((w1 >> 22) & 12u)) * 512u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (4 * g + k >= units)
                o[k] = 0;
        reinterpret_cast<uint4 *>(offs)[g] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncwarp();
}

/* shared memory of both kernels: cost table, then a region each kernel lays out itself */
__global__ void __launch_bounds__(32) wsBlockKernel(const __grid_constant__ WsArgs a)
{
    extern __shared__ __align__(16) uint8_t wsSmem[];
    const int lane = threadIdx.x;
    /* programmatic dependent launch: the join kernel may be scheduled now; its griddepcontrol.wait still holds it until
     * this grid has finished and its writes are visible */
    asm volatile("griddepcontrol.launch_dependents;");
    const int span = a.D + a.len;                                        /* most steps a block runs */
    uint32_t *stage = reinterpret_cast<uint32_t *>(wsSmem + kWfCostBytes);   /* span + 64 bytes */
    uint32_t *offs = stage + ((span + 64 + 15) / 16) * 4;                 /* span / 2 + 12 entries */
    const int total = a.nFrames * a.B;
    uint4 first = make_uint4(0, 0, 0, 0);
    if ((int)blockIdx.x < total) {   /* the first block's symbols are on their way while the cost table is built */
        const int f = blockIdx.x / a.B, c = blockIdx.x - f * a.B;
        const int s = c * a.len, e = min(a.T, s + a.len), lo = max(0, s - a.D);
        first = wsFirstPiece(a.segs + (size_t)f * a.segStride, lo, e - lo, lane, a.packed);
    }
    wfBuildCostTable<true>(a.cost, wsSmem, lane);
    WsForward fwd;
    fwd.setup(lane, wsSmem);
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int f = w / a.B, c = w - f * a.B;
        const int s = c * a.len, e = min(a.T, s + a.len), lo = max(0, s - a.D);
        wsOffsets(a.segs + (size_t)f * a.segStride, lo, e - lo, stage, offs, lane, a.packed, w == (int)blockIdx.x ? &first : nullptr);
        if (lo == 0) {
            fwd.X = a.start[lane].x;
            fwd.Y = a.start[lane].y;
        } else {
            fwd.X = fwd.Y = 0;
        }
        uint2 *rowPtr = a.rows + (size_t)f * a.rowPitch + s;
        uint2 *vec = a.vecs + ((size_t)f * a.B + c) * 64;
        if (s > lo)
            fwd.run<false>(offs, (s - lo) >> 1, rowPtr);
        fwd.normalise();
        vec[lane] = make_uint2(fwd.X, fwd.Y);
        fwd.run<true>(offs + ((s - lo) >> 1), (e - s) >> 1, rowPtr);
        fwd.normalise();
        vec[32 + lane] = make_uint2(fwd.X, fwd.Y);
    }
}

constexpr int kWsJoinThreads = 256;   /* all warps compare the hand-overs and move the rows; warp 0 repairs and walks back */
constexpr int kWsJoinLoads = 6;       /* 16-byte loads a thread keeps in flight */

__global__ void __launch_bounds__(kWsJoinThreads) wsJoinKernel(const __grid_constant__ WsArgs a)
{
    extern __shared__ __align__(16) uint8_t wsSmem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = a.T, B = a.B;
    uint2 *sSurv = reinterpret_cast<uint2 *>(wsSmem + kWfCostBytes);
    uint8_t *sOut = reinterpret_cast<uint8_t *>(sSurv + a.survRows);
    uint32_t *offs = reinterpret_cast<uint32_t *>(sOut + a.outPad);       /* len / 2 + 12 entries */
    uint8_t *sBad = reinterpret_cast<uint8_t *>(offs + ((a.len / 2 + 12 + 3) / 4) * 4);   /* B flags */
    const uint64_t segMagic = ((1ull << 32) + (uint64_t)a.seg - 1) / (uint64_t)a.seg;   /* t / seg = (t * magic) >> 32 for t < 2^16 */
    asm volatile("griddepcontrol.wait;" ::: "memory");   /* the block kernel's rows and vectors (no-op without the launch attribute) */
    bool haveTable = false;
    WsForward fwd;
    for (int f = blockIdx.x; f < a.nFrames; f += gridDim.x) {
        __syncthreads();
        uint2 *vecs = a.vecs + (size_t)f * B * 64;
        /* all hand-overs at once: what block c started with against what block c-1 ended with, 16 bytes per thread and
         * piece (a block that reaches back to step 0 started from the true metrics: nothing to compare) */
        for (int c = tid; c < B; c += kWsJoinThreads)
            sBad[c] = 0;
        __syncthreads();
        const int cFirst = a.D / a.len + 1;   /* first block with c * len - D > 0 */
        const int pieces = B > cFirst ? (B - cFirst) * 16 : 0;
        for (int k0 = tid; k0 < pieces; k0 += kWsJoinLoads * kWsJoinThreads) {
            uint4 u[kWsJoinLoads], v[kWsJoinLoads];
#pragma unroll
            for (int j = 0; j < kWsJoinLoads; j++) {
                const int k = k0 + j * kWsJoinThreads;
                if (k < pieces) {
                    const int c = cFirst + (k >> 4), i = k & 15;
                    u[j] = __ldcg(reinterpret_cast<const uint4 *>(vecs + (size_t)c * 64) + i);
                    v[j] = __ldcg(reinterpret_cast<const uint4 *>(vecs + (size_t)(c - 1) * 64 + 32) + i);
                }
            }
#pragma unroll
            for (int j = 0; j < kWsJoinLoads; j++) {
                const int k = k0 + j * kWsJoinThreads;
                if (k < pieces && ((u[j].x != v[j].x) | (u[j].y != v[j].y) | (u[j].z != v[j].z) | (u[j].w != v[j].w)))
                    sBad[cFirst + (k >> 4)] = 1;
            }
        }
        __syncthreads();
        if (warp == 0) {
            __syncwarp();
            for (int c = 1; c < B; c++) {
                if (!sBad[c])
                    continue;
                /* block c did not start where block c-1 ended: run it again from there */
                if (!haveTable) {
                    wfBuildCostTable<true>(a.cost, wsSmem, lane);
                    fwd.setup(lane, wsSmem);
                    haveTable = true;
                }
                const int s = c * a.len, e = min(T, s + a.len);
                wsOffsets(a.segs + (size_t)f * a.segStride, s, e - s, reinterpret_cast<uint32_t *>(sSurv), offs, lane, a.packed);
                const uint2 from = __ldcg(vecs + (size_t)(c - 1) * 64 + 32 + lane);
                fwd.X = from.x;
                fwd.Y = from.y;
                uint2 *rowPtr = a.rows + (size_t)f * a.rowPitch + s;
                fwd.run<true>(offs, (e - s) >> 1, rowPtr);
                fwd.normalise();
                const uint2 before = __ldcg(vecs + (size_t)c * 64 + 32 + lane);
                const bool changed = __any_sync(0xFFFFFFFFu, before.x != fwd.X || before.y != fwd.Y);
                if (changed) {
                    vecs[(size_t)c * 64 + 32 + lane] = make_uint2(fwd.X, fwd.Y);
                    if (c + 1 < B && (c + 1) * a.len - a.D > 0) {
                        const uint2 next = __ldcg(vecs + (size_t)(c + 1) * 64 + lane);
                        const bool differs = __any_sync(0xFFFFFFFFu, next.x != fwd.X || next.y != fwd.Y);
                        if (lane == 0)
                            sBad[c + 1] = differs ? 1 : 0;
                    }
                }
                __syncwarp();
            }
            __threadfence_block();
        }
        __syncthreads();
        /* the frame's decisions into shared memory, two rows per load, six loads in flight per thread; the rows of
         * segment i are skewed by i slots (wfTraceback) */
        const uint4 *rows2 = reinterpret_cast<const uint4 *>(a.rows + (size_t)f * a.rowPitch);
        const int pairs = (T + 1) >> 1;
        for (int i0 = tid; i0 < pairs; i0 += kWsJoinLoads * kWsJoinThreads) {
            uint4 r[kWsJoinLoads];
#pragma unroll
            for (int k = 0; k < kWsJoinLoads; k++)
                if (i0 + k * kWsJoinThreads < pairs)
                    r[k] = __ldcg(rows2 + i0 + k * kWsJoinThreads);
#pragma unroll
            for (int k = 0; k < kWsJoinLoads; k++) {
                const int i = i0 + k * kWsJoinThreads;
                if (i < pairs) {
                    const int t = 2 * i, sg = (int)(((uint64_t)t * segMagic) >> 32);   /* seg is even: both rows in one segment */
                    sSurv[t + sg] = make_uint2(r[k].x, r[k].y);
                    sSurv[t + sg + 1] = make_uint2(r[k].z, r[k].w);
                }
            }
        }
        __syncthreads();
        if (warp == 0)
            wfTraceback<true>(sSurv, T, 6, a.seg, sOut, lane);
        __syncthreads();
        uint8_t *dst = a.out + (size_t)f * a.outStride;
        const int nOut = (T - 6) >> 3;
        const int nVec = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 ? nOut >> 4 : 0;   /* whole 16-byte pieces of an aligned row */
        for (int i = tid; i < nVec; i += kWsJoinThreads)
            reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(sOut)[i];
        for (int i = 16 * nVec + tid; i < nOut; i += kWsJoinThreads)
            dst[i] = sOut[i];
    }
    if (a.doneFlag) {   /* one frame, one CTA: ring the doorbell once every byte is on its way */
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            *a.doneFlag = atomicAdd(a.doneCounter, 1u) + 1u;
            __threadfence_system();
        }
    }
}

} // namespace ced

/* block geometry, shared memory and scratch of a call; ok = false: not a case for these kernels */
struct WsPlan {
    bool ok;
    int T, len, B, D, seg, survRows, outPad, rowPitch;
    size_t smemBlock, smemJoin, rowBytes, vecBytes;
};

static WsPlan wsPlan(const ced_ctx *c, int nFrames, int T)
{
    WsPlan p = {};
    const char *envOn = getenv("CED_WARP_SPLIT");   /* read per call: tests flip them */
    if (!c || nFrames <= 0 || T < 16 || (T & 1) || ((T - 6) & 7) || (envOn && atoi(envOn) == 0))
        return p;
    const char *envD = getenv("CED_WARP_SPLIT_WARMUP"), *envLen = getenv("CED_WARP_SPLIT_LEN");
    p.T = T;
    p.D = envD && atoi(envD) >= 8 ? atoi(envD) / 8 * 8 : 96;
    const int maxBlocks = std::max(1, 4 * c->sms / nFrames);
    p.len = std::max(64, ((T + maxBlocks - 1) / maxBlocks + 7) / 8 * 8);
    if (envLen && atoi(envLen) >= 8)
        p.len = atoi(envLen) / 8 * 8;
    p.B = (T + p.len - 1) / p.len;
    if (p.B < 2 || (size_t)65 + 2u * (size_t)T > 65535u)
        return p;
    p.seg = (((T + 31) / 32) + 7) & ~7;
    p.survRows = (T + T / p.seg + 9) & ~1;
    p.outPad = ((T >> 3) + 15) / 16 * 16 + 16;
    p.rowPitch = (T + 15) / 8 * 8;
    p.smemBlock = (size_t)ced::kWfCostBytes + (size_t)((p.D + p.len + 64 + 15) / 16) * 16 + (size_t)((p.D + p.len) / 2 + 12 + 3) / 4 * 16;
    p.smemJoin = (size_t)ced::kWfCostBytes + (size_t)p.survRows * sizeof(uint2) + (size_t)p.outPad +
                 (size_t)(p.len / 2 + 12 + 3) / 4 * 16 + (size_t)(p.B + 15) / 16 * 16;
    if (p.smemJoin > 200 * 1024 || p.smemBlock > 200 * 1024 || (size_t)p.survRows * sizeof(uint2) < (size_t)p.len + 64)
        return p;
    p.rowBytes = (size_t)nFrames * p.rowPitch * sizeof(uint2);
    p.vecBytes = (size_t)nFrames * p.B * 64 * sizeof(uint2);
    p.ok = true;
    return p;
}

/* the two launches; `scratch` holds rowBytes + vecBytes */
static int wsLaunch(ced_ctx *c, const WsPlan &p, ced::WsArgs &a, const uint8_t *dSegs, size_t segStride, int nFrames, uint8_t *dOut,
                    size_t outStride, void *scratch, cudaStream_t s)
{
    a.segs = dSegs;
    a.segStride = segStride;
    a.out = dOut;
    a.outStride = outStride;
    a.nFrames = nFrames;
    a.T = p.T;
    a.len = p.len;
    a.B = p.B;
    a.D = p.D;
    a.seg = p.seg;
    a.survRows = p.survRows;
    a.outPad = p.outPad;
    a.rowPitch = p.rowPitch;
    if (nFrames != 1)
        a.doneFlag = nullptr;
    a.rows = reinterpret_cast<uint2 *>(scratch);
    a.vecs = reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(scratch) + p.rowBytes);
    CED_CUDA(cedWarpEnsureSmem(c->device, 2, ced::wsBlockKernel, p.smemBlock));
    CED_CUDA(cedWarpEnsureSmem(c->device, 3, ced::wsJoinKernel, p.smemJoin));
    ced::wsBlockKernel<<<std::min(nFrames * p.B, c->sms * 8), 32, p.smemBlock, s>>>(a);
    /* the join kernel is launched as a programmatic dependent of the block kernel: it is resident and past its
     * prologue when the last block finishes (CED_WARP_SPLIT_PDL=0: an ordinary launch) */
    static const bool pdl = !getenv("CED_WARP_SPLIT_PDL") || atoi(getenv("CED_WARP_SPLIT_PDL")) != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min(nFrames, c->sms * 4));
    cfg.blockDim = dim3(ced::kWsJoinThreads);
    cfg.dynamicSmemBytes = p.smemJoin;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CED_CUDA(cudaLaunchKernelEx(&cfg, ced::wsJoinKernel, a));
    return CED_OK;
}

/*
 * ced_decode_batch for so few frames of a 64-state rate-1/2 code that cutting them in time pays (called by
 * cedDecodeBatchWarpFrame): blocks of >= 64 steps, as many as fill four warps per SM; CED_ERR_UNSUPPORTED = not a case.
 * CED_WARP_SPLIT = 0 switches it off, CED_WARP_SPLIT_WARMUP / CED_WARP_SPLIT_LEN set the warm-up / block length.
 */
int cedDecodeBatchWarpSplit(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                            int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot, bool packed)
{
    if (!c || !code || code->constraintLen != 7 || code->codedBits != 2 || nFrames <= 0 || frameBits <= 0)
        return CED_ERR_UNSUPPORTED;
    const WsPlan p = wsPlan(c, nFrames, frameBits + 6);
    if (!p.ok)
        return CED_ERR_UNSUPPORTED;
    if (segStride < (packed ? (size_t)(p.T + 3) / 4 : (size_t)p.T) || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    ced_ctx::Work &wk = c->work[slot];
    if (wk.scratch.bytes < p.rowBytes + p.vecBytes) {
        CED_CUDA(cudaDeviceSynchronize());   /* growing means freeing: nothing may still use the old block */
        const int rc = wk.scratch.ensure(p.rowBytes + p.vecBytes);
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    ced::WsArgs a;
    a.doneCounter = nullptr;
    a.doneFlag = nullptr;
    a.packed = packed ? 1 : 0;
    cedWarpFrameCosts(code, true, a.cost);
    for (int l = 0; l < 32; l++)   /* :59-67: state 0 starts at 0, every other state at NUM_STATES + 1 */
        a.start[l] = make_uint2((l >> 1) == 0 ? 65u << 16 : 65u | 65u << 16, 65u | 65u << 16);
    const int rc = wsLaunch(c, p, a, dSegs, segStride, nFrames, dOut, outStride, wk.scratch.p, s);
    if (rc != CED_OK)
        return rc;
    c->launches += 2;
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

/*
 * The one-packet call of the reference-named API (ced_stream_decode, last = true from reset-like metrics): the same two
 * kernels on ONE frame whose trellis labels and initial metrics are the caller's (edge[b * 64 + st], metrics[st] <= 65).
 * dSegs / dOut are device-accessible (the pinned mailboxes); nothing is synchronised here, so the pair can be captured
 * in a graph.  CED_ERR_UNSUPPORTED = not a case (odd length, a length that is not whole bytes + 6, too short).
 */
int cedStreamDecodeSplit(ced_ctx *c, const uint8_t *edge, const uint8_t *metrics, const uint8_t *dSegs, int T, uint8_t *dOut,
                         void *scratch, size_t scratchBytes, cudaStream_t s, unsigned int *doneCounter, volatile unsigned int *doneFlag)
{
    const WsPlan p = wsPlan(c, 1, T);
    if (!p.ok || p.rowBytes + p.vecBytes > scratchBytes)
        return CED_ERR_UNSUPPORTED;
    ced::WsArgs a;
    a.doneCounter = doneCounter;
    a.doneFlag = doneFlag;
    a.packed = 0;
    auto hd = [](uint32_t label, uint32_t rx) -> uint32_t { return (uint32_t)__builtin_popcount((label ^ rx) & 3u); };
    memset(a.cost, 0, sizeof(a.cost));
    for (uint32_t rx = 0; rx < 4; rx++)
        for (int l = 0; l < 32; l++) {   /* as cedWarpFrameCosts, labels from the caller's table */
            const int q = l >> 1, h = l & 1;
            a.cost[1][rx][l] = hd(edge[l], rx) | hd(edge[64 + l], rx) << 8 | hd(edge[l + 32], rx) << 16 | hd(edge[64 + l + 32], rx) << 24;
            a.cost[0][rx][l] = hd(edge[h * 64 + q], rx) | hd(edge[h * 64 + q + 16], rx) << 8 | hd(edge[h * 64 + q + 32], rx) << 16 |
                               hd(edge[h * 64 + q + 48], rx) << 24;
        }
    for (int l = 0; l < 32; l++) {
        const int q = l >> 1;
        a.start[l] = make_uint2((uint32_t)metrics[q] | (uint32_t)metrics[q + 16] << 16, (uint32_t)metrics[q + 32] | (uint32_t)metrics[q + 48] << 16);
    }
    return wsLaunch(c, p, a, dSegs, (size_t)T, 1, dOut, 0, scratch, s);
}

bool cedStreamDecodeSplitTakes(const ced_ctx *c, int T)
{
    return wsPlan(c, 1, T).ok;
}

size_t cedStreamDecodeSplitScratchBytes(int maxSteps)
{
    /* one frame: rows + at most maxSteps / 64 + 1 blocks of two 256-byte vectors (smaller blocks only by experiment switch) */
    return (size_t)(maxSteps + 16) * sizeof(uint2) + (size_t)(maxSteps / 8 + 2) * 64 * sizeof(uint2);
}
