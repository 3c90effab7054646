/*
 * swar_generic.cu -- batched decode of any k = 1, n = 2 / 3 code with 4 .. 256 states on the table-driven
 * SIMD-in-word kernels (swar_generic.cuh): persistent forward kernel with the unit scheduler of k7ForwardKernel,
 * thread-per-frame traceback with cp.async-staged survivor rows.  Called by ced_decode_batch for the codes the
 * hand-scheduled K = 7 kernels do not take (ced_abi.cu decodeBatchImpl).
 */
#include "ced_internal.cuh"
#include "swar_generic.cuh"
#include "swar_radix4.cuh"

namespace ced {


template <int W>
struct alignas(W >= 4 ? 16 : 4 * W) SurvRow {
    uint32_t w[W];
};

/* continuous streams (cedDecodeWindowGeneric): the counterpart of FwdWindow (decode_batch.cuh) for any trellis policy */
struct GenWindow {
    const uint4 *metricsIn;  /* [groups][kStateU4][32] or NULL = start of the stream */
    uint4 *metricsOut;       /* renormalised metrics after the last step, or NULL */
    uint32_t *startPos;      /* [groups][32] position (== state: the next phase is 0) of the smallest metric */
    int survSteps;           /* survivor rows per group in `surv` */
    int stepOffset;          /* rows in front of this launch's first step (the carried traceback depth) */
};

/* lowest position whose metric is 0 (after a renormalisation the smallest metric is 0); position == state at phase 0 */
template <class P>
__device__ __forceinline__ uint32_t genBestPosition(const uint32_t (&R)[P::kRegs])
{
    uint32_t best = 0;
#pragma unroll
    for (int r = P::kRegs - 1; r >= 0; r--) {
        const uint32_t z = (R[r] - 0x01010101u) & ~R[r] & 0x80808080u;
        if (z)
            best = 4u * (uint32_t)r + (uint32_t)((__ffs((int)z) - 1) >> 3);
    }
    return best;
}

/* Forward pass: see k7ForwardKernel (decode_batch.cuh) for the scheduler; symbols are byte-per-segment only. */
template <class P, int V, bool ALIGNED, bool CARRY = false>
__global__ void __launch_bounds__(kFwdThreads, P::kRegs >= 64 ? 2 : 3)
genForwardKernel(const uint8_t *__restrict__ segs, size_t stride, int nFrames, int T, SurvRow<P::kWords> *__restrict__ surv,
                 const uint8_t *__restrict__ table, int n, uint32_t minusOne, FwdSched sched, int chunksPerUnit,
                 GenWindow win = GenWindow())
{
    using G = P;
    constexpr int S = P::kPhases;   /* steps per label rotation */
    using TG = TileGeom<ByteSymbols, ALIGNED>;
    constexpr int kPitch = TG::kPitch, kRegs = G::kRegs, kStateU4 = (kRegs + 3) / 4;
    extern __shared__ __align__(16) uint8_t sMem[];
    uint8_t *sTab = sMem;                                              /* G::tableBytes(V) */
    uint8_t *sTile = sMem + (G::tableBytes(V) + 15) / 16 * 16;         /* [4 warps][32 * kPitch] */
    for (int i = threadIdx.x; i < G::tableBytes(V) / 16; i += kFwdThreads)
        reinterpret_cast<uint4 *>(sTab)[i] = reinterpret_cast<const uint4 *>(table)[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile = sTile + warp * 32 * kPitch;
    const uint32_t symMask = 0x01010101u * (uint32_t)(V - 1);
    const unsigned groups = (unsigned)((nFrames + 31) / 32);
    const unsigned chunks = (unsigned)((T + P::kChunk - 1) / P::kChunk);
    const unsigned unitsPerGroup = (chunks + chunksPerUnit - 1) / chunksPerUnit;
    const unsigned total = groups * unitsPerGroup;

    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0)
            v = atomicAdd(sched.counter, 1u);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };
    unsigned u = grab();
    uint4 pre[TG::kPiecesPerRow];
    if (u < total)
        loadTile<ByteSymbols, ALIGNED>(pre, segs, stride, 32LL * (u % groups), nFrames, (int)((u / groups) * chunksPerUnit) * P::kChunk,
                                      T, lane);
    while (u < total) {
        const unsigned g = u % groups, su = u / groups;
        const unsigned cFirst = su * chunksPerUnit, cEnd = min(chunks, cFirst + chunksPerUnit);
        const long long frame0 = 32LL * g;
        const bool live = frame0 + lane < nFrames;
        uint4 *stateSlot = sched.state + ((size_t)g * kStateU4) * 32 + lane;
        uint32_t R[kRegs];
        if (su == 0) {
            P::init(R, n);
            if constexpr (CARRY) {
                if (win.metricsIn) {
#pragma unroll
                    for (int i = 0; i < kStateU4; i++) {
                        const uint4 v = __ldcg(win.metricsIn + ((size_t)g * kStateU4 + i) * 32 + lane);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (4 * i + k < kRegs)
                                R[4 * i + k] = w[k];
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
            for (int i = 0; i < kStateU4; i++) {
                const uint4 v = __ldcg(stateSlot + i * 32);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (4 * i + k < kRegs)
                        R[4 * i + k] = w[k];
            }
        }
        unsigned un = total;
        for (unsigned c = cFirst; c < cEnd; c++) {
            const int t0 = (int)c * P::kChunk;
            __syncwarp();
            storeTile<ByteSymbols, ALIGNED>(tile, pre, lane, symMask);
            if (c + 1 < cEnd) {
                loadTile<ByteSymbols, ALIGNED>(pre, segs, stride, frame0, nFrames, t0 + P::kChunk, T, lane);
            } else {
                un = grab();
                if (un < total)
                    loadTile<ByteSymbols, ALIGNED>(pre, segs, stride, 32LL * (un % groups), nFrames,
                                                  (int)((un / groups) * chunksPerUnit) * P::kChunk, T, lane);
            }
            __syncwarp();
            const uintptr_t rowAddr = reinterpret_cast<uintptr_t>(segs) + (size_t)(frame0 + lane) * stride + (size_t)t0;
            const uint8_t *p = tile + lane * kPitch + (ALIGNED ? 0u : (rowAddr & 15u));
            SurvRow<G::kWords> *o = surv + ((size_t)g * (CARRY ? win.survSteps : T) + (CARRY ? win.stepOffset : 0) + t0) * 32 + lane;
            const int steps = min(P::kChunk, T - t0);
            for (int done = 0; done < steps;) {
                const int nr = min(G::kRenorm, steps - done);
#pragma unroll 1
                for (int base = 0; base < nr; base += S) {
                    const int nn = min(S, nr - base);          /* whole label rotations except at the end of the frame */
                    auto one = [&](auto phTag) {
                        constexpr int PH = decltype(phTag)::value;
                        if (PH < nn) {
                            const uint32_t off = p[PH];          /* rx * 32 */
                            const uint8_t *tab = sTab + G::phaseBase(PH, V) + G::rxOffset(PH, off);
                            uint32_t Tw[G::kWords];
                            P::template step<PH>(R, tab, V, minusOne, Tw);
                            if (live) {
                                SurvRow<G::kWords> row;
#pragma unroll
                                for (int w = 0; w < G::kWords; w++)
                                    row.w[w] = Tw[w];
                                o[(size_t)PH * 32] = row;
                            }
                        }
                    };
                    one(std::integral_constant<int, 0>());
                    if constexpr (S > 1) one(std::integral_constant<int, 1>());
                    if constexpr (S > 2) one(std::integral_constant<int, 2>());
                    if constexpr (S > 3) one(std::integral_constant<int, 3>());
                    if constexpr (S > 4) one(std::integral_constant<int, 4>());
                    if constexpr (S > 5) one(std::integral_constant<int, 5>());
                    if constexpr (S > 6) one(std::integral_constant<int, 6>());
                    if constexpr (S > 7) one(std::integral_constant<int, 7>());
                    p += S;
                    o += (size_t)S * 32;
                }
                done += nr;
                if (done < steps || c + 1 < chunks)
                    P::renorm(R);
            }
        }
        if (cEnd < chunks) {
#pragma unroll
            for (int i = 0; i < kStateU4; i++) {
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (4 * i + k < kRegs)
                        w[k] = R[4 * i + k];
                __stcg(stateSlot + i * 32, make_uint4(w[0], w[1], w[2], w[3]));
            }
            __threadfence();
            __syncwarp();
            if (lane == 0)
                stRelease(sched.done + g, (int)su + 1);
        } else if constexpr (CARRY) {
            if (win.metricsOut) {
                P::renorm(R);   /* a slice is a whole number of label rotations: the next phase is 0, position == state */
                const uint32_t best = genBestPosition<P>(R);
                if (live) {
#pragma unroll
                    for (int i = 0; i < kStateU4; i++) {
                        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (4 * i + k < kRegs)
                                w[k] = R[4 * i + k];
                        __stcg(win.metricsOut + ((size_t)g * kStateU4 + i) * 32 + lane, make_uint4(w[0], w[1], w[2], w[3]));
                    }
                    win.startPos[(size_t)g * 32 + lane] = best;
                }
            }
        }
        u = un;
    }
}

/* Full-frame traceback from state 0 (src/viterbiDecoderButterflyk1.c:205-254), one thread per frame; the rows of the
 * next 8 steps are fetched with cp.async into the thread's own shared-memory slots while the current 8 are walked.
 * A step yields P::kStepBits decoded bits, MSb first within the byte (:249). */
template <class P>
__host__ __device__ constexpr int genTbThreads() { return sizeof(SurvRow<P::kWords>) >= 64 ? 32 : kTbThreads; } /* staging fits 48 KB */

template <class P>
__global__ void __launch_bounds__(kTbThreads)
genTracebackKernel(const SurvRow<P::kWords> *__restrict__ surv, int nFrames, int T, uint8_t *__restrict__ out, size_t outStride)
{
    using Row = SurvRow<P::kWords>;
    constexpr int kb = P::kStepBits, spb = 8 / kb;   /* bits per step, steps per output byte */
    __shared__ Row sW[2][8][genTbThreads<P>()];
    const long long frame = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= nFrames)
        return;
    const Row *s = surv + ((size_t)(frame / 32) * T) * 32 + (frame & 31);
    uint8_t *dst = out + (size_t)frame * outStride;
    const int L = T - P::kTail, nBlocks = L / 8, tid = threadIdx.x;
    auto prefetch = [&](int blk, int buf) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const Row *src = s + (size_t)(8 * blk + i) * 32;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(&sW[buf][i][tid]);
            if constexpr (sizeof(Row) >= 16) {
#pragma unroll
                for (int k = 0; k < (int)sizeof(Row) / 16; k++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * k),
                                 "l"(reinterpret_cast<const uint8_t *>(src) + 16 * k));
            } else if constexpr (sizeof(Row) == 8) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src));
            } else {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    if (nBlocks > 0)
        prefetch(nBlocks - 1, 0);
    uint32_t p = 0; /* state 0 sits at position 0 in every phase */
    uint32_t acc = 0;
    for (int t = T - 1; t >= 8 * nBlocks; t--) {   /* the tail steps carry no output (:208-223); then < 8 ragged steps */
        const Row row = s[(size_t)t * 32];
        const uint32_t v = P::tbStep(p, row.w, t);
        if (t < L) {
            acc = (acc >> kb) | (v << (8 - kb));
            if (t % spb == 0) {
                dst[t / spb] = (uint8_t)acc;
                acc = 0;
            }
        }
    }
    int buf = 0;
    for (int blk = nBlocks - 1; blk >= 0; blk--, buf ^= 1) {
        if (blk > 0) {
            prefetch(blk - 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            const Row row = sW[buf][i][tid];
            const uint32_t v = P::tbStep(p, row.w, 8 * blk + i);
            acc = (acc >> kb) | (v << (8 - kb));
            if (i % spb == 0) {
                dst[(8 * blk + i) / spb] = (uint8_t)acc;
                acc = 0;
            }
        }
    }
}

/*
 * The same walk for narrow survivor rows (<= 16 bytes per step), in blocks of 24 steps like k7TracebackKernel: 24 is a
 * multiple of every label rotation (S = 1, 2, 3, 4, 6, 8), so inside a block the phase of a step -- and with it every shift
 * count -- is a compile-time constant, and a block is a whole number of output bytes (3 for k = 1, 6 for k = 2).  The rows
 * of the next block are fetched with cp.async while the current one is walked.  (The 8-step kernel above computed
 * t mod S and variable shifts per step: for K = 3 its 0.42 ms were as long as the forward pass.)
 */
template <class P>
__host__ __device__ constexpr int genTb24Threads() { return sizeof(SurvRow<P::kWords>) >= 16 ? 32 : 64; }

/* CHECK: the block reaches above step L (steps >= L carry no output): test every step instead of none */
template <class P, int I, bool CHECK>
__device__ __forceinline__ void genWalk24(uint32_t &p, uint32_t &acc, const SurvRow<P::kWords> *rows, int stride, uint8_t *dst, int blk,
                                          int L)
{
    if constexpr (I >= 0) {
        constexpr int kb = P::kStepBits, spb = 8 / kb;
        const SurvRow<P::kWords> row = rows[(size_t)I * stride];
        const uint32_t v = P::template tbStepC<I % P::kPhases>(p, row.w);
        if (!CHECK || 24 * blk + I < L) {
            acc = (acc >> kb) | (v << (8 - kb));
            if constexpr (I % spb == 0) {
                dst[(24 * blk + I) / spb] = (uint8_t)acc;
                acc = 0;
            }
        }
        genWalk24<P, I - 1, CHECK>(p, acc, rows, stride, dst, blk, L);
    }
}

/* startPos (continuous streams): position to start the walk from instead of state 0; `skip` top steps carry no output
 * (the tail of a terminated frame, or the traceback depth of a slice); steps below `emitLo` (a multiple of 24) are not
 * walked: their bits were emitted by an earlier slice.  Output byte 0 holds step emitLo. */
template <class P>
__global__ void __launch_bounds__(genTb24Threads<P>())
genTraceback24Kernel(const SurvRow<P::kWords> *__restrict__ surv, int nFrames, int T, uint8_t *__restrict__ out, size_t outStride,
                     const uint32_t *__restrict__ startPos = nullptr, int skip = P::kTail, int emitLo = 0)
{
    using Row = SurvRow<P::kWords>;
    constexpr int kb = P::kStepBits, spb = 8 / kb, kThreads = genTb24Threads<P>();
    static_assert(24 % P::kPhases == 0, "a block holds whole label rotations");
    __shared__ Row sW[2][24][kThreads];
    const long long frame = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (frame >= nFrames)
        return;
    const Row *s = surv + ((size_t)(frame / 32) * T) * 32 + (frame & 31);
    uint8_t *dst = out + (size_t)frame * outStride - emitLo / spb;   /* dst[t / spb] for t >= emitLo */
    const int L = T - skip, nBlocks = T / 24, blkLo = emitLo / 24, tid = threadIdx.x;   /* output steps [emitLo, L) */
    auto prefetch = [&](int blk, int buf) {
#pragma unroll
        for (int i = 0; i < 24; i++) {
            const Row *src = s + (size_t)(24 * blk + i) * 32;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(&sW[buf][i][tid]);
            if constexpr (sizeof(Row) == 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
            else if constexpr (sizeof(Row) == 8)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src));
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    };
    if (nBlocks > blkLo)
        prefetch(nBlocks - 1, 0);
    uint32_t p = startPos ? startPos[frame] : 0u, acc = 0;
    for (int t = T - 1; t >= 24 * nBlocks; t--) {   /* tail steps (no output, :208-223) and the ragged top */
        const Row row = s[(size_t)t * 32];
        const uint32_t v = P::tbStep(p, row.w, t);
        if (t < L) {
            acc = (acc >> kb) | (v << (8 - kb));
            if (t % spb == 0) {
                dst[t / spb] = (uint8_t)acc;
                acc = 0;
            }
        }
    }
    int buf = 0;
    for (int blk = nBlocks - 1; blk >= blkLo; blk--, buf ^= 1) {
        if (blk > blkLo) {
            prefetch(blk - 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (24 * blk + 24 <= L)
            genWalk24<P, 23, false>(p, acc, &sW[buf][0][tid], kThreads, dst, blk, L);
        else
            genWalk24<P, 23, true>(p, acc, &sW[buf][0][tid], kThreads, dst, blk, L);
    }
}

} // namespace ced

template <class P>
using GenKernelPtr = void (*)(const uint8_t *, size_t, int, int, ced::SurvRow<P::kWords> *, const uint8_t *, int, uint32_t,
                              ced::FwdSched, int, ced::GenWindow);

template <class P, bool CARRY>
static GenKernelPtr<P> genKernelFor(int n, bool aligned)
{
    if (n == 2)
        return aligned ? ced::genForwardKernel<P, 4, true, CARRY> : ced::genForwardKernel<P, 4, false, CARRY>;
    return aligned ? ced::genForwardKernel<P, 8, true, CARRY> : ced::genForwardKernel<P, 8, false, CARRY>;
}

/* table and launch for one call of a trellis policy P; T = trellis steps per frame, V = 2^n received symbols */
template <class P, class BuildTable>
static int launchGen(ced_ctx *c, int n, int T, BuildTable buildTable, const uint8_t *dSegs, size_t segStride, int nFrames,
                     uint8_t *dOut, size_t outStride, cudaStream_t s, int slot)
{
    using Row = ced::SurvRow<P::kWords>;
    const int V = 1 << n;
    ced_ctx::Work &wk = c->work[slot];
    const size_t perFrame = (size_t)T * sizeof(Row);
    size_t waveMax = std::min<size_t>(c->maxWaveFrames, std::max<size_t>(64, kMaxScratchBytes / perFrame)) / 64 * 64;
    const size_t firstWave = std::min<size_t>((size_t)nFrames, waveMax), g0 = (firstWave + 31) / 32;
    constexpr int kStateU4 = (P::kRegs + 3) / 4;
    const size_t tabBytes = ((size_t)P::tableBytes(V) + 15) / 16 * 16;
    const size_t stateBytes = g0 * kStateU4 * 32 * sizeof(uint4) + tabBytes + 512, flagBytes = (g0 + 1) * sizeof(int);
    if (wk.scratch.bytes < g0 * 32 * perFrame || wk.schedState.bytes < stateBytes || wk.schedFlags.bytes < flagBytes) {
        CED_CUDA(cudaDeviceSynchronize());
        int rc = wk.scratch.ensure(g0 * 32 * perFrame);
        if (rc == CED_OK) rc = wk.schedState.ensure(stateBytes);
        if (rc == CED_OK) rc = wk.schedFlags.ensure(flagBytes);
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    /* the step table of this code: built on the host (a few KB), kept behind the hand-off slots of the working set */
    std::vector<uint8_t> table(tabBytes);
    buildTable(table.data());
    uint8_t *dTable = reinterpret_cast<uint8_t *>(wk.schedState.p) + (g0 * kStateU4 * 32 * sizeof(uint4) + 255) / 256 * 256;
    CED_CUDA(cudaMemcpyAsync(dTable, table.data(), tabBytes, cudaMemcpyHostToDevice, s));   /* pageable source: staged */
    /* a tile that is not a whole number of 16-byte pieces (S = 5, 7: 95 / 91 steps) starts at any alignment */
    const bool aligned16 = (reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0 && P::kChunk % 16 == 0;
    const size_t pitchA = ced::TileGeom<ced::ByteSymbols, true>::kPitch, pitchU = ced::TileGeom<ced::ByteSymbols, false>::kPitch;
    const size_t smem = tabBytes + 4 * 32 * (aligned16 ? pitchA : pitchU);
    auto kernel = genKernelFor<P, false>(n, aligned16);
    CED_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int resident = 0;
    CED_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, ced::kFwdThreads, smem));
    resident = std::max(1, resident);
    for (size_t f0 = 0; f0 < (size_t)nFrames; f0 += waveMax) {
        const int wave = (int)std::min<size_t>(waveMax, (size_t)nFrames - f0);
        const int groups = (wave + 31) / 32;
        const int perSm = std::max(1, std::min({4, resident, std::max(3, groups / (4 * c->sms))}));
        const int blocks = std::max(1, std::min(c->sms * perSm, (groups + 3) / 4));
        ced::FwdSched sched;
        sched.counter = reinterpret_cast<unsigned int *>(wk.schedFlags.p);
        sched.done = wk.schedFlags.p + 1;
        sched.state = wk.schedState.p;
        CED_CUDA(cudaMemsetAsync(wk.schedFlags.p, 0, (size_t)(groups + 1) * sizeof(int), s));
        kernel<<<blocks, ced::kFwdThreads, smem, s>>>(dSegs + f0 * segStride, segStride, wave, T, reinterpret_cast<Row *>(wk.scratch.p),
                                                     dTable, n, c->bm0113.minusOne, sched, 2, ced::GenWindow());
        if constexpr (sizeof(Row) <= 16 && 24 % P::kPhases == 0) {
            constexpr int tbT = ced::genTb24Threads<P>();
            ced::genTraceback24Kernel<P><<<(wave + tbT - 1) / tbT, tbT, 0, s>>>(reinterpret_cast<const Row *>(wk.scratch.p), wave, T,
                                                                             dOut + f0 * outStride, outStride);
        } else {
            constexpr int tbT = ced::genTbThreads<P>();
            ced::genTracebackKernel<P><<<(wave + tbT - 1) / tbT, tbT, 0, s>>>(reinterpret_cast<const Row *>(wk.scratch.p), wave, T,
                                                                           dOut + f0 * outStride, outStride);
        }
        c->launches += 2;
    }
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int cedDecodeBatchSwarGeneric(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                              int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot)
{
    if (!code || code->codedBits < 2 || code->codedBits > 3)
        return CED_ERR_UNSUPPORTED;
    const int K = code->constraintLen, S = K - 1;
    if (S < 2 || S > 8)
        return CED_ERR_UNSUPPORTED;
    static const bool off = getenv("CED_SWAR_GENERIC") && atoi(getenv("CED_SWAR_GENERIC")) == 0;
    if (off)
        return CED_ERR_UNSUPPORTED;
    if (segStride < (size_t)(frameBits + S) || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    ced::GenCode gc;
    gc.S = S;
    gc.n = code->codedBits;
    for (int i = 0; i < 3; i++)
        gc.tap[i] = i < gc.n ? reverseBits(code->gen[i], K) : 0u;
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const int T = frameBits + S;
#define CED_GEN_CASE(S_)                                                                                                 \
    case S_:                                                                                                             \
        return launchGen<ced::GenPolicy<S_>>(c, gc.n, T, [&](uint8_t *t) { ced::buildGenTable<S_>(gc, t); }, dSegs, segStride, \
                                             nFrames, dOut, outStride, s, slot)
    switch (S) {
        CED_GEN_CASE(2);
        CED_GEN_CASE(3);
        CED_GEN_CASE(4);
        CED_GEN_CASE(5);
        CED_GEN_CASE(6);
        CED_GEN_CASE(7);
    default:
        return launchGen<ced::GenPolicy<8>>(c, gc.n, T, [&](uint8_t *t) { ced::buildGenTable<8>(gc, t); }, dSegs, segStride, nFrames,
                                            dOut, outStride, s, slot);
    }
#undef CED_GEN_CASE
}

/*
 * Continuous streams with windowed traceback (ced_decode_window_batch) for the codes of these kernels with K <= 7: the
 * procedure of decodeWindowImpl (ced_abi.cu) -- carried metrics, the newest `depth` survivor rows kept between calls,
 * traceback from the best-metric state with the top `depth` steps discarded -- on GenPolicy<S> kernels.
 * Carry block per 32 streams: kStateU4 x 32 uint4 metrics, 32 start positions, depth x 32 survivor rows.
 */
template <int S>
static size_t genWindowCarryGroupBytes(int depth)
{
    using P = ced::GenPolicy<S>;
    constexpr int kStateU4 = (P::kRegs + 3) / 4;
    return (size_t)kStateU4 * 32 * sizeof(uint4) + 32 * sizeof(uint32_t) + (size_t)depth * 32 * sizeof(ced::SurvRow<P::kWords>);
}

template <int S>
static int windowGen(ced_ctx *c, const ced::GenCode &gc, const uint8_t *dSegs, size_t segStride, int nStreams, int nSegments,
                     uint64_t streamPos, int depth, int last, void *dCarry, uint8_t *dOut, size_t outStride, cudaStream_t s)
{
    using P = ced::GenPolicy<S>;
    using Row = ced::SurvRow<P::kWords>;
    static_assert(sizeof(Row) <= 16 && 24 % S == 0, "windowed decoding: K <= 7");
    constexpr int kStateU4 = (P::kRegs + 3) / 4;
    const int n = gc.n, V = 1 << n;
    const int emitLo = (int)std::max<int64_t>(0, (int64_t)depth - (int64_t)streamPos);
    const int Tl = depth + nSegments;
    const int emitHi = last ? Tl - S : nSegments;
    const int bytesOut = emitHi > emitLo ? (emitHi - emitLo) / 8 : 0;
    if (outStride < (size_t)bytesOut) {
        setError("ced_decode_window_batch: stride shorter than a slice");
        return CED_ERR_ARG;
    }
    if (nStreams == 0)
        return bytesOut;
    ced_ctx::Work &wk = c->work[0];
    const size_t perFrame = (size_t)Tl * sizeof(Row);
    const size_t waveMax = std::max<size_t>(64, std::min<size_t>(c->maxWaveFrames, kMaxScratchBytes / perFrame) / 64 * 64);
    const size_t g0n = (std::min<size_t>((size_t)nStreams, waveMax) + 31) / 32;
    const size_t tabBytes = ((size_t)P::tableBytes(V) + 15) / 16 * 16;
    const size_t stateBytes = g0n * kStateU4 * 32 * sizeof(uint4) + tabBytes + 512, flagBytes = (g0n + 1) * sizeof(int);
    if (wk.scratch.bytes < g0n * 32 * perFrame || wk.schedState.bytes < stateBytes || wk.schedFlags.bytes < flagBytes) {
        CED_CUDA(cudaDeviceSynchronize());
        int rc = wk.scratch.ensure(g0n * 32 * perFrame);
        if (rc == CED_OK) rc = wk.schedState.ensure(stateBytes);
        if (rc == CED_OK) rc = wk.schedFlags.ensure(flagBytes);
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    std::vector<uint8_t> table(tabBytes);
    ced::buildGenTable<S>(gc, table.data());
    uint8_t *dTable = reinterpret_cast<uint8_t *>(wk.schedState.p) + (g0n * kStateU4 * 32 * sizeof(uint4) + 255) / 256 * 256;
    CED_CUDA(cudaMemcpyAsync(dTable, table.data(), tabBytes, cudaMemcpyHostToDevice, s));
    const bool aligned16 = (reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0 && P::kChunk % 16 == 0;
    const size_t smem = tabBytes + 4 * 32 * (aligned16 ? ced::TileGeom<ced::ByteSymbols, true>::kPitch
                                                       : ced::TileGeom<ced::ByteSymbols, false>::kPitch);
    auto kernel = genKernelFor<P, true>(n, aligned16);
    CED_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t allGroups = (size_t)(nStreams + 31) / 32;
    uint8_t *carry = static_cast<uint8_t *>(dCarry);
    uint4 *carryMetrics = reinterpret_cast<uint4 *>(carry);
    uint32_t *carryStart = reinterpret_cast<uint32_t *>(carry + allGroups * kStateU4 * 32 * sizeof(uint4));
    uint8_t *carrySurv = carry + allGroups * ((size_t)kStateU4 * 32 * sizeof(uint4) + 32 * sizeof(uint32_t));
    const size_t tailBytes = (size_t)depth * 32 * sizeof(Row);   /* per group */
    const size_t rowBytes = (size_t)Tl * 32 * sizeof(Row);       /* per group in the scratch */
    for (size_t f0 = 0; f0 < (size_t)nStreams; f0 += waveMax) {
        const int wave = (int)std::min<size_t>(waveMax, (size_t)nStreams - f0);
        const int groups = (wave + 31) / 32;
        const size_t gFirst = f0 / 32;
        if (streamPos > 0)
            CED_CUDA(cudaMemcpy2DAsync(wk.scratch.p, rowBytes, carrySurv + gFirst * tailBytes, tailBytes, tailBytes, (size_t)groups,
                                       cudaMemcpyDeviceToDevice, s));
        const int blocks = std::max(1, std::min(c->sms * 3, (groups + 3) / 4));
        ced::FwdSched sched;
        sched.counter = reinterpret_cast<unsigned int *>(wk.schedFlags.p);
        sched.done = wk.schedFlags.p + 1;
        sched.state = wk.schedState.p;
        CED_CUDA(cudaMemsetAsync(wk.schedFlags.p, 0, (size_t)(groups + 1) * sizeof(int), s));
        ced::GenWindow win;
        win.metricsIn = streamPos > 0 ? carryMetrics + gFirst * kStateU4 * 32 : nullptr;
        win.metricsOut = last ? nullptr : carryMetrics + gFirst * kStateU4 * 32;
        win.startPos = carryStart + gFirst * 32;
        win.survSteps = Tl;
        win.stepOffset = depth;
        kernel<<<blocks, ced::kFwdThreads, smem, s>>>(dSegs + f0 * segStride, segStride, wave, nSegments,
                                                     reinterpret_cast<Row *>(wk.scratch.p), dTable, n, c->bm0113.minusOne, sched, 2, win);
        c->launches += 1;
        if (bytesOut > 0) {
            constexpr int tbT = ced::genTb24Threads<P>();
            ced::genTraceback24Kernel<P><<<(wave + tbT - 1) / tbT, tbT, 0, s>>>(
                reinterpret_cast<const Row *>(wk.scratch.p), wave, Tl, dOut + f0 * outStride, outStride,
                last ? nullptr : win.startPos, last ? S : depth, emitLo);
            c->launches += 1;
        }
        if (!last)
            CED_CUDA(cudaMemcpy2DAsync(carrySurv + gFirst * tailBytes, tailBytes,
                                       reinterpret_cast<uint8_t *>(wk.scratch.p) + (size_t)nSegments * 32 * sizeof(Row), rowBytes,
                                       tailBytes, (size_t)groups, cudaMemcpyDeviceToDevice, s));
    }
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return bytesOut;
}

/* bytes of carry block for `code` on this path, 0 if the code is not one of these kernels' (K <= 7, n = 2 or 3) */
size_t cedWindowCarryBytesGeneric(const ced_code_t *code, int nStreams, int depth)
{
    if (!code || code->codedBits < 2 || code->codedBits > 3 || nStreams <= 0 || depth < 24 || depth % 24)
        return 0;
    const size_t groups = (size_t)(nStreams + 31) / 32;
    switch (code->constraintLen - 1) {
    case 2: return groups * genWindowCarryGroupBytes<2>(depth);
    case 3: return groups * genWindowCarryGroupBytes<3>(depth);
    case 4: return groups * genWindowCarryGroupBytes<4>(depth);
    case 6: return groups * genWindowCarryGroupBytes<6>(depth);
    default: return 0;
    }
}

int cedDecodeWindowGeneric(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nStreams, int nSegments,
                           uint64_t streamPos, int depth, int last, void *dCarry, uint8_t *dOut, size_t outStride, void *stream)
{
    const int K = code ? code->constraintLen : 0, S = K - 1;
    if (!code || code->codedBits < 2 || code->codedBits > 3 || !(S == 2 || S == 3 || S == 4 || S == 6))
        return CED_ERR_UNSUPPORTED;
    if (!c || nStreams < 0 || nSegments < 0 || depth < 24 || depth % 24 || depth > 8184 || streamPos % 96 ||
        (nStreams > 0 && (!dSegs || !dOut || !dCarry)) || (reinterpret_cast<uintptr_t>(dCarry) & 15u)) {
        setError("ced_decode_window_batch: bad argument (depth and streamPos must be multiples of 24 / 96)");
        return CED_ERR_ARG;
    }
    if (last ? (nSegments < S || (streamPos + (uint64_t)nSegments - (uint64_t)S) % 8 != 0) : (nSegments == 0 || nSegments % 96 != 0)) {
        setError("ced_decode_window_batch: a slice must be a positive multiple of 96 segments; the last one must end the stream "
                 "on a byte boundary plus K-1 tail segments");
        return CED_ERR_ARG;
    }
    if (nSegments > (int)kStreamMaxSteps * 4 || segStride < (size_t)nSegments) {
        setError("ced_decode_window_batch: slice too long or stride shorter than a slice");
        return CED_ERR_ARG;
    }
    ced::GenCode gc;
    gc.S = S;
    gc.n = code->codedBits;
    for (int i = 0; i < 3; i++)
        gc.tap[i] = i < gc.n ? reverseBits(code->gen[i], K) : 0u;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    switch (S) {
    case 2: return windowGen<2>(c, gc, dSegs, segStride, nStreams, nSegments, streamPos, depth, last, dCarry, dOut, outStride, s);
    case 3: return windowGen<3>(c, gc, dSegs, segStride, nStreams, nSegments, streamPos, depth, last, dCarry, dOut, outStride, s);
    case 4: return windowGen<4>(c, gc, dSegs, segStride, nStreams, nSegments, streamPos, depth, last, dCarry, dOut, outStride, s);
    default: return windowGen<6>(c, gc, dSegs, segStride, nStreams, nSegments, streamPos, depth, last, dCarry, dOut, outStride, s);
    }
}

/* rate-2/n codes (k = 2), n = 2 or 3, 4 .. 256 states: radix-4 SIMD-in-word kernels; CED_ERR_UNSUPPORTED = not a code
 * they take (ced_decode_batch_k then runs the one-warp-per-frame kernel of radix_k.cu) */
int cedDecodeBatchSwarRadix4(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                             int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    if (!code || code->codedBits < 2 || code->codedBits > 3 || code->constraintLen < 2 || code->constraintLen > 5 || (frameBits & 7))
        return CED_ERR_UNSUPPORTED;
    static const bool off = getenv("CED_SWAR_GENERIC") && atoi(getenv("CED_SWAR_GENERIC")) == 0;
    if (off)
        return CED_ERR_UNSUPPORTED;
    const int K = code->constraintLen, S = K - 1, T = frameBits / 2 + S;
    if (segStride < (size_t)T || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch_k: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    ced::R4Code rc;
    rc.S = S;
    rc.n = code->codedBits;
    for (int i = 0; i < 3; i++)
        rc.tap[i] = i < rc.n ? reverseBits(code->gen[i], 2 * K) : 0u;
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
#define CED_R4_CASE(S_)                                                                                                  \
    case S_:                                                                                                             \
        return launchGen<ced::R4Policy<S_>>(c, rc.n, T, [&](uint8_t *t) { ced::buildR4Table<S_>(rc, t); }, dSegs, segStride,  \
                                            nFrames, dOut, outStride, s, 0)
    switch (S) {
        CED_R4_CASE(1);
        CED_R4_CASE(2);
        CED_R4_CASE(3);
    default:
        return launchGen<ced::R4Policy<4>>(c, rc.n, T, [&](uint8_t *t) { ced::buildR4Table<4>(rc, t); }, dSegs, segStride, nFrames,
                                           dOut, outStride, s, 0);
    }
#undef CED_R4_CASE
}
