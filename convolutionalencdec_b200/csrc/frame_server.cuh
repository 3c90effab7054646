/*
 * frame_server.cuh -- the frame-parallel packet decoder of frame_parallel.cuh as a RESIDENT kernel that is handed one
 * packet after the other through a mailbox in page-locked host memory, instead of one graph launch per packet.
 *
 * Why.  The reference's drivers call VITERBI_DECODER_HARD(.., last = true) once per packet and wait for the answer
 * (speedDecode/speedDecode.c:79, berTestK7/berTestK7.c:157).  With one launch per call, ~20 of the 32 us a 2048-bit
 * packet takes are launch + copy node + stream synchronise -- a one-kernel call that does nothing takes 20 us on the
 * same box -- which leaves the GPU level with ONE host core.  Here the host writes the packet into
 * the mailbox and bumps a sequence number; the kernel, already running, sees it over PCIe (~1.5 us), decodes with the
 * same three phases, writes the bytes back into the mailbox and bumps `done`; the host spins on that word.
 *
 *   phase 1   passes: virtual CTA v = (block c, 8 start states) -- exactly fpBlockKernel's work -- spread over the grid
 *   barrier, then CTA 0 runs the min-plus chain over the blocks (fpChain)
 *   barrier
 *   phase 2   select: virtual CTA v = (block c, 8 end states) -- fpSelectKernel's reduction
 *   barrier, then CTA 0 walks the table back from state 0 and writes the bytes to the mailbox
 *
 * One CTA per SM (cooperative launch: the grid barriers need every CTA resident); 55 KB of shared memory and 256
 * threads per SM, so other kernels of the library still fit beside it.  The kernel leaves by itself after kIdleNs
 * without a request (or when asked to), and every wait in it is bounded (kHangNs): it cannot outlive its caller or hang
 * the device.  Inter-CTA data of earlier packets may sit in L1, so everything another CTA wrote is read with ld.cg and
 * the mailbox with ld.volatile.
 *
 * Exactness: the arithmetic is frame_parallel.cuh's (same device functions); tests/test_gpu_parity.py runs the same
 * packets through both paths and against the sequential CPU decoder.
 */
#pragma once
#include "frame_parallel.cuh"

namespace ced {

constexpr unsigned long long kFsIdleNs = 2000000ull;   /* leave after 2 ms without a request */
constexpr unsigned long long kFsHangNs = 50000000ull;  /* no wait inside a request may take 50 ms */
constexpr uint32_t kFsExit = 0xFFFFFFFFu;
constexpr int kFsMaxSegs = 16384 + 8 + 2 * kFpBlock;

/* page-locked, mapped host memory; host-written and device-written words sit in different cache lines */
struct FsMailbox {
    volatile uint32_t seq;       /* host -> device: number of the request in the mailbox; kFsExit = leave now */
    volatile uint32_t T;         /* segments of that packet */
    uint32_t pad0[30];
    volatile uint32_t done;      /* device -> host: number of the last request answered */
    volatile uint32_t state;     /* 1 running, 2 left (idle / asked to), 3 gave up on a wait */
    uint32_t pad1[14];
    volatile unsigned long long stamp[8]; /* %globaltimer of CTA 0: request seen, staged, passes, chain, select, answered */
    uint8_t edge[128];           /* [2][64] edge labels */
    uint8_t metrics[64];         /* path metrics before the packet */
    uint8_t pad2[64];
    uint8_t segs[kFsMaxSegs];    /* 16-byte aligned, readable in whole 128-byte blocks */
    uint8_t out[kFsMaxSegs / 8 + 64];
};

/* device memory; the first four words are initialised before every launch */
struct FsCtl {
    unsigned int cmd;            /* request number CTA 0 saw, or kFsExit */
    unsigned int T;
    unsigned int barrier;        /* monotonic arrival counter of the grid barriers */
    unsigned int gaveUp;
    /* CTA 0 copies the request out of the mailbox once (a few wide PCIe reads); the other CTAs read this copy --
     * 148 CTAs fetching their own pieces from host memory meant ~6000 small PCIe reads per packet (168 us per call) */
    __align__(16) uint8_t edge[128];
    __align__(16) uint8_t metrics[64];
    __align__(16) uint8_t segs[kFsMaxSegs];
};
constexpr size_t kFsCtlInitBytes = 4 * sizeof(unsigned int);

__device__ __forceinline__ uint32_t fsLdVolatile(const volatile uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 fsLdVolatile4(const void *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int fsLdAcquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fsStRelease(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct FsShared {
    uint2 dist[5 * 4 * 32];
    __align__(16) uint8_t edge[128];
    __align__(16) uint8_t seg[kFpBlock];
    __align__(16) uint8_t outCost[64][kFpThreads / 32];
    __align__(16) uint32_t outBits[kFpWords][64][kFpThreads / 32];
    __align__(16) uint32_t v[2][32];
    union {
        uint4 cost[kFpAhead][256];                                    /* ring of the min-plus chain */
        __align__(16) uint32_t best[kFpChainBlocks * 64 * kFpBestWords]; /* table rows of the final walk */
    } u;
    uint32_t word[kFpChainBlocks * kFpWords];
    __align__(16) uint32_t out[kFpThreads / 32][kFpBestWords];
    int state;
    unsigned int cmd, T;
    int bad;
};

/* all CTAs arrive, then all leave; `target` counts arrivals since the launch (wrap-safe compare) */
__device__ __forceinline__ bool fsGridBarrier(FsCtl *ctl, unsigned int &target, FsShared &sm)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(&ctl->barrier, 1u);
        const unsigned long long t0 = fpNow();
        int bad = 0;
        while ((int)(fsLdAcquire(&ctl->barrier) - target) < 0) {
            __nanosleep(40);
            if (fpNow() - t0 > kFsHangNs || fsLdAcquire(&ctl->gaveUp)) {
                bad = 1;
                atomicExch(&ctl->gaveUp, 1u);
                break;
            }
        }
        __threadfence();
        sm.bad = bad;
    }
    __syncthreads();
    return sm.bad == 0;
}

__global__ void __launch_bounds__(kFpThreads, 1) fpServerKernel(FsMailbox *mb, FsCtl *ctl, FpArgs a, uint32_t lastSeq)
{
    extern __shared__ __align__(16) uint8_t fsRaw[];
    FsShared &sm = *reinterpret_cast<FsShared *>(fsRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kFpThreads / 32;
    uint32_t seq = lastSeq;
    unsigned int target = 0;

    for (;;) {
        /* ---- wait for the next packet: CTA 0 watches the mailbox, the others watch CTA 0 ---- */
        if (tid == 0) {
            unsigned int cmd = seq, T = 0;
            const unsigned long long t0 = fpNow();
            if (blockIdx.x == 0) {
                for (;;) {
                    const uint32_t s = fsLdVolatile(&mb->seq);
                    if (s != seq) {
                        cmd = s;
                        T = fsLdVolatile(&mb->T);
                        break;
                    }
                    if (fpNow() - t0 > kFsIdleNs || fsLdAcquire(&ctl->gaveUp)) {
                        cmd = kFsExit;
                        break;
                    }
                    __nanosleep(100);
                }
            } else {
                while ((cmd = fsLdAcquire(&ctl->cmd)) == seq) {
                    __nanosleep(100);
                    if (fpNow() - t0 > kFsIdleNs + kFsHangNs) {
                        cmd = kFsExit;
                        break;
                    }
                }
                T = ctl->T;
            }
            sm.cmd = cmd;
            sm.T = T;
        }
        __syncthreads();
        if (blockIdx.x == 0) {
            /* the request travels host -> device once, as 16-byte pieces all in flight together */
            if (tid == 0)
                mb->stamp[0] = fpNow();
            if (sm.cmd != kFsExit) {
                const int nSeg = ((int)sm.T + 15) / 16, nAll = nSeg + (128 + 64) / 16;
                for (int i = tid; i < nAll; i += kFpThreads) {
                    if (i < nSeg)
                        __stcg(reinterpret_cast<uint4 *>(ctl->segs) + i, fsLdVolatile4(mb->segs + 16 * i));
                    else if (i < nSeg + 8)
                        __stcg(reinterpret_cast<uint4 *>(ctl->edge) + (i - nSeg), fsLdVolatile4(mb->edge + 16 * (i - nSeg)));
                    else
                        __stcg(reinterpret_cast<uint4 *>(ctl->metrics) + (i - nSeg - 8), fsLdVolatile4(mb->metrics + 16 * (i - nSeg - 8)));
                }
                __threadfence();
            }
            __syncthreads();
            if (tid == 0) {
                ctl->T = sm.T;
                fsStRelease(&ctl->cmd, sm.cmd);
                mb->stamp[1] = fpNow();
            }
        }
        if (sm.cmd == kFsExit) {
            if (blockIdx.x == 0 && tid == 0) {
                mb->state = fsLdAcquire(&ctl->gaveUp) ? 3u : 2u;
                __threadfence_system();
            }
            return;
        }
        seq = sm.cmd;
        a.T = (int)sm.T;
        a.nBlocks = (a.T + kFpBlock - 1) / kFpBlock;
        const int nVirtual = a.nBlocks * (64 / kWarps);

        /* ---- branch-cost table of this code (the edge labels travel with every call) ---- */
        if (tid < 32)
            reinterpret_cast<uint32_t *>(sm.edge)[tid] = __ldcg(reinterpret_cast<const uint32_t *>(ctl->edge) + tid);
        __syncthreads();
        for (int i = tid; i < 5 * 4 * 32; i += kFpThreads) {
            const int l = i & 31, r = i >> 7, q = 4 - r;
            const uint32_t rx = (i >> 5) & 3;
            int j = (l >> q) & 1;
            for (int b = 0; b < 5; b++)
                if (b != q)
                    j |= ((l >> b) & 1) << fpLaneBitRole(b, r);
            auto hd = [rx](uint32_t e) {
                const uint32_t x = (e ^ rx) & 3u;
                return x - (x >> 1);
            };
            sm.dist[i] = make_uint2(hd(sm.edge[j]) | hd(sm.edge[j + 32]) << 16, hd(sm.edge[64 + j]) | hd(sm.edge[64 + j + 32]) << 16);
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid == 0)
            mb->stamp[6] = fpNow();

        /* ---- phase 1: the passes (fpBlockKernel) ---- */
        for (int vb = blockIdx.x; vb < nVirtual; vb += gridDim.x) {
            const int wid = vb * kWarps + warp;
            const int c = wid >> 6, s = wid & 63;
            if (tid < kFpBlock / 16)
                reinterpret_cast<uint4 *>(sm.seg)[tid] = __ldcg(reinterpret_cast<const uint4 *>(ctl->segs + (size_t)c * kFpBlock) + tid);
            __syncthreads();
            const int steps = min(kFpBlock, a.T - c * kFpBlock);
            uint32_t seg[kFpBlock / 4];
#pragma unroll
            for (int i = 0; i < kFpBlock / 4; i++)
                seg[i] = reinterpret_cast<const uint32_t *>(sm.seg)[i];
            uint32_t M = ((2 * lane == s) ? 0u : (uint32_t)kFpUnreach) | ((2 * lane + 1 == s) ? 0u : (uint32_t)kFpUnreach) << 16;
            uint32_t p0[kFpWords] = {}, p1[kFpWords] = {};
            uint32_t selLH[5], upMask[5];
#pragma unroll
            for (int ph = 0; ph < 5; ph++) {
                const bool up = (lane >> (4 - ph)) & 1;
                selLH[ph] = up ? 0x3276u : 0x5410u;
                upMask[ph] = up ? 0xFFFFFFFFu : 0u;
            }
            if (steps == kFpBlock) {
                fpSteps<0>(seg, sm.dist + lane, selLH, upMask, M, p0, p1);
            } else {
                for (int t = 0; t < steps; t++)
                    fpStepDyn(t, lane, sm.seg[t], sm.dist + lane, M, p0, p1);
            }
            if (blockIdx.x == 0 && tid == 0 && vb == 0)
                mb->stamp[7] = fpNow();
            const uint32_t m0 = M & 0xFFFFu, m1 = M >> 16;
            const int r = steps % 5;
            int e = 0;
            for (int b = 0; b < 5; b++)
                e |= ((lane >> b) & 1) << fpLaneBitRole(b, r);
            sm.outCost[e][warp] = (uint8_t)min(m0, (uint32_t)kFpNoPath);
            sm.outCost[e + 1][warp] = (uint8_t)min(m1, (uint32_t)kFpNoPath);
#pragma unroll
            for (int w = 0; w < kFpWords; w++) {
                sm.outBits[w][e][warp] = p0[w];
                sm.outBits[w][e + 1][warp] = p1[w];
            }
            __syncthreads();
            const size_t row = (size_t)c * 64, sBase = (size_t)(vb * kWarps) & 63;
            for (int w = tid; w < 64 + 128 * kFpWords; w += kFpThreads) {
                if (w < 64) {
                    __stcg(reinterpret_cast<uint2 *>(a.cost + (row + w) * 64 + sBase), *reinterpret_cast<const uint2 *>(sm.outCost[w]));
                } else {
                    const int hw = (w - 64) >> 7, e2 = ((w - 64) >> 1) & 63, h = w & 1;
                    __stcg(reinterpret_cast<uint4 *>(a.bits[hw] + (row + e2) * 64 + sBase + 4 * h),
                           *reinterpret_cast<const uint4 *>(&sm.outBits[hw][e2][4 * h]));
                }
            }
            __syncthreads();
        }
        if (!fsGridBarrier(ctl, target, sm))
            break;

        /* ---- the sequential part on CTA 0 (fpChain), everybody else waits at the next barrier ---- */
        if (blockIdx.x == 0) {
            if (tid == 0)
                mb->stamp[2] = fpNow();
            if (tid < 64) {
                const int metric0 = (int)__ldcg(ctl->metrics + tid);
                reinterpret_cast<uint16_t *>(sm.v[0])[tid] = (uint16_t)metric0;
                a.v[tid] = metric0;
            }
            __syncthreads();
            fpChain<4, kFpAhead>(a, sm.v, &sm.u.cost[0][0], tid);
            __syncthreads();
            if (tid == 0)
                mb->stamp[3] = fpNow();
        }
        if (!fsGridBarrier(ctl, target, sm))
            break;

        /* ---- phase 2: which start state every survivor came from (fpSelectKernel) ---- */
        for (int vb = blockIdx.x; vb < nVirtual; vb += gridDim.x) {
            const int wid = vb * kWarps + warp;
            const int c = wid >> 6, e = wid & 63;
            constexpr int kKeys = kFpWords / 2 + 1;
            const size_t row = ((size_t)c * 64 + e) * 64;
            unsigned long long k[kKeys];
            auto less = [](const unsigned long long (&x)[kKeys], const unsigned long long (&y)[kKeys]) {
                bool lt = false, eq = true;
#pragma unroll
                for (int i = 0; i < kKeys; i++) {
                    lt = lt || (eq && x[i] < y[i]);
                    eq = eq && x[i] == y[i];
                }
                return lt;
            };
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int s = lane + 32 * h;
                const int x = (int)__ldcg(a.cost + row + s);
                const unsigned long long tot = (unsigned long long)(__ldcg(a.v + c * 64 + s) + x + (x >= kFpNoPath ? kFpBig : 0));
                unsigned long long n[kKeys];
                n[0] = tot << 32 | __ldcg(a.bits[kFpWords - 1] + row + s);
#pragma unroll
                for (int i = 1; i < kKeys - 1; i++)
                    n[i] = (unsigned long long)__ldcg(a.bits[kFpWords - 2 * i] + row + s) << 32 | __ldcg(a.bits[kFpWords - 2 * i - 1] + row + s);
                n[kKeys - 1] = (unsigned long long)__ldcg(a.bits[0] + row + s) << 6 | (__brev(s) >> 26);
                if (h == 0 || less(n, k)) {
#pragma unroll
                    for (int i = 0; i < kKeys; i++)
                        k[i] = n[i];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                unsigned long long n[kKeys];
#pragma unroll
                for (int i = 0; i < kKeys; i++)
                    n[i] = __shfl_xor_sync(0xFFFFFFFFu, k[i], o);
                if (less(n, k)) {
#pragma unroll
                    for (int i = 0; i < kKeys; i++)
                        k[i] = n[i];
                }
            }
            if (lane == 0) {
                sm.out[warp][kFpWords - 1] = (uint32_t)k[0];
#pragma unroll
                for (int i = 1; i < kKeys - 1; i++) {
                    sm.out[warp][kFpWords - 2 * i] = (uint32_t)(k[i] >> 32);
                    sm.out[warp][kFpWords - 2 * i - 1] = (uint32_t)k[i];
                }
                sm.out[warp][0] = (uint32_t)(k[kKeys - 1] >> 6);
                sm.out[warp][kFpWords] = __brev((uint32_t)k[kKeys - 1] & 63u) >> 26;
            }
            __syncthreads();
            if (tid < kWarps * kFpBestWords / 4)
                __stcg(reinterpret_cast<uint4 *>(a.best + (size_t)vb * kWarps * kFpBestWords) + tid,
                       reinterpret_cast<const uint4 *>(&sm.out[0][0])[tid]);
            __syncthreads();
        }
        if (!fsGridBarrier(ctl, target, sm))
            break;

        /* ---- CTA 0 walks the table back from state 0 (src/viterbiDecoderButterflyk1.c:205) and answers ---- */
        if (blockIdx.x == 0) {
            if (tid == 0) {
                sm.state = 0;
                mb->stamp[4] = fpNow();
            }
            __syncthreads();
            const int L = a.T - 6, outBytes = (L - 1) / 8 + 1;
            for (int hi = a.nBlocks; hi > 0; hi -= kFpChainBlocks) {
                const int lo = max(0, hi - kFpChainBlocks);
                const uint4 *src = reinterpret_cast<const uint4 *>(a.best + (size_t)lo * 64 * kFpBestWords);
                uint4 *dst = reinterpret_cast<uint4 *>(sm.u.best);
                const int n4 = (hi - lo) * 64 * kFpBestWords / 4;
                for (int base = 0; base < n4; base += 10 * kFpThreads) {
                    uint4 r[10];
#pragma unroll
                    for (int k2 = 0; k2 < 10; k2++)
                        if (base + k2 * kFpThreads + tid < n4)
                            r[k2] = __ldcg(src + base + k2 * kFpThreads + tid);
#pragma unroll
                    for (int k2 = 0; k2 < 10; k2++)
                        if (base + k2 * kFpThreads + tid < n4)
                            dst[base + k2 * kFpThreads + tid] = r[k2];
                }
                __syncthreads();
                if (tid == 0) {
                    int st = sm.state;
                    for (int cb = hi - 1; cb >= lo; cb--) {
                        const uint32_t *b = &sm.u.best[((cb - lo) * 64 + st) * kFpBestWords];
#pragma unroll
                        for (int w = 0; w < kFpWords; w++)
                            sm.word[kFpWords * (cb - lo) + w] = b[w];
                        st = (int)b[kFpWords];
                    }
                    sm.state = st;
                }
                __syncthreads();
                for (int j = tid; j < (hi - lo) * kFpBlock / 32; j += kFpThreads) {
                    /* four output bytes per store: bit t of a block is bit t % 32 of its word t / 32; a byte goes out
                     * MSb first (:249), the last partial byte zero-filled (:226-227) */
                    const int idx = lo * (kFpBlock / 8) + 4 * j;
                    if (idx < outBytes) {
                        uint32_t w4 = __byte_perm(__brev(sm.word[j]), 0, 0x0123);
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int valid = L - (idx + b) * 8;
                            if (valid < 8)
                                w4 &= ~(0xFFu << (8 * b)) | ((valid > 0 ? (0xFF00u >> valid) & 0xFFu : 0u) << (8 * b));
                        }
                        reinterpret_cast<uint32_t *>(mb->out)[idx >> 2] = w4;
                    }
                }
                __syncthreads();
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                mb->stamp[5] = fpNow();
                mb->done = seq;
                __threadfence_system();
            }
        }
    }
    /* a wait ran into kFsHangNs: say so and leave */
    if (blockIdx.x == 0 && tid == 0) {
        mb->state = 3u;
        __threadfence_system();
    }
}

} // namespace ced
