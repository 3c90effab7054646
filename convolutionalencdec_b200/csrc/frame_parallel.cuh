/*
 * frame_parallel.cuh -- ONE packet decoded by the whole GPU: the latency path behind a one-shot
 * VITERBI_DECODER_HARD(.., last = true) call (src/viterbiDecoderButterflyk1.c:82-263) for the
 * 64-state, n = 2 code, as issued per packet by speedDecode.c:79 and berTestK7.c:157.
 *
 * The add-compare-select recursion is sequential in time (2054 dependent steps for the reference's
 * speedDecode packet), which a single warp cannot run faster than ~100 ns per step.  It is however a
 * (min,+) matrix product chain, so the packet is cut into blocks of kFpBlock = 128 steps:
 *
 *   fpBlockKernel   one warp per (block c, start state s): a forward pass over the block that starts
 *                   with metric 0 in s only and carries the input bits of each survivor in kFpWords registers
 *                   (register exchange), giving cost_c[s][e] and bits_c[s][e] for all 64 end states e.
 *                   The last CTA to finish then runs the short sequential part: v_{c+1}[e] =
 *                   min_s v_c[s] + cost_c[s][e], 64 x 64 candidates per block on 256 threads.
 *   fpSelectKernel  one warp per (c, e): which start state the survivor into e came from, the minimum of
 *                   the key (v_c[s] + cost_c[s][e], bits_c[s][e], rev6(s)).  The last CTA walks the
 *                   resulting table back from state 0 (:205) block by block and writes the packed bits.
 *
 * Exactness.  The reference keeps the path from the LOWER predecessor on equal metrics (strict `>`,
 * :129-130); the predecessors j and j+32 differ in their oldest input bit, so every survivor is the
 * minimum over paths of (cost, input bits read as a number with LATER bits more significant).  The key
 * above is that order written per block: in-block bits first, then the six bits before the block
 * (the start state, newest bit first = rev6).  Metrics are plain ints here; the reference's uint8
 * metrics never wrap for this code, so decisions coincide.  tests/frame_parallel_model.py
 * is the same procedure in numpy, compared on the CPU with the sequential decoder.
 *
 * Lane mapping of fpBlockKernel.  Position = (5 lane bits, 1 slot bit); a lane holds two states.  Before
 * step t (phase r = t mod 5) the slot bit holds the newest state bit b0 and lane bit i holds state bit
 * ((i + r) mod 5) + 1.  The oldest bit b5 (the one a butterfly pairs on) sits in lane bit q = 4 - r:
 * one __shfl_xor_sync swaps it with the slot bit, the butterfly is then local to the lane (slot 0 = lower
 * predecessor, slot 1 = upper), writes its successors 2j / 2j+1 back into slots 0 / 1, and the labelling
 * has advanced to phase r + 1 without any further data movement.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ced {

constexpr int kFpBlock = 128;       /* trellis steps per block */
constexpr int kFpWords = kFpBlock / 32; /* registers holding the input bits of one survivor */
constexpr int kFpThreads = 256;     /* 8 warps per CTA */
constexpr int kFpUnreach = 0x1000;  /* starting metric of the 63 states a pass does not start in (u16 lanes) */
constexpr int kFpNoPath = 255;      /* stored cost of an (s, e) pair without a path; real in-block costs stay below
                                     * (kFpBlock - 6) + 12: a free step costs at most 1 when the two branches out of
                                     * a state carry complementary labels, which the host checks */
static_assert(kFpBlock + 6 < kFpNoPath, "costs are stored as bytes");
constexpr int kFpBig = 1 << 20;     /* added to candidates that do not start in s (short last block) */
constexpr int kFpAhead = 8;        /* blocks of costs in flight ahead of the sequential min-plus chain */
constexpr int kFpBestWords = kFpWords < 4 ? 4 : 8; /* words per entry of the survivor table: bits, start state, pad */
constexpr int kFpChainBlocks = 160 / kFpBestWords;  /* blocks of the survivor table staged per pass of the final walk */

struct FpArgs {
    int T;                     /* segments of the packet */
    int nBlocks;               /* ceil(T / kFpBlock) */
    const uint8_t *edge;       /* [2][64] edge labels, as StreamArgs.edge */
    const uint8_t *metricsIn;  /* [64] path metrics before the packet */
    const uint8_t *segs;       /* 16-byte aligned, readable up to nBlocks * kFpBlock bytes */
    uint8_t *cost;             /* [nBlocks][64 e][64 s] */
    uint32_t *bits[kFpWords];  /* each [nBlocks][64 e][64 s]: input bits of steps 32w .. 32w+31 of the block */
    int *v;                    /* [nBlocks + 1][64] */
    uint32_t *best;            /* [nBlocks][64 e][kFpBestWords] = {bits words, start state, -} */
    unsigned int *tickets;     /* [2], zero between calls */
    uint8_t *out;              /* (T - 6 - 1) / 8 + 1 decoded bytes, MSb first (:249) */
    int stampAll;              /* also the grid-wide (atomic, slow over PCIe) stamps */
    unsigned long long *stamps; /* optional: %globaltimer at the phase boundaries (CED_FP_STAMPS=1), else null */
};

__device__ __forceinline__ unsigned long long fpNow()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fpStamp(const FpArgs &a, int slot, int mode)
{
    if (a.stamps && threadIdx.x == 0 && (mode == 0 || a.stampAll)) {
        if (mode < 0)
            atomicMin(a.stamps + slot, fpNow());
        else if (mode > 0)
            atomicMax(a.stamps + slot, fpNow());
        else
            a.stamps[slot] = fpNow();
    }
}

/* state bit held by lane bit `b` in phase r, see above */
__device__ __forceinline__ int fpLaneBitRole(int b, int r)
{
    return ((b + r) % 5) + 1;
}

/*
 * One trellis step of a pass.  M = m0 | m1 << 16 (the lane's two metrics), p0 / p1 the survivors' input bits.
 * Both lanes of a pair send their whole M; one PRMT with a per-lane selector picks (lower, upper) predecessor
 * = (own.lo, partner.lo) or (partner.hi, own.hi).  The two candidates of a successor then sit in the two halves
 * of one register: one add forms both, VIMNMX.U16x2 against the half-swapped copy gives the minimum in both halves
 * and the predicate a0 <= a1, i.e. the reference's tie rule (keep the lower predecessor unless strictly greater,
 * :129-130).  21-33 instructions + 2-5 shuffles per step, more as the survivors' bit words fill up (the first,
 * scalar form needed 35 with one bit word: the passes are ALU-pipe bound, profiles/r1_packet_kernels_ncu.txt).
 */
template <int T0>
__device__ __forceinline__ void fpStep(const uint32_t (&seg)[kFpBlock / 4], const uint2 *dist, const uint32_t (&selLH)[5],
                                       const uint32_t (&upMask)[5], uint32_t &M, uint32_t (&p0)[kFpWords],
                                       uint32_t (&p1)[kFpWords])
{
    constexpr int r = T0 % 5, q = 4 - r;
    constexpr int kLive = T0 / 32 + 1; /* words above are still zero: nothing to move */
    const uint32_t um = upMask[r];
    const uint32_t recvM = __shfl_xor_sync(0xFFFFFFFFu, M, 1 << q);
    uint32_t pLo[kLive], pHi[kLive];
#pragma unroll
    for (int w = 0; w < kLive; w++) {
        const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, (p0[w] & um) | (p1[w] & ~um), 1 << q);
        pLo[w] = (recv & um) | (p0[w] & ~um);
        pHi[w] = (p1[w] & um) | (recv & ~um);
    }
    const uint32_t LH = __byte_perm(M, recvM, selLH[r]);
    const uint32_t rx = (seg[T0 >> 2] >> (8 * (T0 & 3))) & 3u; /* calcHammingDist(.., n = 2) looks at two bits */
    const uint2 d = dist[(r * 4 + rx) * 32];
    const uint32_t A = LH + d.x, B = LH + d.y;
    bool aHi, aLo, bHi, bLo;
    const uint32_t RA = __vibmin_u16x2(A, __byte_perm(A, 0, 0x1032), &aHi, &aLo);
    const uint32_t RB = __vibmin_u16x2(B, __byte_perm(B, 0, 0x1032), &bHi, &bLo);
    M = __byte_perm(RA, RB, 0x5410);
#pragma unroll
    for (int w = 0; w < kLive; w++) {
        p0[w] = aLo ? pLo[w] : pHi[w];
        p1[w] = (bLo ? pLo[w] : pHi[w]) | (w == T0 / 32 ? 1u << (T0 & 31) : 0u);
    }
}

template <int T0>
__device__ __forceinline__ void fpSteps(const uint32_t (&seg)[kFpBlock / 4], const uint2 *dist, const uint32_t (&selLH)[5],
                                        const uint32_t (&upMask)[5], uint32_t &M, uint32_t (&p0)[kFpWords],
                                        uint32_t (&p1)[kFpWords])
{
    if constexpr (T0 < kFpBlock) {
        fpStep<T0>(seg, dist, selLH, upMask, M, p0, p1);
        fpSteps<T0 + 1>(seg, dist, selLH, upMask, M, p0, p1);
    }
}

/* the same step with the step index at run time: the last, shorter block of a packet */
__device__ __forceinline__ void fpStepDyn(int t, int lane, uint32_t rx, const uint2 *dist, uint32_t &M,
                                          uint32_t (&p0)[kFpWords], uint32_t (&p1)[kFpWords])
{
    const int r = t % 5, q = 4 - r;
    const bool up = (lane >> q) & 1;
    const uint32_t recvM = __shfl_xor_sync(0xFFFFFFFFu, M, 1 << q);
    uint32_t pLo[kFpWords], pHi[kFpWords];
#pragma unroll
    for (int w = 0; w < kFpWords; w++) {
        const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, up ? p0[w] : p1[w], 1 << q);
        pLo[w] = up ? recv : p0[w];
        pHi[w] = up ? p1[w] : recv;
    }
    const uint32_t LH = __byte_perm(M, recvM, up ? 0x3276u : 0x5410u);
    const uint2 d = dist[(r * 4 + (rx & 3u)) * 32];
    const uint32_t A = LH + d.x, B = LH + d.y;
    bool aHi, aLo, bHi, bLo;
    const uint32_t RA = __vibmin_u16x2(A, __byte_perm(A, 0, 0x1032), &aHi, &aLo);
    const uint32_t RB = __vibmin_u16x2(B, __byte_perm(B, 0, 0x1032), &bHi, &bLo);
    M = __byte_perm(RA, RB, 0x5410);
#pragma unroll
    for (int w = 0; w < kFpWords; w++) {
        p0[w] = aLo ? pLo[w] : pHi[w];
        p1[w] = (bLo ? pLo[w] : pHi[w]) | (w == (t >> 5) ? 1u << (t & 31) : 0u);
    }
}

/*
 * The sequential part: v_{c+1}[e] = min_s v_c[s] + cost_c[s][e] on 64 * TPE threads (TPE threads share an
 * end state e and split the 64 start states).  The costs do not depend on v, so each thread streams its
 * own 64 / TPE bytes per block through a private shared-memory ring kFpAhead blocks ahead (cp.async, no
 * registers held).  v stays below 65 + 64 per block < 2^16 for any packet the API takes, so candidates are
 * formed two at a time: VIADDMNMX.U16x2 = min(v + cost, acc) on both halves of a register.
 */
template <int TPE, int AHEAD>
__device__ __forceinline__ void fpChain(const FpArgs &a, uint32_t (&sV)[2][32], uint4 *ring, int tid)
{
    constexpr int kThreads = 64 * TPE, kPieces = 4 / TPE; /* 16-byte pieces of cost per thread and block */
    if (tid >= kThreads)
        return;
    const int e = tid / TPE, h = tid % TPE;
    const uint4 *cp = reinterpret_cast<const uint4 *>(a.cost) + e * 4 + h * kPieces;
    const int nb = a.nBlocks;
    auto fetch = [&](int c2) {
        if (c2 < nb) {
#pragma unroll
            for (int p = 0; p < kPieces; p++) {
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(ring + (c2 % AHEAD) * 256 + p * kThreads + tid);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(cp + (size_t)c2 * 256 + p));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    for (int k = 0; k < AHEAD; k++)
        fetch(k);
    for (int c2 = 0; c2 < nb; c2++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD - 1) : "memory");
        uint4 cur[kPieces];
#pragma unroll
        for (int p = 0; p < kPieces; p++)
            cur[p] = ring[(c2 % AHEAD) * 256 + p * kThreads + tid];
        fetch(c2 + AHEAD);
        const uint32_t *vs = &sV[c2 & 1][h * (32 / TPE)];
        uint32_t m;
        if (c2 < nb - 1) {
            uint32_t acc[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
            for (int p = 0; p < kPieces; p++) {
                const uint4 va = *reinterpret_cast<const uint4 *>(vs + 8 * p);
                const uint4 vb = *reinterpret_cast<const uint4 *>(vs + 8 * p + 4);
                acc[0] = __viaddmin_u16x2(va.x, __byte_perm(cur[p].x, 0, 0x4140), acc[0]);
                acc[1] = __viaddmin_u16x2(va.y, __byte_perm(cur[p].x, 0, 0x4342), acc[1]);
                acc[2] = __viaddmin_u16x2(va.z, __byte_perm(cur[p].y, 0, 0x4140), acc[2]);
                acc[3] = __viaddmin_u16x2(va.w, __byte_perm(cur[p].y, 0, 0x4342), acc[3]);
                acc[0] = __viaddmin_u16x2(vb.x, __byte_perm(cur[p].z, 0, 0x4140), acc[0]);
                acc[1] = __viaddmin_u16x2(vb.y, __byte_perm(cur[p].z, 0, 0x4342), acc[1]);
                acc[2] = __viaddmin_u16x2(vb.z, __byte_perm(cur[p].w, 0, 0x4140), acc[2]);
                acc[3] = __viaddmin_u16x2(vb.w, __byte_perm(cur[p].w, 0, 0x4342), acc[3]);
            }
            const uint32_t x = __vminu2(__vminu2(acc[0], acc[1]), __vminu2(acc[2], acc[3]));
            m = min(x & 0xFFFFu, x >> 16);
        } else { /* only a last block shorter than 6 steps has (s, e) pairs without a path */
            m = 0x7FFFFFFFu;
#pragma unroll
            for (int p = 0; p < kPieces; p++) {
                const uint32_t w[4] = {cur[p].x, cur[p].y, cur[p].z, cur[p].w};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t x = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                    x += x >= (uint32_t)kFpNoPath ? (uint32_t)kFpBig : 0u;
                    m = min(m, ((vs[8 * p + (i >> 1)] >> (16 * (i & 1))) & 0xFFFFu) + x);
                }
            }
        }
        if (TPE >= 2)
            m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
        if (TPE >= 4)
            m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
        if (h == 0) {
            reinterpret_cast<uint16_t *>(sV[(c2 + 1) & 1])[e] = (uint16_t)m;
            a.v[(c2 + 1) * 64 + e] = (int)m;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
    }
}

__global__ void __launch_bounds__(kFpThreads) fpBlockKernel(FpArgs a)
{
    __shared__ uint2 sDist[5 * 4 * 32]; /* [phase][rx][lane] -> {d00 | d0h << 16, d10 | d1h << 16} */
    __shared__ __align__(16) uint8_t sEdge[128];
    __shared__ __align__(16) uint8_t sOutCost[64][kFpThreads / 32];  /* [e][s - s0] of this CTA's 8 passes */
    __shared__ __align__(16) uint32_t sOutBits[kFpWords][64][kFpThreads / 32];
    __shared__ __align__(16) uint32_t sV[2][32];                     /* v as u16x2 pairs (s = 2k, 2k + 1) */
    __shared__ uint4 sCost[kFpAhead][256];                           /* ring of the sequential part */
    __shared__ int sLast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kFpThreads / 32;

    /* a CTA = 8 consecutive start states of ONE block c (64 % kWarps == 0) */
    const int wid = blockIdx.x * kWarps + warp;
    const int c = wid >> 6, s = wid & 63;
    const uint4 *sp = reinterpret_cast<const uint4 *>(a.segs + (size_t)c * kFpBlock);
    fpStamp(a, 0, -1);
    uint4 sq[kFpBlock / 16]; /* in flight while the table is built */
#pragma unroll
    for (int i = 0; i < kFpBlock / 16; i++)
        sq[i] = sp[i];
    const int metric0 = tid < 64 ? a.metricsIn[tid] : 0; /* only the CTA that runs the sequential part uses it */
    if (tid < 32)
        reinterpret_cast<uint32_t *>(sEdge)[tid] = reinterpret_cast<const uint32_t *>(a.edge)[tid];
    __syncthreads();
    for (int i = tid; i < 5 * 4 * 32; i += kFpThreads) {
        const int l = i & 31, r = i >> 7, q = 4 - r;
        const uint32_t rx = (i >> 5) & 3;
        int j = (l >> q) & 1; /* after the swap lane bit q holds b0 */
        for (int b = 0; b < 5; b++)
            if (b != q)
                j |= ((l >> b) & 1) << fpLaneBitRole(b, r);
        auto hd = [rx](uint32_t e) {
            const uint32_t x = (e ^ rx) & 3u;
            return x - (x >> 1);
        };
        sDist[i] = make_uint2(hd(sEdge[j]) | hd(sEdge[j + 32]) << 16, hd(sEdge[64 + j]) | hd(sEdge[64 + j + 32]) << 16);
    }
    __syncthreads();

    {
        const int steps = min(kFpBlock, a.T - c * kFpBlock);
        uint32_t seg[kFpBlock / 4];
#pragma unroll
        for (int i = 0; i < kFpBlock / 16; i++) {
            seg[4 * i] = sq[i].x;
            seg[4 * i + 1] = sq[i].y;
            seg[4 * i + 2] = sq[i].z;
            seg[4 * i + 3] = sq[i].w;
        }
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
            fpSteps<0>(seg, sDist + lane, selLH, upMask, M, p0, p1);
        } else {
            for (int t = 0; t < steps; t++) /* uniform trip count */
                fpStepDyn(t, lane, a.segs[(size_t)c * kFpBlock + t], sDist + lane, M, p0, p1);
        }
        const uint32_t m0 = M & 0xFFFFu, m1 = M >> 16;
        const int r = steps % 5;
        int e = 0;
        for (int b = 0; b < 5; b++)
            e |= ((lane >> b) & 1) << fpLaneBitRole(b, r);
        sOutCost[e][warp] = (uint8_t)min(m0, (uint32_t)kFpNoPath);
        sOutCost[e + 1][warp] = (uint8_t)min(m1, (uint32_t)kFpNoPath);
#pragma unroll
        for (int w = 0; w < kFpWords; w++) {
            sOutBits[w][e][warp] = p0[w];
            sOutBits[w][e + 1][warp] = p1[w];
        }
    }
    __syncthreads();
    { /* rows of 8 start states: per end state one 8-byte piece of costs and whole 32-byte sectors of bits */
        const size_t row = (size_t)c * 64, sBase = (size_t)(blockIdx.x * kWarps) & 63;
        for (int w = tid; w < 64 + 128 * kFpWords; w += kFpThreads) {
            if (w < 64) {
                *reinterpret_cast<uint2 *>(a.cost + (row + w) * 64 + sBase) = *reinterpret_cast<const uint2 *>(sOutCost[w]);
            } else {
                const int hw = (w - 64) >> 7, e = ((w - 64) >> 1) & 63, h = w & 1;
                *reinterpret_cast<uint4 *>(a.bits[hw] + (row + e) * 64 + sBase + 4 * h) =
                    *reinterpret_cast<const uint4 *>(&sOutBits[hw][e][4 * h]);
            }
        }
    }

    /* the last CTA to get here runs the sequential part */
    fpStamp(a, 1, 1);
    __threadfence();
    __syncthreads();
    if (tid == 0)
        sLast = atomicAdd(&a.tickets[0], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!sLast)
        return;
    fpStamp(a, 2, 0);
    __threadfence();
    if (tid < 64) {
        reinterpret_cast<uint16_t *>(sV[0])[tid] = (uint16_t)metric0;
        a.v[tid] = metric0;
    }
    __syncthreads();
    fpStamp(a, 3, 0);
    if (a.stamps && tid == 0)
        a.stamps[10] = (unsigned long long)clock64();
    fpChain<4, kFpAhead>(a, sV, &sCost[0][0], tid);
    __syncthreads();
    if (a.stamps && tid == 0)
        a.stamps[11] = (unsigned long long)clock64();
    fpStamp(a, 4, 0);
    if (tid == 0)
        a.tickets[0] = 0;
}

__global__ void __launch_bounds__(kFpThreads) fpSelectKernel(FpArgs a)
{
    __shared__ __align__(16) uint32_t sBest[kFpChainBlocks * 64 * kFpBestWords];
    __shared__ uint32_t sWord[kFpChainBlocks * kFpWords];
    __shared__ __align__(16) uint32_t sOut[kFpThreads / 32][kFpBestWords];
    __shared__ int sState, sLast;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wid = blockIdx.x * (kFpThreads / 32) + warp;
    const int c = wid >> 6, e = wid & 63;
    if (a.stampAll)
        fpStamp(a, 5, -1);
    else if (blockIdx.x == 0)
        fpStamp(a, 5, 0);
    if (c < a.nBlocks) {
        /* key = (total cost, in-block bits with later steps more significant, rev6(start state)), most significant
         * word first: k[0] = total | top bits word, then pairs of bits words, k[kKeys - 1] = lowest bits word | rev6 */
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
            const int x = a.cost[row + s];
            const unsigned long long tot = (unsigned long long)(a.v[c * 64 + s] + x + (x >= kFpNoPath ? kFpBig : 0));
            unsigned long long n[kKeys];
            n[0] = tot << 32 | a.bits[kFpWords - 1][row + s];
#pragma unroll
            for (int i = 1; i < kKeys - 1; i++)
                n[i] = (unsigned long long)a.bits[kFpWords - 2 * i][row + s] << 32 | a.bits[kFpWords - 2 * i - 1][row + s];
            n[kKeys - 1] = (unsigned long long)a.bits[0][row + s] << 6 | (__brev(s) >> 26);
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
            sOut[warp][kFpWords - 1] = (uint32_t)k[0];
#pragma unroll
            for (int i = 1; i < kKeys - 1; i++) {
                sOut[warp][kFpWords - 2 * i] = (uint32_t)(k[i] >> 32);
                sOut[warp][kFpWords - 2 * i - 1] = (uint32_t)k[i];
            }
            sOut[warp][0] = (uint32_t)(k[kKeys - 1] >> 6);
            sOut[warp][kFpWords] = __brev((uint32_t)k[kKeys - 1] & 63u) >> 26;
        }
    }
    __syncthreads();
    /* the CTA's 8 entries as whole sectors */
    if (tid < (kFpThreads / 32) * kFpBestWords / 4 && c < a.nBlocks)
        reinterpret_cast<uint4 *>(a.best + (size_t)blockIdx.x * (kFpThreads / 32) * kFpBestWords)[tid] =
            reinterpret_cast<const uint4 *>(&sOut[0][0])[tid];

    fpStamp(a, 6, 1);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        sLast = atomicAdd(&a.tickets[1], 1u) == gridDim.x - 1;
        sState = 0; /* the traceback starts in state 0 (:205) */
    }
    __syncthreads();
    if (!sLast)
        return;
    __threadfence();
    fpStamp(a, 7, 0);
    const int L = a.T - 6, outBytes = (L - 1) / 8 + 1;
    for (int hi = a.nBlocks; hi > 0; hi -= kFpChainBlocks) {
        const int lo = max(0, hi - kFpChainBlocks);
        { /* 10 loads per thread and round, all loads of a round in flight together */
            const uint4 *src = reinterpret_cast<const uint4 *>(a.best + (size_t)lo * 64 * kFpBestWords);
            uint4 *dst = reinterpret_cast<uint4 *>(sBest);
            const int n4 = (hi - lo) * 64 * kFpBestWords / 4;
            for (int base = 0; base < n4; base += 10 * kFpThreads) {
                uint4 r[10];
#pragma unroll
                for (int k = 0; k < 10; k++)
                    if (base + k * kFpThreads + tid < n4)
                        r[k] = __ldcg(src + base + k * kFpThreads + tid);
#pragma unroll
                for (int k = 0; k < 10; k++)
                    if (base + k * kFpThreads + tid < n4)
                        dst[base + k * kFpThreads + tid] = r[k];
            }
        }
        __syncthreads();
        if (tid == 0) {
            int st = sState;
            for (int cb = hi - 1; cb >= lo; cb--) {
                const uint32_t *b = &sBest[((cb - lo) * 64 + st) * kFpBestWords];
#pragma unroll
                for (int w = 0; w < kFpWords; w++)
                    sWord[kFpWords * (cb - lo) + w] = b[w];
                st = (int)b[kFpWords];
            }
            sState = st;
        }
        __syncthreads();
        for (int i = tid; i < (hi - lo) * kFpBlock / 8; i += kFpThreads) {
            const int idx = lo * (kFpBlock / 8) + i; /* output byte; input bit t of a block is bit t % 32 of its word t / 32 */
            if (idx < outBytes) {
                uint32_t by = __brev((sWord[i >> 2] >> (8 * (i & 3))) & 0xFFu) >> 24; /* MSb first (:249) */
                const int valid = L - idx * 8;
                if (valid < 8)
                    by &= (0xFF00u >> valid) & 0xFFu; /* last partial byte is zero-filled (:226-227) */
                a.out[idx] = (uint8_t)by;
            }
        }
        __syncthreads();
    }
    fpStamp(a, 8, 0);
    if (tid == 0)
        a.tickets[1] = 0;
}

/* bytes of device scratch for packets of up to maxSteps segments, and the carve-up */
struct FpScratch {
    size_t cost, bits[kFpWords], v, best, tickets, total;
};
inline FpScratch fpScratchLayout(int maxSteps)
{
    const size_t nb = (size_t)(maxSteps + kFpBlock - 1) / kFpBlock;
    FpScratch s;
    s.cost = 0;
    size_t at = nb * 4096;
    for (int w = 0; w < kFpWords; w++) {
        s.bits[w] = at;
        at += nb * 4096 * 4;
    }
    s.v = at;
    s.best = s.v + (nb + 1) * 64 * 4;
    s.tickets = s.best + nb * 64 * kFpBestWords * 4;
    s.total = s.tickets + 16;
    return s;
}

} // namespace ced
