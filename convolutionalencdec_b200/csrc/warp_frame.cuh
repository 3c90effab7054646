/*
 * warp_frame.cuh -- device pieces shared by the warp-per-frame kernels (warp_frame.cu: one warp decodes a whole frame;
 * warp_split.cu: the frame is cut into blocks in time as well).  See the header of warp_frame.cu.
 */
#pragma once
#include "ced_internal.cuh"

#include <type_traits>

namespace ced {

constexpr int kWfMaxV = 8;          /* received symbols: n <= 3 */
constexpr int kWfCostBytes = 16 * 32 * (int)sizeof(uint4);   /* radix 4: [rx1 | rx2 << 2][lane] uint4; radix 2: [rx][lane] uint2 */

struct WfArgs {
    const uint8_t *segs;
    size_t segStride;
    uint8_t *out;
    size_t outStride;
    int nFrames, T, S, n;
    int packed;            /* symbols four to a byte (ced_decode_batch_packed; radix 4 only) */
    int seg;               /* traceback: steps per lane, a multiple of 8 */
    int survRows;          /* >= T + T / seg, even */
    int outPad;            /* bytes of the output row in shared memory, a multiple of 16 */
    uint32_t initMetric;   /* (uint8_t)(NUM_STATES + 1), :59-67 */
    uint32_t cost[2][kWfMaxV][32]; /* [second step of a radix-4 pair / radix-2 step = 1, first = 0][rx][lane]: 4 costs as bytes */
};

/* one step back from state s over a row in the radix-2 format (x: successors 2j, y: successors 2j+1; bit j) */
template <int S>
__device__ __forceinline__ void wfBackOdd(uint32_t &s, const uint2 w)
{
    const uint32_t word = (s & 1u) ? w.y : w.x;
    const uint32_t dec = (word >> (s >> 1)) & 1u;
    s = (s >> 1) | (dec << (S - 1));
}
__device__ __forceinline__ void wfBackOddRt(uint32_t &s, const uint2 w, int S)
{
    const uint32_t word = (s & 1u) ? w.y : w.x;
    const uint32_t dec = (word >> (s >> 1)) & 1u;
    s = (s >> 1) | (dec << (S - 1));
}
/* ... over a row in the state-indexed format of the first step of a radix-4 pair (x: states 0..31, y: 32..63) */
__device__ __forceinline__ void wfBackEven(uint32_t &s, const uint2 w)
{
    const uint32_t word = (s & 32u) ? w.y : w.x;
    const uint32_t dec = (word >> (s & 31u)) & 1u;
    s = (s >> 1) | (dec << 5);
}

/* walk steps [tLo, tHi) downwards from state s (tLo a multiple of 8); EMIT: bits of steps < L go to sOut, MSb first */
template <bool R4, bool EMIT>
__device__ __forceinline__ uint32_t wfWalk(uint32_t s, int tHi, int tLo, const uint2 *rows, int S, int L, uint8_t *sOut)
{
    int t = tHi - 1;
    uint32_t acc = 0;
    for (; t >= tLo && (t & 7) != 7; t--) {   /* ragged top: only where a walk starts at the end of the frame */
        const uint32_t bit = s & 1u;
        if (R4 && !(t & 1))
            wfBackEven(s, rows[t - tLo]);
        else
            wfBackOddRt(s, rows[t - tLo], S);
        if (EMIT && t < L) {
            acc = (acc >> 1) | (bit << 7);
            if ((t & 7) == 0) {
                sOut[t >> 3] = (uint8_t)acc;
                acc = 0;
            }
        }
    }
    for (; t >= tLo; t -= 8) {   /* t + 1 is a multiple of 8: whole bytes; the rows do not depend on the path */
        const int tb = t - 7;
        uint2 w[8];
#pragma unroll
        for (int q = 0; q < 8; q++)
            w[q] = rows[tb - tLo + q];
        uint32_t byte = 0;
#pragma unroll
        for (int q = 7; q >= 0; q--) {
            byte |= (s & 1u) << (7 - q);   /* step tb + q is bit 7 - q of byte tb / 8 (:249) */
            if (R4 && !(q & 1))
                wfBackEven(s, w[q]);
            else if (R4)
                wfBackOdd<6>(s, w[q]);
            else
                wfBackOddRt(s, w[q], S);
        }
        if (EMIT && tb < L)
            sOut[tb >> 3] = (uint8_t)byte;
    }
    return s;
}

/* Warp-parallel exact traceback over the rows of one frame in shared memory (row of step t at t + t / seg): see the
 * header of warp_frame.cu.  Packed bytes land in sOut; the caller synchronises the warp before reading them. */
template <bool R4>
__device__ __forceinline__ void wfTraceback(const uint2 *sSurv, int T, int S, int seg, uint8_t *sOut, int lane)
{
    const int L = T - S;
    const int top = (T - 1) / seg;
    const int lo = lane * seg, hi = min(T, lo + seg);
    const uint2 *myRows = sSurv + lo + lane;
    uint32_t sIn = 0;
    if (lane < top) {
        const int wHi = min(T, hi + seg);      /* warm-up over the steps of the lane above, from state 0 */
        sIn = wfWalk<R4, false>(0u, wHi, hi, sSurv + hi + lane + 1, S, L, sOut);
    }
    uint32_t sLeave = 0;
    if (lane <= top)
        sLeave = wfWalk<R4, true>(sIn, hi, lo, myRows, S, L, sOut);
    for (;;) {
        const uint32_t above = __shfl_down_sync(0xFFFFFFFFu, sLeave, 1);
        const bool redo = lane < top && above != sIn;
        if (!__any_sync(0xFFFFFFFFu, redo))
            break;
        if (redo) {
            sIn = above;
            sLeave = wfWalk<R4, true>(sIn, hi, lo, myRows, S, L, sOut);
        }
    }
}

/* the per-lane cost table in shared memory from the byte form in the kernel arguments */
template <bool R4>
__device__ __forceinline__ void wfBuildCostTable(const uint32_t (&cost)[2][kWfMaxV][32], uint8_t *smem, int lane)
{
    auto spread = [](uint32_t u) { return make_uint2((u & 0xFFu) | ((u >> 8 & 0xFFu) << 16), (u >> 16 & 0xFFu) | ((u >> 24) << 16)); };
    if (R4) {
#pragma unroll 4
        for (int i = lane; i < 16 * 32; i += 32) {
            const uint2 c1 = spread(cost[0][(i >> 5) & 3][lane]), c2 = spread(cost[1][i >> 7][lane]);
            reinterpret_cast<uint4 *>(smem)[i] = make_uint4(c1.x, c1.y, c2.x, c2.y);
        }
    } else {
        for (int i = lane; i < kWfMaxV * 32; i += 32)
            reinterpret_cast<uint2 *>(smem)[i] = spread(cost[1][i >> 5][lane]);
    }
}

} // namespace ced

/* host: cudaFuncAttributeMaxDynamicSharedMemorySize once per (device, kernel, size reached) instead of on every call;
 * `slot` names the kernel (0 .. 3) */
#include <atomic>
template <class Kernel>
static inline cudaError_t cedWarpEnsureSmem(int device, int slot, Kernel kernel, size_t bytes)
{
    static std::atomic<int> reached[32][4];
    std::atomic<int> &r = reached[device & 31][slot & 3];
    if ((int)bytes <= r.load(std::memory_order_relaxed))
        return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess)
        r.store((int)bytes, std::memory_order_relaxed);
    return e;
}

/* host: branch costs of `code` in the byte form the kernels take (warp_frame.cu) */
void cedWarpFrameCosts(const ced_code_t *code, bool r4, uint32_t (&cost)[2][ced::kWfMaxV][32]);
