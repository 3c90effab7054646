/*
 * encode_batch.cuh -- bit-parallel rate-1/n convolutional encoder, sm_100a.
 *
 * Replaces convEnc + computeEncOutputSegment (src/convEncode.c:46-161): the
 * reference shifts one bit at a time and takes two popcount parities per bit;
 * here one thread produces 16 consecutive segments from a 24-bit window of the
 * message with shift/XOR (generator i's tap d contributes window << d), then
 * spreads each 4-bit group to 4 bytes with one multiply.  Output is the
 * reference's wire format: one byte per segment, generator i in bit i.
 *
 * The same kernel serves the per-frame streaming call (nFrames == 1): `hist`
 * carries the previous call's shift register (convEncoderState_t.tappedDelay)
 * and `tailSegs` is K-1 when `last` (src/convEncode.c:108-119), else 0.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ced {

struct EncTaps {
    uint32_t tap[8]; /* bit d taps the input bit d steps back (bit 0 = newest) */
};

constexpr int kEncThreads = 256;
constexpr int kEncFramesPerBlock = 4;

/* 16 coded segments (one uint4) from the 24-bit window `win` (bit i = input bit u[16c - 8 + i]). */
/* FIXED = the production code K=7 n=2 g={0113,0171}: taps 0x69 / 0x4F are compile-time constants, so the
 * tap loop folds to 3-4 shift/XORs per generator (src/convEncode.c:13-17 derives them at run time). */
constexpr uint32_t kFixedTap0 = 0x69u, kFixedTap1 = 0x4Fu;

template <int KK, int NN, bool FIXED = false>
__device__ __forceinline__ uint4 encode16(uint32_t win, const EncTaps &taps, int K, int n)
{
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    const int kk = KK ? KK : K, nn = NN ? NN : n;
#pragma unroll
    for (int g = 0; g < (NN ? NN : 8); g++) {
        if (g >= nn)
            break;
        const uint32_t tap = FIXED ? (g == 0 ? kFixedTap0 : kFixedTap1) : taps.tap[g];
        uint32_t c16 = 0;
#pragma unroll
        for (int d = 0; d < (KK ? KK : 9); d++)
            if (d < kk)
                c16 ^= (0u - ((tap >> d) & 1u)) & (win << d);   /* tap d contributes u[t-d] */
        c16 >>= 8; /* bit s = coded bit g of segment 16c + s */
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t nib = (c16 >> (4 * q)) & 0xFu;
            w[q] |= ((nib * 0x00204081u) & 0x01010101u) << g;   /* bit i of nib -> byte i */
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

/* n = 2 only: the same 16 segments as 4 packed bytes (segment s in bits 2s, 2s+1). */
__device__ __forceinline__ uint32_t spreadBits16(uint32_t x)
{
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
template <int KK, bool FIXED = false>
__device__ __forceinline__ uint32_t encode16Packed(uint32_t win, const EncTaps &taps, int K)
{
    const int kk = KK ? KK : K;
    uint32_t c[2] = {0u, 0u};
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const uint32_t tap = FIXED ? (g == 0 ? kFixedTap0 : kFixedTap1) : taps.tap[g];
#pragma unroll
        for (int d = 0; d < (KK ? KK : 9); d++)
            if (d < kk)
                c[g] ^= (0u - ((tap >> d) & 1u)) & (win << d);
        c[g] >>= 8;
    }
    return spreadBits16(c[0]) | (spreadBits16(c[1]) << 1);
}

/*
 * A CTA encodes kEncFramesPerBlock consecutive frames; its threads stride over the 16-segment
 * chunks of those frames (32-bit index math only).  KK/NN != 0 fix the constraint length and the
 * number of generators at compile time (the K=7 n=2 production code), 0 = run-time values.
 */
template <int KK, int NN, bool PACKED = false, bool FIXED = false>
__device__ __forceinline__ void
encodeBatchBody(const uint8_t *__restrict__ msg, size_t msgStride, int nFrames, int frameBytes,
                uint8_t *__restrict__ segs, size_t segStride, int tailSegs, int K, int n, const EncTaps &taps,
                uint32_t hist, int aligned16)
{
    const int T = 8 * frameBytes + tailSegs;
    const unsigned chunksPerFrame = (unsigned)(T + 15) / 16;
    const long long frameBase = (long long)blockIdx.x * kEncFramesPerBlock;
    const unsigned framesHere = (unsigned)min((long long)kEncFramesPerBlock, (long long)nFrames - frameBase);
    const unsigned items = framesHere * chunksPerFrame;
    for (unsigned item = threadIdx.x; item < items; item += kEncThreads) {
        const unsigned fl = item / chunksPerFrame;
        const int c = (int)(item - fl * chunksPerFrame);
        const long long f = frameBase + fl;
        const uint8_t *m = msg + (size_t)f * msgStride;
        /* message bytes 2c-1, 2c, 2c+1; byte -1 is the carried shift register, bytes past the end are
         * the zero tail (src/convEncode.c:108-119) */
        const int i0 = 2 * c - 1;
        const uint32_t b0 = (i0 < 0) ? (hist & 0xFFu) : (i0 < frameBytes ? __ldg(m + i0) : 0u);
        const uint32_t b1 = (i0 + 1 < frameBytes) ? __ldg(m + i0 + 1) : 0u;
        const uint32_t b2 = (i0 + 2 < frameBytes) ? __ldg(m + i0 + 2) : 0u;
        /* window bit i = input bit u[16c - 8 + i]  (bytes are sent MSb first, src/convEncode.c:91) */
        const uint32_t win = __brev(((b0 << 16) | (b1 << 8) | b2) << 8);
        if (PACKED) { /* segStride / aligned16 then describe the packed rows (4-byte alignment suffices) */
            const uint32_t pk = encode16Packed<KK, FIXED>(win, taps, K);
            uint8_t *dstp = segs + (size_t)f * segStride + 4 * (size_t)c;
            const int cnt = min(16, T - 16 * c);
            if (aligned16 && cnt == 16) {
                *reinterpret_cast<uint32_t *>(dstp) = pk;
            } else {
                const uint32_t keep = cnt == 16 ? pk : (pk & ((1u << (2 * cnt)) - 1u));
                for (int b2 = 0; b2 < (cnt + 3) / 4; b2++)
                    dstp[b2] = (uint8_t)(keep >> (8 * b2));
            }
            continue;
        }
        const uint4 v = encode16<KK, NN, FIXED>(win, taps, K, n);
        uint8_t *dst = segs + (size_t)f * segStride + 16 * (size_t)c;
        if (aligned16 && 16 * c + 16 <= T) {
            *reinterpret_cast<uint4 *>(dst) = v;
        } else {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            for (int s2 = 0; s2 < 16 && 16 * c + s2 < T; s2++)
                dst[s2] = (uint8_t)(w[s2 >> 2] >> (8 * (s2 & 3)));
        }
    }
}

template <int KK, int NN, bool PACKED = false, bool FIXED = false>
__global__ void __launch_bounds__(kEncThreads)
encodeBatchKernel(const uint8_t *__restrict__ msg, size_t msgStride, int nFrames, int frameBytes,
                  uint8_t *__restrict__ segs, size_t segStride, int tailSegs, int K, int n, EncTaps taps,
                  uint32_t hist, int aligned16)
{
    encodeBatchBody<KK, NN, PACKED, FIXED>(msg, msgStride, nFrames, frameBytes, segs, segStride, tailSegs, K, n, taps, hist, aligned16);
}

/* The per-packet call (ced_stream_encode): one CTA, message and segments in the pinned mailboxes; when the segments
 * are out it writes a running count to a doorbell word the host spins on instead of synchronising the stream. */
__global__ void __launch_bounds__(kEncThreads)
encodeStreamKernel(const uint8_t *__restrict__ msg, int frameBytes, uint8_t *__restrict__ segs, size_t segStride, int tailSegs,
                   int K, int n, EncTaps taps, uint32_t hist, unsigned int *doneCounter, volatile unsigned int *doneFlag)
{
    encodeBatchBody<0, 0>(msg, (size_t)max(frameBytes, 1), 1, frameBytes, segs, segStride, tailSegs, K, n, taps, hist, 1);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *doneFlag = atomicAdd(doneCounter, 1u) + 1u;
        __threadfence_system();
    }
}

/*
 * Table-spread encoder for n = 2 codes with byte-per-segment output and 16-byte aligned segment rows.
 *
 * ncu on the kernel above: ~100 instructions per STG.128, issue-bound at 3.9 TB/s.  Two changes cut that to ~40:
 *   - a warp owns whole frames (no division per item) and lane l of iteration k encodes the 16 segments of chunk
 *     32k + l, so one STG.128 of the warp covers 512 contiguous bytes -- full 32-byte sectors.  tools/write_probe.cu:
 *     this HBM takes 7.1 TB/s of full-sector writes, but a store instruction that leaves sectors half written (e.g. 32
 *     contiguous bytes per lane as two STG.128) halves the rate, because the L2 fetches a sector on a partial write;
 *   - the 1-bit -> 1-byte spread (shift, mask, multiply, mask per nibble and generator) becomes a table lookup: the
 *     nibbles of the two coded-bit words are interleaved with two LOP3 (index = c0 nibble | c1 nibble << 4) and a
 *     256-entry shared-memory table returns the four output bytes of four segments.
 * Message bytes come straight from global memory through L1 (two 2-byte loads per chunk, four chunks' loads in
 * flight per lane before the first use; the 128-byte line is shared by the whole warp).  The ragged end of a row is
 * written with at most four stores.
 */
constexpr int kEncLutThreads = 256;

/* 16 coded segments of chunk c from the window x = byte(2c-1) << 24 | byte(2c) << 16 | byte(2c+1) << 8 (low byte: don't
 * care), spread to one byte per segment through the table */
template <bool FIXED>
__device__ __forceinline__ uint4 encodeChunkLut(uint32_t x, uint32_t tap0, uint32_t tap1, const uint32_t *lut)
{
    /* window bit i = input bit u[16c - 8 + i]  (bytes are sent MSb first, src/convEncode.c:91) */
    const uint32_t win = __brev(x);
    uint32_t c0, c1;
    if (FIXED) { /* taps 0x69 = {0,3,5,6}, 0x4F = {0,1,2,3,6}  (src/convEncode.c:13-17 on 0113 / 0171) */
        const uint32_t s3 = win << 3, s6 = win << 6;
        c0 = (win ^ s3 ^ (win << 5) ^ s6) >> 8;
        c1 = (win ^ (win << 1) ^ (win << 2) ^ s3 ^ s6) >> 8;
    } else {
        c0 = c1 = 0u;
#pragma unroll
        for (int d = 0; d <= 8; d++) {
            c0 ^= (0u - ((tap0 >> d) & 1u)) & (win << d);
            c1 ^= (0u - ((tap1 >> d) & 1u)) & (win << d);
        }
        c0 >>= 8;
        c1 >>= 8;
    }
    const uint32_t ze = (c0 & 0x0F0Fu) | ((c1 << 4) & 0xF0F0u); /* segments 8k .. 8k+3   in byte k */
    const uint32_t zo = ((c0 >> 4) & 0x0F0Fu) | (c1 & 0xF0F0u); /* segments 8k+4 .. 8k+7 in byte k */
    return make_uint4(lut[ze & 0xFFu], lut[zo & 0xFFu], lut[ze >> 8], lut[zo >> 8]);
}

template <bool FIXED, bool SHORT>
__global__ void __launch_bounds__(kEncLutThreads)
encodeBatchLutKernel(const uint8_t *__restrict__ msg, size_t msgStride, int nFrames, int frameBytes,
                     uint8_t *__restrict__ segs, size_t segStride, int tailSegs, uint32_t tap0, uint32_t tap1)
{
    __shared__ uint32_t lut[256];
    {
        const uint32_t idx = threadIdx.x;
        uint32_t e = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            e |= (((idx >> i) & 1u) | (((idx >> (4 + i)) & 1u) << 1)) << (8 * i); /* generator i in bit i */
        lut[idx] = e;
    }
    __syncthreads();
    const int T = 8 * frameBytes + tailSegs;
    const int chunks = (T + 15) / 16;
    /* chunks whose three message bytes 2c-1 .. 2c+1 all exist (c = 0 apart): whole warp iterations of them run
     * without bounds checks; frameBytes is even on this path */
    const int interiorIters = (frameBytes / 2) / 32;
    const int lane = threadIdx.x & 31;
    /* SHORT frames (at most 16 chunks): a warp encodes 32 / cp2 frames side by side, cp2 lanes each; a separate
     * instantiation because per-lane frame pointers cost the long-frame loop 5 % */
    const int cp2 = !SHORT ? 32 : chunks > 8 ? 16 : chunks > 4 ? 8 : chunks > 2 ? 4 : chunks > 1 ? 2 : 1;
    const int framesPerWarp = 32 / cp2, cl = lane & (cp2 - 1);
    const long long warps = (long long)gridDim.x * (kEncLutThreads / 32);
    for (long long f = ((long long)blockIdx.x * (kEncLutThreads / 32) + (threadIdx.x >> 5)) * framesPerWarp + lane / cp2;
         f < nFrames; f += warps * framesPerWarp) {
        const uint8_t *m = msg + (size_t)f * msgStride;
        uint8_t *out = segs + (size_t)f * segStride;
        const uint16_t *m16 = reinterpret_cast<const uint16_t *>(m);
        uint4 *out16 = reinterpret_cast<uint4 *>(out);
        int k = 0;
        for (; k + 4 <= interiorIters; k += 4) {
            uint32_t a[4], h[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c = 32 * (k + u) + lane; /* whole-warp iterations exist only when cp2 == 32 */
                a[u] = __ldg(m16 + max(c - 1, 0)); /* bytes 2c-2, 2c-1 */
                h[u] = __ldg(m16 + c);             /* bytes 2c, 2c+1   */
            }
            if (k == 0 && lane == 0)
                a[0] = 0u; /* nothing precedes the first byte: the register starts at STARTING_STATE 0 */
#pragma unroll
            for (int u = 0; u < 4; u++)
                out16[32 * (k + u) + lane] = encodeChunkLut<FIXED>(__byte_perm(a[u], h[u], 0x1450), tap0, tap1, lut);
        }
        for (; k < interiorIters; k++) { /* whole-warp iterations left over by the unrolled loop */
            const int c = 32 * k + lane;
            uint32_t a = __ldg(m16 + max(c - 1, 0));
            const uint32_t h = __ldg(m16 + c);
            if (c == 0)
                a = 0u;
            out16[c] = encodeChunkLut<FIXED>(__byte_perm(a, h, 0x1450), tap0, tap1, lut);
        }
        /* the remaining chunks: ragged message end, zero tail (src/convEncode.c:108-119), ragged row end */
        for (int c = 32 * k + cl; c < chunks; c += cp2) {
            const int i1 = 2 * c;
            const uint32_t b0 = (c > 0 && i1 - 1 < frameBytes) ? __ldg(m + i1 - 1) : 0u;
            const uint32_t b1 = (i1 < frameBytes) ? __ldg(m + i1) : 0u;
            const uint32_t b2 = (i1 + 1 < frameBytes) ? __ldg(m + i1 + 1) : 0u;
            const uint4 v = encodeChunkLut<FIXED>((b0 << 24) | (b1 << 16) | (b2 << 8), tap0, tap1, lut);
            uint8_t *dst = out + 16 * c;
            const int cnt = T - 16 * c;
            if (cnt >= 16) {
                *reinterpret_cast<uint4 *>(dst) = v;
            } else { /* ragged end: 8 + 4 + 2 + 1 bytes as needed (dst is 16-byte aligned) */
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                int b = 0;
                if (cnt & 8) {
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(v.x, v.y);
                    b = 8;
                }
                if (cnt & 4) {
                    *reinterpret_cast<uint32_t *>(dst + b) = w[b >> 2];
                    b += 4;
                }
                if (cnt & 2) {
                    *reinterpret_cast<uint16_t *>(dst + b) = (uint16_t)(w[b >> 2] >> (8 * (b & 3)));
                    b += 2;
                }
                if (cnt & 1)
                    dst[b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
            }
        }
    }
}

} // namespace ced
