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
__global__ void __launch_bounds__(kEncThreads)
encodeBatchKernel(const uint8_t *__restrict__ msg, size_t msgStride, int nFrames, int frameBytes,
                  uint8_t *__restrict__ segs, size_t segStride, int tailSegs, int K, int n, EncTaps taps,
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

/*
 * Fast path for the production code (K=7, g={0113,0171}, byte-per-segment output, 16-byte aligned
 * segment rows).  ncu on the kernel above showed 82 % issue-slot utilisation at 31 % DRAM throughput
 * (profiles/r1_final_ncu_full_summary.txt) and each thread's load -> compute -> store chain was exposed
 * once per item.  Here a CTA first stages the messages of kEncSmemFrames frames in shared memory
 * (one global latency per CTA), then every warp emits whole 512-segment spans: lane l produces
 * segments [512*span + 16*l, +16) from three staged bytes and the warp's STG.128 covers 512
 * contiguous bytes.  Index math is a couple of adds per item.
 */
constexpr int kEncSmemFrames = 8;

__device__ __forceinline__ uint4 encode16Fixed(uint32_t win)
{
    /* taps 0x69 = bits 0,3,5,6 and 0x4F = bits 0,1,2,3,6 (src/convEncode.c:13-17 on 0113 / 0171) */
    const uint32_t s3 = win << 3, s6 = win << 6;
    const uint32_t c0 = (win ^ s3 ^ (win << 5) ^ s6) >> 8;
    const uint32_t c1 = (win ^ (win << 1) ^ (win << 2) ^ s3 ^ s6) >> 8;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t n0 = (c0 >> (4 * q)) & 0xFu, n1 = (c1 >> (4 * q)) & 0xFu;
        w[q] = ((n0 * 0x00204081u) & 0x01010101u) + 2u * ((n1 * 0x00204081u) & 0x01010101u);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(kEncThreads)
encodeBatchSmemKernel(const uint8_t *__restrict__ msg, size_t msgStride, int nFrames, int frameBytes,
                      uint8_t *__restrict__ segs, size_t segStride, int tailSegs, int msgAligned16)
{
    extern __shared__ __align__(16) uint8_t sMsg[];   /* [frames][16 zero bytes | message | >= 16 zero bytes] */
    const int rowPitch = (frameBytes + 47) / 16 * 16; /* 16 in front + message + tail padding */
    const int T = 8 * frameBytes + tailSegs;
    const int spans = (T + 511) / 512;
    const long long frameBase = (long long)blockIdx.x * kEncSmemFrames;
    const int framesHere = (int)min((long long)kEncSmemFrames, (long long)nFrames - frameBase);

    /* stage: zero fill, then the messages */
    for (int i = threadIdx.x; i < framesHere * rowPitch / 16; i += kEncThreads)
        reinterpret_cast<uint4 *>(sMsg)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (msgAligned16 && (frameBytes & 15) == 0) {
        const int vecPerRow = frameBytes / 16;
        for (int i = threadIdx.x; i < framesHere * vecPerRow; i += kEncThreads) {
            const int fl = i / vecPerRow, v = i - fl * vecPerRow;
            reinterpret_cast<uint4 *>(sMsg + fl * rowPitch + 16)[v] =
                __ldg(reinterpret_cast<const uint4 *>(msg + (size_t)(frameBase + fl) * msgStride) + v);
        }
    } else {
        for (int i = threadIdx.x; i < framesHere * frameBytes; i += kEncThreads) {
            const int fl = i / frameBytes, b = i - fl * frameBytes;
            sMsg[fl * rowPitch + 16 + b] = __ldg(msg + (size_t)(frameBase + fl) * msgStride + b);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int unit = warp; unit < framesHere * spans; unit += kEncThreads / 32) {
        const int fl = unit / spans, span = unit - fl * spans;
        const int seg0 = 512 * span + 16 * lane;
        if (seg0 >= T)
            continue;
        /* message bytes 2c-1, 2c, 2c+1 with c = seg0/16; byte -1 and bytes past the end are zero */
        const uint8_t *p = sMsg + fl * rowPitch + 16 + 64 * span + 2 * lane;
        const uint32_t b0 = p[-1];
        const uint32_t b12 = *reinterpret_cast<const uint16_t *>(p);   /* b1 | b2 << 8 */
        const uint32_t win = __brev((b0 << 24) | ((b12 & 0xFFu) << 16) | ((b12 >> 8) << 8));
        const uint4 v = encode16Fixed(win);
        uint8_t *dst = segs + (size_t)(frameBase + fl) * segStride + seg0;
        if (seg0 + 16 <= T) {
            *reinterpret_cast<uint4 *>(dst) = v;
        } else {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            for (int s2 = 0; s2 < T - seg0; s2++)
                dst[s2] = (uint8_t)(w[s2 >> 2] >> (8 * (s2 & 3)));
        }
    }
}

} // namespace ced
