/*
 * channel_kernels.cuh -- the BER-mode helpers around the decoder:
 *   berCountKernel   <- bitErrors()          berTestK7/berTestK7.c:45-53
 *   bscChannelKernel <- corruptCodedArray()  berTestK7/berTestK7.c:29-43 (IID flips;
 *                       the generator is counter-based instead of rand(), so the
 *                       stream is independent of batch sharding)
 *   randomBytesKernel<- the `(uint8_t) rand()` message fill, berTestK7.c:135-138
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ced {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ uint64_t keyed(uint64_t seed, uint64_t frame, uint64_t idx)
{
    return mix64(mix64(seed * 0x9E3779B97F4A7C15ULL + frame) + idx * 0xD1B54A32D192ED03ULL);
}

__global__ void __launch_bounds__(256)
berCountKernel(const uint8_t *__restrict__ a, size_t strideA, const uint8_t *__restrict__ b, size_t strideB,
               int nFrames, int bytesPerFrame, unsigned long long *counters, int aligned4)
{
    unsigned long long errs = 0;
    const long long total = (long long)nFrames * (aligned4 ? bytesPerFrame / 4 : bytesPerFrame);
    const int per = aligned4 ? bytesPerFrame / 4 : bytesPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / per;
        const int c = (int)(i - f * per);
        if (aligned4) {
            const uint32_t x = __ldg(reinterpret_cast<const uint32_t *>(a + (size_t)f * strideA) + c);
            const uint32_t y = __ldg(reinterpret_cast<const uint32_t *>(b + (size_t)f * strideB) + c);
            errs += __popc(x ^ y);
        } else {
            errs += __popc((uint32_t)(a[(size_t)f * strideA + c] ^ b[(size_t)f * strideB + c]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        errs += __shfl_xor_sync(0xFFFFFFFFu, errs, o);
    if ((threadIdx.x & 31) == 0 && errs)
        atomicAdd(&counters[0], errs);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&counters[1], (unsigned long long)nFrames * (unsigned long long)bytesPerFrame * 8ULL);
}

__device__ __forceinline__ uint32_t bscFlips(uint64_t seed, uint64_t frame, int t, int n, uint32_t threshold)
{
    uint32_t flips = 0;
    for (int j = 0; j < n; j += 2) {
        const uint64_t h = keyed(seed, frame, (uint64_t)t * 4u + (uint64_t)(j >> 1));
        flips |= ((uint32_t)h < threshold ? 1u : 0u) << j;
        if (j + 1 < n)
            flips |= ((uint32_t)(h >> 32) < threshold ? 1u : 0u) << (j + 1);
    }
    return flips;
}

/* One thread handles 16 consecutive segments of one frame (one 128-bit load/store when the rows are
 * 16-byte aligned, byte accesses otherwise). */
__global__ void __launch_bounds__(256)
bscChannelKernel(uint8_t *segs, size_t segStride, int nFrames, int segsPerFrame, int n, uint32_t threshold,
                 uint64_t seed, uint64_t firstFrame, unsigned long long *counters, int aligned16)
{
    unsigned long long flipsTotal = 0;
    const int chunksPerFrame = (segsPerFrame + 15) / 16;
    const long long total = (long long)nFrames * chunksPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / chunksPerFrame;
        const int t0 = (int)(i - f * chunksPerFrame) * 16;
        uint8_t *p = segs + (size_t)f * segStride + t0;
        const int cnt = min(16, segsPerFrame - t0);
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (int s = 0; s < cnt; s++) {
            const uint32_t fl = bscFlips(seed, firstFrame + (uint64_t)f, t0 + s, n, threshold);
            w[s >> 2] |= fl << (8 * (s & 3));
            flipsTotal += __popc(fl);
        }
        if (aligned16 && (size_t)(t0 + 16) <= segStride) {
            uint4 v = *reinterpret_cast<uint4 *>(p);
            v.x ^= w[0];
            v.y ^= w[1];
            v.z ^= w[2];
            v.w ^= w[3];
            *reinterpret_cast<uint4 *>(p) = v;   /* bytes past segsPerFrame are XORed with 0 */
        } else {
            for (int s = 0; s < cnt; s++)
                p[s] ^= (uint8_t)(w[s >> 2] >> (8 * (s & 3)));
        }
    }
    if (counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            flipsTotal += __shfl_xor_sync(0xFFFFFFFFu, flipsTotal, o);
        if ((threadIdx.x & 31) == 0 && flipsTotal)
            atomicAdd(&counters[0], flipsTotal);
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicAdd(&counters[1], (unsigned long long)nFrames * (unsigned long long)segsPerFrame *
                                        (unsigned long long)n);
    }
}

/* byte-per-segment symbols -> packed symbols (4 per byte, segment t in bits 2*(t%4).. of byte t/4).
 * One thread packs 16 segments into one 32-bit word. */
__global__ void __launch_bounds__(256)
packSymbolsKernel(const uint8_t *__restrict__ segs, size_t segStride, int nFrames, int segsPerFrame,
                  uint8_t *__restrict__ packed, size_t packedStride, int aligned)
{
    const int chunksPerFrame = (segsPerFrame + 15) / 16;
    const long long total = (long long)nFrames * chunksPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / chunksPerFrame;
        const int c = (int)(i - f * chunksPerFrame);
        const uint8_t *src = segs + (size_t)f * segStride + 16 * (size_t)c;
        const int cnt = min(16, segsPerFrame - 16 * c);
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (aligned && (size_t)(16 * c + 16) <= segStride) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            for (int s2 = 0; s2 < cnt; s2++)
                w[s2 >> 2] |= (uint32_t)src[s2] << (8 * (s2 & 3));
        }
        uint32_t out = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t m = w[q] & 0x03030303u;                 /* 4 segments, 2 bits each, one per byte */
            const uint32_t b = (m | (m >> 6) | (m >> 12) | (m >> 18)) & 0xFFu;
            out |= b << (8 * q);
        }
        uint8_t *dst = packed + (size_t)f * packedStride + 4 * (size_t)c;
        const int bytes = (cnt + 3) / 4;
        if (aligned && bytes == 4) {
            *reinterpret_cast<uint32_t *>(dst) = out;
        } else {
            for (int b2 = 0; b2 < bytes; b2++)
                dst[b2] = (uint8_t)(out >> (8 * b2));
        }
    }
}

/* Soft symbols -> packed hard symbols.  Soft format: two int8 per segment (soft value of coded bit 0,
 * then of coded bit 1), BPSK convention bit 0 -> +, bit 1 -> - ; the hard decision is the sign bit
 * (value 0 slices to bit 0).  The reference decodes hard decisions only, so soft input is defined by
 * this slicing (no reference behaviour to match).  One thread slices 16 segments (32 bytes). */
__global__ void __launch_bounds__(256)
sliceSoftSymbolsKernel(const int8_t *__restrict__ soft, size_t softStride, int nFrames, int segsPerFrame,
                       uint8_t *__restrict__ packed, size_t packedStride, int aligned)
{
    const int chunksPerFrame = (segsPerFrame + 15) / 16;
    const long long total = (long long)nFrames * chunksPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / chunksPerFrame;
        const int c = (int)(i - f * chunksPerFrame);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(soft) + (size_t)f * softStride + 32 * (size_t)c;
        const int cnt = min(16, segsPerFrame - 16 * c);
        uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (aligned && (size_t)(32 * c + 32) <= softStride) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(src));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
            for (int b2 = 0; b2 < 2 * cnt; b2++)
                w[b2 >> 2] |= (uint32_t)src[b2] << (8 * (b2 & 3));
        }
        uint32_t out = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {   /* one word = 2 segments -> 4 bits */
            const uint32_t m = (w[q] >> 7) & 0x01010101u;
            out |= ((m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xFu) << (4 * q);
        }
        if (cnt < 16)
            out &= (1u << (2 * cnt)) - 1u;
        uint8_t *dst = packed + (size_t)f * packedStride + 4 * (size_t)c;
        const int bytes = (cnt + 3) / 4;
        if (aligned && bytes == 4) {
            *reinterpret_cast<uint32_t *>(dst) = out;
        } else {
            for (int b2 = 0; b2 < bytes; b2++)
                dst[b2] = (uint8_t)(out >> (8 * b2));
        }
    }
}

__global__ void __launch_bounds__(256)
randomBytesKernel(uint8_t *msg, size_t msgStride, int nFrames, int frameBytes, uint64_t seed, uint64_t firstFrame)
{
    const int wordsPerFrame = (frameBytes + 7) / 8;
    const long long total = (long long)nFrames * wordsPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / wordsPerFrame;
        const int w = (int)(i - f * wordsPerFrame);
        const uint64_t h = keyed(seed ^ 0x6D657373616765ULL, firstFrame + (uint64_t)f, (uint64_t)w);
        uint8_t *dst = msg + (size_t)f * msgStride + 8 * (size_t)w;
        for (int b = 0; b < 8 && 8 * w + b < frameBytes; b++)
            dst[b] = (uint8_t)(h >> (8 * b));
    }
}

} // namespace ced
