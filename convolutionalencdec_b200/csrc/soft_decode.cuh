/*
 * soft_decode.cuh -- batched SOFT-decision Viterbi decode of the K=7 rate-1/2 code, sm_100a.
 *
 * The reference decodes hard decisions only (src/viterbiDecoderButterflyk1.c:104-115 uses calcHammingDist,
 * src/viterbiDecoder.c:260-285); north_star kernel (1) asks for "soft or hard symbols".  This kernel is the
 * hard forward kernel (decode_batch.cuh) with 16-bit metrics: the same persistent unit scheduler, the same
 * survivor stream (64 decision bits per step, two steps per 128-bit store, group-major) and therefore the
 * same traceback kernel, instantiated for the 16-bit lane layout.  Definition of the result:
 * the weighted Hamming cost stated in trellis_swar16.cuh (the test suite holds a CPU restatement of it); inputs of
 * constant magnitude reproduce the hard decoder bit for bit.
 *
 * Wire format: two int8 per segment (generator 0 first), rows of 2*(frameBits+6) bytes, base and stride
 * multiples of 16 bytes.  A warp stages tiles of 32 frames x 48 segments: 6 coalesced 128-bit loads per lane,
 * prefetched one tile ahead, converted while staging to one 32-bit word per step (trellis_swar16.cuh
 * softWord) -- so the consumer pays one LDS.64 per two steps plus 2 PRMT + 7 FMA-pipe ops per step for the
 * branch words, instead of a table lookup.
 */
#pragma once
#include "decode_batch.cuh"
#include "trellis_swar16.cuh"

namespace ced {

constexpr int kSoftChunk = 48;                 /* steps per staged tile row = renormalisation period */
constexpr int kSoftPieces = kSoftChunk * 2 / 16; /* 16-byte pieces per tile row */
constexpr int kSoftPitch = 200;                /* bytes per tile row in shared memory: 48 words + 2; pitch / 8 odd ->
                                                  LDS.64 of 32 rows is conflict-free */
constexpr int kSoftStateUint4 = 8;             /* 32 metric registers = 8 uint4 per lane in the hand-off slot */

__device__ __forceinline__ void softLoadTile(uint4 (&v)[kSoftPieces], const int8_t *__restrict__ soft, size_t stride,
                                             long long frame0, int nFrames, int t0, int T, int lane)
{
    const uint8_t *base = reinterpret_cast<const uint8_t *>(soft);
    const uint8_t *bufHi = base + (size_t)(nFrames - 1) * stride + (size_t)2 * T;
#pragma unroll
    for (int i = 0; i < kSoftPieces; i++) {
        const int piece = i * 32 + lane;
        const int row = piece / kSoftPieces, pc = piece % kSoftPieces;
        const long long f = frame0 + row;
        v[i] = make_uint4(0, 0, 0, 0);
        if (f < nFrames) {
            const uint8_t *src = base + (size_t)f * stride + (size_t)2 * t0 + 16 * pc;
            if (src + 16 <= bufHi) {
                v[i] = __ldg(reinterpret_cast<const uint4 *>(src));
            } else {
                uint32_t w[4] = {0, 0, 0, 0};
                for (int b = 0; b < 16; b++)
                    if (src + b < bufHi)
                        w[b >> 2] |= (uint32_t)src[b] << (8 * (b & 3));
                v[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

/* two segments (s0a s1a s0b s1b as bytes of w) -> their two staged words, computed two lanes at a time:
 * u = s + 128 in [0,255], cost'(0) = 512 - u0 - u1, cost'(1) = 256 + u0 - u1 (trellis_swar16.cuh softWord) */
__device__ __forceinline__ uint2 softWordsOfPair(uint32_t w)
{
    const uint32_t u = w ^ 0x80808080u;
    const uint32_t U0 = prmt(u, 0u, 0x4240u), U1 = prmt(u, 0u, 0x4341u); /* (u0a, u0b), (u1a, u1b) as 16-bit lanes */
    const uint32_t C0 = 0x02000200u - U0 - U1;
    const uint32_t C1 = 0x01000100u + U0 - U1;
    return make_uint2(prmt(C0, C1, 0x5410u), prmt(C0, C1, 0x7632u));
}

__device__ __forceinline__ void softStoreTile(uint8_t *tile, const uint4 (&v)[kSoftPieces], int lane)
{
#pragma unroll
    for (int i = 0; i < kSoftPieces; i++) {
        const int piece = i * 32 + lane;
        const int row = piece / kSoftPieces, pc = piece % kSoftPieces;
        uint2 *dst = reinterpret_cast<uint2 *>(tile + row * kSoftPitch + pc * 32); /* 8 segments x 4 bytes */
        dst[0] = softWordsOfPair(v[i].x);
        dst[1] = softWordsOfPair(v[i].y);
        dst[2] = softWordsOfPair(v[i].z);
        dst[3] = softWordsOfPair(v[i].w);
    }
}

template <class Code, int PH>
__device__ __forceinline__ void softStep(uint32_t (&R)[32], uint32_t W, uint32_t minusOne, uint32_t minusTwo,
                                         uint32_t &t0, uint32_t &t1)
{
    uint32_t X[4], E[4];
    softBranchWords<Code, PH>(W, minusTwo, X, E);
    acsStep16<Code, PH>(R, X, E, minusOne, t0, t1);
}

/*
 * Persistent forward kernel, soft symbols: see k7ForwardKernel for the unit scheduler (units = 32-frame group
 * x chunksPerUnit chunks, handed out chunk-major by one atomic counter; metrics travel between units through
 * an L2-resident slot guarded by an acquire/release counter).
 */
template <class Code>
__global__ void __launch_bounds__(kFwdThreads)
k7SoftForwardKernel(const int8_t *__restrict__ soft, size_t stride, int nFrames, int T, uint4 *__restrict__ surv,
                    uint32_t minusOne, FwdSched sched, int chunksPerUnit)
{
    __shared__ __align__(16) uint8_t sTile[kFwdThreads / 32][32 * kSoftPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile = sTile[warp];
    const uint32_t minusTwo = minusOne << 1;
    const size_t pairs = (size_t)(T / 2);
    const unsigned groups = (unsigned)((nFrames + 31) / 32);
    const unsigned chunks = (unsigned)((T + kSoftChunk - 1) / kSoftChunk);
    const unsigned unitsPerGroup = (chunks + chunksPerUnit - 1) / chunksPerUnit;
    const unsigned total = groups * unitsPerGroup;

    auto grab = [&]() -> unsigned {
        unsigned v = 0;
        if (lane == 0)
            v = atomicAdd(sched.counter, 1u);
        return __shfl_sync(0xFFFFFFFFu, v, 0);
    };

    unsigned u = grab();
    uint4 pre[kSoftPieces];
    if (u < total)
        softLoadTile(pre, soft, stride, 32LL * (u % groups), nFrames, (int)((u / groups) * chunksPerUnit) * kSoftChunk,
                     T, lane);

    while (u < total) {
        const unsigned g = u % groups, su = u / groups;
        const unsigned cFirst = su * chunksPerUnit, cEnd = min(chunks, cFirst + chunksPerUnit);
        const long long frame0 = 32LL * g;
        const bool live = frame0 + lane < nFrames;
        uint4 *stateSlot = sched.state + ((size_t)g * kSoftStateUint4) * 32 + lane;

        uint32_t R[32];
        if (su == 0) {
            initMetrics16(R);
        } else {
            if (lane == 0)
                while (ldAcquire(sched.done + g) < (int)su)
                    __nanosleep(200);
            __syncwarp();
            __threadfence();
#pragma unroll
            for (int i = 0; i < kSoftStateUint4; i++) {
                const uint4 v = __ldcg(stateSlot + i * 32);
                R[4 * i] = v.x;
                R[4 * i + 1] = v.y;
                R[4 * i + 2] = v.z;
                R[4 * i + 3] = v.w;
            }
        }
        unsigned un = total;
        for (unsigned c = cFirst; c < cEnd; c++) {
            const int t0 = (int)c * kSoftChunk;
            __syncwarp();
            softStoreTile(tile, pre, lane);
            if (c + 1 < cEnd) {
                softLoadTile(pre, soft, stride, frame0, nFrames, t0 + kSoftChunk, T, lane);
            } else {
                un = grab();
                if (un < total)
                    softLoadTile(pre, soft, stride, 32LL * (un % groups), nFrames,
                                 (int)((un / groups) * chunksPerUnit) * kSoftChunk, T, lane);
            }
            __syncwarp();
            const uint2 *p = reinterpret_cast<const uint2 *>(tile + lane * kSoftPitch);
            uint4 *o = surv + ((size_t)g * pairs + (size_t)(t0 / 2)) * 32 + lane;
            const int steps = min(kSoftChunk, T - t0);
            const int full = steps / 6;
            for (int it = 0; it < full; it++) {
                uint4 s;
                const uint2 w01 = p[0], w23 = p[1], w45 = p[2];
                softStep<Code, 0>(R, w01.x, minusOne, minusTwo, s.x, s.y);
                softStep<Code, 1>(R, w01.y, minusOne, minusTwo, s.z, s.w);
                if (live) o[0] = s;
                softStep<Code, 2>(R, w23.x, minusOne, minusTwo, s.x, s.y);
                softStep<Code, 3>(R, w23.y, minusOne, minusTwo, s.z, s.w);
                if (live) o[32] = s;
                softStep<Code, 4>(R, w45.x, minusOne, minusTwo, s.x, s.y);
                softStep<Code, 5>(R, w45.y, minusOne, minusTwo, s.z, s.w);
                if (live) o[64] = s;
                p += 3;
                o += 96;
            }
            /* T is even, so the remainder is 0, 2 or 4 steps (end of the frame only) */
            const int rem = steps - 6 * full;
            if (rem >= 2) {
                uint4 s;
                const uint2 w01 = p[0];
                softStep<Code, 0>(R, w01.x, minusOne, minusTwo, s.x, s.y);
                softStep<Code, 1>(R, w01.y, minusOne, minusTwo, s.z, s.w);
                if (live) o[0] = s;
            }
            if (rem >= 4) {
                uint4 s;
                const uint2 w23 = p[1];
                softStep<Code, 2>(R, w23.x, minusOne, minusTwo, s.x, s.y);
                softStep<Code, 3>(R, w23.y, minusOne, minusTwo, s.z, s.w);
                if (live) o[32] = s;
            }
            if (c + 1 < chunks)
                renorm16(R); /* every 48 steps: candidates stay below 2^15 (trellis_swar16.cuh) */
        }
        if (cEnd < chunks) {
#pragma unroll
            for (int i = 0; i < kSoftStateUint4; i++)
                __stcg(stateSlot + i * 32, make_uint4(R[4 * i], R[4 * i + 1], R[4 * i + 2], R[4 * i + 3]));
            __threadfence();
            __syncwarp();
            if (lane == 0)
                stRelease(sched.done + g, (int)su + 1);
        }
        u = un;
    }
}

/* BPSK over AWGN, quantised to int8 (BER-sweep input, berTestK7's channel generalised to soft output):
 * coded bit b -> amplitude * (1 - 2b) + sigma * amplitude * N(0,1), rounded, clamped to [-127, 127].  The
 * normal deviates come from a counter-based hash of (seed, frame index, segment, bit) through Box-Muller,
 * so the result does not depend on how frames are sharded.  counters (may be NULL): [0] += hard-decision
 * errors (sign of the soft value disagrees with the coded bit), [1] += coded bits. */
__device__ __forceinline__ uint32_t softHash(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (uint32_t)x;
}

__global__ void awgnChannelKernel(const uint8_t *__restrict__ segs, size_t segStride, int nFrames, int segsPerFrame,
                                  int8_t *__restrict__ soft, size_t softStride, float amplitude, float sigma,
                                  uint64_t seed, uint64_t firstFrame, unsigned long long *counters)
{
    const long long total = (long long)nFrames * segsPerFrame;
    unsigned long long flips = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / segsPerFrame;
        const int t = (int)(i - f * segsPerFrame);
        const uint32_t seg = segs[(size_t)f * segStride + t];
        const uint64_t key = (seed * 0x9E3779B97F4A7C15ULL) ^ ((firstFrame + (uint64_t)f) << 20) ^ (uint64_t)t;
        const uint32_t h0 = softHash(key), h1 = softHash(key ^ 0xD1B54A32D192ED03ULL);
        /* one Box-Muller pair serves the two coded bits of the segment */
        const float u0 = ((float)(h0 >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = (float)(h1 >> 8) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u0));
        float sn, cs;
        sincospif(2.0f * u1, &sn, &cs);
        const float z[2] = {rad * cs, rad * sn};
        int8_t out[2];
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int bit = (seg >> b) & 1;
            const float y = amplitude * ((bit ? -1.0f : 1.0f) + sigma * z[b]);
            const int q = max(-127, min(127, __float2int_rn(y)));
            out[b] = (int8_t)q;
            flips += (unsigned)((q < 0) != (bit != 0));
        }
        *reinterpret_cast<char2 *>(soft + (size_t)f * softStride + 2 * (size_t)t) = make_char2(out[0], out[1]);
    }
    if (counters) {
        for (int o = 16; o > 0; o >>= 1)
            flips += __shfl_down_sync(0xFFFFFFFFu, flips, o);
        if ((threadIdx.x & 31) == 0 && flips)
            atomicAdd(counters, flips);
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicAdd(counters + 1, (unsigned long long)total * 2ull);
    }
}

/* soft values -> hard byte-per-segment symbols (sign bit; 0 slices to bit 0), the input of the hard decoder
 * on the same channel output: what a receiver without soft information would see */
__global__ void sliceSoftToBytesKernel(const int8_t *__restrict__ soft, size_t softStride, int nFrames, int segsPerFrame,
                                       uint8_t *__restrict__ segs, size_t segStride)
{
    const long long total = (long long)nFrames * segsPerFrame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / segsPerFrame;
        const int t = (int)(i - f * segsPerFrame);
        const char2 s = *reinterpret_cast<const char2 *>(soft + (size_t)f * softStride + 2 * (size_t)t);
        segs[(size_t)f * segStride + t] = (uint8_t)((s.x < 0 ? 1 : 0) | (s.y < 0 ? 2 : 0));
    }
}

} // namespace ced
