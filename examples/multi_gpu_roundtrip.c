/*
 * Plain-C, single-process use of every GPU in the box (include/ced_abi.h "one process, several GPUs"):
 *
 *   1. host batch:  encode nFrames frames, corrupt them, decode them, all through HOST arrays that
 *      ced_{encode,decode}_batch_host_multi shard over the GPUs -- the loops of speedEncode.c:65-67 and
 *      speedDecode.c:78-79 for a batch larger than one GPU should take; every decoded byte is checked;
 *   2. BER mode:    every GPU generates, encodes, corrupts (BSC) and decodes its own frames on the device,
 *      counts bit errors there (berTestK7.c:45-53) and ONE NCCL all-reduce sums the four counters.
 *
 *   usage: multi_gpu_roundtrip [frames (default 2^19)] [frame bits (default 4096)]
 *   exit code 0 = no wrong byte and consistent counters.
 */
#include "ced_abi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define CHECK(call)                                                    \
    do {                                                               \
        if ((call) != CED_OK) {                                        \
            printf("%s failed: %s\n", #call, ced_last_error());        \
            return 1;                                                  \
        }                                                              \
    } while (0)

int main(int argc, char **argv)
{
    const int nFrames = argc > 1 ? atoi(argv[1]) : 1 << 19;
    const int frameBits = argc > 2 ? atoi(argv[2]) : 4096, frameBytes = frameBits / 8;
    const int T = frameBits + 6;                          /* K-1 = 6 tail segments */
    const size_t segStride = ((size_t)T + 15) / 16 * 16;  /* rows padded to 16 bytes */
    ced_multi *m = NULL;
    if (ced_multi_create(NULL, 0, &m) != CED_OK) {
        printf("no GPU: %s\n", ced_last_error());
        return 2;
    }
    const int G = ced_multi_device_count(m);
    ced_code_t code;
    memset(&code, 0, sizeof(code));
    code.constraintLen = 7;
    code.codedBits = 2;
    code.gen[0] = 0113;
    code.gen[1] = 0171;

    /* ---- 1. one host batch over all GPUs ---- */
    void *pm, *ps, *pd;
    CHECK(ced_host_alloc((size_t)nFrames * frameBytes, &pm));
    CHECK(ced_host_alloc((size_t)nFrames * segStride, &ps));
    CHECK(ced_host_alloc((size_t)nFrames * frameBytes, &pd));
    uint8_t *msg = pm, *segs = ps, *dec = pd;
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < (size_t)nFrames * frameBytes; i += 8) {
        x ^= x << 13, x ^= x >> 7, x ^= x << 17;          /* xorshift64 */
        memcpy(msg + i, &x, 8);
    }
    CHECK(ced_encode_batch_host_multi(m, &code, msg, frameBytes, nFrames, frameBytes, segs, segStride));
    /* flip one coded bit in every 37th segment: well inside the code's correcting power */
    for (int f = 0; f < nFrames; f++)
        for (int t = f % 37; t < T; t += 37)
            segs[(size_t)f * segStride + t] ^= (uint8_t)(1u << (t & 1));
    CHECK(ced_decode_batch_host_multi(m, &code, segs, segStride, nFrames, frameBits, dec, frameBytes)); /* warm-up */
    memset(dec, 0, (size_t)nFrames * frameBytes);
    const double t0 = now();
    CHECK(ced_decode_batch_host_multi(m, &code, segs, segStride, nFrames, frameBits, dec, frameBytes));
    const double dt = now() - t0;
    size_t wrong = 0;
    for (size_t i = 0; i < (size_t)nFrames * frameBytes; i++)
        wrong += dec[i] != msg[i];
    double up = 0, down = 0;
    CHECK(ced_multi_probe_copy_ceiling(m, 256u << 20, 3, &up, &down));
    const double h2dBytes = (double)nFrames * (double)segStride;
    printf("%d GPUs: decoded %d frames x %d bits from host memory in %.2f ms = %.1f Gbit/s, %zu wrong bytes\n", G,
           nFrames, frameBits, dt * 1e3, (double)nFrames * frameBits / dt / 1e9, wrong);
    printf("raw pinned-copy ceiling of these GPUs together: H2D %.1f GB/s, D2H %.1f GB/s; the decode moved its "
           "symbols at %.1f GB/s = %.0f %% of it\n", up / 1e9, down / 1e9, h2dBytes / dt / 1e9,
           100.0 * h2dBytes / dt / up);

    /* ---- 2. BER mode: device-resident batches, NCCL-summed counters ---- */
    const int perGpu = 1 << 14, berBits = 2048, berBytes = berBits / 8, berT = berBits + 6;
    const size_t berStride = ((size_t)berT + 15) / 16 * 16;
    uint64_t *counters[64] = {0};
    for (int g = 0; g < G && g < 64; g++) {
        ced_ctx *c = ced_multi_ctx(m, g);
        void *dMsg, *dSegs, *dOut, *dCnt;
        CHECK(ced_device_alloc(c, (size_t)perGpu * berBytes, &dMsg));
        CHECK(ced_device_alloc(c, (size_t)perGpu * berStride, &dSegs));
        CHECK(ced_device_alloc(c, (size_t)perGpu * berBytes, &dOut));
        CHECK(ced_device_alloc(c, 4 * sizeof(uint64_t), &dCnt));
        counters[g] = dCnt;
        const uint64_t first = (uint64_t)g * perGpu;       /* frames are keyed by their global index */
        CHECK(ced_random_bytes(c, dMsg, berBytes, perGpu, berBytes, 9865, first, NULL));
        CHECK(ced_encode_batch(c, &code, dMsg, berBytes, perGpu, berBytes, dSegs, berStride, NULL));
        CHECK(ced_bsc_channel(c, dSegs, berStride, perGpu, berT, 2, 3.716174e-02, 1, first, counters[g], NULL));
        CHECK(ced_decode_batch(c, &code, dSegs, berStride, perGpu, berBits, dOut, berBytes, NULL));
        CHECK(ced_ber_count(c, dOut, berBytes, dMsg, berBytes, perGpu, berBytes, counters[g] + 2, NULL));
        CHECK(ced_sync(c, NULL));
        ced_device_free(c, dMsg);
        ced_device_free(c, dSegs);
        ced_device_free(c, dOut);
    }
    uint64_t before[4] = {0, 0, 0, 0}, sum[4];
    for (int g = 0; g < G; g++) {
        uint64_t v[4];
        CHECK(ced_copy_to_host(ced_multi_ctx(m, g), v, counters[g], sizeof(v)));
        for (int i = 0; i < 4; i++)
            before[i] += v[i];
    }
    CHECK(ced_ber_allreduce(m, counters, 4));
    int consistent = 1;
    for (int g = 0; g < G; g++) {
        CHECK(ced_copy_to_host(ced_multi_ctx(m, g), sum, counters[g], sizeof(sum)));
        consistent &= memcmp(sum, before, sizeof(sum)) == 0;
        ced_device_free(ced_multi_ctx(m, g), counters[g]);
    }
    printf("BER mode, %d GPUs x %d packets of %d bits at channel p = 3.716e-2 (berTestK7.c:96): NCCL %d all-reduce -> "
           "%llu flips / %llu coded bits, %llu errors / %llu decoded bits (BER %.3e; reference expects ~5.18e-4 +- 10 %%); "
           "every GPU holds the sum: %s\n", G, perGpu, berBits, ced_nccl_version(),
           (unsigned long long)sum[0], (unsigned long long)sum[1], (unsigned long long)sum[2], (unsigned long long)sum[3],
           (double)sum[2] / (double)sum[3], consistent ? "yes" : "NO");
    ced_host_free(msg);
    ced_host_free(segs);
    ced_host_free(dec);
    ced_multi_destroy(m);
    return (wrong == 0 && consistent) ? 0 : 1;
}
