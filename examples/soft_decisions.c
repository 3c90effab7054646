/*
 * Plain-C user of the soft-decision entry points (include/ced_abi.h): one batch of K=7 rate-1/2 frames goes through a
 * BPSK + AWGN channel on the device and is decoded three ways from the SAME channel output -- hard decisions
 * (ced_decode_batch), int8 reliabilities (ced_decode_batch_soft) and 3-bit soft decisions in the hard format's wire size
 * (ced_quantize_soft + ced_decode_batch_softq) -- and the decoded bit-error rates are printed.  No CUDA headers needed.
 *
 *   gcc -O2 -std=gnu11 -Iinclude -o examples/soft_decisions examples/soft_decisions.c \
 *       -Lconvolutionalencdec_b200 -lced_cuda -lm -Wl,-rpath,'$ORIGIN/../convolutionalencdec_b200'
 */
#include "ced_abi.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call)                                                     \
    do {                                                                \
        if ((call) != CED_OK) {                                         \
            printf("%s failed: %s\n", #call, ced_last_error());         \
            return 1;                                                   \
        }                                                               \
    } while (0)

int main(int argc, char **argv)
{
    const double ebn0Db = argc > 1 ? atof(argv[1]) : 3.0;
    const int nFrames = 32768, frameBits = 2048, frameBytes = frameBits / 8, T = frameBits + 6;
    const size_t segStride = 2064, softStride = 4112;    /* rows padded to 16 bytes */
    ced_ctx *ctx = NULL;
    if (ced_ctx_create(0, &ctx) != CED_OK) {
        printf("no GPU context: %s\n", ced_last_error());
        return 2;
    }
    ced_code_t code;
    memset(&code, 0, sizeof(code));
    code.constraintLen = 7;
    code.codedBits = 2;
    code.gen[0] = 0113;
    code.gen[1] = 0171;

    void *dMsg, *dSegs, *dSoft, *dSyms, *dHard, *dOut, *dCnt;
    CHECK(ced_device_alloc(ctx, (size_t)nFrames * frameBytes, &dMsg));
    CHECK(ced_device_alloc(ctx, (size_t)nFrames * segStride, &dSegs));
    CHECK(ced_device_alloc(ctx, (size_t)nFrames * softStride, &dSoft));
    CHECK(ced_device_alloc(ctx, (size_t)nFrames * segStride, &dSyms));
    CHECK(ced_device_alloc(ctx, (size_t)nFrames * segStride, &dHard));
    CHECK(ced_device_alloc(ctx, (size_t)nFrames * frameBytes, &dOut));
    CHECK(ced_device_alloc(ctx, 8 * sizeof(uint64_t), &dCnt));          /* zero-filled */
    uint64_t *cnt = (uint64_t *)dCnt;

    const double amplitude = 32.0, sigma = pow(10.0, -ebn0Db / 20.0);  /* rate 1/2: sigma = 1 / sqrt(Eb/N0) */
    CHECK(ced_random_bytes(ctx, dMsg, frameBytes, nFrames, frameBytes, 314, 0, NULL));
    CHECK(ced_encode_batch(ctx, &code, dMsg, frameBytes, nFrames, frameBytes, dSegs, segStride, NULL));
    CHECK(ced_awgn_channel(ctx, dSegs, segStride, nFrames, T, dSoft, softStride, amplitude, sigma, 2718, 0, cnt, NULL));

    CHECK(ced_slice_soft_to_bytes(ctx, dSoft, softStride, nFrames, T, dHard, segStride, NULL));
    CHECK(ced_decode_batch(ctx, &code, dHard, segStride, nFrames, frameBits, dOut, frameBytes, NULL));
    CHECK(ced_ber_count(ctx, dOut, frameBytes, dMsg, frameBytes, nFrames, frameBytes, cnt + 2, NULL));

    CHECK(ced_decode_batch_soft(ctx, &code, dSoft, softStride, nFrames, frameBits, dOut, frameBytes, NULL));
    CHECK(ced_ber_count(ctx, dOut, frameBytes, dMsg, frameBytes, nFrames, frameBytes, cnt + 4, NULL));

    CHECK(ced_quantize_soft(ctx, dSoft, softStride, nFrames, T, 0.6 * sigma * amplitude, dSyms, segStride, NULL));
    CHECK(ced_decode_batch_softq(ctx, &code, dSyms, segStride, nFrames, frameBits, dOut, frameBytes, NULL));
    CHECK(ced_ber_count(ctx, dOut, frameBytes, dMsg, frameBytes, nFrames, frameBytes, cnt + 6, NULL));

    uint64_t h[8];
    CHECK(ced_copy_to_host(ctx, h, dCnt, sizeof(h)));
    printf("Eb/N0 %.1f dB, %d frames x %d bits: channel BER %.4e | decoded BER hard %.3e, int8 soft %.3e, 3-bit soft %.3e\n",
           ebn0Db, nFrames, frameBits, (double)h[0] / (double)h[1], (double)h[2] / (double)h[3], (double)h[4] / (double)h[5],
           (double)h[6] / (double)h[7]);
    const int ok = h[4] <= h[6] && h[6] * 4 < h[2];   /* int8 soft <= 3-bit soft << hard */
    ced_device_free(ctx, dMsg); ced_device_free(ctx, dSegs); ced_device_free(ctx, dSoft); ced_device_free(ctx, dSyms);
    ced_device_free(ctx, dHard); ced_device_free(ctx, dOut); ced_device_free(ctx, dCnt);
    ced_ctx_destroy(ctx);
    return ok ? 0 : 1;
}
