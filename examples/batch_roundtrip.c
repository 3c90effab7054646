/*
 * Plain-C user of the batched entry points (include/ced_abi.h): the GPU equivalent of the loops in
 * speedEncode.c:65-67 and speedDecode.c:78-79 over a whole batch, using HOST buffers only (no CUDA
 * headers needed on the caller's side).
 *
 *   gcc -O2 -std=gnu11 -Iinclude -o examples/batch_roundtrip examples/batch_roundtrip.c \
 *       -Lconvolutionalencdec_b200 -lced_cuda -Wl,-rpath,'$ORIGIN/../convolutionalencdec_b200'
 */
#include "ced_abi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(void)
{
    const int nFrames = 20000, frameBits = 4096, frameBytes = frameBits / 8;
    const int segsPerFrame = frameBits + 6;              /* K-1 = 6 tail segments */
    const size_t segStride = 4112;                       /* rows padded to 16 bytes */
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

    uint8_t *msg = malloc((size_t)nFrames * frameBytes), *dec = malloc((size_t)nFrames * frameBytes);
    uint8_t *segs = calloc((size_t)nFrames, segStride);
    srand(314);
    for (size_t i = 0; i < (size_t)nFrames * frameBytes; i++)
        msg[i] = (uint8_t)rand();
    if (ced_encode_batch_host(ctx, &code, msg, frameBytes, nFrames, frameBytes, segs, segStride) != CED_OK) {
        printf("encode failed: %s\n", ced_last_error());
        return 1;
    }
    /* flip one coded bit in every 40th segment: well inside the code's correcting power */
    long flips = 0;
    for (int f = 0; f < nFrames; f++)
        for (int t = 7; t < segsPerFrame; t += 40, flips++)
            segs[(size_t)f * segStride + t] ^= 1u << ((t / 40) & 1);
    if (ced_decode_batch_host(ctx, &code, segs, segStride, nFrames, frameBits, dec, frameBytes) != CED_OK) {
        printf("decode failed: %s\n", ced_last_error());
        return 1;
    }
    long wrong = 0;
    for (size_t i = 0; i < (size_t)nFrames * frameBytes; i++)
        wrong += msg[i] != dec[i];
    printf("%d frames x %d bits, %ld channel bit flips, %ld wrong bytes after decoding, %llu kernel launches\n",
           nFrames, frameBits, flips, wrong, (unsigned long long)ced_launch_count(ctx));
    ced_ctx_destroy(ctx);
    free(msg);
    free(dec);
    free(segs);
    return wrong == 0 ? 0 : 1;
}
