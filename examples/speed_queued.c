/*
 * The reference's speed loops (speedEncode/speedEncode.c:55-90 and speedDecode/speedDecode.c:72-110: 2048-bit
 * packets, encode / decode them over and over, report information Mbps) with the two-line change of
 * include/viterbiDecoderQueue.h: packets are submitted one by one exactly as before, the library runs them on the
 * GPU a batch at a time.  The first packets are checked against the synchronous convEnc, every decoded packet
 * against its message.
 *
 *   gcc -O2 -std=gnu11 -Iinclude/params/default -Iinclude -o examples/_bin/speed_queued \
 *       examples/speed_queued.c -Lconvolutionalencdec_b200 -lconvencdec_k7 -lced_cuda -pthread
 */
#include "convEncode.h"
#include "viterbiDecoder.h"
#include "viterbiDecoderQueue.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PKT_BYTES 256
#define PKT_SEGS (8 * PKT_BYTES / k + S)
#define PKTS 16384

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    const double seconds = argc > 1 ? atof(argv[1]) : 2.0;
    const int perBatch = argc > 2 ? atoi(argv[2]) : 8192;
    uint8_t (*msg)[PKT_BYTES] = malloc((size_t)PKTS * PKT_BYTES);
    uint8_t (*coded)[PKT_SEGS] = malloc((size_t)PKTS * PKT_SEGS);
    uint8_t (*decoded)[PKT_BYTES] = calloc(PKTS, PKT_BYTES);

    convEncoderState_t enc;
    resetConvEncoder(&enc);
    initConvEncoder(&enc);
    srand(314);
    for (int i = 0; i < PKTS; i++)
        for (int j = 0; j < PKT_BYTES; j++)
            msg[i][j] = (uint8_t)rand();

    convEncQueue_t *eq = convEncQueueCreate(PKT_BYTES, perBatch);
    long encoded = 0;
    const double e0 = now();
    double e1 = e0;
    do {
        for (int i = 0; i < PKTS; i++)
            convEncQueueSubmit(eq, msg[i], coded[i]);      /* was: convEnc(&enc, msg[i], coded[i], PKT_BYTES, true) */
        encoded += convEncQueueFlush(eq);
        e1 = now();
    } while (e1 - e0 < seconds);
    convEncQueueDestroy(eq);
    long encWrong = 0;
    for (int i = 0; i < 64; i++) {
        uint8_t want[PKT_SEGS];
        if (convEnc(&enc, msg[i], want, PKT_BYTES, true) != PKT_SEGS)
            return 1;
        encWrong += memcmp(want, coded[i], PKT_SEGS) != 0;
    }
    printf("Encode rate: %f Mbps (%ld packets)\n", (double)encoded * 8.0 * PKT_BYTES / (e1 - e0) / 1e6, encoded);

    for (int i = 0; i < PKTS; i++)
        for (int j = 7 + i % 5; j < PKT_SEGS; j += 61) /* a correctable sprinkle of channel errors */
            coded[i][j] ^= (uint8_t)(1 + (j & 1));

    viterbiQueue_t *q = viterbiQueueCreate(PKT_SEGS, perBatch);
    for (int i = 0; i < PKTS; i++)  /* warm-up pass */
        viterbiQueueSubmit(q, coded[i], decoded[i]);
    viterbiQueueFlush(q);

    long packets = 0;
    const double t0 = now();
    double t1 = t0;
    do {
        for (int i = 0; i < PKTS; i++)
            viterbiQueueSubmit(q, coded[i], decoded[i]);   /* was: VITERBI_DECODER_HARD(&state, coded[i], decoded[i], PKT_SEGS, true) */
        packets += viterbiQueueFlush(q);
        t1 = now();
    } while (t1 - t0 < seconds);
    viterbiQueueDestroy(q);

    long wrong = 0;
    for (int i = 0; i < PKTS; i++)
        wrong += memcmp(decoded[i], msg[i], PKT_BYTES) != 0;
    printf("Packets: %ld, Packet bits: %d, Batch: %d\n", packets, 8 * PKT_BYTES, perBatch);
    printf("Rate: %f Mbps\n", (double)packets * 8.0 * PKT_BYTES / (t1 - t0) / 1e6);
    printf("%s\n", (wrong || encWrong) ? "Failed: results differ from the synchronous API / the messages" : "Success!");
    return (wrong || encWrong) ? 1 : 0;
}
