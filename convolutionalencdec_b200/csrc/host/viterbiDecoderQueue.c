/*
 * include/viterbiDecoderQueue.h: collects whole packets into page-locked batches and decodes (or encodes) a batch
 * with one ced_decode_batch_host / ced_encode_batch_host call on a worker thread while the caller fills the other
 * batch.
 */
#include "viterbiDecoderQueue.h"
#include "ced_abi.h"
#include "convCodeParams.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint8_t *segs;      /* [maxPackets][stride], page-locked */
    uint8_t *bytes;     /* [maxPackets][outBytes], page-locked */
    uint8_t **dest;     /* where packet i's bytes go */
    int count;
    int busy;           /* handed to the worker, not yet delivered */
} queueBatch_t;

struct viterbiQueue {
    int encode;         /* 0: segments in, bytes out; 1: bytes in, segments out */
    int segments, outBytes, maxPackets;   /* outBytes = information bytes per packet */
    size_t stride;
    queueBatch_t batch[2];
    int fill;           /* batch the producer writes into */
    long delivered;
    int stop;
    pthread_t worker;
    pthread_mutex_t mu;
    pthread_cond_t wake, done;
    ced_ctx *ctx;
    ced_code_t code;
};

static void queueFail(const char *what)
{
    printf("viterbiQueue: %s: %s\n", what, ced_last_error());
    exit(1);
}

static void *queueWorker(void *arg)
{
    viterbiQueue_t *q = (viterbiQueue_t *)arg;
    pthread_mutex_lock(&q->mu);
    for (int next = 0;;) {
        while (!q->batch[next].busy && !q->stop)
            pthread_cond_wait(&q->wake, &q->mu);
        if (!q->batch[next].busy)
            break;
        queueBatch_t *b = &q->batch[next];
        pthread_mutex_unlock(&q->mu);
        if (q->encode) {
            if (ced_encode_batch_host(q->ctx, &q->code, b->bytes, (size_t)q->outBytes, b->count, q->outBytes, b->segs,
                                      q->stride) != CED_OK)
                queueFail("encode");
            for (int i = 0; i < b->count; i++)
                memcpy(b->dest[i], b->segs + (size_t)i * q->stride, (size_t)q->segments);
        } else {
            if (ced_decode_batch_host(q->ctx, &q->code, b->segs, q->stride, b->count, 8 * q->outBytes, b->bytes,
                                      (size_t)q->outBytes) != CED_OK)
                queueFail("decode");
            for (int i = 0; i < b->count; i++)
                memcpy(b->dest[i], b->bytes + (size_t)i * q->outBytes, (size_t)q->outBytes);
        }
        pthread_mutex_lock(&q->mu);
        q->delivered += b->count;
        b->count = 0;
        b->busy = 0;
        pthread_cond_broadcast(&q->done);
        next ^= 1; /* batches are handed over alternately */
    }
    pthread_mutex_unlock(&q->mu);
    return NULL;
}

static viterbiQueue_t *queueCreate(int segmentsPerPacket, int maxPackets, int encode)
{
    const int bits = (segmentsPerPacket - S) * k;
    if (segmentsPerPacket <= S || bits % 8 != 0 || maxPackets <= 0) {
        printf("viterbiQueue: packets must be a whole number of bytes plus %d tail segments\n", S);
        exit(1);
    }
    viterbiQueue_t *q = (viterbiQueue_t *)calloc(1, sizeof(*q));
    if (!q)
        return NULL;
    q->encode = encode;
    q->segments = segmentsPerPacket;
    q->outBytes = bits / 8;
    q->maxPackets = maxPackets;
    q->stride = ((size_t)segmentsPerPacket + 15) / 16 * 16;
    q->ctx = ced_default_ctx();
    if (!q->ctx)
        queueFail("no GPU context");
    q->code.constraintLen = K;
    q->code.codedBits = n;
    for (int i = 0; i < n; i++)
        q->code.gen[i] = g[i];
    for (int i = 0; i < 2; i++) {
        queueBatch_t *b = &q->batch[i];
        if (ced_host_alloc((size_t)maxPackets * q->stride, (void **)&b->segs) != CED_OK ||
            ced_host_alloc((size_t)maxPackets * (size_t)q->outBytes, (void **)&b->bytes) != CED_OK)
            queueFail("page-locked allocation");
        memset(b->segs, 0, (size_t)maxPackets * q->stride);
        b->dest = (uint8_t **)calloc((size_t)maxPackets, sizeof(uint8_t *));
        if (!b->dest)
            queueFail("allocation");
    }
    pthread_mutex_init(&q->mu, NULL);
    pthread_cond_init(&q->wake, NULL);
    pthread_cond_init(&q->done, NULL);
    if (pthread_create(&q->worker, NULL, queueWorker, q) != 0)
        queueFail("worker thread");
    return q;
}

viterbiQueue_t *viterbiQueueCreate(int segmentsPerPacket, int maxPackets)
{
    return queueCreate(segmentsPerPacket, maxPackets, 0);
}

convEncQueue_t *convEncQueueCreate(int bytesPerPacket, int maxPackets)
{
    if (bytesPerPacket <= 0) {
        printf("convEncQueue: packets must have at least one byte\n");
        exit(1);
    }
    return queueCreate(8 * bytesPerPacket / k + S, maxPackets, 1);
}

/* hand the filling batch to the worker and continue in the other one once it is free */
static void queueHandOver(viterbiQueue_t *q)
{
    pthread_mutex_lock(&q->mu);
    q->batch[q->fill].busy = 1;
    pthread_cond_signal(&q->wake);
    q->fill ^= 1;
    while (q->batch[q->fill].busy)
        pthread_cond_wait(&q->done, &q->mu);
    pthread_mutex_unlock(&q->mu);
}

int viterbiQueueSubmit(viterbiQueue_t *q, const uint8_t *codedSegments, uint8_t *uncoded)
{
    queueBatch_t *b = &q->batch[q->fill];
    memcpy(b->segs + (size_t)b->count * q->stride, codedSegments, (size_t)q->segments);
    b->dest[b->count] = uncoded;
    if (++b->count == q->maxPackets)
        queueHandOver(q);
    return q->outBytes;
}

int convEncQueueSubmit(convEncQueue_t *q, const uint8_t *uncoded, uint8_t *codedSegments)
{
    queueBatch_t *b = &q->batch[q->fill];
    memcpy(b->bytes + (size_t)b->count * q->outBytes, uncoded, (size_t)q->outBytes);
    b->dest[b->count] = codedSegments;
    if (++b->count == q->maxPackets)
        queueHandOver(q);
    return q->segments;
}

long convEncQueueFlush(convEncQueue_t *q)
{
    return viterbiQueueFlush(q);
}

void convEncQueueDestroy(convEncQueue_t *q)
{
    viterbiQueueDestroy(q);
}

long viterbiQueueFlush(viterbiQueue_t *q)
{
    if (q->batch[q->fill].count > 0)
        queueHandOver(q);
    pthread_mutex_lock(&q->mu);
    while (q->batch[0].busy || q->batch[1].busy)
        pthread_cond_wait(&q->done, &q->mu);
    const long nDone = q->delivered;
    q->delivered = 0;
    pthread_mutex_unlock(&q->mu);
    return nDone;
}

void viterbiQueueDestroy(viterbiQueue_t *q)
{
    if (!q)
        return;
    viterbiQueueFlush(q);
    pthread_mutex_lock(&q->mu);
    q->stop = 1;
    pthread_cond_signal(&q->wake);
    pthread_mutex_unlock(&q->mu);
    pthread_join(q->worker, NULL);
    for (int i = 0; i < 2; i++) {
        ced_host_free(q->batch[i].segs);
        ced_host_free(q->batch[i].bytes);
        free(q->batch[i].dest);
    }
    pthread_mutex_destroy(&q->mu);
    pthread_cond_destroy(&q->wake);
    pthread_cond_destroy(&q->done);
    free(q);
}
