/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
 *
 * Thin batch / threading loops around the UNMODIFIED reference objects
 * (/root/reference/src/{convEncode,convHelpers,viterbiDecoder}.c +
 * src/defaultParams/convCodeParams.c), compiled where they lie and linked into
 * oracle/_ref/libced_ref_*.so.  This file only calls the reference's public
 * API (convEncode.h / viterbiDecoder.h); it contains no decoding logic itself.
 * The loops mirror the reference drivers: speedDecode/speedDecode.c:72-110 and
 * berTestK7/berTestK7.c:109-165.
 */
#define _GNU_SOURCE
#include "convEncode.h"
#include "viterbiDecoder.h"
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

int refh_K(void) { return K; }
int refh_n(void) { return n; }
int refh_states(void) { return (int)NUM_STATES; }
size_t refh_state_bytes(void) { return sizeof(viterbiHardState_t); }
uint64_t refh_g(int i) { return g[i]; }

void refh_polys(uint8_t *out)
{
    convEncoderState_t e;
    resetConvEncoder(&e);
    initConvEncoder(&e);
    for (int i = 0; i < n; i++)
        out[i] = (uint8_t)e.polynomials[i];
}

static viterbiHardState_t *refh_new_decoder(void)
{
    viterbiHardState_t *st = (viterbiHardState_t *)aligned_alloc(64, (sizeof(viterbiHardState_t) + 63) & ~(size_t)63);
    VITERBI_RESET(st);
    VITERBI_INIT(st);
    return st;
}

void refh_edge_symm(uint8_t *out)
{
    viterbiHardState_t *st = refh_new_decoder();
    memcpy(out, st->edgeCodedBitsSymm, NUM_STATES / 2);
    free(st);
}

int refh_encode(const uint8_t *in, int bytesIn, uint8_t *segs)
{
    convEncoderState_t e;
    resetConvEncoder(&e);
    initConvEncoder(&e);
    return convEnc(&e, (uint8_t *)in, segs, bytesIn, true);
}

/* chunked encode: `chunk` bytes per call, last call carries last=true */
int refh_encode_chunked(const uint8_t *in, int bytesIn, uint8_t *segs, int chunk)
{
    convEncoderState_t e;
    resetConvEncoder(&e);
    initConvEncoder(&e);
    int done = 0, outSegs = 0;
    while (done < bytesIn) {
        int m = bytesIn - done < chunk ? bytesIn - done : chunk;
        outSegs += convEnc(&e, (uint8_t *)in + done, segs + outSegs, m, done + m == bytesIn);
        done += m;
    }
    return outSegs;
}

int refh_decode_batch(const uint8_t *segs, size_t stride, int nFrames, int segsPerFrame, uint8_t *out,
                      size_t outStride)
{
    viterbiHardState_t *st = refh_new_decoder();
    int bytes = 0;
    for (int f = 0; f < nFrames; f++)
        bytes = VITERBI_DECODER_HARD(st, (uint8_t *)segs + (size_t)f * stride, out + (size_t)f * outStride,
                                     segsPerFrame, true);
    free(st);
    return bytes;
}

/* streaming: `chunk` segments per call with last=false, then a final last=true
 * call; metricsOut (if non-NULL) receives the 64 node metrics after every call
 * ([calls][NUM_STATES]).  Returns bytes written by the last call. */
int refh_decode_chunked(const uint8_t *segs, int segsPerFrame, int chunk, uint8_t *out, uint8_t *metricsOut)
{
    viterbiHardState_t *st = refh_new_decoder();
    int done = 0, call = 0, bytes = 0;
    while (done < segsPerFrame) {
        int m = segsPerFrame - done < chunk ? segsPerFrame - done : chunk;
        bytes = VITERBI_DECODER_HARD(st, (uint8_t *)segs + done, out, m, false);
        if (metricsOut)
            memcpy(metricsOut + (size_t)call * NUM_STATES, *st->nodeMetricsCur, NUM_STATES);
        done += m;
        call++;
    }
    bytes = VITERBI_DECODER_HARD(st, (uint8_t *)segs + done, out, 0, true);
    free(st);
    return bytes;
}

/* berTestK7.c:22-53,109-165 driven through the reference's own functions.
 * counts = {channel flips, coded bits, decoded bit errors, decoded bits}. */
int refh_bertest(unsigned seedOrZero, int pkts, int pktBytes, double p, int64_t *counts, uint8_t *pktsOut,
                 uint8_t *msgOut)
{
    if (seedOrZero)
        srand(seedOrZero);
    const int segsPerPkt = 8 * pktBytes / k + S;
    convEncoderState_t e;
    resetConvEncoder(&e);
    initConvEncoder(&e);
    viterbiHardState_t *st = refh_new_decoder();
    uint8_t *msg = malloc((size_t)pktBytes), *dec = malloc((size_t)pktBytes);
    uint8_t *clean = malloc((size_t)segsPerPkt), *noisy = malloc((size_t)segsPerPkt);
    memset(counts, 0, 4 * sizeof(int64_t));
    for (int it = 0; it < pkts; it++) {
        for (int j = 0; j < pktBytes; j++)
            msg[j] = (uint8_t)rand();
        int ns = convEnc(&e, msg, clean, pktBytes, true);
        counts[1] += (int64_t)ns * n;
        for (int i = 0; i < ns; i++) {
            uint8_t flips = 0;
            for (int j = 0; j < n; j++) {
                uint8_t f = ((double)rand() / RAND_MAX) > p ? 0 : 1;
                flips = (uint8_t)((flips << 1) | f);
                counts[0] += f;
            }
            noisy[i] = clean[i] ^ flips;
        }
        int nb = VITERBI_DECODER_HARD(st, noisy, dec, ns, true);
        counts[3] += (int64_t)nb * 8;
        for (int j = 0; j < pktBytes; j++)
            counts[2] += calcHammingDist(msg[j], dec[j], 8);
        if (pktsOut)
            memcpy(pktsOut + (size_t)it * (size_t)segsPerPkt, noisy, (size_t)segsPerPkt);
        if (msgOut)
            memcpy(msgOut + (size_t)it * (size_t)pktBytes, msg, (size_t)pktBytes);
    }
    free(msg); free(dec); free(clean); free(noisy); free(st);
    return 0;
}

/* ---- multi-core speedDecode loop (bench.py cpu_baseline kind="reference") ---- */
typedef struct {
    const uint8_t *segs;
    size_t stride;
    int nFrames, segsPerFrame, first;
    double budget;
    int64_t bits;
    uint8_t sink;
} refh_speed_arg_t;

static double refh_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *refh_speed_thread(void *p)
{
    refh_speed_arg_t *a = (refh_speed_arg_t *)p;
    viterbiHardState_t *st = refh_new_decoder();
    const int frameBytes = (a->segsPerFrame - S) / 8;
    uint8_t *out = malloc((size_t)frameBytes + 1);
    int f = a->first % a->nFrames;
    const double t0 = refh_now();
    a->bits = 0;
    do {
        for (int rep = 0; rep < 8; rep++) {
            VITERBI_DECODER_HARD(st, (uint8_t *)a->segs + (size_t)f * a->stride, out, a->segsPerFrame, true);
            a->sink ^= out[0];
            a->bits += (int64_t)frameBytes * 8;
            f = (f + 1 == a->nFrames) ? 0 : f + 1;
        }
    } while (refh_now() - t0 < a->budget);
    free(out);
    free(st);
    return NULL;
}

int64_t refh_speed_decode(const uint8_t *segs, size_t stride, int nFrames, int segsPerFrame, int nThreads,
                          double budgetSeconds, double *seconds)
{
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nThreads);
    refh_speed_arg_t *args = calloc((size_t)nThreads, sizeof(refh_speed_arg_t));
    const double t0 = refh_now();
    /* one thread per host core, each pinned to its own CPU of this process's affinity mask (the way
     * speedDecode.c:23,147 pins its worker); with more threads than CPUs the extra ones share round-robin */
    cpu_set_t allowed;
    int cpus[CPU_SETSIZE], nCpus = 0;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &allowed))
                cpus[nCpus++] = c;
    for (int i = 0; i < nThreads; i++) {
        args[i] = (refh_speed_arg_t){segs, stride, nFrames, segsPerFrame, i * 7, budgetSeconds, 0, 0};
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        if (nCpus > 0) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[i % nCpus], &one);
            pthread_attr_setaffinity_np(&attr, sizeof(one), &one);
        }
        if (pthread_create(&th[i], &attr, refh_speed_thread, &args[i]) != 0)
            pthread_create(&th[i], NULL, refh_speed_thread, &args[i]);
        pthread_attr_destroy(&attr);
    }
    int64_t bits = 0;
    for (int i = 0; i < nThreads; i++) {
        pthread_join(th[i], NULL);
        bits += args[i].bits;
    }
    *seconds = refh_now() - t0;
    free(th);
    free(args);
    return bits;
}

/* speedEncode.c:65-104 loop shape */
typedef struct {
    const uint8_t *msgs;
    int nFrames, frameBytes, first;
    double budget;
    int64_t bits;
    uint8_t sink;
} refh_enc_arg_t;

static void *refh_enc_thread(void *p)
{
    refh_enc_arg_t *a = (refh_enc_arg_t *)p;
    convEncoderState_t e;
    resetConvEncoder(&e);
    initConvEncoder(&e);
    uint8_t *segs = malloc((size_t)a->frameBytes * 8 + S);
    int f = a->first % a->nFrames;
    const double t0 = refh_now();
    a->bits = 0;
    do {
        for (int rep = 0; rep < 32; rep++) {
            convEnc(&e, (uint8_t *)a->msgs + (size_t)f * (size_t)a->frameBytes, segs, a->frameBytes, true);
            a->sink ^= segs[5];
            a->bits += (int64_t)a->frameBytes * 8;
            f = (f + 1 == a->nFrames) ? 0 : f + 1;
        }
    } while (refh_now() - t0 < a->budget);
    free(segs);
    return NULL;
}

int64_t refh_speed_encode(const uint8_t *msgs, int nFrames, int frameBytes, int nThreads, double budgetSeconds,
                          double *seconds)
{
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nThreads);
    refh_enc_arg_t *args = calloc((size_t)nThreads, sizeof(refh_enc_arg_t));
    const double t0 = refh_now();
    for (int i = 0; i < nThreads; i++) {
        args[i] = (refh_enc_arg_t){msgs, nFrames, frameBytes, i * 3, budgetSeconds, 0, 0};
        pthread_create(&th[i], NULL, refh_enc_thread, &args[i]);
    }
    int64_t bits = 0;
    for (int i = 0; i < nThreads; i++) {
        pthread_join(th[i], NULL);
        bits += args[i].bits;
    }
    *seconds = refh_now() - t0;
    free(th);
    free(args);
    return bits;
}

/* whole-batch decode on several threads (full-size parity test: every frame of a 2^16-frame batch) */
typedef struct {
    const uint8_t *segs;
    size_t stride, outStride;
    int first, count, segsPerFrame;
    uint8_t *out;
} refh_mt_arg_t;

static void *refh_mt_thread(void *p)
{
    refh_mt_arg_t *a = (refh_mt_arg_t *)p;
    viterbiHardState_t *st = refh_new_decoder();
    for (int f = a->first; f < a->first + a->count; f++)
        VITERBI_DECODER_HARD(st, (uint8_t *)a->segs + (size_t)f * a->stride, a->out + (size_t)f * a->outStride,
                             a->segsPerFrame, true);
    free(st);
    return NULL;
}

int refh_decode_batch_mt(const uint8_t *segs, size_t stride, int nFrames, int segsPerFrame, uint8_t *out,
                         size_t outStride, int nThreads)
{
    if (nThreads < 1)
        nThreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nThreads);
    refh_mt_arg_t *args = calloc((size_t)nThreads, sizeof(refh_mt_arg_t));
    for (int i = 0; i < nThreads; i++) {
        const int lo = (int)((long long)nFrames * i / nThreads), hi = (int)((long long)nFrames * (i + 1) / nThreads);
        args[i] = (refh_mt_arg_t){segs, stride, outStride, lo, hi - lo, segsPerFrame, out};
        pthread_create(&th[i], NULL, refh_mt_thread, &args[i]);
    }
    for (int i = 0; i < nThreads; i++)
        pthread_join(th[i], NULL);
    free(th);
    free(args);
    return 0;
}
