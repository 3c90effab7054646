/*
 * ced_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, run-time parameterised restatement of the algorithm implemented
 * by ucb-cyarp/ConvolutionalEncDec (k = 1 codes), used as the parity checker
 * for the CUDA path.  Nothing under convolutionalencdec_b200/ may link, load
 * or call this file: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the handTraced literals     (handTracedTest/handTraced.c:29,38,55,66,72-111)
 *   - the K=7 KATs of SURVEY 8(c) (poly 0x69/0x4f, edge table, 38-segment vector)
 *   - the berTestK7 golden counts (berTestK7/berTestK7.c with srand(9865))
 *   - the unmodified reference objects in oracle/_ref (random frames, bit-exact)
 * Exception -- PARITY UNPINNED: orc_decode_window() (windowed traceback for continuous streams) has no
 * runnable counterpart in the reference; it defines the semantics of ced_decode_window_batch and says
 * so in its own comment.
 *
 * Each function cites the reference lines it restates.  The reference fixes the
 * code at compile time (src/defaultParams/convCodeParams.h:8-17); here K, n and
 * the generators are run-time fields so one object serves K=3 and K=7.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_MAX_STATES 256
#define ORC_MAX_N 8

typedef struct {
    int K;                 /* constraint length                                  */
    int n;                 /* coded bits per segment                             */
    int S;                 /* K-1 tail segments                                  */
    int N;                 /* 2^(K-1) trellis states                             */
    int symmetric;         /* 1: butterfly file's shortcut, 0: general butterfly */
    uint32_t taps[ORC_MAX_N];            /* generators, bit 0 = newest input bit */
    uint8_t edgeSymm[ORC_MAX_STATES / 2];
    uint8_t edge[2][ORC_MAX_STATES];
    uint8_t metric[ORC_MAX_STATES];
    uint32_t iteration;
    uint32_t renormCounter;
    uint32_t capacity;     /* survivor rows allocated */
    uint8_t *surv;         /* [capacity][N] one decision byte per state and step */
} orc_decoder_t;

/* src/convEncode.c:163-175 -- reverse the K generator bits so that the LSb taps
 * the bit shifted in last. */
static uint32_t orc_reverse_generator(uint64_t gen, int K)
{
    uint32_t r = 0;
    for (int i = 0; i < K; i++) {
        r = (r << 1) | (uint32_t)(gen & 1u);
        gen >>= 1;
    }
    return r;
}

/* src/convEncode.c:132-161 -- one output segment: generator i contributes bit i. */
static uint8_t orc_segment(uint32_t reg, const uint32_t *taps, int n)
{
    uint8_t seg = 0;
    for (int i = 0; i < n; i++)
        seg |= (uint8_t)((__builtin_popcount(reg & taps[i]) & 1) << i);
    return seg;
}

/* src/viterbiDecoder.c:260-285 with FORCE_NO_POPCNT_DECODER: only the low
 * `bits` bits of a^b are counted. */
uint8_t orc_hamming(uint8_t a, uint8_t b, int bits)
{
    uint8_t x = a ^ b, d = 0;
    for (int i = 0; i < bits; i++) {
        d += x & 1u;
        x >>= 1;
    }
    return d;
}

void orc_taps(int K, int n, const uint64_t *g, uint32_t *taps_out)
{
    for (int i = 0; i < n; i++)
        taps_out[i] = orc_reverse_generator(g[i], K);
}

/*
 * src/convEncode.c:46-130.  Bytes ascending, MSb first; reg = (reg<<1)|bit; when
 * `last`, S zeros are pushed and the register returns to state 0.  `reg` carries
 * the shift register between calls (convEncoderState_t.tappedDelay).
 */
int orc_encode(int K, int n, const uint64_t *g, uint32_t *reg, const uint8_t *in,
               int bytesIn, uint8_t *segs, int last)
{
    uint32_t taps[ORC_MAX_N];
    const uint32_t keep = (1u << K) - 1u;
    int out = 0;
    orc_taps(K, n, g, taps);
    for (int b = 0; b < bytesIn; b++) {
        for (int bit = 7; bit >= 0; bit--) {
            *reg = ((*reg << 1) | ((in[b] >> bit) & 1u)) & keep;
            segs[out++] = orc_segment(*reg, taps, n);
        }
    }
    if (last) {
        for (int i = 0; i < K - 1; i++) {
            *reg = (*reg << 1) & keep;
            segs[out++] = orc_segment(*reg, taps, n);
        }
        *reg = 0; /* src/convEncode.c:122 */
    }
    return out;
}

/* src/viterbiDecoderButterflyk1.c:46-80 */
void orc_dec_reset(orc_decoder_t *d)
{
    d->metric[0] = 0;
    for (int i = 1; i < d->N; i++)
        d->metric[i] = (uint8_t)(d->N + 1);
    d->iteration = 0;
    d->renormCounter = 0;
}

/*
 * src/viterbiDecoderButterflyk1.c:8-44 (symmetric table: output of the 0-edge
 * leaving state j < N/2) and src/viterbiDecoder.c:32-50 (general table: output
 * of edge b leaving state s).
 */
orc_decoder_t *orc_dec_new(int K, int n, const uint64_t *g, int symmetric, int maxSegments)
{
    orc_decoder_t *d = (orc_decoder_t *)calloc(1, sizeof(*d));
    if (!d)
        return NULL;
    d->K = K;
    d->n = n;
    d->S = K - 1;
    d->N = 1 << (K - 1);
    d->symmetric = symmetric;
    orc_taps(K, n, g, d->taps);
    const uint32_t keep = (1u << K) - 1u;
    for (int j = 0; j < d->N / 2; j++)
        d->edgeSymm[j] = orc_segment(((uint32_t)j << 1) & keep, d->taps, n);
    for (int b = 0; b < 2; b++)
        for (int s = 0; s < d->N; s++)
            d->edge[b][s] = orc_segment((((uint32_t)s << 1) | (uint32_t)b) & keep, d->taps, n);
    d->capacity = (uint32_t)maxSegments;
    d->surv = (uint8_t *)malloc((size_t)maxSegments * (size_t)d->N);
    if (!d->surv) {
        free(d);
        return NULL;
    }
    orc_dec_reset(d);
    return d;
}

void orc_dec_free(orc_decoder_t *d)
{
    if (d) {
        free(d->surv);
        free(d);
    }
}

void orc_dec_metrics(const orc_decoder_t *d, uint8_t *out)
{
    memcpy(out, d->metric, (size_t)d->N);
}

void orc_dec_edge_symm(const orc_decoder_t *d, uint8_t *out)
{
    memcpy(out, d->edgeSymm, (size_t)d->N / 2);
}

void orc_dec_edge(const orc_decoder_t *d, uint8_t *out)
{
    memcpy(out, d->edge[0], (size_t)d->N);
    memcpy(out + d->N, d->edge[1], (size_t)d->N);
}

/* survivors of step t, one byte per state (testing aid for the CUDA survivor store) */
void orc_dec_survivors(const orc_decoder_t *d, uint32_t t, uint8_t *out)
{
    memcpy(out, d->surv + (size_t)t * (size_t)d->N, (size_t)d->N);
}

/*
 * Forward recursion: src/viterbiDecoderButterflyk1.c:85-196.
 *   butterfly j: predecessors j and j+N/2, successors 2j and 2j+1       (:101-149)
 *   ties keep the path from the lower predecessor (strict '>')           (:129-130)
 *   when renormCounter >= 120 the minimum new metric is subtracted       (:159-183)
 *   metrics are uint8_t and wrap on store                                (:109-115)
 * The general (non-symmetric) branch costs follow the generic decoder's ACS,
 * src/viterbiDecoder.c:95-128 with argmin2's '<=' (:425-430) -- identical tie rule.
 * Traceback: :200-260 -- start in state 0, drop S tail steps, emit bit 0 of the
 * survivor state of every remaining step MSb-first.
 */
int orc_dec_step(orc_decoder_t *d, const uint8_t *segs, int segmentsIn, uint8_t *uncoded, int last)
{
    const int N = d->N, H = N / 2, n = d->n;
    uint8_t next[ORC_MAX_STATES];
    for (int i = 0; i < segmentsIn; i++) {
        if (d->iteration >= d->capacity)
            return -1;
        const uint8_t rx = segs[i];
        uint8_t hd[1 << ORC_MAX_N]; /* calcHammingDist(c, rx, n) for every n-bit pattern c */
        for (int c = 0; c < (1 << n); c++)
            hd[c] = orc_hamming((uint8_t)c, rx, n);
        uint8_t *row = d->surv + (size_t)d->iteration * (size_t)N;
        for (int j = 0; j < H; j++) {
            uint8_t a0, a1, b0, b1;
            if (d->symmetric) {
                const uint8_t e = hd[d->edgeSymm[j]];
                const uint8_t ec = (uint8_t)(n - e);
                a0 = (uint8_t)(d->metric[j] + e);
                a1 = (uint8_t)(d->metric[j + H] + ec);
                b0 = (uint8_t)(d->metric[j] + ec);
                b1 = (uint8_t)(d->metric[j + H] + e);
            } else {
                a0 = (uint8_t)(d->metric[j] + hd[d->edge[0][j]]);
                a1 = (uint8_t)(d->metric[j + H] + hd[d->edge[0][j + H]]);
                b0 = (uint8_t)(d->metric[j] + hd[d->edge[1][j]]);
                b1 = (uint8_t)(d->metric[j + H] + hd[d->edge[1][j + H]]);
            }
            const uint8_t da = a0 > a1, db = b0 > b1;
            next[2 * j] = da ? a1 : a0;
            next[2 * j + 1] = db ? b1 : b0;
            row[2 * j] = da;
            row[2 * j + 1] = db;
        }
        if (d->renormCounter >= 120) {
            uint8_t lo = next[0];
            for (int s = 1; s < N; s++)
                if (next[s] < lo)
                    lo = next[s];
            for (int s = 0; s < N; s++)
                next[s] = (uint8_t)(next[s] - lo);
            d->renormCounter = 0;
        } else {
            d->renormCounter++;
        }
        memcpy(d->metric, next, (size_t)N);
        d->iteration++;
    }
    if (!last)
        return 0;

    const uint32_t T = d->iteration, S = (uint32_t)d->S;
    uint32_t state = 0;
    for (uint32_t i = 0; i < S; i++) {
        const uint32_t t = T - 1 - i;
        const uint32_t dec = d->surv[(size_t)t * (size_t)N + state];
        state = (state >> 1) | (dec << (S - 1));
    }
    uncoded[(T - S - 1) / 8] = 0;
    for (uint32_t i = S; i < T; i++) {
        const uint32_t t = T - 1 - i;
        const uint32_t dec = d->surv[(size_t)t * (size_t)N + state];
        uncoded[t / 8] = (uint8_t)((uncoded[t / 8] >> 1) | ((state & 1u) << 7));
        state = (state >> 1) | (dec << (S - 1));
    }
    const int bytesOut = (int)((T - S - 1) / 8 + 1);
    orc_dec_reset(d);
    return bytesOut;
}

/*
 * Windowed traceback over a long (continuous) terminated stream, SURVEY 8(f)4.  PARITY UNPINNED against the
 * reference: its only sliding-window decoder (TRACEBACK_LEN = 5K, src/viterbiDecoder.h:19, generic
 * viterbiDecoderHard src/viterbiDecoder.c:32-258) aborts at HEAD and no test reaches it; the behavioural
 * model is MATLAB's vitdec(..., tblen, ..., 'hard') (scripts/matlab/viterbiBEREstimate.m:17,99).  This
 * function DEFINES the semantics the CUDA window path is tested against:
 *   - the forward recursion is the reference's (orc_dec_step, bit-exact decisions);
 *   - the stream arrives in calls of `callSegs` segments (the last call takes the rest);
 *   - after a call that ends at step P (not the last) the decoder starts in the state with the smallest
 *     metric (lowest state index on ties), walks the survivors back and emits the bits of the steps
 *     below P - depth that were not emitted before -- every bit is decided by a traceback of at
 *     least `depth` steps;
 *   - the last call starts in state 0 (terminated stream, :205-223), drops the S tail steps and emits
 *     everything left.
 * Output bit t is bit 0 of the survivor state after step t, MSb-first (:244-249).  Returns the number
 * of bits written (total - S) or -1.
 */
int orc_decode_window2(int K, int n, const uint64_t *g, int symmetric, const uint8_t *segs, int totalSegs, int callSegs,
                       int depth, uint8_t *out);

int orc_decode_window(int K, int n, const uint64_t *g, const uint8_t *segs, int totalSegs, int callSegs,
                      int depth, uint8_t *out)
{
    return orc_decode_window2(K, n, g, 1, segs, totalSegs, callSegs, depth, out);
}

/* the same with the general (non-symmetric) branch costs of orc_dec_step for any code */
int orc_decode_window2(int K, int n, const uint64_t *g, int symmetric, const uint8_t *segs, int totalSegs, int callSegs,
                       int depth, uint8_t *out)
{
    if (callSegs <= 0 || depth < 0 || totalSegs <= K - 1)
        return -1;
    orc_decoder_t *d = orc_dec_new(K, n, g, symmetric, totalSegs);
    if (!d)
        return -1;
    const int N = d->N, S = d->S;
    memset(out, 0, (size_t)(totalSegs - S + 7) / 8);
    int emitted = 0; /* bits of steps [0, emitted) are final */
    for (int P0 = 0; P0 < totalSegs;) {
        const int P = (totalSegs - P0 <= callSegs) ? totalSegs : P0 + callSegs;
        const int lastCall = P == totalSegs;
        orc_dec_step(d, segs + P0, P - P0, NULL, 0);
        uint32_t state = 0;
        if (!lastCall)
            for (int s2 = 1; s2 < N; s2++)
                if (d->metric[s2] < d->metric[state])
                    state = (uint32_t)s2;
        const int hi = lastCall ? totalSegs - S : P - depth; /* emit steps [emitted, hi) */
        for (int t = P - 1; t >= emitted; t--) {
            const uint32_t dec = d->surv[(size_t)t * (size_t)N + state];
            if (t < hi && (state & 1u))
                out[t / 8] |= (uint8_t)(0x80u >> (t % 8));
            state = (state >> 1) | (dec << (S - 1));
        }
        if (hi > emitted)
            emitted = hi;
        P0 = P;
    }
    orc_dec_free(d);
    return emitted;
}

/* whole frames, one-shot (speedDecode.c:79 call shape) */
int orc_decode_batch(int K, int n, const uint64_t *g, int symmetric, const uint8_t *segs,
                     size_t segStride, int nFrames, int segsPerFrame, uint8_t *out, size_t outStride)
{
    orc_decoder_t *d = orc_dec_new(K, n, g, symmetric, segsPerFrame);
    if (!d)
        return -1;
    for (int f = 0; f < nFrames; f++)
        orc_dec_step(d, segs + (size_t)f * segStride, segsPerFrame, out + (size_t)f * outStride, 1);
    orc_dec_free(d);
    return 0;
}

int orc_encode_batch(int K, int n, const uint64_t *g, const uint8_t *in, size_t inStride, int nFrames,
                     int bytesPerFrame, uint8_t *segs, size_t segStride)
{
    for (int f = 0; f < nFrames; f++) {
        uint32_t reg = 0;
        orc_encode(K, n, g, &reg, in + (size_t)f * inStride, bytesPerFrame, segs + (size_t)f * segStride, 1);
    }
    return 0;
}

/*
 * Soft-decision decoding (north_star kernel (1) "soft or hard symbols").  The reference is a
 * hard-decision decoder, so this restatement EXTENDS it; it is pinned to the reference by reduction:
 *   - trellis, butterfly order, tie rule (strict '>' keeps the lower predecessor, :129-130), traceback from
 *     state 0 and MSb-first packing are exactly orc_dec_step's (src/viterbiDecoderButterflyk1.c:101-149,200-260);
 *   - a received segment is n int8 values s_0..s_{n-1} (s_i for generator i; BPSK bit 0 -> +, bit 1 -> -).
 *     The cost of an edge labelled c is  sum_i |s_i| * [hard(s_i) != c_i]  with hard(s) = (s < 0), i.e. the
 *     reference's calcHammingDist(c, rx, n) (src/viterbiDecoder.c:260-285) with every disagreeing bit
 *     weighted by its reliability.  When all |s_i| equal one constant A this is A * calcHammingDist(c,
 *     hard(rx), n): every comparison (ties included) is the hard decoder's, so the output must equal
 *     orc_dec_step's on the sliced symbols bit for bit -- tests/test_oracle.py checks that, and the CUDA
 *     path is held to both this function and that reduction;
 *   - metrics are 32-bit ints and never renormalised (16390 steps * n * 128 < 2^31); decisions depend on
 *     metric differences only, so any non-overflowing renormalisation schedule is equivalent.
 * Start metrics: 0 for state 0, "never wins" for the rest (the reference's NUM_STATES+1 plays that role for
 * costs <= n; here n * 128 * K).
 */
typedef struct {
    int K, n, S, N;
    uint8_t edge[2][ORC_MAX_STATES];
    int32_t metric[ORC_MAX_STATES];
    uint32_t iteration, capacity;
    uint8_t *surv;
} orc_soft_decoder_t;

orc_soft_decoder_t *orc_soft_new(int K, int n, const uint64_t *g, int maxSegments)
{
    orc_soft_decoder_t *d = (orc_soft_decoder_t *)calloc(1, sizeof(*d));
    uint32_t taps[ORC_MAX_N];
    if (!d)
        return NULL;
    d->K = K;
    d->n = n;
    d->S = K - 1;
    d->N = 1 << (K - 1);
    orc_taps(K, n, g, taps);
    const uint32_t keep = (1u << K) - 1u;
    for (int b = 0; b < 2; b++)
        for (int s = 0; s < d->N; s++)
            d->edge[b][s] = orc_segment((((uint32_t)s << 1) | (uint32_t)b) & keep, taps, n);
    d->capacity = (uint32_t)maxSegments;
    d->surv = (uint8_t *)malloc((size_t)maxSegments * (size_t)d->N);
    if (!d->surv) {
        free(d);
        return NULL;
    }
    d->metric[0] = 0;
    for (int i = 1; i < d->N; i++)
        d->metric[i] = n * 128 * K;
    return d;
}

void orc_soft_free(orc_soft_decoder_t *d)
{
    if (d) {
        free(d->surv);
        free(d);
    }
}

/* soft: n int8 per segment, segment after segment */
int orc_dec_step_soft(orc_soft_decoder_t *d, const int8_t *soft, int segmentsIn, uint8_t *uncoded, int last)
{
    const int N = d->N, H = N / 2, n = d->n;
    int32_t next[ORC_MAX_STATES];
    for (int i = 0; i < segmentsIn; i++) {
        if (d->iteration >= d->capacity)
            return -1;
        int32_t cost[1 << ORC_MAX_N];
        for (int c = 0; c < (1 << n); c++) {
            cost[c] = 0;
            for (int b = 0; b < n; b++) {
                const int s = soft[(size_t)i * (size_t)n + (size_t)b];
                const int hard = s < 0, mag = s < 0 ? -s : s;
                if (hard != ((c >> b) & 1))
                    cost[c] += mag;
            }
        }
        uint8_t *row = d->surv + (size_t)d->iteration * (size_t)N;
        for (int j = 0; j < H; j++) {
            const int32_t a0 = d->metric[j] + cost[d->edge[0][j]];
            const int32_t a1 = d->metric[j + H] + cost[d->edge[0][j + H]];
            const int32_t b0 = d->metric[j] + cost[d->edge[1][j]];
            const int32_t b1 = d->metric[j + H] + cost[d->edge[1][j + H]];
            const uint8_t da = a0 > a1, db = b0 > b1; /* ties keep the lower predecessor (:129-130) */
            next[2 * j] = da ? a1 : a0;
            next[2 * j + 1] = db ? b1 : b0;
            row[2 * j] = da;
            row[2 * j + 1] = db;
        }
        memcpy(d->metric, next, (size_t)N * sizeof(int32_t));
        d->iteration++;
    }
    if (!last)
        return 0;
    const uint32_t T = d->iteration, S = (uint32_t)d->S;
    uint32_t state = 0;
    memset(uncoded, 0, (size_t)((T - S - 1) / 8 + 1));
    for (uint32_t i = 0; i < T; i++) {
        const uint32_t t = T - 1 - i;
        const uint32_t dec = d->surv[(size_t)t * (size_t)N + state];
        if (i >= S && (state & 1u))
            uncoded[t / 8] |= (uint8_t)(0x80u >> (t % 8));
        state = (state >> 1) | (dec << (S - 1));
    }
    return (int)((T - S - 1) / 8 + 1);
}

/*
 * orc_decode_window with soft inputs: the windowing procedure of orc_decode_window (same PARITY status: it defines the
 * semantics, anchored by depth >= length == full traceback) around the soft forward recursion orc_dec_step_soft.
 * soft: n int8 per segment.  Returns the number of bits written (total - S) or -1.
 */
int orc_decode_window_soft(int K, int n, const uint64_t *g, const int8_t *soft, int totalSegs, int callSegs, int depth,
                           uint8_t *out)
{
    if (callSegs <= 0 || depth < 0 || totalSegs <= K - 1)
        return -1;
    orc_soft_decoder_t *d = orc_soft_new(K, n, g, totalSegs);
    if (!d)
        return -1;
    const int N = d->N, S = d->S;
    memset(out, 0, (size_t)(totalSegs - S + 7) / 8);
    int emitted = 0;
    for (int P0 = 0; P0 < totalSegs;) {
        const int P = (totalSegs - P0 <= callSegs) ? totalSegs : P0 + callSegs;
        const int lastCall = P == totalSegs;
        orc_dec_step_soft(d, soft + (size_t)P0 * (size_t)n, P - P0, NULL, 0);
        uint32_t state = 0;
        if (!lastCall)
            for (int s2 = 1; s2 < N; s2++)
                if (d->metric[s2] < d->metric[state])
                    state = (uint32_t)s2;
        const int hi = lastCall ? totalSegs - S : P - depth;
        for (int t = P - 1; t >= emitted; t--) {
            const uint32_t dec = d->surv[(size_t)t * (size_t)N + state];
            if (t < hi && (state & 1u))
                out[t / 8] |= (uint8_t)(0x80u >> (t % 8));
            state = (state >> 1) | (dec << (S - 1));
        }
        if (hi > emitted)
            emitted = hi;
        P0 = P;
    }
    orc_soft_free(d);
    return emitted;
}

int orc_decode_soft_batch(int K, int n, const uint64_t *g, const int8_t *soft, size_t softStride, int nFrames,
                          int segsPerFrame, uint8_t *out, size_t outStride)
{
    for (int f = 0; f < nFrames; f++) {
        orc_soft_decoder_t *d = orc_soft_new(K, n, g, segsPerFrame);
        if (!d)
            return -1;
        orc_dec_step_soft(d, soft + (size_t)f * softStride, segsPerFrame, out + (size_t)f * outStride, 1);
        orc_soft_free(d);
    }
    return 0;
}

/*
 * berTestK7/berTestK7.c:22-53,109-165 restated: glibc rand() stream, BSC with
 * `flip = frand() > p ? 0 : 1`, MSb of the segment drawn first.  counts =
 * {channel flips, coded bits, decoded bit errors, decoded bits}.  `pktsOut`, if
 * non-NULL, receives the corrupted segments of every packet ([pkts][segs]) and
 * `msgOut` the messages ([pkts][pktBytes]) so the CUDA path can be fed the very
 * same data.  The caller seeds with srand() once (berTestK7.c:66) -- use
 * orc_srand() -- because the three configurations share one rand() stream.
 */
void orc_srand(unsigned seed)
{
    srand(seed);
}

int orc_bertest(int K, int n, const uint64_t *g, int pkts, int pktBytes, double p, int64_t *counts,
                uint8_t *pktsOut, uint8_t *msgOut)
{
    const int segsPerPkt = 8 * pktBytes + (K - 1);
    orc_decoder_t *d = orc_dec_new(K, n, g, 1, segsPerPkt);
    uint8_t *msg = (uint8_t *)malloc((size_t)pktBytes);
    uint8_t *dec = (uint8_t *)malloc((size_t)pktBytes);
    uint8_t *clean = (uint8_t *)malloc((size_t)segsPerPkt);
    uint8_t *noisy = (uint8_t *)malloc((size_t)segsPerPkt);
    if (!d || !msg || !dec || !clean || !noisy)
        return -1;
    memset(counts, 0, 4 * sizeof(int64_t));
    for (int it = 0; it < pkts; it++) {
        for (int j = 0; j < pktBytes; j++)
            msg[j] = (uint8_t)rand();
        uint32_t reg = 0;
        const int ns = orc_encode(K, n, g, &reg, msg, pktBytes, clean, 1);
        counts[1] += (int64_t)ns * n;
        for (int i = 0; i < ns; i++) {
            uint8_t flips = 0;
            for (int j = 0; j < n; j++) {
                const double u = (double)rand() / RAND_MAX;
                const uint8_t f = u > p ? 0 : 1;
                flips = (uint8_t)((flips << 1) | f);
                counts[0] += f;
            }
            noisy[i] = clean[i] ^ flips;
        }
        const int nb = orc_dec_step(d, noisy, ns, dec, 1);
        counts[3] += (int64_t)nb * 8;
        for (int j = 0; j < pktBytes; j++)
            counts[2] += orc_hamming(msg[j], dec[j], 8);
        if (pktsOut)
            memcpy(pktsOut + (size_t)it * (size_t)segsPerPkt, noisy, (size_t)segsPerPkt);
        if (msgOut)
            memcpy(msgOut + (size_t)it * (size_t)pktBytes, msg, (size_t)pktBytes);
    }
    free(msg);
    free(dec);
    free(clean);
    free(noisy);
    orc_dec_free(d);
    return 0;
}

/*
 * Statistical anchor of orc_decode_window to numbers the reference holds.  berTestK7/berTestK7.c:98 keeps MATLAB's
 * expectations for vitdec(..., tblen = 5*K = 35, 'term', 'hard') (scripts/matlab/viterbiBEREstimate.m:17,99;
 * generators 133/171): decoded BER 5.295410e-03 / 5.421997e-04 / 3.385010e-05 at channel BER 5.585640e-02 /
 * 3.716174e-02 / 2.262231e-02.  orc_decode_window with callSegs = 1 IS that decoder: after every step it starts
 * in the best state, walks back `depth` steps and emits the one bit that became final, and the terminated end is
 * flushed from state 0.  This loop measures its BER on a BSC (berTestK7.c:29-43 channel, own xorshift stream per
 * packet so threads are independent) so a test can apply the reference's own +-10 % rule (berTestK7.c:167-172).
 * counts = {channel flips, coded bits, decoded bit errors, decoded bits}.
 */
typedef struct {
    int K, n, pktBytes, callSegs, depth, first, count;
    const uint64_t *g;
    double p;
    uint64_t seed;
    int64_t counts[4];
} orc_wber_arg_t;

static uint64_t orc_xorshift(uint64_t *s)
{
    uint64_t x = *s;
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return *s = x;
}

static void *orc_wber_thread(void *vp)
{
    orc_wber_arg_t *a = (orc_wber_arg_t *)vp;
    const int T = 8 * a->pktBytes + a->K - 1;
    uint8_t *msg = (uint8_t *)malloc((size_t)a->pktBytes), *dec = (uint8_t *)malloc((size_t)a->pktBytes + 1);
    uint8_t *segs = (uint8_t *)malloc((size_t)T);
    const uint64_t thr = (uint64_t)(a->p * 18446744073709551616.0);
    for (int it = a->first; it < a->first + a->count; it++) {
        uint64_t s = (a->seed + 1) * 0x9E3779B97F4A7C15ull + (uint64_t)it * 0xD1B54A32D192ED03ull;
        for (int w = 0; w < 4; w++)
            orc_xorshift(&s);
        for (int j = 0; j < a->pktBytes; j++)
            msg[j] = (uint8_t)(orc_xorshift(&s) >> 32);
        uint32_t reg = 0;
        orc_encode(a->K, a->n, a->g, &reg, msg, a->pktBytes, segs, 1);
        for (int i = 0; i < T; i++)
            for (int b = 0; b < a->n; b++)
                if (orc_xorshift(&s) < thr) {
                    segs[i] ^= (uint8_t)(1u << b);
                    a->counts[0]++;
                }
        a->counts[1] += (int64_t)T * a->n;
        orc_decode_window(a->K, a->n, a->g, segs, T, a->callSegs, a->depth, dec);
        for (int j = 0; j < a->pktBytes; j++)
            a->counts[2] += __builtin_popcount((unsigned)(msg[j] ^ dec[j]));
        a->counts[3] += 8 * (int64_t)a->pktBytes;
    }
    free(msg);
    free(dec);
    free(segs);
    return NULL;
}

int orc_window_ber(int K, int n, const uint64_t *g, int pkts, int pktBytes, double p, int callSegs, int depth,
                   uint64_t seed, int nThreads, int64_t *counts)
{
    if (nThreads < 1)
        nThreads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nThreads);
    orc_wber_arg_t *args = (orc_wber_arg_t *)calloc((size_t)nThreads, sizeof(orc_wber_arg_t));
    for (int i = 0; i < nThreads; i++) {
        const int lo = (int)((int64_t)pkts * i / nThreads), hi = (int)((int64_t)pkts * (i + 1) / nThreads);
        args[i] = (orc_wber_arg_t){K, n, pktBytes, callSegs, depth, lo, hi - lo, g, p, seed, {0, 0, 0, 0}};
        pthread_create(&th[i], NULL, orc_wber_thread, &args[i]);
    }
    memset(counts, 0, 4 * sizeof(int64_t));
    for (int i = 0; i < nThreads; i++) {
        pthread_join(th[i], NULL);
        for (int c = 0; c < 4; c++)
            counts[c] += args[i].counts[c];
    }
    free(th);
    free(args);
    return 0;
}

/*
 * CPU-baseline loop for bench.py ("port" kind): speedDecode.c:72-110 semantics
 * (frames pre-encoded, decode only, one-shot last=true calls, CLOCK_MONOTONIC),
 * one thread per requested core, each with a private decoder.  Returns decoded
 * information bits summed over threads; *seconds receives the wall time.
 */

typedef struct {
    int K, n, segsPerFrame, nFrames, first;
    const uint64_t *g;
    const uint8_t *segs;
    size_t stride;
    double budget;
    int64_t bits;
    uint8_t sink;
} orc_speed_arg_t;

static double orc_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *orc_speed_thread(void *p)
{
    orc_speed_arg_t *a = (orc_speed_arg_t *)p;
    orc_decoder_t *d = orc_dec_new(a->K, a->n, a->g, 1, a->segsPerFrame);
    const int frameBytes = (a->segsPerFrame - (a->K - 1)) / 8;
    uint8_t *out = (uint8_t *)malloc((size_t)frameBytes + 1);
    int f = a->first % a->nFrames;
    const double t0 = orc_now();
    a->bits = 0;
    do {
        for (int rep = 0; rep < 8; rep++) {
            orc_dec_step(d, a->segs + (size_t)f * a->stride, a->segsPerFrame, out, 1);
            a->sink ^= out[0];
            a->bits += (int64_t)frameBytes * 8;
            f = (f + 1 == a->nFrames) ? 0 : f + 1;
        }
    } while (orc_now() - t0 < a->budget);
    free(out);
    orc_dec_free(d);
    return NULL;
}

int64_t orc_speed_decode(int K, int n, const uint64_t *g, const uint8_t *segs, size_t stride, int nFrames,
                         int segsPerFrame, int nThreads, double budgetSeconds, double *seconds)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nThreads);
    orc_speed_arg_t *args = (orc_speed_arg_t *)calloc((size_t)nThreads, sizeof(orc_speed_arg_t));
    const double t0 = orc_now();
    for (int i = 0; i < nThreads; i++) {
        args[i] = (orc_speed_arg_t){K, n, segsPerFrame, nFrames, i * 7, g, segs, stride, budgetSeconds, 0, 0};
        pthread_create(&th[i], NULL, orc_speed_thread, &args[i]);
    }
    int64_t bits = 0;
    for (int i = 0; i < nThreads; i++) {
        pthread_join(th[i], NULL);
        bits += args[i].bits;
    }
    *seconds = orc_now() - t0;
    free(th);
    free(args);
    return bits;
}
