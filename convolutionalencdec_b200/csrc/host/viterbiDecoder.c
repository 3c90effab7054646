/*
 * Host side of include/viterbiDecoder.h and viterbiDecoderButterflyk1.h: state
 * bookkeeping, trellis-label tables and the exit(1) error convention of the
 * reference (src/viterbiDecoder.c:10-13).  The add-compare-select recursion and
 * the traceback are NOT here: they run in streamDecodeKernel via
 * ced_stream_decode (include/ced_abi.h).
 */
#include "viterbiDecoder.h"
#include "ced_abi.h"
#include "convEncode.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int viterbiConfigCheck()
{
    if (STARTING_STATE != 0) {
        printf("Currently only support starting state of 0\n");
        exit(1);
    }
    /* The reference additionally rejects generators that do not tap both the
     * newest and the oldest bit (src/viterbiDecoder.c:15-27) because its butterfly
     * relies on that symmetry -- which makes its own handTracedTest abort.  The
     * GPU streaming kernel evaluates all four branch metrics of a butterfly, so
     * every k=1 code is accepted here. */
    return 0;
}

uint8_t calcHammingDist(uint8_t a, uint8_t b, int bits)
{
    /* reference: src/viterbiDecoder.c:260-285 -- only the low `bits` bits count */
    const unsigned diff = (unsigned)(a ^ b) & ((bits >= 8) ? 0xFFu : ((1u << bits) - 1u));
    return (uint8_t)__builtin_popcount(diff);
}

static void labelEdges(viterbiHardState_t *state)
{
    convEncoderState_t enc;
    resetConvEncoder(&enc);
    initConvEncoder(&enc);
    for (unsigned long s = 0; s < NUM_STATES; s++)
        for (int b = 0; b < 2; b++) {
            enc.tappedDelay = (TAPPED_DELAY_TYPE)s;
            state->edgeCodedBits[b][s] = (EDGE_METRIC_INDEX_TYPE)convEncOneInput(&enc, (uint8_t)b);
        }
    for (unsigned long j = 0; j < NUM_STATES / 2; j++)
        state->edgeCodedBitsSymm[j] = state->edgeCodedBits[0][j];
}

void viterbiInitButterflyk1(viterbiHardState_t *state)
{
    printf("Specialized Viterbi Decoder for k=1\n"); /* reference: src/viterbiDecoderButterflyk1.c:16 */
    labelEdges(state);
}

void viterbiInit(viterbiHardState_t *state)
{
    labelEdges(state);
}

void resetViterbiDecoderHardButterflyk1(viterbiHardState_t *state)
{
    state->nodeMetricsCur = &state->nodeMetricsA;
    state->nodeMetricsNext = &state->nodeMetricsB;
    state->traceBackCur = &state->traceBackA;
    state->traceBackNext = &state->traceBackB;
    state->nodeMetricsA[0] = 0;
    for (unsigned long s = 1; s < NUM_STATES; s++)
        state->nodeMetricsA[s] = (METRIC_TYPE)(NUM_STATES + 1);
    state->iteration = 0;
    state->renormCounter = 0;
    state->decodeCarryOver = 0;
    state->decodeCarryOverCount = 0;
}

void resetViterbiDecoderHard(viterbiHardState_t *state)
{
    resetViterbiDecoderHardButterflyk1(state);
}

void swapViterbiArrays(viterbiHardState_t *state)
{
    METRIC_TYPE(*m)[NUM_STATES] = state->nodeMetricsCur;
    state->nodeMetricsCur = state->nodeMetricsNext;
    state->nodeMetricsNext = m;
    TRACEBACK_TYPE(*t)[NUM_STATES] = state->traceBackCur;
    state->traceBackCur = state->traceBackNext;
    state->traceBackNext = t;
}

int viterbiDecoderHardButterflyk1(viterbiHardState_t *restrict state, uint8_t *restrict codedSegments,
                                  uint8_t *restrict uncoded, int segmentsIn, bool last)
{
    const int bytes = ced_stream_decode(K, n, &state->edgeCodedBits[0][0], state->nodeMetricsA, &state->iteration,
                                        &state->renormCounter, &state->survivorWords[0][0], MAX_PKT_LEN_SEGMENTS,
                                        codedSegments, segmentsIn, uncoded, last ? 1 : 0);
    if (bytes < 0) {
        printf("viterbiDecoderHardButterflyk1: GPU decoder failed: %s\n", ced_last_error());
        exit(1);
    }
    if (last)
        resetViterbiDecoderHardButterflyk1(state); /* reference: src/viterbiDecoderButterflyk1.c:259 */
    return bytes;
}

int viterbiDecoderHard(viterbiHardState_t *restrict state, uint8_t *restrict codedSegments,
                       uint8_t *restrict uncoded, int segmentsIn, bool last)
{
    return viterbiDecoderHardButterflyk1(state, codedSegments, uncoded, segmentsIn, last);
}

/* ---- small reductions kept as exported helpers (src/viterbiDecoder.h:170-186) ---- */
static int argminOver(const METRIC_TYPE *m, int count)
{
    int best = 0;
    for (int i = 1; i < count; i++)
        if (m[i] < m[best])
            best = i;
    return best;
}

int argmin2(const METRIC_TYPE (*metrics)[2]) { return argminOver(*metrics, 2); }
int argmin4(const METRIC_TYPE (*metrics)[4]) { return argminOver(*metrics, 4); }
int argmin8(const METRIC_TYPE (*metrics)[8]) { return argminOver(*metrics, 8); }
int argmin16(const METRIC_TYPE (*metrics)[16]) { return argminOver(*metrics, 16); }
int argmin32(const METRIC_TYPE (*metrics)[32]) { return argminOver(*metrics, 32); }
int argmin64(const METRIC_TYPE (*metrics)[64]) { return argminOver(*metrics, 64); }
int argminPathMetrics(const METRIC_TYPE (*metrics)[POW2(k)]) { return argminOver(*metrics, (int)POW2(k)); }
int argminNodeMetrics(const METRIC_TYPE (*metrics)[NUM_STATES]) { return argminOver(*metrics, (int)NUM_STATES); }

METRIC_TYPE minMetricGeneric(const METRIC_TYPE (*metrics)[NUM_STATES])
{
    return (*metrics)[argminOver(*metrics, (int)NUM_STATES)];
}
