/*
 * viterbiDecoder.h -- hard-decision Viterbi decoder, drop-in for the reference's
 * src/viterbiDecoder.h: the VITERBI_DECODER_HARD / VITERBI_INIT / VITERBI_RESET
 * dispatch macros (:87-95), the option macros drivers print, the caller-owned
 * viterbiHardState_t with the fields drivers touch (nodeMetricsCur,
 * handTracedTest/handTraced.c:72-111), and the 27 exported symbols of SURVEY 8(b).
 *
 * The forward recursion and traceback run on the GPU (include/ced_abi.h).  The
 * struct differs from the reference's in one deliberate way: survivor decisions
 * are kept bit-packed (CED_SURV_WORDS uint32 per step, 128 KiB for K=7) instead
 * of one byte per state (1 MiB), because the device produces them ballot-packed.
 */
#ifndef CED_VITERBI_DECODER_H
#define CED_VITERBI_DECODER_H

#include "convCodeParams.h"
#include "convHelpers.h"
#include <stdbool.h>

/* ---- options (src/viterbiDecoder.h:17-39) ---- */
#define MAX_PKT_LEN_UNCODED_BITS (1024 * 16)
#define TRACEBACK_LEN (5 * K)
#define TRACEBACK_BUFFER_LEN (MAX_PKT_LEN_UNCODED_BITS)
#define NUM_STATES (POW2(k * S))
#define MAX_EDGE_WEIGHT (n)
#define MAX_PKT_LEN_SEGMENTS (MAX_PKT_LEN_UNCODED_BITS + S)
/* Kept for source compatibility.  This library never needs the symmetric
 * shortcut to be legal: the streaming kernel evaluates all four branches of a
 * butterfly, the batch kernel checks symmetry at compile time. */
#define USE_POLY_SYMMETRY
#define FORCE_NO_POPCNT_DECODER

/* ---- metric / table element types (src/viterbiDecoder.h:47-76) ---- */
#if k == 1
#define METRIC_TYPE uint8_t /* renormalised every 121 steps, never exceeds 135 */
#define METRIC_MAX UINT8_MAX
#else
#error this library accelerates k == 1 codes only (the reference routes k > 1 to its generic decoder)
#endif
#define EDGE_METRIC_INDEX_TYPE uint8_t
#define TRACEBACK_TYPE uint8_t
#define TRACEBACK_BITS 8
#define TRACEBACK_BYTES ((TRACEBACK_BUFFER_LEN + S * k) / TRACEBACK_BITS + 1)

#define VITERBI_DECODER_HARD viterbiDecoderHardButterflyk1
#define VITERBI_INIT viterbiInitButterflyk1
#define VITERBI_RESET resetViterbiDecoderHardButterflyk1

/* uint32 words of ballot-packed decisions per trellis step: for each group of 32
 * butterflies one word for the even successors and one for the odd ones */
#define CED_SURV_WORDS (2 * ((NUM_STATES / 2 + 31) / 32))

/* Caller-owned decoder state (stack-allocated by every reference driver). */
typedef struct {
    /* trellis labels: coded segment on the 0-edge of butterfly j, and on every edge */
    EDGE_METRIC_INDEX_TYPE edgeCodedBitsSymm[NUM_STATES / 2];
    EDGE_METRIC_INDEX_TYPE edgeCodedBits[POW2(k)][NUM_STATES];

    METRIC_TYPE nodeMetricsA[NUM_STATES] __attribute__((aligned(64)));
    TRACEBACK_TYPE traceBackA[NUM_STATES] __attribute__((aligned(64)));
    METRIC_TYPE nodeMetricsB[NUM_STATES];
    TRACEBACK_TYPE traceBackB[NUM_STATES];
    METRIC_TYPE (*restrict nodeMetricsCur)[NUM_STATES];   /* always &nodeMetricsA */
    TRACEBACK_TYPE (*restrict traceBackCur)[NUM_STATES];
    METRIC_TYPE (*restrict nodeMetricsNext)[NUM_STATES];
    TRACEBACK_TYPE (*restrict traceBackNext)[NUM_STATES];

    unsigned int iteration;      /* trellis steps taken in the current packet        */
    unsigned int renormCounter;  /* steps since the last metric renormalisation      */
    uint8_t decodeCarryOver;
    uint8_t decodeCarryOverCount;

    /* host mirror of the survivor decisions of the current packet, so a packet
     * can be fed in chunks while the device keeps no state between calls */
    uint32_t survivorWords[MAX_PKT_LEN_SEGMENTS][CED_SURV_WORDS] __attribute__((aligned(64)));
} viterbiHardState_t;

/* Generic k>=1 decoder entry points of the reference (src/viterbiDecoder.h:160-186).
 * Exported for link compatibility; for k == 1 they forward to the butterfly
 * implementation (no driver reaches them, the reference's own version is broken
 * at HEAD -- SURVEY 0.3). */
int viterbiDecoderHard(viterbiHardState_t *restrict state, uint8_t *restrict codedSegments,
                       uint8_t *restrict uncoded, int segmentsIn, bool last);
void swapViterbiArrays(viterbiHardState_t *state);
/* exit(1) with a message unless STARTING_STATE == 0 (src/viterbiDecoder.c:9-14). */
int viterbiConfigCheck();
void viterbiInit(viterbiHardState_t *state);
void resetViterbiDecoderHard(viterbiHardState_t *state);

/* number of differing bits among the low `bits` bits of a and b */
uint8_t calcHammingDist(uint8_t a, uint8_t b, int bits);

/* index of the smallest entry, first one on ties */
int argminPathMetrics(const METRIC_TYPE (*metrics)[POW2(k)]);
int argminNodeMetrics(const METRIC_TYPE (*metrics)[NUM_STATES]);
int argmin2(const METRIC_TYPE (*metrics)[2]);
int argmin4(const METRIC_TYPE (*metrics)[4]);
int argmin8(const METRIC_TYPE (*metrics)[8]);
int argmin16(const METRIC_TYPE (*metrics)[16]);
int argmin32(const METRIC_TYPE (*metrics)[32]);
int argmin64(const METRIC_TYPE (*metrics)[64]);

#include "viterbiDecoderButterflyk1.h"

#endif
