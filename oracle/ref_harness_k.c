/*
 * ref_harness_k.c -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile, target refk).
 *
 * The UNMODIFIED reference sources compiled with k = 2 parameter headers (oracle/params/<name>/, written in the
 * reference's own parameter format).  With k != 1 the reference's macros route to its generic decoder
 * viterbiDecoderHard (src/viterbiDecoder.h:87-95), whose register-exchange traceback does not work at HEAD
 * (TRACEBACK_TYPE is uint8_t, src/viterbiDecoder.h:75-76, shifted by (5K-1)k bits, src/viterbiDecoder.c:148) --
 * but its encoder (src/convEncode.c:46-130), its trellis labels (viterbiInit, src/viterbiDecoder.c:32-50) and its
 * add-compare-select (src/viterbiDecoder.c:95-128: the path metrics after every step) do, and those are what this
 * file exposes for pinning the k > 1 restatement.  It only calls the reference's public API.
 */
#include "convEncode.h"
#include "viterbiDecoder.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int refk_K(void) { return K; }
int refk_k(void) { return k; }
int refk_n(void) { return n; }
int refk_states(void) { return (int)NUM_STATES; }
uint64_t refk_g(int i) { return g[i]; }

int refk_encode(const uint8_t *in, int bytesIn, uint8_t *segs)
{
    convEncoderState_t e;
    resetConvEncoder(&e);
    initConvEncoder(&e);
    return convEnc(&e, (uint8_t *)in, segs, bytesIn, true);
}

/* edgeCodedBits[edgeInd][stateInd] after viterbiInit, as uint8 [2^k][NUM_STATES] */
void refk_edges(uint8_t *out)
{
    viterbiHardState_t *st = (viterbiHardState_t *)aligned_alloc(64, (sizeof(viterbiHardState_t) + 63) & ~(size_t)63);
    VITERBI_RESET(st);
    VITERBI_INIT(st);
    for (int e = 0; e < (int)POW2(k); e++)
        for (int s = 0; s < (int)NUM_STATES; s++)
            out[e * (int)NUM_STATES + s] = (uint8_t)st->edgeCodedBits[e][s];
    free(st);
}

/* path metrics after every trellis step: metrics[step][state] (uint32), one viterbiDecoderHard call per segment */
void refk_metrics(const uint8_t *segs, int nSegs, uint32_t *metrics)
{
    viterbiHardState_t *st = (viterbiHardState_t *)aligned_alloc(64, (sizeof(viterbiHardState_t) + 63) & ~(size_t)63);
    uint8_t *sink = (uint8_t *)calloc((size_t)nSegs + 64, 1);
    VITERBI_RESET(st);
    VITERBI_INIT(st);
    for (int i = 0; i < nSegs; i++) {
        VITERBI_DECODER_HARD(st, (uint8_t *)segs + i, sink, 1, false);
        for (int s = 0; s < (int)NUM_STATES; s++)
            metrics[(size_t)i * NUM_STATES + s] = (uint32_t)(*st->nodeMetricsCur)[s];
    }
    free(sink);
    free(st);
}
