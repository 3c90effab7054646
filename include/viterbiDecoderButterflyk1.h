/*
 * viterbiDecoderButterflyk1.h -- the k=1 butterfly decoder entry points, drop-in
 * for the reference's src/viterbiDecoderButterflyk1.h:6-19.
 */
#ifndef CED_VITERBI_DECODER_BUTTERFLY_K1_H
#define CED_VITERBI_DECODER_BUTTERFLY_K1_H

#include "viterbiDecoder.h"

/*
 * Feed segmentsIn coded segments (one byte each, low n bits used) of the current
 * packet.  Nothing is emitted until last=true; then the whole packet is traced
 * back from state 0, the decoded bits are written MSb-first to `uncoded`, the
 * state is reset and the number of bytes written is returned (0 otherwise).
 * A packet may be split over any number of calls (last=true may come with
 * segmentsIn == 0).  CUDA failures: message on stdout, exit(1).
 */
int viterbiDecoderHardButterflyk1(viterbiHardState_t *restrict state, uint8_t *restrict codedSegments,
                                  uint8_t *restrict uncoded, int segmentsIn, bool last);

/* Build the trellis edge labels; prints the reference's banner line. */
void viterbiInitButterflyk1(viterbiHardState_t *state);

/* Start a new packet: metric 0 for state 0, NUM_STATES+1 elsewhere.  Writes but
 * never reads the struct (drivers call it on uninitialised memory before INIT). */
void resetViterbiDecoderHardButterflyk1(viterbiHardState_t *state);

METRIC_TYPE minMetricGeneric(const METRIC_TYPE (*metrics)[NUM_STATES]);

#endif
