/*
 * convEncode.h -- rate-k/n convolutional encoder, drop-in for the reference's
 * src/convEncode.h (same type, field and function names, same argument
 * meaning).  convEnc() runs on the GPU through ced_stream_encode()
 * (include/ced_abi.h); the bit-level helpers that only build tables
 * (convEncOneInput, computeEncOutputSegment, bitReverseGenerator) are host code.
 */
#ifndef CED_CONV_ENCODE_H
#define CED_CONV_ENCODE_H

#include "convCodeParams.h"
#include "convHelpers.h"
#include <stdbool.h>

/* smallest unsigned type holding the k*K-bit shift register (src/convEncode.h:8-18) */
#if k * K <= 8
#define TAPPED_DELAY_TYPE uint8_t
#elif k * K <= 16
#define TAPPED_DELAY_TYPE uint16_t
#elif k * K <= 32
#define TAPPED_DELAY_TYPE uint32_t
#elif k * K <= 64
#define TAPPED_DELAY_TYPE uint64_t
#else
#error constraint lengths with k*K > 64 are not supported
#endif

/* Caller-owned encoder state (src/convEncode.h:29-38).  tappedDelay bit 0 is the
 * bit shifted in last; polynomials[i] is g[i] bit-reversed to that order. */
typedef struct {
    TAPPED_DELAY_TYPE tappedDelay;
    TAPPED_DELAY_TYPE polynomials[n];
    uint8_t remainingUncoded;      /* only used when 8 % k != 0 */
    uint8_t remainingUncodedCount;
} convEncoderState_t;

TAPPED_DELAY_TYPE bitReverseGenerator(TAPPED_DELAY_TYPE packed);

/* Return the shift register to STARTING_STATE. */
void resetConvEncoder(convEncoderState_t *state);

/* Fill state->polynomials from g[]. */
void initConvEncoder(convEncoderState_t *state);

/*
 * Encode bytesIn bytes (ascending index, MSb of each byte first) into one byte per
 * n-bit coded segment, generator i in bit i.  With last=true, S zero chunks are
 * appended and the encoder resets.  codedSegments must hold 8*bytesIn/k + S
 * entries.  Returns the number of segments written.  On a CUDA failure the
 * reference's error convention applies: message on stdout, exit(1).
 */
int convEnc(convEncoderState_t *state, uint8_t *uncoded, uint8_t *codedSegments, int bytesIn, bool last);

/* Output segment for the current shift-register contents. */
uint8_t computeEncOutputSegment(convEncoderState_t *state);

/* Shift k bits in (MSb first) and return the resulting segment; used to derive
 * the trellis edge labels. */
int convEncOneInput(convEncoderState_t *state, uint8_t bitsToShiftIn);

#endif
