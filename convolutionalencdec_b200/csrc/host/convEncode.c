/*
 * Host side of include/convEncode.h.  convEnc() marshals one call to the GPU
 * encoder kernel (ced_stream_encode, include/ced_abi.h); the remaining
 * functions are table-building helpers that touch a handful of bits and stay
 * on the host (they are what viterbiInit* uses to label trellis edges,
 * reference: src/viterbiDecoderButterflyk1.c:24-29).
 */
#include "convEncode.h"
#include "ced_abi.h"
#include <stdio.h>
#include <stdlib.h>

#if k != 1
#error the GPU encoder path covers k == 1 codes
#endif

TAPPED_DELAY_TYPE bitReverseGenerator(TAPPED_DELAY_TYPE packed)
{
    /* reference: src/convEncode.c:163-175 -- K*k bits, bit i <-> bit K*k-1-i */
    TAPPED_DELAY_TYPE out = 0;
    for (int i = 0; i < k * K; i++)
        if ((packed >> i) & 1u)
            out |= (TAPPED_DELAY_TYPE)((TAPPED_DELAY_TYPE)1 << (k * K - 1 - i));
    return out;
}

void resetConvEncoder(convEncoderState_t *state)
{
    state->tappedDelay = STARTING_STATE;
    state->remainingUncoded = 0;
    state->remainingUncodedCount = 0;
}

void initConvEncoder(convEncoderState_t *state)
{
    for (int i = 0; i < n; i++)
        state->polynomials[i] = bitReverseGenerator((TAPPED_DELAY_TYPE)g[i]);
}

uint8_t computeEncOutputSegment(convEncoderState_t *state)
{
    /* reference: src/convEncode.c:132-161 -- parity of the tapped bits, generator i -> bit i */
    uint8_t segment = 0;
    for (int i = 0; i < n; i++) {
        const unsigned long long tapped = (unsigned long long)(state->tappedDelay & state->polynomials[i]);
        segment |= (uint8_t)((__builtin_parityll(tapped) & 1) << i);
    }
    return segment;
}

int convEncOneInput(convEncoderState_t *state, uint8_t bitsToShiftIn)
{
    /* reference: src/convEncode.c:19-44 with k == 1 */
    state->tappedDelay = (TAPPED_DELAY_TYPE)((state->tappedDelay << 1) | (bitsToShiftIn & 1u));
    return computeEncOutputSegment(state);
}

int convEnc(convEncoderState_t *state, uint8_t *uncoded, uint8_t *codedSegments, int bytesIn, bool last)
{
    uint32_t taps[n];
    for (int i = 0; i < n; i++)
        taps[i] = (uint32_t)state->polynomials[i];
    uint32_t reg = (uint32_t)state->tappedDelay;
    const int segments = ced_stream_encode(K, n, taps, &reg, uncoded, bytesIn, codedSegments, last ? 1 : 0);
    if (segments < 0) {
        printf("convEnc: GPU encoder failed: %s\n", ced_last_error());
        exit(1);
    }
    if (last)
        resetConvEncoder(state); /* reference: src/convEncode.c:122 */
    else
        state->tappedDelay = (TAPPED_DELAY_TYPE)reg;
    return segments;
}
