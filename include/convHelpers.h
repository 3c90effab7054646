/*
 * convHelpers.h -- small helpers shared by the encoder and decoder headers.
 * Drop-in for the reference's src/convHelpers.h: POW2 stays `unsigned long`
 * (drivers print NUM_STATES with %lu, speedDecode/speedDecode.c:130) and the
 * two unpack helpers keep their signatures (:8-9).  The reference's ARGMIN_*
 * macro family (:11-102) only served its generic register-exchange decoder,
 * which this library does not accelerate; it is not reproduced.
 */
#ifndef CED_CONV_HELPERS_H
#define CED_CONV_HELPERS_H

#include <stdint.h>

#define POW2(X) (1ul << (X))

/* rotate the low BITS bits of VAL by SHIFT_AMT (src/convHelpers.h:115-116) */
#define ROTATE_RIGHT(VAL, SHIFT_AMT, BITS) \
    ((((VAL) >> (SHIFT_AMT)) | (((VAL) & (POW2(SHIFT_AMT) - 1)) << ((BITS) - (SHIFT_AMT)))))
#define ROTATE_LEFT(VAL, SHIFT_AMT, BITS) \
    (((((VAL) & (POW2((BITS) - (SHIFT_AMT)) - 1)) << (SHIFT_AMT)) | ((VAL) >> ((BITS) - (SHIFT_AMT)))))

/* one bit per output byte; "BigToLittle": unpackArray[len-1] receives the LSb */
void unpackBigToLittleEndian(uint8_t *unpackArray, int unpackArrayLen, uint64_t packed);
/* unpackArray[0] receives the LSb */
void unpackLittleToLittleEndian(uint8_t *unpackArray, int unpackArrayLen, uint64_t packed);

#endif
