/* Host helpers behind include/convHelpers.h (reference: src/convHelpers.c:3-15). */
#include "convHelpers.h"

void unpackBigToLittleEndian(uint8_t *unpackArray, int unpackArrayLen, uint64_t packed)
{
    for (int pos = 0; pos < unpackArrayLen; pos++)
        unpackArray[unpackArrayLen - 1 - pos] = (uint8_t)((packed >> pos) & 1u);
}

void unpackLittleToLittleEndian(uint8_t *unpackArray, int unpackArrayLen, uint64_t packed)
{
    for (int pos = 0; pos < unpackArrayLen; pos++)
        unpackArray[pos] = (uint8_t)((packed >> pos) & 1u);
}
