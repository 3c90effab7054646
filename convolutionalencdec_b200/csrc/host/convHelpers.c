/*
 * Host helpers behind include/convHelpers.h.  Neither function is used by the encoder or decoder
 * (the same is true in the reference); they are exported because the reference's objects export them.
 */
#include "convHelpers.h"

/* bit `pos` of `packed` goes to slot pos (ascending) or to slot len-1-pos (descending) */
static void spreadBits(uint8_t *slots, int len, uint64_t packed, int descending)
{
    for (int pos = 0; pos < len; pos++) {
        const uint8_t bit = (uint8_t)((packed >> pos) & 1u);
        slots[descending ? len - 1 - pos : pos] = bit;
    }
}

void unpackBigToLittleEndian(uint8_t *unpackArray, int unpackArrayLen, uint64_t packed)
{
    spreadBits(unpackArray, unpackArrayLen, packed, 1);
}

void unpackLittleToLittleEndian(uint8_t *unpackArray, int unpackArrayLen, uint64_t packed)
{
    spreadBits(unpackArray, unpackArrayLen, packed, 0);
}
