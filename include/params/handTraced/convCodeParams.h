/*
 * Code parameters of the hand-traced known-answer test: K=3, rate 1/2,
 * generators 0b111 / 0b110 (handTracedTest/testParams/convCodeParams.h:8-17,
 * convCodeParams.c:6 in the reference).
 */
#ifndef CED_PARAMS_HANDTRACED_CONV_CODE_PARAMS_H
#define CED_PARAMS_HANDTRACED_CONV_CODE_PARAMS_H

#include <stdint.h>

#define k (1)
#define n (2)
#define K (3)
#define S ((K) - 1)
#define Rc ((double) k / n)
#define STARTING_STATE (0)

extern const uint64_t g[n];

#endif
