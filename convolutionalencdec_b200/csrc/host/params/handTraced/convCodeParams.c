#include "convCodeParams.h"

/* K=3 hand-traced test code (handTracedTest/testParams/convCodeParams.c:6). */
const uint64_t g[n] = {0x7, 0x6};
