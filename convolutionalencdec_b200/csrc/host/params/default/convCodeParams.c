#include "convCodeParams.h"

/* K=7 rate-1/2 generators, octal, MSb = newest input bit
 * (same values as the reference's src/defaultParams/convCodeParams.c:6). */
const uint64_t g[n] = {0113, 0171};
