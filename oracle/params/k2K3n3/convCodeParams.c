#include "convCodeParams.h"
/* Proakis convention: MSb = most recent input bit (src/defaultParams/convCodeParams.c:3-6) */
const uint64_t g[n] = {027,075,072};
