/* TEST INFRASTRUCTURE: code parameters in the reference's own format (src/defaultParams/convCodeParams.h:8-17)
 * for building the UNMODIFIED reference sources with k = 2 (oracle/Makefile, target refk).  Rate-2/3 code of
 * constraint length K = 3 from Proakis, Digital Communications 4th ed., table 8.2-8. */
#ifndef _CONV_CODE_PARAMS_H_
#define _CONV_CODE_PARAMS_H_
#include <stdint.h>
#define K (3)
#define k (2)
#define S ((K)-1)
#define n (3)
#define Rc ((double) k/n)
#define STARTING_STATE (0)
extern const uint64_t g[n];
#endif
