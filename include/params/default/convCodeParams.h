/*
 * Code parameters of the production decoder: K=7, rate 1/2, generators 0113 /
 * 0171 (octal, Proakis convention: MSb = newest input bit).  Same macro names
 * and values as the reference's src/defaultParams/convCodeParams.h:8-17, so
 * drivers written against the reference compile unchanged.  Pick a parameter
 * set by putting its directory first on the include path (-Iinclude/params/default).
 */
#ifndef CED_PARAMS_DEFAULT_CONV_CODE_PARAMS_H
#define CED_PARAMS_DEFAULT_CONV_CODE_PARAMS_H

#include <stdint.h>

#define k (1)                 /* input bits per trellis step                  */
#define n (2)                 /* coded bits per trellis step                  */
#define K (7)                 /* constraint length                            */
#define S ((K) - 1)           /* shift-register memory = tail segments        */
#define Rc ((double) k / n)   /* code rate                                    */
#define STARTING_STATE (0)    /* encoder starts (and is terminated) in state 0 */

extern const uint64_t g[n];   /* generator polynomials, params/default/convCodeParams.c */

#endif
