/* Layout queries for language bindings (ctypes) that cannot see the C headers. */
#include "convEncode.h"
#include "viterbiDecoder.h"
#include <stddef.h>

size_t ced_sizeof_encoder_state(void) { return sizeof(convEncoderState_t); }
size_t ced_sizeof_decoder_state(void) { return sizeof(viterbiHardState_t); }
size_t ced_offsetof_node_metrics_cur(void) { return offsetof(viterbiHardState_t, nodeMetricsCur); }
size_t ced_offsetof_edge_symm(void) { return offsetof(viterbiHardState_t, edgeCodedBitsSymm); }
size_t ced_offsetof_polynomials(void) { return offsetof(convEncoderState_t, polynomials); }
int ced_param_K(void) { return K; }
int ced_param_n(void) { return n; }
unsigned long ced_param_num_states(void) { return NUM_STATES; }
