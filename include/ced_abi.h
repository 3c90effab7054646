/*
 * ced_abi.h -- the C ABI between the host-side C library (the drop-in for
 * convEncode.h / viterbiDecoder.h / viterbiDecoderButterflyk1.h) and the
 * sm_100a CUDA kernels.  Plain pointers and sizes only; no C++ or torch types.
 *
 * The reference (ucb-cyarp/ConvolutionalEncDec) has no FFI layer: its boundary
 * is the three public headers plus the link line.  Every entry
 * point below therefore names the reference interface it stands behind:
 *
 *   ced_stream_encode   <- convEnc                         src/convEncode.h:74
 *   ced_stream_decode   <- viterbiDecoderHardButterflyk1   src/viterbiDecoderButterflyk1.h:6
 *                          (VITERBI_DECODER_HARD,           src/viterbiDecoder.h:87-95)
 *   ced_encode_batch    <- the loop `for pkt: convEnc(..., last=true)`
 *                                                          speedEncode/speedEncode.c:65-67
 *   ced_decode_batch    <- the loop `for pkt: VITERBI_DECODER_HARD(..., last=true)`
 *                                                          speedDecode/speedDecode.c:78-79
 *   ced_ber_count       <- bitErrors()                     berTestK7/berTestK7.c:45-53
 *   ced_bsc_channel     <- corruptCodedArray()             berTestK7/berTestK7.c:29-43
 *
 * There is no CPU implementation behind any of these: every call either runs
 * on the GPU or returns an error code (text in ced_last_error()).
 */
#ifndef CED_ABI_H
#define CED_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CED_OK 0
#define CED_ERR_CUDA (-1)        /* a CUDA runtime call or kernel failed           */
#define CED_ERR_ARG (-2)         /* bad pointer / size / alignment argument        */
#define CED_ERR_UNSUPPORTED (-3) /* code parameters outside what the kernels cover */
#define CED_ERR_NOMEM (-4)

#define CED_MAX_N 8        /* coded bits per segment (one byte per segment on the wire) */
#define CED_MAX_STATES 256 /* K <= 9 for the streaming kernel                            */

/* Code parameters, Proakis convention as in src/defaultParams/convCodeParams.c:6:
 * g[i] has its MSb (bit K-1) on the newest input bit.  k is always 1. */
typedef struct {
    int32_t constraintLen;      /* K */
    int32_t codedBits;          /* n */
    uint64_t gen[CED_MAX_N];    /* g[] */
} ced_code_t;

/* NOTE: no identifier in this header is a bare K, k, n, S or g -- the reference's
 * parameter headers #define those (src/defaultParams/convCodeParams.h:8-17). */

typedef struct ced_ctx ced_ctx; /* one per GPU: stream, scratch, staging buffers */

/* ---------------------------------------------------------------- contexts */
int ced_device_count(void);
const char *ced_last_error(void);
int ced_ctx_create(int device, ced_ctx **out);
void ced_ctx_destroy(ced_ctx *ctx);
int ced_ctx_device(const ced_ctx *ctx);
/* Lazily-created context on device $CED_DEVICE (default 0) used by the
 * streaming entry points; NULL on failure. */
ced_ctx *ced_default_ctx(void);
/* Block until everything queued on `stream` (NULL: the context's stream) ran. */
int ced_sync(ced_ctx *ctx, void *stream);
/* Number of kernel launches issued through this context so far. */
uint64_t ced_launch_count(const ced_ctx *ctx);

/* ------------------------------------------------------------ measurement
 * ced_ctx_set_profiling(ctx, 1): ced_decode_batch brackets its forward (ACS)
 * and traceback kernels with CUDA events on the launching stream;
 * ced_ctx_last_kernel_ms then returns their summed durations for the most
 * recent call ([0] forward, [1] traceback) after synchronising those events.
 * ced_probe_int_peak runs a dependent-free LOP3 stream (mode 0: ALU pipe only,
 * mode 1: LOP3 + IMAD co-issued) and reports 32-bit lane-operations per second
 * of the LOP3 stream -- the INT roofline denominator of SURVEY 8(d). */
int ced_ctx_set_profiling(ced_ctx *ctx, int enable);
int ced_ctx_last_kernel_ms(ced_ctx *ctx, float *ms2);
int ced_probe_int_peak(ced_ctx *ctx, int mode, double *laneOpsPerSecond);
/* With CED_FUSED=1 (or 2) in the environment ced_decode_batch runs batches of at least 16384 frames of the compiled-in
 * codes as ONE kernel with the traceback inside it (k7FusedKernel / k7FusedWsKernel; then
 * ced_ctx_last_kernel_ms: [0] that kernel, [1] the re-decode below).  A frame whose in-kernel traceback cannot be
 * proven equal to the reference's full traceback is decoded again by the two-kernel path in the same call; this
 * returns how many frames of the (last wave of the) most recent call that was.  Synchronises the device. */
int ced_ctx_last_fallback_frames(ced_ctx *ctx, int *frames);

/* ------------------------------------------------------- batched hot path
 * Frames are independent packets, each encoded from state 0 and terminated
 * with K-1 zero bits.  A frame of frameBits information bits (multiple of 8)
 * is frameBits+K-1 coded segments, one byte per segment holding c0 | c1<<1
 * (src/viterbiDecoder.h:154), at dSegs[frame*segStride + t].  Decoded / message
 * bits are MSb-first bytes at d*[frame*stride + t/8].
 *
 * Pointers are DEVICE pointers; `stream` is a cudaStream_t.  NULL selects the
 * context's own stream, which is non-blocking: it does NOT synchronise with
 * the legacy default stream -- pass cudaStreamLegacy ((cudaStream_t)0x1) to
 * launch there.  Calls are asynchronous with respect to the host.
 * Fastest when base pointers and strides are multiples of 16 bytes; any
 * alignment is accepted.
 * Codes: any k=1 code with constraintLen 2..9 and 1..8 coded bits.  K=7 codes with
 * 2 or 3 generators that all tap the newest and the oldest bit (src/viterbiDecoder.c:20-24)
 * run on the hand-scheduled SIMD-in-word kernel -- 0113/0171 and 0133/0171 compiled in, any
 * other set through a table built on first use; K = 3 .. 9 with 2 or 3 generators of
 * ANY shape on table-driven SIMD-in-word kernels (swar_generic.cu); the rest on a
 * one-warp-per-frame kernel.  (k > 1: ced_decode_batch_k.)
 */
int ced_decode_batch(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dSegs, size_t segStride,
                     int nFrames, int frameBits, uint8_t *dOut, size_t outStride, void *stream);

int ced_encode_batch(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dMsg, size_t msgStride,
                     int nFrames, int frameBytes, uint8_t *dSegs, size_t segStride, void *stream);

/* Packed wire format (not in the reference): four 2-bit segments per byte, segment t
 * in bits 2*(t%4)..2*(t%4)+1 of byte t/4, rows of ceil((frameBits+K-1)/4) valid bytes.  A quarter of
 * the HBM and PCIe bytes of the byte-per-segment format; decoded output is identical.
 * ced_pack_symbols converts byte-per-segment symbols (low 2 bits) to this format on the device. */
int ced_decode_batch_packed(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dPacked, size_t packedStride,
                            int nFrames, int frameBits, uint8_t *dOut, size_t outStride, void *stream);
int ced_decode_batch_packed_host(ced_ctx *ctx, const ced_code_t *code, const uint8_t *hPacked, size_t packedStride,
                                 int nFrames, int frameBits, uint8_t *hOut, size_t outStride);
int ced_pack_symbols(ced_ctx *ctx, const uint8_t *dSegs, size_t segStride, int nFrames, int segsPerFrame,
                     uint8_t *dPacked, size_t packedStride, void *stream);
/* The same packing on the HOST (no GPU involved; `threads` worker threads, AVX2 when available), e.g. for a
 * receiver that wants to hand ced_decode_batch_packed_host a quarter of the bytes.  ced_decode_batch_host
 * uses it internally as transfer compression: for page-locked buffers only on the chunks the copy engine is not
 * ready for, for pageable buffers on every chunk (CED_HOST_PACK = 0 / 1 / 2 forces never / always / adaptive;
 * DESIGN.md 6). */
int ced_host_pack_symbols(const uint8_t *segs, size_t segStride, int nFrames, int segsPerFrame, uint8_t *packed,
                          size_t packedStride, int threads);
/* Encoder writing the packed format directly (n = 2 codes). */
int ced_encode_batch_packed(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dMsg, size_t msgStride,
                            int nFrames, int frameBytes, uint8_t *dPacked, size_t packedStride, void *stream);
/* Soft symbols: two int8 per segment (coded bit 0 then coded bit 1; BPSK bit 0 -> +, bit 1 -> -), rows of
 * 2 * (frameBits + K - 1) bytes.
 *
 * ced_decode_batch_soft is a true soft-decision decoder (north_star kernel (1) "soft or hard symbols"): the
 * branch cost of an edge is the reference's calcHammingDist (src/viterbiDecoder.c:260-285, used at
 * src/viterbiDecoderButterflyk1.c:104-115) with every disagreeing coded bit weighted by its reliability |s|;
 * trellis, tie rule, traceback and output format are the hard decoder's.  Definition: oracle/ced_oracle.c
 * orc_dec_step_soft.  Inputs of one constant magnitude give exactly the hard decoder's output.  16-bit path
 * metrics (k7SoftForwardKernel); codes 0113/0171 and 0133/0171; dSoft and softStride multiples of 16 bytes.
 *
 * ced_slice_soft_symbols / ced_slice_soft_to_bytes only keep the sign (0 slices to bit 0) and produce the
 * packed / byte-per-segment hard format -- what a receiver without soft information would decode. */
int ced_decode_batch_soft(ced_ctx *ctx, const ced_code_t *code, const int8_t *dSoft, size_t softStride, int nFrames,
                          int frameBits, uint8_t *dOut, size_t outStride, void *stream);
int ced_slice_soft_symbols(ced_ctx *ctx, const int8_t *dSoft, size_t softStride, int nFrames, int segsPerFrame,
                           uint8_t *dPacked, size_t packedStride, void *stream);
int ced_slice_soft_to_bytes(ced_ctx *ctx, const int8_t *dSoft, size_t softStride, int nFrames, int segsPerFrame,
                            uint8_t *dSegs, size_t segStride, void *stream);
/* 3-bit soft decisions at (nearly) the hard decoder's speed and wire size: one byte per segment x0 | x1 << 3, x in 0..7
 * (0 = surely bit 0 ... 7 = surely bit 1; generator 0 in the low field), i.e. the reliabilities s = 7 - 2x of
 * ced_decode_batch_soft's definition; the decoded bytes equal ced_decode_batch_soft's on those int8 values.  Byte path
 * metrics, the hard kernel's survivor stream and traceback (k7SoftQForwardKernel).  Codes 0113/0171 and
 * 0133/0171; any base / stride (16-byte aligned rows are fastest).  ced_quantize_soft turns int8 reliabilities (two per
 * segment) into this format with a uniform 8-level quantiser of step `delta` (thresholds 0, +-delta, +-2 delta,
 * +-3 delta; about 0.6 sigma of the channel noise in int8 units is the usual choice). */
int ced_decode_batch_softq(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dSyms, size_t symStride, int nFrames,
                           int frameBits, uint8_t *dOut, size_t outStride, void *stream);
int ced_quantize_soft(ced_ctx *ctx, const int8_t *dSoft, size_t softStride, int nFrames, int segsPerFrame, double delta,
                      uint8_t *dSyms, size_t symStride, void *stream);
/* the same decode on HOST buffers through the chunked H2D / kernel / D2H pipeline of ced_decode_batch_host (synchronous) */
int ced_decode_batch_softq_host(ced_ctx *ctx, const ced_code_t *code, const uint8_t *hSyms, size_t symStride, int nFrames,
                                int frameBits, uint8_t *hOut, size_t outStride);
/* BPSK over AWGN with int8 soft output (the BER sweep's channel, berTestK7/berTestK7.c:29-43 generalised):
 * soft = clamp(round(amplitude * ((1 - 2 bit) + sigma * N(0,1))), -127, 127) for both coded bits of every
 * byte-per-segment symbol; for rate 1/2, sigma = 1 / sqrt(Eb/N0).  Counter-based generator keyed by (seed,
 * frame index, segment), independent of sharding.  dCounters (may be NULL): [0] += coded bits whose sign came
 * out wrong, [1] += coded bits. */
int ced_awgn_channel(ced_ctx *ctx, const uint8_t *dSegs, size_t segStride, int nFrames, int segsPerFrame, int8_t *dSoft,
                     size_t softStride, double amplitude, double sigma, uint64_t seed, uint64_t firstFrameIndex,
                     uint64_t *dCounters, void *stream);

/* Same operations on HOST buffers: chunked H2D / kernel / D2H pipelined on a copy-in stream, four compute
 * streams and a copy-out stream; host worker threads pack part of the symbol chunks to 2 bits while the copy
 * engine moves the others.  Synchronous: returns when hOut / hSegs is complete. */
int ced_decode_batch_host(ced_ctx *ctx, const ced_code_t *code, const uint8_t *hSegs, size_t segStride,
                          int nFrames, int frameBits, uint8_t *hOut, size_t outStride);

int ced_encode_batch_host(ced_ctx *ctx, const ced_code_t *code, const uint8_t *hMsg, size_t msgStride,
                          int nFrames, int frameBytes, uint8_t *hSegs, size_t segStride);

/* Page-locked host memory for the *_host calls (pageable buffers work too, at a fraction of the PCIe rate). */
int ced_host_alloc(size_t bytes, void **out);
void ced_host_free(void *p);
/* Page-lock memory the caller already owns (malloc'ed arrays it will pass to the *_host calls many times);
 * takes about as long as one copy of the buffer, so it pays from the second call on. */
int ced_host_register(void *p, size_t bytes);
int ced_host_unregister(void *p);

/* Bytes of survivor scratch ced_decode_batch keeps inside the context for a
 * batch of this shape (grown on demand, reused across calls). */
size_t ced_decode_scratch_bytes(int nFrames, int frameBits);

/*
 * Continuous streams with windowed traceback (bounded survivor memory for streams longer than the
 * reference's MAX_PKT_LEN_SEGMENTS; TRACEBACK_LEN src/viterbiDecoder.h:19 and vitdec(..., tblen, ...) in
 * scripts/matlab/viterbiBEREstimate.m:99 are the models, the reference's own windowed decoder does not
 * run at HEAD).  nStreams independent K=7 streams that started in state 0 are decoded a slice at a time:
 * every call consumes the next nSegments segments of each stream (row i of dSegs) and writes the bits
 * that became final -- each decided by a traceback of at least `depth` steps from the best-metric state --
 * at the start of row i of dOut, MSb-first.  The caller concatenates the rows of successive calls.
 *   streamPos  segments of each stream consumed by earlier calls (0 on the first call)
 *   depth      traceback depth in steps, a multiple of 24 (48 covers the usual 5*K = 35)
 *   last       the stream ends with this slice, terminated by K-1 zero bits: traceback starts in state 0
 *              and everything left is flushed (total length - (K-1) must be a multiple of 8)
 *   dCarry     ced_window_carry_bytes(nStreams, depth) bytes of device memory, 16-byte aligned, owned by
 *              the caller and passed unchanged from call to call (metrics + the last `depth` decisions)
 * nSegments and streamPos must be multiples of 96 except nSegments of the last call.
 * Returns the number of decoded BYTES written per stream by this call (>= 0), or a negative CED_ERR_*.
 *
 * ced_decode_window_batch_packed takes the packed wire format (4 segments per byte); its slices and streamPos are
 * multiples of 192 segments (48 bytes), the last slice excepted.
 *
 * Pinning: the reference's own windowed decoder does not run at HEAD, so the semantics are those of
 * oracle/ced_oracle.c:orc_decode_window (bit-exact tests).  That definition, run with one-step slices and depth 35,
 * IS vitdec(..., tblen = 35, 'term', 'hard') and reproduces the MATLAB expectations the reference keeps in
 * berTestK7/berTestK7.c:98 within its own +-10 % rule (tests/test_oracle.py); with depth >= the stream length it is the
 * reference's full traceback.
 */
size_t ced_window_carry_bytes(int nStreams, int depth);
/* ... for any code ced_decode_window_batch takes: the K=7 codes above, and -- byte format -- every other code with K = 3, 4, 5
 * or 7 and 2 or 3 generators of any shape (table-driven kernels; their carry block depends on the number of states).
 * 0 = not a code the windowed decoder takes. */
size_t ced_window_carry_bytes_code(const ced_code_t *code, int nStreams, int depth);
int ced_decode_window_batch(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dSegs, size_t segStride,
                            int nStreams, int nSegments, uint64_t streamPos, int depth, int last, void *dCarry,
                            uint8_t *dOut, size_t outStride, void *stream);
int ced_decode_window_batch_packed(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dPacked, size_t packedStride,
                                   int nStreams, int nSegments, uint64_t streamPos, int depth, int last, void *dCarry,
                                   uint8_t *dOut, size_t outStride, void *stream);
/* the same for 3-bit soft symbols (one byte per segment x0 | x1 << 3, see ced_decode_batch_softq): slices and streamPos
 * multiples of 96 segments; codes 0113/0171 and 0133/0171.  Semantics: orc_decode_window_soft (the windowing of
 * orc_decode_window around the soft recursion) on the reliabilities s = 7 - 2x. */
int ced_decode_window_batch_softq(ced_ctx *ctx, const ced_code_t *code, const uint8_t *dSyms, size_t symStride, int nStreams,
                                  int nSegments, uint64_t streamPos, int depth, int last, void *dCarry, uint8_t *dOut,
                                  size_t outStride, void *stream);

/* dCounters[0] += popcount(dA ^ dB) over nFrames x bytesPerFrame; dCounters[1] +=
 * bits compared.  Device-side uint64 counters, so a BER sweep can all-reduce
 * them (NCCL sum) without a host round trip. */
int ced_ber_count(ced_ctx *ctx, const uint8_t *dA, size_t strideA, const uint8_t *dB, size_t strideB,
                  int nFrames, int bytesPerFrame, uint64_t *dCounters, void *stream);

/* Binary symmetric channel on byte-per-segment symbols: each of the codedBits coded
 * bits of every segment is flipped with probability p using a counter-based
 * generator keyed by (seed, frame index, segment index), so the result does
 * not depend on how frames are sharded.  dCounters (may be NULL):
 * [0] += flips, [1] += coded bits. */
int ced_bsc_channel(ced_ctx *ctx, uint8_t *dSegs, size_t segStride, int nFrames, int segsPerFrame, int codedBits,
                    double p, uint64_t seed, uint64_t firstFrameIndex, uint64_t *dCounters, void *stream);

/* Uniform random message bytes keyed by (seed, frame index, byte index). */
int ced_random_bytes(ced_ctx *ctx, uint8_t *dMsg, size_t msgStride, int nFrames, int frameBytes,
                     uint64_t seed, uint64_t firstFrameIndex, void *stream);

/* ------------------------------------------------ rate-k/n codes with k > 1
 * The reference's headers describe a code by K, k, n and n generators of k*K bits: ONE shift register of k*K bits that
 * takes inputBits = k message bits per coded segment (src/convEncode.h:8-18, src/convEncode.c:46-130), decoded on a
 * trellis of 2^(k*(K-1)) states with 2^k branches into every state (src/viterbiDecoder.c:95-128; the lowest edgeIn keeps
 * a tie).  code->gen[i] is generator i as the reference writes it (MSb = newest input bit, k*K bits wide).
 * A frame of frameBytes*8 message bits is frameBytes*8/k + K-1 segments, one byte per segment (low n bits).
 * The decoder is the reference's add-compare-select followed by the full traceback from state 0 that its butterfly
 * decoder performs (src/viterbiDecoderButterflyk1.c:200-256, formulas for general k) -- the reference's own k > 1
 * traceback does not run at HEAD; encoder, trellis labels and per-step path metrics are pinned to the unmodified
 * reference built with k = 2 parameters (tests/test_oracle_k.py).
 * inputBits in {1, 2, 4} (1 forwards to ced_encode_batch / ced_decode_batch), k*(K-1) <= 8, k*K <= 32.  k = 2 with 2 or 3
 * generators runs on thread-per-frame SIMD-in-word kernels with radix-4 butterflies, the rest on a one-warp-per-frame
 * kernel. */
int ced_encode_batch_k(ced_ctx *ctx, const ced_code_t *code, int inputBits, const uint8_t *dMsg, size_t msgStride,
                       int nFrames, int frameBytes, uint8_t *dSegs, size_t segStride, void *stream);
int ced_decode_batch_k(ced_ctx *ctx, const ced_code_t *code, int inputBits, const uint8_t *dSegs, size_t segStride,
                       int nFrames, int frameBits, uint8_t *dOut, size_t outStride, void *stream);

/* ------------------------------------------------ one process, several GPUs
 * Frames are independent packets (the reference resets its state per packet,
 * src/viterbiDecoderButterflyk1.c:259), so a batch shards over the GPUs of a box with no traffic between
 * them: device g of G takes the contiguous range ced_shard_range(nFrames, G, g) of the caller's HOST arrays
 * and runs its own host pipeline (ced_decode_batch_host) on a worker thread that is pinned to the CPUs of the
 * device's NUMA node, so its page-locked staging buffers are NUMA-local.  The loop these calls replace is
 * `for pkt: VITERBI_DECODER_HARD(..., last=true)` of speedDecode/speedDecode.c:78-79 / the convEnc loop of
 * speedEncode/speedEncode.c:65-67, for a batch that is larger than one GPU should take.
 *
 * ced_ber_allreduce is the only collective: the uint64 counters that ced_ber_count / ced_bsc_channel /
 * ced_awgn_channel accumulated on every device (dCounters[g] = device pointer on device g, `count` words each)
 * are summed in place over NCCL (communicator from ncclCommInitAll on first use, one ncclAllReduce per device
 * inside a group, NVLink / NVSwitch).  NCCL is opened with dlopen at that point; without it the call returns
 * CED_ERR_UNSUPPORTED -- there is no host-side substitute.
 */
typedef struct ced_multi ced_multi;
void ced_shard_range(int nFrames, int nShards, int shard, int *first, int *count);
int ced_multi_create(const int *devices, int nDevices, ced_multi **out); /* nDevices = 0: every visible GPU */
void ced_multi_destroy(ced_multi *m);
int ced_multi_device_count(const ced_multi *m);
ced_ctx *ced_multi_ctx(ced_multi *m, int i);                             /* context of the i-th device */
int ced_decode_batch_host_multi(ced_multi *m, const ced_code_t *code, const uint8_t *hSegs, size_t segStride,
                                int nFrames, int frameBits, uint8_t *hOut, size_t outStride);
int ced_encode_batch_host_multi(ced_multi *m, const ced_code_t *code, const uint8_t *hMsg, size_t msgStride,
                                int nFrames, int frameBytes, uint8_t *hSegs, size_t segStride);
int ced_ber_allreduce(ced_multi *m, uint64_t *const *dCounters, int count);
int ced_nccl_version(void); /* 0 if NCCL could not be opened */
/* Device memory for C callers that keep batches resident (BER mode): zero-filled allocation on the context's
 * device, synchronous copies on its stream. */
int ced_device_alloc(ced_ctx *ctx, size_t bytes, void **out);
void ced_device_free(ced_ctx *ctx, void *p);
int ced_copy_to_device(ced_ctx *ctx, void *dDst, const void *hSrc, size_t bytes);
int ced_copy_to_host(ced_ctx *ctx, void *hDst, const void *dSrc, size_t bytes);
/* Raw copy rate between page-locked host memory and the device(s), no kernels: the ceiling for every
 * host-buffer figure (best of `reps`, bytes per second; the multi form runs all devices at once and sums). */
int ced_probe_copy_ceiling(ced_ctx *ctx, size_t bytes, int reps, double *h2dBytesPerSecond, double *d2hBytesPerSecond);
int ced_multi_probe_copy_ceiling(ced_multi *m, size_t bytesPerDevice, int reps, double *h2dBytesPerSecond,
                                 double *d2hBytesPerSecond);

/* ------------------------------------------------ per-frame streaming path
 * One frame, fed in arbitrary chunks exactly like the reference API
 * (SURVEY A.6).  All decoder state lives in caller memory (the host struct
 * viterbiHardState_t); the device side is stateless between calls.
 *
 *   edge[b*nStates + s] : coded segment on the edge leaving state s with input b
 *   metrics             : nStates path metrics, updated in place, with the
 *                         reference's uint8 arithmetic and renormalisation
 *                         schedule (src/viterbiDecoderButterflyk1.c:159-183)
 *   iteration, renorm   : viterbiHardState_t.iteration / .renormCounter
 *   surv                : host mirror of the packed survivor decisions,
 *                         survWordsPerStep uint32 per trellis step
 *                         (ced_stream_surv_words(nStates)), capacity in steps
 * Returns the number of decoded bytes written (0 unless last), <0 on error.
 */
int ced_stream_surv_words(int nStates);
/* With CED_STREAM_SERVER=1 one-shot K=7 n=2 packets (last = true on the first call of a packet) are decoded by a
 * RESIDENT kernel that takes packet after packet from a mailbox in page-locked host memory -- no launch, copy or stream
 * synchronise per call (frame_server.cuh; it leaves by itself after 2 ms without a request and whenever another entry
 * point of the library needs the GPU).  Default: one graph launch per packet (measured equally fast at the reference's
 * packet length).  Returns 1 if the resident path has been used and is healthy; *requests = packets it
 * has answered, *launches = times the kernel was started. */
int ced_stream_server_stats(uint64_t *requests, uint64_t *launches);

int ced_stream_decode(int constraintLen, int codedBits, const uint8_t *edge, uint8_t *metrics, uint32_t *iteration,
                      uint32_t *renormCounter, uint32_t *surv, uint32_t survCapacitySteps,
                      const uint8_t *segs, int segmentsIn, uint8_t *uncoded, int last);

/* taps[i]: generator i with bit 0 on the newest input bit (convEncoderState_t.
 * polynomials).  *reg is convEncoderState_t.tappedDelay.  Returns segments written. */
int ced_stream_encode(int constraintLen, int codedBits, const uint32_t *taps, uint32_t *reg, const uint8_t *in, int bytesIn,
                      uint8_t *segs, int last);

#ifdef __cplusplus
}
#endif
#endif /* CED_ABI_H */
