/*
 * ced_internal.cuh -- what the translation units behind include/ced_abi.h share: the per-GPU context
 * (ced_ctx), the device / pinned buffer helpers and the error convention.  Not part of the ABI.
 */
#pragma once
#include "../../include/ced_abi.h"
#include "decode_batch.cuh"

#include <algorithm>
#include <cstdlib>
#include <cuda_runtime.h>
#include <mutex>
#include <vector>

namespace ced_host {
struct Packer;
Packer *packerCreate(int threads);
void packerDestroy(Packer *p);
int packerThreads(const Packer *p);
void packerRun(Packer *p, const uint8_t *in, size_t inStride, int nRows, int segs, uint8_t *out, size_t outStride);
} // namespace ced_host

/* text behind ced_last_error(), one buffer per thread (defined in ced_abi.cu) */
void cedSetError(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#define setError cedSetError

#define CED_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            setError("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return CED_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

constexpr size_t kMaxScratchBytes = 12ull << 30;  /* survivor scratch per wave            */
constexpr size_t kMaxWaveFrames = 1u << 20;       /* bounds the scheduler state (2 KB per 32 frames) */
constexpr int kHostChunkFrames = 8192;            /* frames per H2D/kernel/D2H pipeline stage */
constexpr int kPipeDepth = 4;                     /* chunks in flight in the host pipeline      */
constexpr int kStreamSplitMinSegments = 134;       /* one-shot packets from this length on: warp_split.cu (ced_stream_decode) */
constexpr uint32_t kStreamMaxSteps = 16384 + 8;   /* MAX_PKT_LEN_SEGMENTS (src/viterbiDecoder.h:18,45) */

template <typename T>
struct DeviceBuf {
    T *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes)
            return CED_OK;
        if (p)
            cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), need);
        if (e != cudaSuccess) {
            setError("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
            return CED_ERR_NOMEM;
        }
        bytes = need;
        return CED_OK;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct PinnedBuf {
    uint8_t *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes)
            return CED_OK;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&p), need);
        if (e != cudaSuccess) {
            setError("cudaMallocHost(%zu) failed: %s", need, cudaGetErrorString(e));
            return CED_ERR_NOMEM;
        }
        bytes = need;
        return CED_OK;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

enum class CodeId { Unsupported, K7_0113_0171, K7_0133_0171, K7_Runtime, K7_RuntimeN3 };

inline CodeId classify(const ced_code_t *c)
{
    /* generators that tap the newest and the oldest bit (the butterfly symmetry the reference itself
     * requires, src/viterbiDecoder.c:20-24) */
    auto bothEnds = [](uint64_t g) { return g < 128 && (g & 1u) && ((g >> 6) & 1u); };
    if (c && c->constraintLen == 7 && c->codedBits == 3 && bothEnds(c->gen[0]) && bothEnds(c->gen[1]) &&
        bothEnds(c->gen[2]))
        return CodeId::K7_RuntimeN3;
    if (!c || c->constraintLen != 7 || c->codedBits != 2)
        return CodeId::Unsupported;
    if (c->gen[0] == 0113 && c->gen[1] == 0171)
        return CodeId::K7_0113_0171;
    if (c->gen[0] == 0133 && c->gen[1] == 0171)
        return CodeId::K7_0133_0171;
    /* any other symmetric pair: SWAR kernel driven by a step table */
    if (bothEnds(c->gen[0]) && bothEnds(c->gen[1]))
        return CodeId::K7_Runtime;
    return CodeId::Unsupported;
}

inline uint32_t reverseBits(uint64_t g, int K)
{
    uint32_t r = 0;
    for (int i = 0; i < K; i++)
        r |= (uint32_t)((g >> i) & 1u) << (K - 1 - i);
    return r;
}


struct ced_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;   /* compute */
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t inReady[kPipeDepth] = {}, inFree[kPipeDepth] = {}, outReady[kPipeDepth] = {}, outFree[kPipeDepth] = {};
    /* decode working set: slot 0 serves direct calls, slots 1-2 the two chunks the host pipeline keeps
     * in flight on its two compute streams */
    struct Work {
        DeviceBuf<uint4> scratch;    /* survivor words of the wave in flight */
        DeviceBuf<uint4> schedState; /* FwdSched.state */
        DeviceBuf<int> schedFlags;   /* [0] unit counter, [1 + g] FwdSched.done */
        /* k7FusedKernel (decode_fused.cuh): decision rings, pass-to-pass state + flags, and the dense buffers the
         * frames it hands back are gathered into */
        DeviceBuf<uint4> ring;
        DeviceBuf<int> fusedAux;
        DeviceBuf<int> wsAux;        /* k7FusedWsKernel: pass-done flags and per-pass start / arrival states */
        DeviceBuf<uint8_t> gatherIn, gatherOut;
        cudaEvent_t idle = nullptr;  /* recorded after the last kernel that used this working set */
        cudaStream_t lastStream = nullptr;
    } work[1 + kPipeDepth];
    cudaStream_t pipe[kPipeDepth] = {}; /* compute streams of the host pipeline and of the wave pipeline below */
    cudaEvent_t waveFork = nullptr, waveJoin[kPipeDepth] = {}; /* ced_decode_batch: waves of a large batch in flight */
    /* host-side transfer compression (host_pack.cpp): pinned packed staging + worker threads */
    ced_host::Packer *packer = nullptr;
    PinnedBuf packStage[kPipeDepth];
    PinnedBuf outStage[kPipeDepth];  /* results on their way to a pageable caller buffer */
    cudaEvent_t stageFree[kPipeDepth] = {};
    int packHoldoff = 0, packPenalty = 1; /* adaptive transfer compression: back-off after the link ran dry */
    int fwdBlocks = 0;               /* persistent grid of k7ForwardKernel (0 = adaptive) */
    int sms = 0, fwdResident = 0;
    size_t maxWaveFrames = 0;        /* frames per wave cap (CED_MAX_WAVE_FRAMES overrides, for tests) */
    DeviceBuf<uint8_t> hostIn[kPipeDepth], hostOut[kPipeDepth];
    /* streaming path */
    DeviceBuf<uint8_t> sIn, sOut;
    DeviceBuf<uint32_t> sSurv;
    DeviceBuf<uint8_t> sParallel;    /* frame_parallel.cuh scratch (one-shot K=7 packets) */
    DeviceBuf<uint8_t> sSplit;       /* warp_split.cu scratch of the same calls */
    DeviceBuf<unsigned int> sSplitSeq;   /* its doorbell: device counter ... */
    PinnedBuf sDoorbell;             /* ... and the host-visible word the join kernel writes the count to */
    unsigned int splitSeq = 0;       /* count the host expects next */
    uint8_t fpGraphTables[192] = {}; /* edge labels + start metrics baked into a graph of the split kernels */
    bool fpGraphSplit = false, fpGraphDoorbell = false;
    cudaGraphExec_t fpGraph = nullptr; /* copy in + fpBlockKernel + fpSelectKernel for packets of fpGraphSegs segments */
    int fpGraphSegs = 0;
    const void *fpGraphKey[5] = {};  /* staging pointers baked into fpGraph */
    PinnedBuf sPinIn, sPinOut;
    /* resident packet decoder (frame_server.cuh): mailbox in mapped page-locked memory, control words on the device */
    void *fsMailbox = nullptr, *fsMailboxDev = nullptr, *fsCtl = nullptr;
    cudaStream_t fsStream = nullptr;
    uint32_t fsSeq = 0;
    bool fsLaunched = false, fsDisabled = false;
    uint64_t fsRequests = 0, fsLaunches = 0;
    std::recursive_mutex mu;
    uint64_t launches = 0;
    ced::BmTable bm0113, bm0133, bm0113s, bm0133s;   /* ...s: the same code with its generators exchanged */
    /* step tables of run-time K=7 codes (ced::buildStepTable), built on first use and kept */
    struct StepTable {
        int n;
        uint64_t g[3];
        uint2 *dev;
    };
    std::vector<StepTable> stepTables;
    DeviceBuf<uint4> softqTable[2];  /* softq_decode.cuh cost tables of 0113/0171 and 0133/0171, built on first use */
    /* optional kernel timing (ced_ctx_set_profiling) */
    bool profiling = false;
    static constexpr int kMaxProfWaves = 64;
    cudaEvent_t prof[kMaxProfWaves][3] = {};
    int profWaves = 0;
    const int *lastFusedCount = nullptr; /* device counter of the frames the last fused decode handed to the two-kernel path */
    bool counted = false;            /* this context is in the per-device census (activeGpus) */
};

using Code0113 = ced::K7Code<0113, 0171>;
using Code0133 = ced::K7Code<0133, 0171>;
static_assert(Code0113::symmetric && Code0133::symmetric, "SWAR butterflies need symmetric generators");
static_assert(Code0113::tap0 == 0x69 && Code0113::tap1 == 0x4F, "SURVEY 8(c) KAT: taps of 0113/0171");

/* Frames per wave and the working set of one wave of the SWAR decoder: the same computation sizes the
 * buffers in decodeBatchImpl and answers ced_decode_scratch_bytes. */
struct DecodeWorkingSet {
    size_t waveMax, firstGroups, perFrame, scratchBytes, stateBytes, flagBytes;
};
inline size_t maxWaveFramesSetting()
{
    const char *waveEnv = getenv("CED_MAX_WAVE_FRAMES"); /* tests force several waves with it */
    return (waveEnv && atoll(waveEnv) >= 64) ? (size_t)atoll(waveEnv) / 64 * 64 : kMaxWaveFrames;
}
inline DecodeWorkingSet decodeWorkingSet(size_t nFrames, int T, size_t maxWaveFrames)
{
    DecodeWorkingSet w;
    w.perFrame = (size_t)(T / 2) * sizeof(uint4);
    w.waveMax = std::min<size_t>(maxWaveFrames, std::max<size_t>(64, kMaxScratchBytes / w.perFrame)) / 64 * 64;
    w.firstGroups = (std::min<size_t>(nFrames, w.waveMax) + 31) / 32;
    w.scratchBytes = w.firstGroups * 32 * w.perFrame;
    w.stateBytes = w.firstGroups * 4 * 32 * sizeof(uint4);
    w.flagBytes = (w.firstGroups + 1) * sizeof(int);
    return w;
}


/* ced_abi.cu: asks the resident packet decoder of the default context (if one is running) to leave; batch entry
 * points call it so that their kernels and allocations do not wait for its idle time-out */
extern "C" void cedStopPacketServer();

/* swar_generic.cu: any k = 1, n = 2 / 3 code with 4 .. 256 states on the table-driven SIMD-in-word kernels;
 * CED_ERR_UNSUPPORTED = not a code these kernels take */
int cedDecodeBatchSwarGeneric(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                              int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot);
/* softq_decode.cu: ced_decode_batch_softq on working-set `slot` (the host pipeline keeps several chunks in flight) */
int cedDecodeBatchSoftQ(ced_ctx *c, const ced_code_t *code, const uint8_t *dSyms, size_t symStride, int nFrames,
                        int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot);

int cedSoftQForwardWindow(ced_ctx *c, const ced_code_t *code, bool aligned16, int blocks, cudaStream_t s, const uint8_t *in,
                          size_t symStride, int wave, int nSegments, uint4 *scratch, ced::FwdSched sched, ced::FwdWindow win);

/* warp_frame.cu: small batches, one warp per frame (k = 1 codes with <= 64 states, n <= 3, byte format);
 * CED_ERR_UNSUPPORTED = not a case for it */
int cedDecodeBatchWarpFrame(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                            int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot = 0, bool packed = false);
/* warp_split.cu: the same for so few frames of a 64-state rate-1/2 code that they are cut into blocks in time as well */
int cedDecodeBatchWarpSplit(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                            int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot, bool packed = false);
bool cedWarpFrameTakes(const ced_ctx *c, const ced_code_t *code, int nFrames, int frameBits, bool packed = false);

/* warp_split.cu: the one-packet call of the reference-named API on the same kernels (labels and start metrics are the caller's) */
int cedStreamDecodeSplit(ced_ctx *c, const uint8_t *edge, const uint8_t *metrics, const uint8_t *dSegs, int T, uint8_t *dOut,
                         void *scratch, size_t scratchBytes, cudaStream_t s, unsigned int *doneCounter,
                         volatile unsigned int *doneFlag);
size_t cedStreamDecodeSplitScratchBytes(int maxSteps);
bool cedStreamDecodeSplitTakes(const ced_ctx *c, int T);

/* swar_generic.cu: continuous streams for the table-driven kernels (K <= 7); CED_ERR_UNSUPPORTED / 0 = not their code */
size_t cedWindowCarryBytesGeneric(const ced_code_t *code, int nStreams, int depth);
int cedDecodeWindowGeneric(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nStreams, int nSegments,
                           uint64_t streamPos, int depth, int last, void *dCarry, uint8_t *dOut, size_t outStride, void *stream);

/* swar_generic.cu: rate-2/n codes on the radix-4 SIMD-in-word kernels (swar_radix4.cuh); same convention */
int cedDecodeBatchSwarRadix4(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                             int frameBits, uint8_t *dOut, size_t outStride, void *stream);
