/*
 * ced_abi.cu -- implementation of include/ced_abi.h: per-GPU context, launch
 * logic for the batched hot path and the stateless per-frame streaming path.
 * Everything here runs kernels; there is no host implementation of any
 * encode/decode arithmetic in this file.
 */
#include "ced_internal.cuh"
#include "channel_kernels.cuh"
#include "decode_batch.cuh"
#include "decode_fused.cuh"
#include "encode_batch.cuh"
#include "frame_parallel.cuh"
#include "frame_server.cuh"
#include "probe_kernels.cuh"
#include "stream_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

static std::atomic<ced_ctx *> gServerCtx{nullptr}; /* context that owns the resident packet decoder */
extern "C" {
static void fsStop(ced_ctx *c);
}

static thread_local char gLastError[512] = "";

void cedSetError(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(gLastError, sizeof(gLastError), fmt, ap);
    va_end(ap);
}


/* How many GPUs share this host's cores and memory right now: CED_ACTIVE_GPUS if set, else the ranks of a
 * one-process-per-GPU launch on this node (LOCAL_WORLD_SIZE, set by torchrun), else the devices this process
 * holds contexts on.  (Round 1 divided by cudaGetDeviceCount(): on an 8-GPU node a single rank then saw 4 cores
 * per GPU and never packed -- 53 Gbit/s where the 16-core 1-GPU box reached 76.) */
static std::mutex gActiveMu;
static int gCtxPerDevice[64] = {};

static int activeGpus()
{
    for (const char *name : {"CED_ACTIVE_GPUS", "LOCAL_WORLD_SIZE"}) {
        const char *e = getenv(name);
        if (e && atoi(e) > 0)
            return atoi(e);
    }
    std::lock_guard<std::mutex> lock(gActiveMu);
    int n = 0;
    for (int d = 0; d < 64; d++)
        n += gCtxPerDevice[d] > 0 ? 1 : 0;
    return std::max(1, n);
}

static int hostCoresPerActiveGpu()
{
    return std::max(1, (int)std::thread::hardware_concurrency() / activeGpus());
}

extern "C" {

static int ctxInit(ced_ctx *c, int device);

int ced_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

const char *ced_last_error(void)
{
    return gLastError;
}

int ced_ctx_create(int device, ced_ctx **out)
{
    if (!out)
        return CED_ERR_ARG;
    *out = nullptr;
    int n = 0;
    CED_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        setError("device %d out of range (%d visible)", device, n);
        return CED_ERR_ARG;
    }
    CED_CUDA(cudaSetDevice(device));
    ced_ctx *c = new ced_ctx();
    c->device = device;
    const int rc = ctxInit(c, device);
    if (rc != CED_OK) {
        char keep[512];
        snprintf(keep, sizeof(keep), "%s", ced_last_error());
        ced_ctx_destroy(c);   /* releases whatever was created so far */
        setError("%s", keep);
        return rc;
    }
    {
        std::lock_guard<std::mutex> lock(gActiveMu);
        gCtxPerDevice[device & 63]++;
    }
    c->counted = true;
    *out = c;
    return CED_OK;
}

static int ctxInit(ced_ctx *c, int device)
{
    c->bm0113 = ced::makeBmTable<Code0113>();
    c->bm0133 = ced::makeBmTable<Code0133>();
    c->bm0113s = c->bm0113;
    c->bm0133s = c->bm0133;
    for (ced::BmTable *t : {&c->bm0113s, &c->bm0133s})   /* generators in the other order: received symbols 01 <-> 10 */
        for (int ph = 0; ph < 6; ph++)
            for (int i = 0; i < 2; i++)
                std::swap(t->x[(ph * 4 + 1) * 2 + i], t->x[(ph * 4 + 2) * 2 + i]);
    CED_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CED_CUDA(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CED_CUDA(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < kPipeDepth; i++) {
        CED_CUDA(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
        CED_CUDA(cudaEventCreateWithFlags(&c->inReady[i], cudaEventDisableTiming));
        CED_CUDA(cudaEventCreateWithFlags(&c->inFree[i], cudaEventDisableTiming));
        CED_CUDA(cudaEventCreateWithFlags(&c->outReady[i], cudaEventDisableTiming));
        CED_CUDA(cudaEventCreateWithFlags(&c->outFree[i], cudaEventDisableTiming));
    }
    for (int w = 0; w < ced_ctx::kMaxProfWaves; w++)
        for (int e = 0; e < 3; e++)
            CED_CUDA(cudaEventCreate(&c->prof[w][e]));
    for (auto &w : c->work)
        CED_CUDA(cudaEventCreateWithFlags(&w.idle, cudaEventDisableTiming));
    CED_CUDA(cudaEventCreateWithFlags(&c->waveFork, cudaEventDisableTiming));
    for (int i = 0; i < kPipeDepth; i++)
        CED_CUDA(cudaEventCreateWithFlags(&c->waveJoin[i], cudaEventDisableTiming));
    for (int i = 0; i < kPipeDepth; i++)
        CED_CUDA(cudaEventCreateWithFlags(&c->stageFree[i], cudaEventDisableTiming));
    {
        /* persistent forward grid: CED_FWD_BLOCKS_PER_SM CTAs of 4 warps per SM (default 4 = 4 warps per
         * sub-partition), never more than the kernel's resident capacity */
        int sms = 0, resident = 0;
        CED_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        c->maxWaveFrames = maxWaveFramesSetting();
        CED_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &resident, ced::k7ForwardKernel<Code0113, ced::PackedSymbols, false>, ced::kFwdThreads, 0));
        const char *env = getenv("CED_FWD_BLOCKS_PER_SM");
        c->sms = sms;
        c->fwdResident = std::max(1, resident);
        c->fwdBlocks = env ? sms * std::max(1, std::min(atoi(env), c->fwdResident)) : 0;
    }
    return CED_OK;
}

void ced_ctx_destroy(ced_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (gServerCtx.load() == c)
        gServerCtx.store(nullptr);
    fsStop(c);
    cudaDeviceSynchronize();
    if (c->fsMailbox)
        cudaFreeHost(c->fsMailbox);
    if (c->fsCtl)
        cudaFree(c->fsCtl);
    if (c->fsStream)
        cudaStreamDestroy(c->fsStream);
    for (auto &w : c->work) {
        if (w.idle)
            cudaEventDestroy(w.idle);
        w.scratch.release();
        w.schedState.release();
        w.schedFlags.release();
        w.ring.release();
        w.fusedAux.release();
        w.wsAux.release();
        w.gatherIn.release();
        w.gatherOut.release();
    }
    for (int i = 0; i < kPipeDepth; i++) {
        if (c->pipe[i])
            cudaStreamDestroy(c->pipe[i]);
        if (c->stageFree[i])
            cudaEventDestroy(c->stageFree[i]);
        c->packStage[i].release();
        c->outStage[i].release();
    }
    if (c->waveFork)
        cudaEventDestroy(c->waveFork);
    for (int i = 0; i < kPipeDepth; i++)
        if (c->waveJoin[i])
            cudaEventDestroy(c->waveJoin[i]);
    if (c->packer)
        ced_host::packerDestroy(c->packer);
    for (int i = 0; i < kPipeDepth; i++) {
        c->hostIn[i].release();
        c->hostOut[i].release();
        for (cudaEvent_t e : {c->inReady[i], c->inFree[i], c->outReady[i], c->outFree[i]})
            if (e)
                cudaEventDestroy(e);
    }
    for (auto &t : c->stepTables)
        cudaFree(t.dev);
    c->stepTables.clear();
    c->sIn.release();
    c->sOut.release();
    c->sSurv.release();
    c->sParallel.release();
    c->sSplit.release();
    c->sSplitSeq.release();
    c->sDoorbell.release();
    if (c->fpGraph)
        cudaGraphExecDestroy(c->fpGraph);
    c->sPinIn.release();
    c->sPinOut.release();
    for (int w = 0; w < ced_ctx::kMaxProfWaves; w++)
        for (int e = 0; e < 3; e++)
            if (c->prof[w][e])
                cudaEventDestroy(c->prof[w][e]);
    for (cudaStream_t st : {c->stream, c->h2d, c->d2h})
        if (st)
            cudaStreamDestroy(st);
    if (c->counted) {
        std::lock_guard<std::mutex> lock(gActiveMu);
        gCtxPerDevice[c->device & 63]--;
    }
    cudaGetLastError();
    delete c;
}

int ced_ctx_device(const ced_ctx *c)
{
    return c ? c->device : -1;
}

ced_ctx *ced_default_ctx(void)
{
    static std::once_flag once;
    static ced_ctx *ctx = nullptr;
    std::call_once(once, [] {
        const char *env = getenv("CED_DEVICE");
        int dev = env ? atoi(env) : 0;
        if (ced_ctx_create(dev, &ctx) != CED_OK)
            ctx = nullptr;
    });
    return ctx;
}

int ced_sync(ced_ctx *c, void *stream)
{
    if (!c)
        return CED_ERR_ARG;
    CED_CUDA(cudaSetDevice(c->device));
    CED_CUDA(cudaStreamSynchronize(stream ? (cudaStream_t)stream : c->stream));
    return CED_OK;
}

uint64_t ced_launch_count(const ced_ctx *c)
{
    return c ? c->launches : 0;
}

int ced_host_alloc(size_t bytes, void **out)
{
    if (!out || bytes == 0) {
        setError("ced_host_alloc: bad argument");
        return CED_ERR_ARG;
    }
    ced_ctx *c = ced_default_ctx();   /* makes sure a device is selected */
    if (!c)
        return CED_ERR_CUDA;
    CED_CUDA(cudaSetDevice(c->device));
    cudaError_t e = cudaMallocHost(out, bytes);
    if (e != cudaSuccess) {
        setError("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return CED_ERR_NOMEM;
    }
    return CED_OK;
}

void ced_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

int ced_host_register(void *p, size_t bytes)
{
    if (!p || bytes == 0) {
        setError("ced_host_register: bad argument");
        return CED_ERR_ARG;
    }
    ced_ctx *c = ced_default_ctx();
    if (!c)
        return CED_ERR_CUDA;
    CED_CUDA(cudaSetDevice(c->device));
    CED_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return CED_OK;
}

int ced_host_unregister(void *p)
{
    if (!p)
        return CED_ERR_ARG;
    CED_CUDA(cudaHostUnregister(p));
    return CED_OK;
}

size_t ced_decode_scratch_bytes(int nFrames, int frameBits)
{
    if (nFrames <= 0 || frameBits <= 0)
        return 0;
    const DecodeWorkingSet ws = decodeWorkingSet((size_t)nFrames, frameBits + ced::kTailSteps, maxWaveFramesSetting());
    return ws.scratchBytes + ws.stateBytes + ws.flagBytes;
}

/* device copy of the step table of a run-time K=7 code (caller holds c->mu, device is current) */
static int stepTableFor(ced_ctx *c, const ced_code_t *code, const uint2 **out)
{
    const int n = code->codedBits;
    const uint64_t g2 = n > 2 ? code->gen[2] : 0;
    for (const auto &t : c->stepTables)
        if (t.n == n && t.g[0] == code->gen[0] && t.g[1] == code->gen[1] && t.g[2] == g2) {
            *out = t.dev;
            return CED_OK;
        }
    const uint32_t gens[3] = {(uint32_t)code->gen[0], (uint32_t)code->gen[1], (uint32_t)g2};
    std::vector<ced::Word2> host((size_t)6 * 16 * (1u << n));
    ced::buildStepTable(ced::makeK7Taps(n, gens), host.data());
    uint2 *dev = nullptr;
    CED_CUDA(cudaMalloc(&dev, host.size() * sizeof(ced::Word2)));
    CED_CUDA(cudaMemcpy(dev, host.data(), host.size() * sizeof(ced::Word2), cudaMemcpyHostToDevice));
    c->stepTables.push_back({n, {code->gen[0], code->gen[1], g2}, dev});
    *out = dev;
    return CED_OK;
}

/* Any k=1 code with K <= 9, n <= 8: one warp / CTA per frame (genericBatchDecodeKernel). */
static int decodeBatchGeneric(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                              int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot)
{
    if (!code || code->constraintLen < 2 || code->constraintLen > 9 || code->codedBits < 1 ||
        code->codedBits > CED_MAX_N) {
        setError("ced_decode_batch: K must be 2..9 and n 1..8");
        return CED_ERR_UNSUPPORTED;
    }
    const int K = code->constraintLen, n = code->codedBits, N = 1 << (K - 1), W = ced_stream_surv_words(N);
    const int T = frameBits + K - 1;
    if (segStride < (size_t)T || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    ced_ctx::Work &wk = c->work[slot];
    /* trellis labels and start metrics (table building, as viterbiInit does on the host) */
    uint8_t table[1024];
    uint32_t taps[CED_MAX_N];
    for (int i = 0; i < n; i++)
        taps[i] = reverseBits(code->gen[i], K);
    for (int b = 0; b < 2; b++)
        for (int st = 0; st < N; st++) {
            const uint32_t reg = (((uint32_t)st << 1) | (uint32_t)b) & ((1u << K) - 1u);
            uint8_t seg = 0;
            for (int i = 0; i < n; i++)
                seg |= (uint8_t)((__builtin_popcount(reg & taps[i]) & 1) << i);
            table[b * N + st] = seg;
        }
    for (int st = 0; st < N; st++)
        table[512 + st] = st == 0 ? 0 : (uint8_t)(N + 1);
    const size_t perFrame = (size_t)T * W * sizeof(uint32_t);
    size_t waveMax = std::min<size_t>(c->maxWaveFrames, std::max<size_t>(1, kMaxScratchBytes / perFrame));
    const size_t firstWave = std::min<size_t>((size_t)nFrames, waveMax);
    const int gridMax = c->sms * (N <= 64 ? 32 : 16);   /* one-warp CTAs: 32 per SM (6 KB of shared memory each) */
    const size_t sinkBytes = 1024 + (size_t)gridMax * (256 + 4);
    if (wk.scratch.bytes < firstWave * perFrame || wk.schedState.bytes < sinkBytes) {
        CED_CUDA(cudaDeviceSynchronize());
        int rc = wk.scratch.ensure(firstWave * perFrame);
        if (rc == CED_OK) rc = wk.schedState.ensure(sinkBytes);
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    uint8_t *dTable = reinterpret_cast<uint8_t *>(wk.schedState.p);
    CED_CUDA(cudaMemcpyAsync(dTable, table, 1024, cudaMemcpyHostToDevice, s)); /* pageable source: staged by the driver */
    for (size_t f0 = 0; f0 < (size_t)nFrames; f0 += waveMax) {
        const int wave = (int)std::min<size_t>(waveMax, (size_t)nFrames - f0);
        ced::GenericBatchArgs g;
        g.K = K;
        g.n = n;
        g.N = N;
        g.W = W;
        g.nFrames = wave;
        g.T = T;
        g.edge = dTable;
        g.initMetrics = dTable + 512;
        g.segs = dSegs + f0 * segStride;
        g.segStride = segStride;
        g.surv = reinterpret_cast<uint32_t *>(wk.scratch.p);
        g.out = dOut + f0 * outStride;
        g.outStride = outStride;
        g.metricsSink = dTable + 1024;
        g.stateSink = reinterpret_cast<uint32_t *>(dTable + 1024 + (size_t)gridMax * 256);
        const int blocks = std::min(wave, gridMax);
        if (N <= 64)
            ced::genericBatchDecodeKernel<true><<<blocks, 32, 0, s>>>(g);
        else
            ced::genericBatchDecodeKernel<false><<<blocks, N / 2, 0, s>>>(g);
        c->launches += 1;
    }
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

static int decodeBatchImpl(ced_ctx *c, const ced_code_t *code, bool packed, const uint8_t *dSegs, size_t segStride,
                           int nFrames, int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot = 0)
{
    cedStopPacketServer();
    if (!c || nFrames < 0 || frameBits <= 0 || (frameBits & 7) || (nFrames > 0 && (!dSegs || !dOut))) {
        setError("ced_decode_batch: bad argument (frameBits must be a positive multiple of 8)");
        return CED_ERR_ARG;
    }
    /* small batches: one warp per frame, decisions in shared memory (warp_frame.cu) */
    const CodeId idSmall = classify(code);   /* the packed format keeps the code family it has at any batch size */
    const bool packedOk = idSmall == CodeId::K7_0113_0171 || idSmall == CodeId::K7_0133_0171 || idSmall == CodeId::K7_Runtime;
    if ((!packed || packedOk) && cedWarpFrameTakes(c, code, nFrames, frameBits, packed)) {
        const int rc = cedDecodeBatchWarpFrame(c, code, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream, slot, packed);
        if (rc != CED_ERR_UNSUPPORTED)
            return rc;
    }
    CodeId id = classify(code);
    /* the compiled-in codes with their two generators written the other way round ((0171, 0133) is how the NASA standard
     * code is often given): the same kernel with the rows of its branch-cost table for the received symbols 01 and 10
     * exchanged -- HD(rx, swap(label)) = HD(swap(rx), label) -- instead of the step-table kernel (167 vs 151 Gbit/s) */
    bool swapped = false;
    if (id == CodeId::K7_Runtime && code->gen[0] == 0171 && (code->gen[1] == 0113 || code->gen[1] == 0133)) {
        id = code->gen[1] == 0113 ? CodeId::K7_0113_0171 : CodeId::K7_0133_0171;
        swapped = true;
    }
    if (id == CodeId::Unsupported || (packed && id == CodeId::K7_RuntimeN3)) {
        if (packed) {
            setError("ced_decode_batch_packed: 2-bit packing is for K=7 n=2 codes whose generators tap both ends");
            return CED_ERR_UNSUPPORTED;
        }
        /* thread-per-frame SIMD-in-word kernels driven by a step table (swar_generic.cu): K = 3, 4, 5, 7, 9 with 2 or 3
         * generators of any shape; what they do not take goes to the one-warp-per-frame kernel */
        const int rc = cedDecodeBatchSwarGeneric(c, code, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream, slot);
        if (rc != CED_ERR_UNSUPPORTED)
            return rc;
        return decodeBatchGeneric(c, code, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream, slot);
    }
    const int T = frameBits + ced::kTailSteps;
    const size_t rowBytes = packed ? (size_t)(T + 3) / 4 : (size_t)T;
    if (segStride < rowBytes || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const uint2 *stepTable = nullptr;
    if (id == CodeId::K7_Runtime || id == CodeId::K7_RuntimeN3) {
        int rc = stepTableFor(c, code, &stepTable);
        if (rc != CED_OK)
            return rc;
    }
    const DecodeWorkingSet ws = decodeWorkingSet((size_t)nFrames, T, c->maxWaveFrames);
    const size_t waveMax = ws.waveMax;
    ced_ctx::Work &wk = c->work[slot];
    if (wk.scratch.bytes < ws.scratchBytes || wk.schedState.bytes < ws.stateBytes || wk.schedFlags.bytes < ws.flagBytes) {
        /* growing means freeing: make sure nothing still uses the old blocks */
        CED_CUDA(cudaDeviceSynchronize());
        int rc = wk.scratch.ensure(ws.scratchBytes);
        if (rc == CED_OK) rc = wk.schedState.ensure(ws.stateBytes);
        if (rc == CED_OK) rc = wk.schedFlags.ensure(ws.flagBytes);
        if (rc != CED_OK)
            return rc;
    }
    /* CED_FUSED=1 / 2: compile-time codes and batches that fill the GPU take the fused kernel (decode_fused.cuh:
     * traceback inside the forward warps / on its own warp).  Off by default: measured 1.58 / 1.85 ms against 1.65 ms
     * for the two kernels one call at a time and 1.355 ms with three calls in flight */
    const int envFused = getenv("CED_FUSED") ? atoi(getenv("CED_FUSED")) : 0;   /* read per call: tests flip them */
    const int envFusedMin = getenv("CED_FUSED_MIN_FRAMES") ? atoi(getenv("CED_FUSED_MIN_FRAMES")) : 16384;
    const bool fused = envFused != 0 && !swapped && (id == CodeId::K7_0113_0171 || id == CodeId::K7_0133_0171) &&
                       nFrames >= envFusedMin;
    if (fused) {
        const size_t wave0 = std::min<size_t>((size_t)nFrames, waveMax), g0 = (wave0 + 31) / 32;
        const size_t ringBytes = std::min<size_t>(g0, 2 * ced::kCohortGroups) * ced::WsGeom<384, 96>::kRingPairs * 32 * sizeof(uint4);
        const size_t wsPasses = (size_t)(T + 95) / 96;   /* at most one pass per 96 steps */
        const size_t wsBytes = g0 * wsPasses * (sizeof(int) + 64);
        const size_t auxBytes = (2 * (g0 + 1) + 1 + 2 * wave0 + g0 * 32) * sizeof(int);
        const size_t gInBytes = wave0 * ((rowBytes + 15) / 16 * 16), gOutBytes = wave0 * (size_t)(frameBits / 8);
        if (wk.ring.bytes < ringBytes || wk.fusedAux.bytes < auxBytes || wk.gatherIn.bytes < gInBytes ||
            wk.gatherOut.bytes < gOutBytes || wk.wsAux.bytes < wsBytes) {
            CED_CUDA(cudaDeviceSynchronize());
            int rc = wk.ring.ensure(ringBytes);
            if (rc == CED_OK) rc = wk.wsAux.ensure(wsBytes);
            if (rc == CED_OK) rc = wk.fusedAux.ensure(auxBytes);
            if (rc == CED_OK) rc = wk.gatherIn.ensure(gInBytes);
            if (rc == CED_OK) rc = wk.gatherOut.ensure(gOutBytes);
            if (rc != CED_OK)
                return rc;
        }
    }
    /* the working set is shared by all calls on this context: a call on another stream first waits for
     * the previous one (use one context per stream to keep several decodes in flight) */
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    const bool aligned16 = (reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0;
    c->profWaves = 0;
    for (size_t f0 = 0; f0 < (size_t)nFrames; f0 += waveMax) {
        const bool prof = c->profiling && c->profWaves < ced_ctx::kMaxProfWaves;
        const int pw = c->profWaves;
        const int wave = (int)std::min<size_t>(waveMax, (size_t)nFrames - f0);
        const int groups = (wave + 31) / 32;
        /* persistent grid: as many 4-warp CTAs per SM as there are 32-frame groups per sub-partition,
         * between 3 (2^16 frames: 3.46 groups per SMSP; a 4th warp would mostly spin) and 5 (measured:
         * 2^18 frames 177.8 / 184.7 / 185.8 Gbit/s at 3 / 4 / 5), never more warps than groups */
        static const int envGrid = getenv("CED_FWD_GRID") ? atoi(getenv("CED_FWD_GRID")) : 0; /* experiments */
        int gridBlocks = envGrid > 0 ? envGrid : c->fwdBlocks;
        if (gridBlocks == 0)
            gridBlocks = c->sms * std::max(3, std::min({5, c->fwdResident, groups / (4 * c->sms)}));
        const int blocks = std::max(1, std::min(gridBlocks, (groups + 3) / 4));
        const uint8_t *in = dSegs + f0 * segStride;
        uint8_t *out = dOut + f0 * outStride;
        ced::FwdSched sched;
        sched.counter = reinterpret_cast<unsigned int *>(wk.schedFlags.p);
        sched.done = wk.schedFlags.p + 1;
        sched.state = wk.schedState.p;
        CED_CUDA(cudaMemsetAsync(wk.schedFlags.p, 0, (size_t)(groups + 1) * sizeof(int), s));
        const ced::BmTable &bm = (id == CodeId::K7_0113_0171) ? (swapped ? c->bm0113s : c->bm0113) : (swapped ? c->bm0133s : c->bm0133);
        static const int envCpu = getenv("CED_FWD_CHUNKS_PER_UNIT") ? atoi(getenv("CED_FWD_CHUNKS_PER_UNIT")) : 0;
        const int cpu = envCpu > 0 ? envCpu : 2; /* chunks a warp runs before handing its group on: 1 / 2 / 4 / 8 -> 1.249 / 1.217 / 1.225 / 1.262 ms */
        if (fused) {
            /* one kernel: forward ACS with the traceback inside (decode_fused.cuh); the frames it hands back (failed
             * pass check; none at useful noise levels) are gathered, decoded by the two-kernel path and scattered */
            const size_t gStride = (rowBytes + 15) / 16 * 16, outBytes = (size_t)frameBits / 8;
            int *aux = wk.fusedAux.p;                       /* [counterA doneA[groups]] [counterB doneB[groups]] [count] [flag[wave]] */
            int *auxB = aux + groups + 1, *count = auxB + groups + 1;
            unsigned int *flag = reinterpret_cast<unsigned int *>(count + 1);
            int *list = reinterpret_cast<int *>(flag) + wave;
            uint32_t *expect = reinterpret_cast<uint32_t *>(list + wave);
            CED_CUDA(cudaMemsetAsync(aux, 0, (size_t)(2 * (groups + 1) + 1 + wave) * sizeof(int), s));
            if (prof)
                CED_CUDA(cudaEventRecord(c->prof[pw][0], s));
            ced::FwdSched schedA = {reinterpret_cast<unsigned int *>(aux), aux + 1, wk.schedState.p};
            ced::FwdSched schedB = {reinterpret_cast<unsigned int *>(auxB), auxB + 1, wk.schedState.p};
            bool wsLaunched = false;
            if (envFused == 2 && id == CodeId::K7_0113_0171 && !packed && aligned16) {
                /* warp-specialised form (k7FusedWsKernel): traceback on its own warp, ring streamed with bulk copies */
                int wsE = 192, wsD = 72;
                if (const char *ge = getenv("CED_FUSED_GEOM"))
                    sscanf(ge, "%d,%d", &wsE, &wsD);
                const int passes = (T + wsE - 1) / wsE;
                ced::WsArgs wa;
                wa.ring = wk.ring.p;
                wa.passes = passes;
                wa.out = out;
                wa.outStride = outStride;
                wa.ringSlots = std::min(groups, 2 * ced::kCohortGroups);
                uint8_t *wsAux = reinterpret_cast<uint8_t *>(wk.wsAux.p);
                wa.passDone = reinterpret_cast<int *>(wsAux);
                wa.startState = wsAux + (size_t)groups * passes * sizeof(int);
                wa.arriveState = wa.startState + (size_t)groups * passes * 32;
                CED_CUDA(cudaMemsetAsync(wa.passDone, 0, (size_t)groups * passes * sizeof(int), s));
                static const int envWsPerSm = getenv("CED_FUSED_BLOCKS_PER_SM") ? atoi(getenv("CED_FUSED_BLOCKS_PER_SM")) : 0;
                const int wsBlocks = std::max(1, std::min(c->sms * (envWsPerSm > 0 ? envWsPerSm : 3), (groups + 3) / 4));
#define CED_WS_CASE(E_, D_)                                                                                          \
    if (wsE == E_ && wsD == D_) {                                                                                    \
        ced::k7FusedWsKernel<Code0113, ced::ByteSymbols, true, ced::WsGeom<E_, D_>><<<wsBlocks, ced::kWsThreads, 0, s>>>( \
            in, segStride, wave, T, bm, schedA, cpu, wa);                                                            \
        wsLaunched = true;                                                                                           \
    }
                CED_WS_CASE(96, 72)
                CED_WS_CASE(192, 72)
                CED_WS_CASE(384, 96)
#undef CED_WS_CASE
                if (wsLaunched) {
                    ced::fusedVerifyKernel<<<(wave + 255) / 256, 256, 0, s>>>(wa.startState, wa.arriveState, wave, passes, wsE,
                                                                             wsD, list, count);
                    c->launches += 1;
                }
            }
            if (!wsLaunched) {
            ced::FusedArgs fa;
            fa.ring = wk.ring.p;
            fa.expect = expect;
            fa.flag = flag;
            fa.list = list;
            fa.count = count;
            fa.out = out;
            fa.outStride = outStride;
            fa.ringSlots = std::min(groups, 2 * ced::kCohortGroups);
            static const int envFusedPerSm = getenv("CED_FUSED_BLOCKS_PER_SM") ? atoi(getenv("CED_FUSED_BLOCKS_PER_SM")) : 0;
            const int perSm = envFusedPerSm > 0 ? envFusedPerSm : std::max(3, std::min(4, groups / (4 * c->sms)));
            const int fBlocks = std::max(1, std::min(c->sms * perSm, (groups + 3) / 4));
#define CED_LAUNCH_FUSED_ONE(...)                                                                                  \
    ced::k7FusedKernel<__VA_ARGS__><<<fBlocks, ced::kFwdThreads, 0, s>>>(in, segStride, wave, T, bm, schedA, cpu, fa)
#define CED_LAUNCH_FUSED(CODE, FMT)                                                                                \
    do {                                                                                                           \
        if (aligned16)                                                                                             \
            CED_LAUNCH_FUSED_ONE(CODE, ced::FMT, true);                                                            \
        else                                                                                                       \
            CED_LAUNCH_FUSED_ONE(CODE, ced::FMT, false);                                                           \
    } while (0)
            /* experiments: other ring geometries for the default code, byte format, aligned rows (CED_FUSED_GEOM=E,D) */
            int geoE = 0, geoD = 0;
            if (const char *ge = getenv("CED_FUSED_GEOM"))
                sscanf(ge, "%d,%d", &geoE, &geoD);
            /* keep the decision rings in L2: a persisting access-policy window over them, so that the symbol stream
             * (269 MB per decode, read once) does not evict them (CED_FUSED_PERSIST=0 turns it off) */
            static const int envPersist = getenv("CED_FUSED_PERSIST") ? atoi(getenv("CED_FUSED_PERSIST")) : 1;
            bool windowSet = false;
            if (envPersist) {
                int maxPersist = 0, maxWin = 0;
                cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
                cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
                const int pairs = geoE > 0 ? (geoE + geoD) / 2 : ced::kRingPairs;
                const size_t ringUsed = (size_t)fa.ringSlots * pairs * 32 * sizeof(uint4);
                if (maxPersist > 0 && maxWin > 0) {
                    static bool limitSet = false;
                    if (!limitSet) {
                        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxPersist);
                        limitSet = true;
                        if (getenv("CED_FUSED_DEBUG"))
                            fprintf(stderr, "fused: persisting L2 up to %.1f MB, window up to %.1f MB, rings %.1f MB\n",
                                    maxPersist / 1e6, maxWin / 1e6, ringUsed / 1e6);
                    }
                    cudaStreamAttrValue v = {};
                    v.accessPolicyWindow.base_ptr = wk.ring.p;
                    v.accessPolicyWindow.num_bytes = std::min<size_t>(ringUsed, (size_t)maxWin);
                    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)maxPersist / (double)ringUsed);
                    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    windowSet = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v) == cudaSuccess;
                    cudaGetLastError();
                }
            }
#define CED_FUSED_GEOM_CASE(E_, D_)                                                                                \
    if (id == CodeId::K7_0113_0171 && !packed && aligned16 && geoE == E_ && geoD == D_)                            \
        CED_LAUNCH_FUSED_ONE(Code0113, ced::ByteSymbols, true, ced::FusedGeom<E_, D_>);                            \
    else
            CED_FUSED_GEOM_CASE(96, 48)
            CED_FUSED_GEOM_CASE(192, 72)
            CED_FUSED_GEOM_CASE(192, 96)
            CED_FUSED_GEOM_CASE(384, 96)
#undef CED_FUSED_GEOM_CASE
            if (id == CodeId::K7_0113_0171 && !packed)
                CED_LAUNCH_FUSED(Code0113, ByteSymbols);
            else if (id == CodeId::K7_0113_0171)
                CED_LAUNCH_FUSED(Code0113, PackedSymbols);
            else if (!packed)
                CED_LAUNCH_FUSED(Code0133, ByteSymbols);
            else
                CED_LAUNCH_FUSED(Code0133, PackedSymbols);
#undef CED_LAUNCH_FUSED
#undef CED_LAUNCH_FUSED_ONE
            if (windowSet) { /* the kernels that follow on this stream are not to inherit the window */
                cudaStreamAttrValue v = {};
                v.accessPolicyWindow.num_bytes = 0;
                cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v);
                cudaGetLastError();
            }
            } /* !wsLaunched */
            if (prof)
                CED_CUDA(cudaEventRecord(c->prof[pw][1], s));
            const int auxGrid = c->sms * 4;
            ced::gatherRowsKernel<<<auxGrid, 256, 0, s>>>(in, segStride, list, count, wk.gatherIn.p, gStride, (int)rowBytes);
            if (id == CodeId::K7_0113_0171 && !packed)
                ced::k7ForwardKernel<Code0113, ced::ByteSymbols, true><<<blocks, ced::kFwdThreads, 0, s>>>(
                    wk.gatherIn.p, gStride, wave, T, wk.scratch.p, bm, schedB, cpu, ced::FwdWindow(), nullptr, count);
            else if (id == CodeId::K7_0113_0171)
                ced::k7ForwardKernel<Code0113, ced::PackedSymbols, true><<<blocks, ced::kFwdThreads, 0, s>>>(
                    wk.gatherIn.p, gStride, wave, T, wk.scratch.p, bm, schedB, cpu, ced::FwdWindow(), nullptr, count);
            else if (!packed)
                ced::k7ForwardKernel<Code0133, ced::ByteSymbols, true><<<blocks, ced::kFwdThreads, 0, s>>>(
                    wk.gatherIn.p, gStride, wave, T, wk.scratch.p, bm, schedB, cpu, ced::FwdWindow(), nullptr, count);
            else
                ced::k7ForwardKernel<Code0133, ced::PackedSymbols, true><<<blocks, ced::kFwdThreads, 0, s>>>(
                    wk.gatherIn.p, gStride, wave, T, wk.scratch.p, bm, schedB, cpu, ced::FwdWindow(), nullptr, count);
            ced::k7TracebackKernel<ced::Lanes8><<<(wave + ced::kTbThreads - 1) / ced::kTbThreads, ced::kTbThreads, 0, s>>>(
                wk.scratch.p, wave, T, wk.gatherOut.p, outBytes, nullptr, ced::kTailSteps, 0, count);
            ced::scatterRowsKernel<<<auxGrid, 256, 0, s>>>(wk.gatherOut.p, outBytes, list, count, out, outStride, (int)outBytes);
            if (prof) {
                CED_CUDA(cudaEventRecord(c->prof[pw][2], s));
                c->profWaves++;
            }
            c->launches += 5;
            c->lastFusedCount = count;
            continue;
        }
        if (prof)
            CED_CUDA(cudaEventRecord(c->prof[pw][0], s));
#define CED_LAUNCH_FWD(CODE, FMT)                                                                                  \
    do {                                                                                                           \
        if (aligned16)                                                                                             \
            ced::k7ForwardKernel<CODE, ced::FMT, true><<<blocks, ced::kFwdThreads, 0, s>>>(                         \
                in, segStride, wave, T, wk.scratch.p, bm, sched, cpu, ced::FwdWindow(), stepTable);                \
        else                                                                                                       \
            ced::k7ForwardKernel<CODE, ced::FMT, false><<<blocks, ced::kFwdThreads, 0, s>>>(                        \
                in, segStride, wave, T, wk.scratch.p, bm, sched, cpu, ced::FwdWindow(), stepTable);                \
    } while (0)
        if (id == CodeId::K7_RuntimeN3)
            CED_LAUNCH_FWD(ced::RuntimeK7<3>, ByteSymbols);
        else if (id == CodeId::K7_Runtime && !packed)
            CED_LAUNCH_FWD(ced::RuntimeK7<2>, ByteSymbols);
        else if (id == CodeId::K7_Runtime)
            CED_LAUNCH_FWD(ced::RuntimeK7<2>, PackedSymbols);
        else if (id == CodeId::K7_0113_0171 && !packed)
            CED_LAUNCH_FWD(Code0113, ByteSymbols);
        else if (id == CodeId::K7_0113_0171)
            CED_LAUNCH_FWD(Code0113, PackedSymbols);
        else if (!packed)
            CED_LAUNCH_FWD(Code0133, ByteSymbols);
        else
            CED_LAUNCH_FWD(Code0133, PackedSymbols);
#undef CED_LAUNCH_FWD
        if (prof)
            CED_CUDA(cudaEventRecord(c->prof[pw][1], s));
        ced::k7TracebackKernel<ced::Lanes8><<<(wave + ced::kTbThreads - 1) / ced::kTbThreads, ced::kTbThreads, 0, s>>>(
            wk.scratch.p, wave, T, out, outStride);
        if (prof) {
            CED_CUDA(cudaEventRecord(c->prof[pw][2], s));
            c->profWaves++;
        }
        c->launches += 2;
    }
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

/*
 * A batch of several GPU-fills is decoded as waves of 2^16 frames kept in flight on three internal streams, each with
 * its own working set: the forward (ACS) kernel is issue-bound and the traceback kernel HBM-bound, so the traceback
 * of wave w runs under the forward passes of waves w+1 and w+2 -- what bench.py used to arrange from outside with
 * three contexts (166 -> 198 Gbit/s) now happens inside ONE call.  The caller's stream forks into the
 * internal streams and joins them again, so the call keeps its stream semantics.  CED_WAVE_PIPELINE=0 turns it off.
 */
constexpr int kPipeWaveFrames = 1 << 16;
constexpr int kPipeWaves = 3;

enum class WireKind { Bytes, Packed, SoftQ };   /* hard symbols, 4 hard symbols per byte, 3-bit soft symbols */

static int decodeBatchPipelined(ced_ctx *c, const ced_code_t *code, WireKind kind, const uint8_t *dSegs, size_t segStride,
                                int nFrames, int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    auto one = [&](const uint8_t *in, int cnt, uint8_t *out, void *st, int slot) -> int {
        if (kind == WireKind::SoftQ)
            return cedDecodeBatchSoftQ(c, code, in, segStride, cnt, frameBits, out, outStride, st, slot);
        return decodeBatchImpl(c, code, kind == WireKind::Packed, in, segStride, cnt, frameBits, out, outStride, st, slot);
    };
    static const bool enabled = !getenv("CED_WAVE_PIPELINE") || atoi(getenv("CED_WAVE_PIPELINE")) != 0;
    /* experiments: CED_PIPE_WAVE_FRAMES (frames per wave), CED_PIPE_WAVES (waves in flight, <= kPipeDepth - 1) */
    static const int envWaveFrames = getenv("CED_PIPE_WAVE_FRAMES") ? atoi(getenv("CED_PIPE_WAVE_FRAMES")) : 0;
    static const int envWaves = getenv("CED_PIPE_WAVES") ? atoi(getenv("CED_PIPE_WAVES")) : 0;
    const int kPipeWaveFrames = envWaveFrames >= 4096 ? envWaveFrames / 64 * 64 : ::kPipeWaveFrames;
    const int kPipeWaves = envWaves >= 1 ? std::min(envWaves, kPipeDepth) : ::kPipeWaves;
    if (!c || !enabled || nFrames < 2 * kPipeWaveFrames || classify(code) == CodeId::Unsupported || !dSegs || !dOut ||
        frameBits <= 0 || (frameBits & 7) || frameBits > 8192)
        return one(dSegs, nFrames, dOut, stream, 0);
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CED_CUDA(cudaEventRecord(c->waveFork, s));
    for (int i = 0; i < kPipeWaves; i++)
        CED_CUDA(cudaStreamWaitEvent(c->pipe[i], c->waveFork, 0));
    int w = 0;
    for (long long f0 = 0; f0 < nFrames; f0 += kPipeWaveFrames, w++) {
        const int cnt = (int)std::min<long long>(kPipeWaveFrames, nFrames - f0);
        const int rc = one(dSegs + (size_t)f0 * segStride, cnt, dOut + (size_t)f0 * outStride, c->pipe[w % kPipeWaves],
                           1 + w % kPipeWaves);
        if (rc != CED_OK) {
            for (int i = 0; i < kPipeWaves; i++)
                cudaStreamSynchronize(c->pipe[i]);
            return rc;
        }
    }
    for (int i = 0; i < kPipeWaves; i++) {
        CED_CUDA(cudaEventRecord(c->waveJoin[i], c->pipe[i]));
        CED_CUDA(cudaStreamWaitEvent(s, c->waveJoin[i], 0));
    }
    return CED_OK;
}

int ced_decode_batch(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nFrames,
                     int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    return decodeBatchPipelined(c, code, WireKind::Bytes, dSegs, segStride, nFrames, frameBits, dOut, outStride, stream);
}

int ced_decode_batch_packed(ced_ctx *c, const ced_code_t *code, const uint8_t *dPacked, size_t packedStride,
                            int nFrames, int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    return decodeBatchPipelined(c, code, WireKind::Packed, dPacked, packedStride, nFrames, frameBits, dOut, outStride, stream);
}

int ced_decode_batch_softq(ced_ctx *c, const ced_code_t *code, const uint8_t *dSyms, size_t symStride, int nFrames,
                           int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    return decodeBatchPipelined(c, code, WireKind::SoftQ, dSyms, symStride, nFrames, frameBits, dOut, outStride, stream);
}

/* ------------------------------------------------ continuous streams, windowed traceback */

static inline size_t windowCarryGroupBytes(int depth)
{
    /* per 32-stream group: 4 rows of metrics (uint4 x 32), 32 start positions, depth/2 survivor rows */
    return 4 * 32 * sizeof(uint4) + 32 * sizeof(uint32_t) + (size_t)(depth / 2) * 32 * sizeof(uint4);
}

size_t ced_window_carry_bytes(int nStreams, int depth)
{
    if (nStreams <= 0 || depth < 24 || depth % 24)
        return 0;
    return (size_t)((nStreams + 31) / 32) * windowCarryGroupBytes(depth);
}

size_t ced_window_carry_bytes_code(const ced_code_t *code, int nStreams, int depth)
{
    if (classify(code) != CodeId::Unsupported)
        return ced_window_carry_bytes(nStreams, depth);
    return cedWindowCarryBytesGeneric(code, nStreams, depth);
}

static int decodeWindowImpl(ced_ctx *c, const ced_code_t *code, WireKind kind, const uint8_t *dSegs, size_t segStride,
                            int nStreams, int nSegments, uint64_t streamPos, int depth, int last, void *dCarry,
                            uint8_t *dOut, size_t outStride, void *stream)
{
    const bool packed = kind == WireKind::Packed, softq = kind == WireKind::SoftQ;
    const int sliceUnit = packed ? ced::PackedSymbols::kChunk : ced::ByteSymbols::kChunk; /* 192 / 96 segments */
    if (!c || nStreams < 0 || nSegments < 0 || depth < 24 || depth % 24 || depth > 8184 || streamPos % sliceUnit ||
        (nStreams > 0 && (!dSegs || !dOut || !dCarry)) || (reinterpret_cast<uintptr_t>(dCarry) & 15u)) {
        setError("ced_decode_window_batch: bad argument (depth and streamPos must be multiples of 24 / 96)");
        return CED_ERR_ARG;
    }
    const CodeId id = classify(code);
    if (softq && id != CodeId::K7_0113_0171 && id != CodeId::K7_0133_0171) {
        setError("ced_decode_window_batch_softq: K=7 rate-1/2 codes 0113/0171 and 0133/0171 only");
        return CED_ERR_UNSUPPORTED;
    }
    if (id == CodeId::Unsupported && kind == WireKind::Bytes) {
        /* other code parameters (K <= 7, 2 or 3 generators of any shape): the table-driven kernels (swar_generic.cu) */
        const int rg = cedDecodeWindowGeneric(c, code, dSegs, segStride, nStreams, nSegments, streamPos, depth, last, dCarry, dOut,
                                              outStride, stream);
        if (rg != CED_ERR_UNSUPPORTED)
            return rg;
    }
    if (id == CodeId::Unsupported || (packed && id == CodeId::K7_RuntimeN3)) {
        setError("ced_decode_window_batch: K=7 codes with 2 or 3 generators that tap the newest and the oldest bit only "
                 "(2 generators for the packed format)");
        return CED_ERR_UNSUPPORTED;
    }
    if (last ? (nSegments < ced::kTailSteps || (streamPos + (uint64_t)nSegments - ced::kTailSteps) % 8 != 0)
             : (nSegments == 0 || nSegments % sliceUnit != 0)) {
        setError("ced_decode_window_batch: a slice must be a positive multiple of 96 segments; the last one must "
                 "end the stream on a byte boundary plus K-1 tail segments");
        return CED_ERR_ARG;
    }
    if (nSegments > kStreamMaxSteps * 4) {
        setError("ced_decode_window_batch: slice too long");
        return CED_ERR_ARG;
    }
    /* local step index l = absolute step - (streamPos - depth): rows [0, depth) are the carried decisions */
    const int emitLo = (int)std::max<int64_t>(0, (int64_t)depth - (int64_t)streamPos);
    const int Tl = depth + nSegments;
    const int emitHi = last ? Tl - ced::kTailSteps : nSegments;
    const int bytesOut = emitHi > emitLo ? (emitHi - emitLo) / 8 : 0;
    if (segStride < (packed ? (size_t)(nSegments + 3) / 4 : (size_t)nSegments) || outStride < (size_t)bytesOut) {
        setError("ced_decode_window_batch: stride shorter than a slice");
        return CED_ERR_ARG;
    }
    if (nStreams == 0)
        return bytesOut;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const uint2 *stepTable = nullptr;
    if (id == CodeId::K7_Runtime || id == CodeId::K7_RuntimeN3) {
        int rc = stepTableFor(c, code, &stepTable);
        if (rc != CED_OK)
            return rc;
    }
    const size_t perFrame = (size_t)(Tl / 2) * sizeof(uint4);
    size_t waveMax = std::min<size_t>(c->maxWaveFrames, std::max<size_t>(64, kMaxScratchBytes / perFrame));
    waveMax = waveMax / 64 * 64;
    const size_t firstWave = std::min<size_t>((size_t)nStreams, waveMax);
    const size_t firstGroups = (firstWave + 31) / 32;
    ced_ctx::Work &wk = c->work[0];
    if (wk.scratch.bytes < firstGroups * 32 * perFrame || wk.schedState.bytes < firstGroups * 4 * 32 * sizeof(uint4) ||
        wk.schedFlags.bytes < (firstGroups + 1) * sizeof(int)) {
        CED_CUDA(cudaDeviceSynchronize());
        int rc = wk.scratch.ensure(firstGroups * 32 * perFrame);
        if (rc == CED_OK) rc = wk.schedState.ensure(firstGroups * 4 * 32 * sizeof(uint4));
        if (rc == CED_OK) rc = wk.schedFlags.ensure((firstGroups + 1) * sizeof(int));
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    const bool aligned16 = (reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0;
    const size_t allGroups = (size_t)(nStreams + 31) / 32;
    uint8_t *carry = static_cast<uint8_t *>(dCarry);
    uint4 *carryMetrics = reinterpret_cast<uint4 *>(carry);
    uint32_t *carryStart = reinterpret_cast<uint32_t *>(carry + allGroups * 4 * 32 * sizeof(uint4));
    uint8_t *carrySurv = carry + allGroups * (4 * 32 * sizeof(uint4) + 32 * sizeof(uint32_t));
    const size_t tailBytes = (size_t)(depth / 2) * 32 * sizeof(uint4);   /* per group */
    const size_t rowBytes = (size_t)(Tl / 2) * 32 * sizeof(uint4);       /* per group in the scratch */
    for (size_t f0 = 0; f0 < (size_t)nStreams; f0 += waveMax) {
        const int wave = (int)std::min<size_t>(waveMax, (size_t)nStreams - f0);
        const int groups = (wave + 31) / 32;
        const size_t g0 = f0 / 32;
        if (streamPos > 0)
            CED_CUDA(cudaMemcpy2DAsync(wk.scratch.p, rowBytes, carrySurv + g0 * tailBytes, tailBytes, tailBytes,
                                       (size_t)groups, cudaMemcpyDeviceToDevice, s));
        int gridBlocks = c->fwdBlocks;
        if (gridBlocks == 0)
            gridBlocks = c->sms * std::max(3, std::min({5, c->fwdResident, groups / (4 * c->sms)}));
        const int blocks = std::max(1, std::min(gridBlocks, (groups + 3) / 4));
        ced::FwdSched sched;
        sched.counter = reinterpret_cast<unsigned int *>(wk.schedFlags.p);
        sched.done = wk.schedFlags.p + 1;
        sched.state = wk.schedState.p;
        CED_CUDA(cudaMemsetAsync(wk.schedFlags.p, 0, (size_t)(groups + 1) * sizeof(int), s));
        ced::FwdWindow win;
        win.metricsIn = streamPos > 0 ? carryMetrics + g0 * 4 * 32 : nullptr;
        win.metricsOut = last ? nullptr : carryMetrics + g0 * 4 * 32;
        win.startPos = carryStart + g0 * 32;
        win.survPairs = Tl / 2;
        win.pairOffset = depth / 2;
        const uint8_t *in = dSegs + f0 * segStride;
        const ced::BmTable &bm = (id == CodeId::K7_0113_0171) ? c->bm0113 : c->bm0133;
#define CED_LAUNCH_WIN(CODE, FMT)                                                                                  \
    do {                                                                                                           \
        if (aligned16)                                                                                             \
            ced::k7ForwardKernel<CODE, ced::FMT, true, true><<<blocks, ced::kFwdThreads, 0, s>>>(                  \
                in, segStride, wave, nSegments, wk.scratch.p, bm, sched, 2, win, stepTable);                       \
        else                                                                                                       \
            ced::k7ForwardKernel<CODE, ced::FMT, false, true><<<blocks, ced::kFwdThreads, 0, s>>>(                 \
                in, segStride, wave, nSegments, wk.scratch.p, bm, sched, 2, win, stepTable);                       \
    } while (0)
        if (softq) {
            const int rq = cedSoftQForwardWindow(c, code, aligned16, blocks, s, in, segStride, wave, nSegments, wk.scratch.p,
                                                 sched, win);
            if (rq != CED_OK)
                return rq;
        } else if (id == CodeId::K7_RuntimeN3)
            CED_LAUNCH_WIN(ced::RuntimeK7<3>, ByteSymbols);
        else if (id == CodeId::K7_Runtime && !packed)
            CED_LAUNCH_WIN(ced::RuntimeK7<2>, ByteSymbols);
        else if (id == CodeId::K7_Runtime)
            CED_LAUNCH_WIN(ced::RuntimeK7<2>, PackedSymbols);
        else if (id == CodeId::K7_0113_0171 && !packed)
            CED_LAUNCH_WIN(Code0113, ByteSymbols);
        else if (id == CodeId::K7_0113_0171)
            CED_LAUNCH_WIN(Code0113, PackedSymbols);
        else if (!packed)
            CED_LAUNCH_WIN(Code0133, ByteSymbols);
        else
            CED_LAUNCH_WIN(Code0133, PackedSymbols);
#undef CED_LAUNCH_WIN
        c->launches += 1;
        if (bytesOut > 0) {
            ced::k7TracebackKernel<ced::Lanes8><<<(wave + ced::kTbThreads - 1) / ced::kTbThreads, ced::kTbThreads, 0, s>>>(
                wk.scratch.p, wave, Tl, dOut + f0 * outStride, outStride, last ? nullptr : win.startPos,
                last ? ced::kTailSteps : depth, emitLo);
            c->launches += 1;
        }
        if (!last)
            CED_CUDA(cudaMemcpy2DAsync(carrySurv + g0 * tailBytes, tailBytes,
                                       reinterpret_cast<uint8_t *>(wk.scratch.p) + (size_t)(nSegments / 2) * 32 * sizeof(uint4),
                                       rowBytes, tailBytes, (size_t)groups, cudaMemcpyDeviceToDevice, s));
    }
    CED_CUDA(cudaEventRecord(wk.idle, s));
    wk.lastStream = s;
    CED_CUDA(cudaGetLastError());
    return bytesOut;
}

int ced_decode_window_batch(ced_ctx *c, const ced_code_t *code, const uint8_t *dSegs, size_t segStride, int nStreams,
                            int nSegments, uint64_t streamPos, int depth, int last, void *dCarry, uint8_t *dOut,
                            size_t outStride, void *stream)
{
    return decodeWindowImpl(c, code, WireKind::Bytes, dSegs, segStride, nStreams, nSegments, streamPos, depth, last, dCarry,
                            dOut, outStride, stream);
}

int ced_decode_window_batch_packed(ced_ctx *c, const ced_code_t *code, const uint8_t *dPacked, size_t packedStride,
                                   int nStreams, int nSegments, uint64_t streamPos, int depth, int last, void *dCarry,
                                   uint8_t *dOut, size_t outStride, void *stream)
{
    return decodeWindowImpl(c, code, WireKind::Packed, dPacked, packedStride, nStreams, nSegments, streamPos, depth, last,
                            dCarry, dOut, outStride, stream);
}

int ced_decode_window_batch_softq(ced_ctx *c, const ced_code_t *code, const uint8_t *dSyms, size_t symStride, int nStreams,
                                  int nSegments, uint64_t streamPos, int depth, int last, void *dCarry, uint8_t *dOut,
                                  size_t outStride, void *stream)
{
    return decodeWindowImpl(c, code, WireKind::SoftQ, dSyms, symStride, nStreams, nSegments, streamPos, depth, last, dCarry,
                            dOut, outStride, stream);
}

int ced_pack_symbols(ced_ctx *c, const uint8_t *dSegs, size_t segStride, int nFrames, int segsPerFrame,
                     uint8_t *dPacked, size_t packedStride, void *stream)
{
    if (!c || nFrames < 0 || segsPerFrame <= 0 || (nFrames > 0 && (!dSegs || !dPacked)) ||
        segStride < (size_t)segsPerFrame || packedStride < (size_t)(segsPerFrame + 3) / 4) {
        setError("ced_pack_symbols: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const long long work = (long long)nFrames * ((segsPerFrame + 15) / 16);
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 32);
    const int aligned = ((reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0 &&
                         (reinterpret_cast<uintptr_t>(dPacked) & 3u) == 0 && (packedStride & 3u) == 0) ? 1 : 0;
    ced::packSymbolsKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        dSegs, segStride, nFrames, segsPerFrame, dPacked, packedStride, aligned);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int ced_ctx_set_profiling(ced_ctx *c, int enable)
{
    if (!c)
        return CED_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    c->profiling = enable != 0;
    c->profWaves = 0;
    return CED_OK;
}

int ced_ctx_last_kernel_ms(ced_ctx *c, float *ms2)
{
    if (!c || !ms2)
        return CED_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    ms2[0] = ms2[1] = 0.f;
    for (int w = 0; w < c->profWaves; w++) {
        float a = 0.f, b = 0.f;
        CED_CUDA(cudaEventSynchronize(c->prof[w][2]));
        CED_CUDA(cudaEventElapsedTime(&a, c->prof[w][0], c->prof[w][1]));
        CED_CUDA(cudaEventElapsedTime(&b, c->prof[w][1], c->prof[w][2]));
        ms2[0] += a;
        ms2[1] += b;
    }
    return CED_OK;
}

int ced_ctx_last_fallback_frames(ced_ctx *c, int *frames)
{
    if (!c || !frames)
        return CED_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    *frames = 0;
    if (!c->lastFusedCount)
        return CED_OK;
    CED_CUDA(cudaDeviceSynchronize());
    CED_CUDA(cudaMemcpy(frames, c->lastFusedCount, sizeof(int), cudaMemcpyDeviceToHost));
    return CED_OK;
}

int ced_probe_int_peak(ced_ctx *c, int mode, double *laneOpsPerSecond)
{
    if (!c || !laneOpsPerSecond || mode < 0 || mode > 1)
        return CED_ERR_ARG;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    int sms = 0;
    CED_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    uint32_t *sink = nullptr;
    CED_CUDA(cudaMalloc(reinterpret_cast<void **>(&sink), 64));
    cudaEvent_t e0, e1;
    CED_CUDA(cudaEventCreate(&e0));
    CED_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CED_CUDA(cudaEventRecord(e0, c->stream));
        if (mode == 0)
            ced::intProbeKernel<0><<<blocks, threads, 0, c->stream>>>(sink, iters, 0x9E3779B9u + rep, 0x7F4A7C15u);
        else
            ced::intProbeKernel<1><<<blocks, threads, 0, c->stream>>>(sink, iters, 0x9E3779B9u + rep, 0x7F4A7C15u);
        CED_CUDA(cudaEventRecord(e1, c->stream));
        CED_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CED_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)blocks * threads * (double)iters * ced::kProbeUnroll * ced::kProbeChains;
        if (rep > 0)
            best = std::max(best, ops / (ms * 1e-3));
        c->launches += 1;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *laneOpsPerSecond = best;
    return CED_OK;
}

static int launchEncode(ced_ctx *c, const ced_code_t *code, const uint8_t *dMsg, size_t msgStride, int nFrames,
                        int frameBytes, uint8_t *dSegs, size_t segStride, int tailSegs, uint32_t hist, cudaStream_t s,
                        bool packed = false)
{
    ced::EncTaps taps;
    for (int i = 0; i < 8; i++)
        taps.tap[i] = i < code->codedBits ? reverseBits(code->gen[i], code->constraintLen) : 0u;
    const int T = 8 * frameBytes + tailSegs;
    if (T == 0)
        return CED_OK;
    const int blocks = (nFrames + ced::kEncFramesPerBlock - 1) / ced::kEncFramesPerBlock;
    if (packed) {
        const int aligned4 = ((reinterpret_cast<uintptr_t>(dSegs) & 3u) == 0 && (segStride & 3u) == 0) ? 1 : 0;
        if (code->constraintLen == 7 && taps.tap[0] == ced::kFixedTap0 && taps.tap[1] == ced::kFixedTap1)
            ced::encodeBatchKernel<7, 2, true, true><<<blocks, ced::kEncThreads, 0, s>>>(
                dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, tailSegs, 7, 2, taps, hist, aligned4);
        else if (code->constraintLen == 7)
            ced::encodeBatchKernel<7, 2, true><<<blocks, ced::kEncThreads, 0, s>>>(
                dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, tailSegs, 7, 2, taps, hist, aligned4);
        else
            ced::encodeBatchKernel<0, 2, true><<<blocks, ced::kEncThreads, 0, s>>>(
                dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, tailSegs, code->constraintLen, 2, taps, hist,
                aligned4);
        c->launches += 1;
        CED_CUDA(cudaGetLastError());
        return CED_OK;
    }
    const int aligned16 = ((reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0) ? 1 : 0;
    const bool fixedTaps = code->constraintLen == 7 && code->codedBits == 2 && taps.tap[0] == ced::kFixedTap0 &&
                           taps.tap[1] == ced::kFixedTap1;
    static const bool noLut = getenv("CED_ENC_NO_LUT") != nullptr; /* experiments: the per-item kernel */
    if (code->codedBits == 2 && aligned16 && hist == 0u && !noLut && (reinterpret_cast<uintptr_t>(dMsg) & 1u) == 0 &&
        (msgStride & 1u) == 0 && (frameBytes & 1) == 0) {
        /* a warp encodes one frame at a time; 16 CTAs of 8 warps per SM (8 resident, the rest balance the end of the
         * grid): 8 / 16 per SM measured 4213 / 4399 Gbit/s at 2^20 frames */
        static const int envPerSm = getenv("CED_ENC_CTAS_PER_SM") ? atoi(getenv("CED_ENC_CTAS_PER_SM")) : 0;
        const long long perSm = envPerSm > 0 ? envPerSm : 16;
        const int chunksL = (T + 15) / 16;
        const int fpw = chunksL > 16 ? 1 : chunksL > 8 ? 2 : chunksL > 4 ? 4 : chunksL > 2 ? 8 : chunksL > 1 ? 16 : 32; /* frames per warp */
        const int blocksL = (int)std::min<long long>(((long long)nFrames + 8 * fpw - 1) / (8 * fpw),
                                                     (long long)(c->sms > 0 ? c->sms : 148) * perSm);
        auto launchLut = [&](auto kernel) {
            kernel<<<blocksL, ced::kEncLutThreads, 0, s>>>(dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, tailSegs,
                                                          taps.tap[0], taps.tap[1]);
        };
        if (fixedTaps && fpw > 1)
            launchLut(ced::encodeBatchLutKernel<true, true>);
        else if (fixedTaps)
            launchLut(ced::encodeBatchLutKernel<true, false>);
        else if (fpw > 1)
            launchLut(ced::encodeBatchLutKernel<false, true>);
        else
            launchLut(ced::encodeBatchLutKernel<false, false>);
    } else if (fixedTaps)
        ced::encodeBatchKernel<7, 2, false, true><<<blocks, ced::kEncThreads, 0, s>>>(
            dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, tailSegs, 7, 2, taps, hist, aligned16);
    else if (code->constraintLen == 7 && code->codedBits == 2)
        ced::encodeBatchKernel<7, 2><<<blocks, ced::kEncThreads, 0, s>>>(dMsg, msgStride, nFrames, frameBytes, dSegs,
                                                                         segStride, tailSegs, 7, 2, taps, hist, aligned16);
    else
        ced::encodeBatchKernel<0, 0><<<blocks, ced::kEncThreads, 0, s>>>(dMsg, msgStride, nFrames, frameBytes, dSegs,
                                                                         segStride, tailSegs, code->constraintLen,
                                                                         code->codedBits, taps, hist, aligned16);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int ced_encode_batch(ced_ctx *c, const ced_code_t *code, const uint8_t *dMsg, size_t msgStride, int nFrames,
                     int frameBytes, uint8_t *dSegs, size_t segStride, void *stream)
{
    cedStopPacketServer();
    if (!c || !code || nFrames < 0 || frameBytes <= 0 || (nFrames > 0 && (!dMsg || !dSegs))) {
        setError("ced_encode_batch: bad argument");
        return CED_ERR_ARG;
    }
    if (code->constraintLen < 2 || code->constraintLen > 9 || code->codedBits < 1 || code->codedBits > CED_MAX_N) {
        setError("ced_encode_batch: K must be 2..9 and n 1..8");
        return CED_ERR_UNSUPPORTED;
    }
    if (msgStride < (size_t)frameBytes || segStride < (size_t)(8 * frameBytes + code->constraintLen - 1)) {
        setError("ced_encode_batch: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    return launchEncode(c, code, dMsg, msgStride, nFrames, frameBytes, dSegs, segStride, code->constraintLen - 1, 0u,
                        stream ? (cudaStream_t)stream : c->stream);
}

int ced_encode_batch_packed(ced_ctx *c, const ced_code_t *code, const uint8_t *dMsg, size_t msgStride, int nFrames,
                            int frameBytes, uint8_t *dPacked, size_t packedStride, void *stream)
{
    if (!c || !code || nFrames < 0 || frameBytes <= 0 || (nFrames > 0 && (!dMsg || !dPacked))) {
        setError("ced_encode_batch_packed: bad argument");
        return CED_ERR_ARG;
    }
    if (code->constraintLen < 2 || code->constraintLen > 9 || code->codedBits != 2) {
        setError("ced_encode_batch_packed: the packed format is defined for n = 2 (K 2..9)");
        return CED_ERR_UNSUPPORTED;
    }
    if (msgStride < (size_t)frameBytes ||
        packedStride < (size_t)(8 * frameBytes + code->constraintLen - 1 + 3) / 4) {
        setError("ced_encode_batch_packed: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    return launchEncode(c, code, dMsg, msgStride, nFrames, frameBytes, dPacked, packedStride, code->constraintLen - 1,
                        0u, stream ? (cudaStream_t)stream : c->stream, true);
}

int ced_slice_soft_symbols(ced_ctx *c, const int8_t *dSoft, size_t softStride, int nFrames, int segsPerFrame,
                           uint8_t *dPacked, size_t packedStride, void *stream)
{
    if (!c || nFrames < 0 || segsPerFrame <= 0 || (nFrames > 0 && (!dSoft || !dPacked)) ||
        softStride < (size_t)2 * segsPerFrame || packedStride < (size_t)(segsPerFrame + 3) / 4) {
        setError("ced_slice_soft_symbols: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const long long work = (long long)nFrames * ((segsPerFrame + 15) / 16);
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 32);
    const int aligned = ((reinterpret_cast<uintptr_t>(dSoft) & 15u) == 0 && (softStride & 15u) == 0 &&
                         (reinterpret_cast<uintptr_t>(dPacked) & 3u) == 0 && (packedStride & 3u) == 0) ? 1 : 0;
    ced::sliceSoftSymbolsKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        dSoft, softStride, nFrames, segsPerFrame, dPacked, packedStride, aligned);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

/* H2D -> kernels -> D2H over two buffers; `encode` selects the direction of the sizes. */
enum class HostOp { Encode, Decode, DecodePacked, DecodeViaPack, DecodeAdaptive, DecodeSoftQ };

/* ordinary malloc'ed / stack memory, i.e. neither page-locked by CUDA nor registered */
static bool isPageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static int ensurePacker(ced_ctx *c)
{
    if (!c->packer) {
        const char *envT = getenv("CED_HOST_THREADS");
        int threads = envT ? atoi(envT) : hostCoresPerActiveGpu();
        c->packer = ced_host::packerCreate(std::max(1, std::min(threads, envT ? 64 : 8)));
    }
    return c->packer ? CED_OK : CED_ERR_NOMEM;
}

/*
 * Whether packing chunks on the host pays depends on the machine: on a box with 16 cores per GPU the PCIe link is
 * the bottleneck and the mix gains 45 % (53.6 -> 77 Gbit/s); on the 8-GPU box (32 cores, 4 per GPU, all links
 * sharing the host memory) raw copies already run at a third of the link rate and packing loses 9 % (156 -> 142
 * Gbit/s in total; with every rank calibrating at once the measurement below is also too noisy there: 3 of 8 ranks
 * chose packing, 146 Gbit/s).  So (1) hosts with fewer than 8 cores per ACTIVE GPU (activeGpus()) always copy raw, and (2) elsewhere
 * every process measures it once per device: the first 16 eligible calls alternate four raw,
 * four adaptive (the first of each four is a transition and not counted), then the faster mode is kept -- adaptive
 * only if it is at least 5 % faster -- and re-measured after 512 calls.  CED_HOST_PACK = 0 / 1 / 2 bypasses this.
 */
struct PackTuner {
    std::mutex mu;
    int calls = 0;          /* eligible calls since the last (re)start of the calibration */
    double rate[2] = {0, 0};
    int n[2] = {0, 0};
    int decided = -1;       /* -1 calibrating, 0 raw copies, 1 adaptive packing */
    int sinceDecision = 0;
};
static PackTuner gPackTuner[32];
constexpr int kTunerPhaseCalls = 4, kTunerCalibrationCalls = 16, kTunerRedoAfter = 512;

/* returns the mode for this call (0 raw, 1 adaptive) and its ticket (>= 0 while calibrating, else -1) */
static int packTunerBegin(int device, int *ticket)
{
    PackTuner &t = gPackTuner[device & 31];
    std::lock_guard<std::mutex> lock(t.mu);
    if (t.decided >= 0) {
        *ticket = -1;
        const int mode = t.decided;
        if (++t.sinceDecision > kTunerRedoAfter) { /* measure again: other processes may have come or gone */
            t.calls = t.n[0] = t.n[1] = 0;
            t.rate[0] = t.rate[1] = 0;
            t.decided = -1;
        }
        return mode;
    }
    *ticket = t.calls++;
    return (*ticket / kTunerPhaseCalls) % 2;
}

static void packTunerEnd(int device, int ticket, int mode, double bytesPerSecond)
{
    if (ticket < 0)
        return;
    PackTuner &t = gPackTuner[device & 31];
    std::lock_guard<std::mutex> lock(t.mu);
    if (t.decided >= 0)
        return;
    if (ticket % kTunerPhaseCalls != 0) {
        t.rate[mode] += bytesPerSecond;
        t.n[mode]++;
    }
    if (t.n[0] + t.n[1] >= kTunerCalibrationCalls - kTunerCalibrationCalls / kTunerPhaseCalls && t.n[0] && t.n[1]) {
        const double raw = t.rate[0] / t.n[0], mix = t.rate[1] / t.n[1];
        t.decided = mix > 1.05 * raw ? 1 : 0;
        t.sinceDecision = 0;
        if (getenv("CED_HOST_PACK_TRACE"))
            fprintf(stderr, "ced host pipeline: device %d raw %.1f GB/s, adaptive packing %.1f GB/s per call -> %s\n",
                    device, raw / 1e9, mix / 1e9, t.decided ? "adaptive" : "raw copies");
    }
}

static int hostPipelineBody(ced_ctx *c, const ced_code_t *code, HostOp op, const uint8_t *hIn, size_t inStride,
                            size_t inRowBytes, int nFrames, int frameParam, uint8_t *hOut, size_t outStride,
                            size_t outRowBytes);

/* On failure nothing may still be reading or writing the caller's buffers when the call returns: drain the copy
 * and compute streams first (the error text of the failing call is kept). */
static int hostPipeline(ced_ctx *c, const ced_code_t *code, HostOp op, const uint8_t *hIn, size_t inStride,
                        size_t inRowBytes, int nFrames, int frameParam, uint8_t *hOut, size_t outStride,
                        size_t outRowBytes)
{
    cedStopPacketServer();
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    const int rc = hostPipelineBody(c, code, op, hIn, inStride, inRowBytes, nFrames, frameParam, hOut, outStride,
                                    outRowBytes);
    if (rc != CED_OK) {
        char keep[512];
        snprintf(keep, sizeof(keep), "%s", ced_last_error());
        cudaStreamSynchronize(c->h2d);
        for (int i = 0; i < kPipeDepth; i++)
            cudaStreamSynchronize(c->pipe[i]);
        cudaStreamSynchronize(c->d2h);
        cudaGetLastError();
        setError("%s", keep);
    }
    return rc;
}

static int hostPipelineBody(ced_ctx *c, const ced_code_t *code, HostOp op, const uint8_t *hIn, size_t inStride,
                            size_t inRowBytes, int nFrames, int frameParam, uint8_t *hOut, size_t outStride,
                            size_t outRowBytes)
{
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const bool mayPack = op == HostOp::DecodeViaPack || op == HostOp::DecodeAdaptive;
    const size_t packStride = mayPack ? (((size_t)frameParam + 6 + 3) / 4 + 15) / 16 * 16 : 0;
    /* DecodeAdaptive: a chunk is packed by the host threads only while the copy engine still has `lookback`
     * earlier chunks queued, i.e. while packing costs the link nothing (CED_HOST_PACK_LOOKBACK, 1..kPipeDepth-1;
     * measured 1 / 2 / 3 chunks: 73.5 / 79.2 / 72.8 Gbit/s with two callers, tools/host_pack_modes.py) */
    const int envLook = getenv("CED_HOST_PACK_LOOKBACK") ? atoi(getenv("CED_HOST_PACK_LOOKBACK")) : 0;
    const int lookback = std::max(1, std::min(envLook > 0 ? envLook : 2, kPipeDepth - 1));
    static const int envChunk = getenv("CED_HOST_CHUNK_FRAMES") ? atoi(getenv("CED_HOST_CHUNK_FRAMES")) : 0;
    const int chunk = std::min(nFrames, envChunk >= 32 ? envChunk : kHostChunkFrames);
    for (int b = 0; b < kPipeDepth && mayPack; b++) {
        int rc = c->packStage[b].ensure((size_t)chunk * packStride + 16);
        if (rc != CED_OK)
            return rc;
    }
    for (int b = 0; b < kPipeDepth; b++) {
        int rc = c->hostIn[b].ensure((size_t)chunk * std::max(inStride, packStride) + 16);
        if (rc == CED_OK)
            rc = c->hostOut[b].ensure((size_t)chunk * outStride + 16);
        if (rc != CED_OK)
            return rc;
    }
    /* A D2H copy into pageable memory stalls the calling thread until the chunk's kernels are done, which would
     * serialise the pipeline: such results go to page-locked staging first and are moved by the host one
     * pipeline round later. */
    const bool stageOut = isPageable(hOut);
    struct Pending {
        uint8_t *dst;
        size_t bytes;
    } pending[kPipeDepth] = {};
    for (int b = 0; b < kPipeDepth && stageOut; b++) {
        int rc = c->outStage[b].ensure((size_t)chunk * outStride + 16);
        if (rc != CED_OK)
            return rc;
    }
    auto deliver = [&](int b) -> int {
        if (pending[b].bytes) {
            CED_CUDA(cudaEventSynchronize(c->outFree[b]));
            memcpy(pending[b].dst, c->outStage[b].p, pending[b].bytes);
            pending[b].bytes = 0;
        }
        return CED_OK;
    };
    int idx = 0, nPackedChunks = 0;
    for (int f0 = 0; f0 < nFrames; f0 += chunk, idx++) {
        const int b = idx % kPipeDepth;
        cudaStream_t cs = c->pipe[b]; /* consecutive chunks run on different compute streams, so their
                                         (small, latency-bound) kernels overlap on the GPU */
        const int cnt = std::min(chunk, nFrames - f0);
        const size_t inBytes = (size_t)(cnt - 1) * inStride + inRowBytes;
        const size_t outBytes = (size_t)(cnt - 1) * outStride + outRowBytes;
        if (idx >= kPipeDepth)
            CED_CUDA(cudaStreamWaitEvent(c->h2d, c->inFree[b], 0));
        bool packThis = op == HostOp::DecodeViaPack;
        if (op == HostOp::DecodeAdaptive && idx >= lookback) {
            if (c->packHoldoff > 0) {
                c->packHoldoff--;
            } else {
                packThis = cudaEventQuery(c->inReady[(idx - lookback) % kPipeDepth]) == cudaErrorNotReady;
                cudaGetLastError();
            }
        }
        nPackedChunks += packThis ? 1 : 0;
        if (packThis) {
            /* pack this chunk on the host (4 segments per byte) into pinned staging, then copy a quarter
             * of the bytes; the staging slot is reused once its previous H2D has completed */
            const int T = frameParam + 6;
            const size_t pStride = packStride;
            if (idx >= kPipeDepth)
                CED_CUDA(cudaEventSynchronize(c->stageFree[b]));
            ced_host::packerRun(c->packer, hIn + (size_t)f0 * inStride, inStride, cnt, T, c->packStage[b].p, pStride);
            CED_CUDA(cudaMemcpyAsync(c->hostIn[b].p, c->packStage[b].p, (size_t)cnt * pStride, cudaMemcpyHostToDevice,
                                     c->h2d));
            CED_CUDA(cudaEventRecord(c->stageFree[b], c->h2d));
            if (op == HostOp::DecodeAdaptive) {
                /* did the copy engine run out of work while the host was packing (slow or oversubscribed host
                 * cores)?  Then the next chunks go raw, twice as many after every repeat. */
                const bool ranDry = cudaEventQuery(c->inReady[(idx - 1) % kPipeDepth]) == cudaSuccess;
                cudaGetLastError();
                if (ranDry) {
                    c->packHoldoff = c->packPenalty;
                    c->packPenalty = std::min(2 * c->packPenalty, 16);
                } else {
                    c->packPenalty = std::max(1, c->packPenalty / 2);
                }
            }
        } else
        CED_CUDA(cudaMemcpyAsync(c->hostIn[b].p, hIn + (size_t)f0 * inStride, inBytes, cudaMemcpyHostToDevice, c->h2d));
        CED_CUDA(cudaEventRecord(c->inReady[b], c->h2d));
        CED_CUDA(cudaStreamWaitEvent(cs, c->inReady[b], 0));
        if (idx >= kPipeDepth)
            CED_CUDA(cudaStreamWaitEvent(cs, c->outFree[b], 0));
        int rc;
        if (op == HostOp::Encode)
            rc = ced_encode_batch(c, code, c->hostIn[b].p, inStride, cnt, frameParam, c->hostOut[b].p, outStride, cs);
        else if (op == HostOp::DecodeSoftQ)
            rc = cedDecodeBatchSoftQ(c, code, c->hostIn[b].p, inStride, cnt, frameParam, c->hostOut[b].p, outStride, cs, 1 + b);
        else if (packThis)
            rc = decodeBatchImpl(c, code, true, c->hostIn[b].p, packStride, cnt, frameParam, c->hostOut[b].p, outStride,
                                 cs, 1 + b);
        else
            rc = decodeBatchImpl(c, code, op == HostOp::DecodePacked, c->hostIn[b].p, inStride, cnt, frameParam,
                                 c->hostOut[b].p, outStride, cs, 1 + b);
        if (rc != CED_OK)
            return rc;
        CED_CUDA(cudaEventRecord(c->inFree[b], cs));
        CED_CUDA(cudaEventRecord(c->outReady[b], cs));
        CED_CUDA(cudaStreamWaitEvent(c->d2h, c->outReady[b], 0));
        if (stageOut) {
            int drc = deliver(b);   /* the slot's previous result leaves the staging buffer first */
            if (drc != CED_OK)
                return drc;
            CED_CUDA(cudaMemcpyAsync(c->outStage[b].p, c->hostOut[b].p, outBytes, cudaMemcpyDeviceToHost, c->d2h));
            pending[b].dst = hOut + (size_t)f0 * outStride;
            pending[b].bytes = outBytes;
        } else {
            CED_CUDA(cudaMemcpyAsync(hOut + (size_t)f0 * outStride, c->hostOut[b].p, outBytes, cudaMemcpyDeviceToHost,
                                     c->d2h));
        }
        CED_CUDA(cudaEventRecord(c->outFree[b], c->d2h));
    }
    for (int b = 0; b < kPipeDepth; b++) {
        int drc = deliver(b);
        if (drc != CED_OK)
            return drc;
    }
    CED_CUDA(cudaStreamSynchronize(c->d2h));
    for (int i = 0; i < kPipeDepth; i++)
        CED_CUDA(cudaStreamSynchronize(c->pipe[i]));
    static const bool trace = getenv("CED_HOST_PACK_TRACE") != nullptr;
    if (trace && mayPack)
        fprintf(stderr, "ced host pipeline: %d of %d chunks packed on the host\n", nPackedChunks, idx);
    return CED_OK;
}

int ced_decode_batch_host(ced_ctx *c, const ced_code_t *code, const uint8_t *hSegs, size_t segStride, int nFrames,
                          int frameBits, uint8_t *hOut, size_t outStride)
{
    if (!c || !code || !hSegs || !hOut || nFrames < 0 || frameBits <= 0 || (frameBits & 7)) {
        setError("ced_decode_batch_host: bad argument");
        return CED_ERR_ARG;
    }
    if (segStride < (size_t)frameBits + code->constraintLen - 1 || outStride < (size_t)frameBits / 8) {
        setError("ced_decode_batch_host: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    /* small batches (speedDecode's 16 packets, speedDecode/speedDecode.c:18-19): one copy in, the warp-per-frame kernel
     * (warp_frame.cu), one copy out -- no chunk pipeline, no host-side packing */
    if (cedWarpFrameTakes(c, code, nFrames, frameBits)) {
        std::lock_guard<std::recursive_mutex> lock(c->mu);
        CED_CUDA(cudaSetDevice(c->device));
        const size_t T = (size_t)frameBits + code->constraintLen - 1;
        const size_t inBytes = (size_t)(nFrames - 1) * segStride + T, outBytes = (size_t)(nFrames - 1) * outStride + (size_t)frameBits / 8;
        int rc = c->hostIn[0].ensure(inBytes + 16);
        if (rc == CED_OK)
            rc = c->hostOut[0].ensure(outBytes + 16);
        if (rc != CED_OK)
            return rc;
        CED_CUDA(cudaMemcpyAsync(c->hostIn[0].p, hSegs, inBytes, cudaMemcpyHostToDevice, c->stream));
        rc = cedDecodeBatchWarpFrame(c, code, c->hostIn[0].p, segStride, nFrames, frameBits, c->hostOut[0].p, outStride, c->stream);
        if (rc == CED_OK) {
            CED_CUDA(cudaMemcpyAsync(hOut, c->hostOut[0].p, outBytes, cudaMemcpyDeviceToHost, c->stream));
            CED_CUDA(cudaStreamSynchronize(c->stream));
            return CED_OK;
        }
        CED_CUDA(cudaStreamSynchronize(c->stream));   /* not a case for that kernel: the general pipeline below */
        if (rc != CED_ERR_UNSUPPORTED)
            return rc;
    }
    /* Transfer compression: worker threads pack the symbols to 2 bits into page-locked staging and a quarter of
     * the bytes crosses PCIe.  CED_HOST_PACK forces 0 = never, 1 = every chunk, 2 = adaptive; the default is
     *  - page-locked caller buffers: adaptive, if this machine gains from it (PackTuner above measures that during
     *    the first 16 calls of at least four chunks; shorter calls copy raw).  The copy engine moves raw chunks at
     *    the PCIe rate (~54 GB/s) while the host threads pack the chunks it is not ready for, so both read the
     *    caller's buffer at once: 53.6 -> 77 Gbit/s with two callers, 47 -> 56-63 with one, close to the ~93 GB/s
     *    at which that host reads its own memory at all (packing every chunk: 48.7 / 66.8 -- the host alone is
     *    slower than the link);
     *  - pageable caller buffers (what a program written against the reference owns): every chunk -- a direct
     *    copy from pageable memory runs at ~10 GB/s, the packing threads read it at several times that.
     * CED_HOST_THREADS sets the pool size (default: host cores / visible GPUs, at most 8). */
    HostOp op = HostOp::Decode;
    bool tuned = false;
    int ticket = -1, tunedMode = 0;
    const CodeId id = classify(code);
    if (id == CodeId::K7_0113_0171 || id == CodeId::K7_0133_0171 || id == CodeId::K7_Runtime) {
        const char *envP = getenv("CED_HOST_PACK");
        const bool pageable = isPageable(hSegs);
        int mode = envP ? atoi(envP) : (pageable ? 1 : 2);
        if (!envP && !pageable) {
            /* page-locked buffers: adaptive packing only where this machine gains from it (PackTuner) */
            const int coresPerGpu = hostCoresPerActiveGpu();
            if (nFrames >= 4 * kHostChunkFrames && coresPerGpu >= 8)
                tuned = true;
            else
                mode = 0;
        }
        if (tuned)
            mode = packTunerBegin(c->device, &ticket) ? 2 : 0;
        tunedMode = mode == 2 ? 1 : 0;
        if (mode != 0) {
            std::lock_guard<std::recursive_mutex> lock(c->mu);
            int rc = ensurePacker(c);
            if (rc != CED_OK)
                return rc;
            /* 2: pack only the chunks the copy engine is not ready for (page-locked buffers) */
            op = mode == 2 && !pageable ? HostOp::DecodeAdaptive : HostOp::DecodeViaPack;
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = hostPipeline(c, code, op, hSegs, segStride, (size_t)frameBits + code->constraintLen - 1, nFrames,
                                frameBits, hOut, outStride, (size_t)frameBits / 8);
    if (tuned && rc == CED_OK) {
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        packTunerEnd(c->device, ticket, tunedMode, (double)nFrames * (double)(frameBits + 6) / std::max(sec, 1e-9));
    }
    return rc;
}

int ced_decode_batch_packed_host(ced_ctx *c, const ced_code_t *code, const uint8_t *hPacked, size_t packedStride,
                                 int nFrames, int frameBits, uint8_t *hOut, size_t outStride)
{
    if (!c || !code || !hPacked || !hOut || nFrames < 0 || frameBits <= 0 || (frameBits & 7)) {
        setError("ced_decode_batch_packed_host: bad argument");
        return CED_ERR_ARG;
    }
    if (packedStride < ((size_t)frameBits + code->constraintLen - 1 + 3) / 4 || outStride < (size_t)frameBits / 8) {
        setError("ced_decode_batch_packed_host: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    return hostPipeline(c, code, HostOp::DecodePacked, hPacked, packedStride,
                        ((size_t)frameBits + code->constraintLen - 1 + 3) / 4, nFrames, frameBits, hOut, outStride,
                        (size_t)frameBits / 8);
}

int ced_decode_batch_softq_host(ced_ctx *c, const ced_code_t *code, const uint8_t *hSyms, size_t symStride, int nFrames,
                                int frameBits, uint8_t *hOut, size_t outStride)
{
    if (!c || !code || !hSyms || !hOut || nFrames < 0 || frameBits <= 0 || (frameBits & 7) || code->constraintLen != 7) {
        setError("ced_decode_batch_softq_host: bad argument");
        return CED_ERR_ARG;
    }
    if (symStride < (size_t)frameBits + 6 || outStride < (size_t)frameBits / 8) {
        setError("ced_decode_batch_softq_host: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    return hostPipeline(c, code, HostOp::DecodeSoftQ, hSyms, symStride, (size_t)frameBits + 6, nFrames, frameBits, hOut,
                        outStride, (size_t)frameBits / 8);
}

int ced_encode_batch_host(ced_ctx *c, const ced_code_t *code, const uint8_t *hMsg, size_t msgStride, int nFrames,
                          int frameBytes, uint8_t *hSegs, size_t segStride)
{
    if (!c || !code || !hMsg || !hSegs || nFrames < 0 || frameBytes <= 0) {
        setError("ced_encode_batch_host: bad argument");
        return CED_ERR_ARG;
    }
    if (msgStride < (size_t)frameBytes || segStride < (size_t)frameBytes * 8 + code->constraintLen - 1) {
        setError("ced_encode_batch_host: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    return hostPipeline(c, code, HostOp::Encode, hMsg, msgStride, (size_t)frameBytes, nFrames, frameBytes, hSegs, segStride,
                        (size_t)frameBytes * 8 + code->constraintLen - 1);
}

int ced_ber_count(ced_ctx *c, const uint8_t *dA, size_t strideA, const uint8_t *dB, size_t strideB, int nFrames,
                  int bytesPerFrame, uint64_t *dCounters, void *stream)
{
    if (!c || !dA || !dB || !dCounters || nFrames < 0 || bytesPerFrame <= 0) {
        setError("ced_ber_count: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const int aligned4 = ((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB) | strideA | strideB |
                           (size_t)bytesPerFrame) & 3u) == 0;
    const long long work = (long long)nFrames * (aligned4 ? bytesPerFrame / 4 : bytesPerFrame);
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 16);
    ced::berCountKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        dA, strideA, dB, strideB, nFrames, bytesPerFrame, reinterpret_cast<unsigned long long *>(dCounters), aligned4);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int ced_bsc_channel(ced_ctx *c, uint8_t *dSegs, size_t segStride, int nFrames, int segsPerFrame, int codedBits, double p,
                    uint64_t seed, uint64_t firstFrameIndex, uint64_t *dCounters, void *stream)
{
    if (!c || !dSegs || nFrames < 0 || segsPerFrame <= 0 || codedBits < 1 || codedBits > CED_MAX_N || !(p >= 0.0) ||
        !(p < 1.0)) {
        setError("ced_bsc_channel: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const uint32_t threshold = (uint32_t)(p * 4294967296.0);
    const long long work = (long long)nFrames * ((segsPerFrame + 15) / 16);
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 32);
    const int aligned16 = ((reinterpret_cast<uintptr_t>(dSegs) & 15u) == 0 && (segStride & 15u) == 0) ? 1 : 0;
    ced::bscChannelKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        dSegs, segStride, nFrames, segsPerFrame, codedBits, threshold, seed, firstFrameIndex,
        reinterpret_cast<unsigned long long *>(dCounters), aligned16);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int ced_random_bytes(ced_ctx *c, uint8_t *dMsg, size_t msgStride, int nFrames, int frameBytes, uint64_t seed,
                     uint64_t firstFrameIndex, void *stream)
{
    if (!c || !dMsg || nFrames < 0 || frameBytes <= 0 || msgStride < (size_t)frameBytes) {
        setError("ced_random_bytes: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const long long work = (long long)nFrames * ((frameBytes + 7) / 8);
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 16);
    ced::randomBytesKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(dMsg, msgStride, nFrames,
                                                                                         frameBytes, seed,
                                                                                         firstFrameIndex);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

/* ------------------------------------------------------------- streaming */

/* CED_STREAM_ZEROCOPY: bit 0 per-frame encoder, bit 1 per-frame decoder input, bit 2 per-frame decoder output */
static int streamZeroCopyMask()
{
    static const int mask = [] {
        const char *e = getenv("CED_STREAM_ZEROCOPY");
        return e ? atoi(e) : 5; /* measured: encoder -5 us, decoder output -5..10 us per call; decoder input slower */
    }();
    return mask;
}

/* CED_STREAM_PARALLEL=0 sends one-shot K=7 packets through the single-warp kernel as well */
/* CED_STREAM_GRAPH=0 issues the copy and the two kernels of that path one by one instead of as one graph launch */
static bool streamGraphEnabled()
{
    static const bool on = [] {
        const char *e = getenv("CED_STREAM_GRAPH");
        return !e || atoi(e) != 0;
    }();
    return on;
}

/* CED_FP_STAMPS=1: the frame-parallel kernels record %globaltimer at their phase boundaries into host-visible
 * memory and every call prints the intervals to stderr (a measurement aid) */
static unsigned long long *fpStamps()
{
    static unsigned long long *p = [] {
        const char *e = getenv("CED_FP_STAMPS");
        unsigned long long *q = nullptr;
        if (e && atoi(e) != 0 && cudaMallocHost(reinterpret_cast<void **>(&q), 16 * sizeof(unsigned long long)) != cudaSuccess)
            q = nullptr;
        return q;
    }();
    return p;
}

static bool streamParallelEnabled()
{
    static const bool on = [] {
        const char *e = getenv("CED_STREAM_PARALLEL");
        return !e || atoi(e) != 0;
    }();
    return on;
}

/* CED_STREAM_SPLIT: one-shot K=7 packets of at least that many segments run on the time-split kernels (warp_split.cu)
 * instead of the frame-parallel ones; 0 = never.  Read per call: tests flip it. */
static int streamSplitMinSegments()
{
    const char *e = getenv("CED_STREAM_SPLIT");
    return e ? atoi(e) : kStreamSplitMinSegments;
}

/* ------------------------------------------------ resident packet decoder (frame_server.cuh) */
/* CED_STREAM_SERVER=1: the resident kernel instead of one graph launch per packet.  Off by default: measured on B200 it
 * takes the host-visible overhead of a call from ~20 us to 2.5 us, but the device-side phases of a 2048-bit packet then
 * add up to 27 us (request staged over PCIe 4.4, passes 6.7 -- they are bound by the ALU work of 64 start states per
 * block --, three grid barriers ~1.3 each, chain 4, select 1.5, walk 2.2): 31.7 us per call either way */
static bool streamServerEnabled()
{
    const char *e = getenv("CED_STREAM_SERVER"); /* read per call: tests flip it */
    return e && atoi(e) != 0;
}

int ced_stream_server_stats(uint64_t *requests, uint64_t *launches)
{
    ced_ctx *c = gServerCtx.load(std::memory_order_acquire);
    if (requests)
        *requests = c ? c->fsRequests : 0;
    if (launches)
        *launches = c ? c->fsLaunches : 0;
    return c && !c->fsDisabled ? 1 : 0;
}

static void fsStop(ced_ctx *c)
{
    ced::FsMailbox *mb = static_cast<ced::FsMailbox *>(c->fsMailbox);
    if (!mb || !c->fsLaunched)
        return;
    mb->seq = ced::kFsExit;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    cudaStreamSynchronize(c->fsStream);
    mb->seq = c->fsSeq;
    c->fsLaunched = false;
}

void cedStopPacketServer()
{
    ced_ctx *c = gServerCtx.load(std::memory_order_acquire);
    if (!c || !c->fsLaunched)
        return;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    cudaSetDevice(c->device);
    fsStop(c);
}

static int fsLaunch(ced_ctx *c, const ced::FpArgs &f, uint32_t lastSeq)
{
    ced::FsMailbox *mb = static_cast<ced::FsMailbox *>(c->fsMailbox);
    unsigned int init[4] = {lastSeq, 0u, 0u, 0u}; /* FsCtl: cmd, T, barrier, gaveUp */
    mb->state = 1;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    static_assert(sizeof(init) == ced::kFsCtlInitBytes, "FsCtl header");
    CED_CUDA(cudaMemcpyAsync(c->fsCtl, init, sizeof(init), cudaMemcpyHostToDevice, c->fsStream));
    ced::FsMailbox *mbDev = static_cast<ced::FsMailbox *>(c->fsMailboxDev);
    ced::FsCtl *ctl = static_cast<ced::FsCtl *>(c->fsCtl);
    ced::FpArgs args = f;
    void *params[] = {&mbDev, &ctl, &args, &lastSeq};
    CED_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(ced::fpServerKernel), dim3(c->sms), dim3(ced::kFpThreads),
                                         params, sizeof(ced::FsShared), c->fsStream));
    c->fsLaunched = true;
    c->fsLaunches++;
    c->launches++;
    return CED_OK;
}

/* one packet through the resident kernel; CED_ERR_UNSUPPORTED = not available (the caller takes the launch path) */
static int fsDecode(ced_ctx *c, const ced::FpArgs &scratch, const uint8_t *edge, const uint8_t *metrics, const uint8_t *segs,
                    int T, uint8_t *uncoded, size_t decodedBytes)
{
    if (c->fsDisabled || T > ced::kFsMaxSegs - 2 * ced::kFpBlock)
        return CED_ERR_UNSUPPORTED;
    if (!c->fsMailbox) {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        void *mb = nullptr;
        if (!coop || cudaHostAlloc(&mb, sizeof(ced::FsMailbox), cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            c->fsDisabled = true;
            return CED_ERR_UNSUPPORTED;
        }
        memset(mb, 0, sizeof(ced::FsMailbox));
        c->fsMailbox = mb;
        if (cudaHostGetDevicePointer(&c->fsMailboxDev, mb, 0) != cudaSuccess || cudaMalloc(&c->fsCtl, sizeof(ced::FsCtl)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->fsStream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaFuncSetAttribute(ced::fpServerKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ced::FsShared)) !=
                cudaSuccess) {
            cudaGetLastError();
            c->fsDisabled = true;
            return CED_ERR_UNSUPPORTED;
        }
        gServerCtx.store(c, std::memory_order_release);
    }
    ced::FsMailbox *mb = static_cast<ced::FsMailbox *>(c->fsMailbox);
    memcpy(mb->edge, edge, 128);
    memcpy(mb->metrics, metrics, 64);
    memcpy(mb->segs, segs, (size_t)T);
    mb->T = (uint32_t)T;
    uint32_t s = c->fsSeq + 1;
    if (s == ced::kFsExit)
        s = 1;
    const uint32_t lastSeq = c->fsSeq;
    c->fsSeq = s;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    mb->seq = s;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    auto giveUp = [&](const char *why) {
        mb->seq = ced::kFsExit;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        cudaStreamSynchronize(c->fsStream);
        cudaGetLastError();
        c->fsLaunched = false;
        c->fsDisabled = true;
        fprintf(stderr, "convolutionalencdec: resident packet decoder switched off (%s); using one launch per packet\n", why);
        return CED_ERR_UNSUPPORTED;
    };
    if (!c->fsLaunched || mb->state != 1) {
        if (c->fsLaunched)
            CED_CUDA(cudaStreamSynchronize(c->fsStream));
        if (fsLaunch(c, scratch, lastSeq) != CED_OK) {
            cudaGetLastError();
            c->fsDisabled = true;
            return CED_ERR_UNSUPPORTED;
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; spins++) {
        if (mb->done == s)
            break;
        const uint32_t st = mb->state;
        if (st == 3)
            return giveUp("a wait inside the kernel timed out");
        if (st == 2) { /* it left (idle time-out) while this request was on its way: start it again */
            CED_CUDA(cudaStreamSynchronize(c->fsStream));
            if (mb->done == s)
                break;
            if (fsLaunch(c, scratch, lastSeq) != CED_OK)
                return giveUp("relaunch failed");
        }
        if ((spins & 1023u) == 1023u) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2))
                return giveUp("no answer within 2 s");
            if (cudaStreamQuery(c->fsStream) != cudaErrorNotReady && mb->done != s && mb->state == 1)
                return giveUp("kernel ended without an answer");
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);
    memcpy(uncoded, mb->out, decodedBytes);
    c->fsRequests++;
    if (getenv("CED_FP_STAMPS") && c->fsRequests % 1000 == 2) {
        const double host = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        const volatile unsigned long long *st = mb->stamp;
        fprintf(stderr, "resident T=%d: host post->answer %.1f us | device: stage %.1f  passes+barrier %.1f  chain %.1f  "
                        "barrier+select+barrier %.1f  walk+write %.1f  (seen->answered %.1f; table built +%.1f, first pass done +%.1f after staging)\n",
                T, host, (st[1] - st[0]) * 1e-3, (st[2] - st[1]) * 1e-3, (st[3] - st[2]) * 1e-3, (st[4] - st[3]) * 1e-3,
                (st[5] - st[4]) * 1e-3, (st[5] - st[0]) * 1e-3, (st[6] - st[1]) * 1e-3, (st[7] - st[1]) * 1e-3);
    }
    return CED_OK;
}

int ced_stream_surv_words(int nStates)
{
    const int H = nStates / 2;
    return 2 * ((H + 31) / 32);
}

/* The doorbell of the per-packet calls: a device counter and a pinned word the last kernel of a call writes the new
 * count to once its results are in the pinned mailbox; the host spins on the word instead of synchronising the stream. */
static int ensureDoorbell(ced_ctx *c)
{
    if (c->sSplitSeq.p)
        return CED_OK;
    int rc = c->sSplitSeq.ensure(64);
    if (rc == CED_OK) rc = c->sDoorbell.ensure(64);
    if (rc != CED_OK)
        return rc;
    CED_CUDA(cudaMemsetAsync(c->sSplitSeq.p, 0, 64, c->stream));
    memset(c->sDoorbell.p, 0, 64);
    c->splitSeq = 0;
    return CED_OK;
}

/* wait for the next count; after 5 ms without it the stream is asked instead, so a failed launch is still reported */
static int waitDoorbell(ced_ctx *c)
{
    const unsigned int expect = ++c->splitSeq;
    volatile unsigned int *bell = reinterpret_cast<volatile unsigned int *>(c->sDoorbell.p);
    const auto t0 = std::chrono::steady_clock::now();
    bool rung = false;
    for (unsigned int spins = 0; !(rung = *bell == expect);) {
        if ((++spins & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(5))
            break;
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (!rung) {
        const cudaError_t es = cudaStreamSynchronize(c->stream);
        c->splitSeq = *bell;   /* whatever was counted is what the next call starts from */
        CED_CUDA(es);
    }
    return CED_OK;
}

static bool streamDoorbellEnabled()
{
    static const bool on = !getenv("CED_STREAM_DOORBELL") || atoi(getenv("CED_STREAM_DOORBELL")) != 0;
    return on;
}

int ced_stream_decode(int K, int n, const uint8_t *edge, uint8_t *metrics, uint32_t *iteration,
                      uint32_t *renormCounter, uint32_t *surv, uint32_t survCapacitySteps, const uint8_t *segs,
                      int segmentsIn, uint8_t *uncoded, int last)
{
    if (K < 2 || K > 9 || n < 1 || n > CED_MAX_N || !edge || !metrics || !iteration || !renormCounter || !surv ||
        segmentsIn < 0 || (segmentsIn > 0 && !segs) || (last && !uncoded)) {
        setError("ced_stream_decode: bad argument");
        return CED_ERR_ARG;
    }
    ced_ctx *c = ced_default_ctx();
    if (!c) {
        setError("ced_stream_decode: no CUDA context (%s)", gLastError);
        return CED_ERR_CUDA;
    }
    const int N = 1 << (K - 1), W = ced_stream_surv_words(N), S = K - 1;
    const uint32_t it0 = *iteration;
    const uint64_t total = (uint64_t)it0 + (uint64_t)segmentsIn;
    if (total > survCapacitySteps || total > kStreamMaxSteps) {
        setError("ced_stream_decode: packet longer than the survivor buffer (%llu steps)", (unsigned long long)total);
        return CED_ERR_ARG;
    }
    if (last && total <= (uint64_t)S) {
        setError("ced_stream_decode: last=true with no information bits");
        return CED_ERR_ARG;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const size_t inBytes = 1024 + (size_t)segmentsIn;
    const size_t survBytes = (size_t)kStreamMaxSteps * W * sizeof(uint32_t);
    const size_t outBytes = 4096 + std::max<size_t>(survBytes, kStreamMaxSteps / 8 + 8);
    /* A whole 64-state n = 2 packet in one call (what speedDecode.c:79 and berTestK7.c:157 issue): every
     * block of 128 steps is worked on at once (frame_parallel.cuh).  The reference's uint8 metrics cannot
     * wrap from reset-like starting values, which is what makes plain ints equivalent. */
    bool parallel = K == 7 && n == 2 && last && it0 == 0 && *renormCounter == 0 && segmentsIn > 2 * S &&
                    streamParallelEnabled();
    for (int i = 0; parallel && i < N; i++) /* + the two branches out of a state carry complementary labels (both
                                             * generators tap the newest bit): a free step then costs at most 1, which
                                             * bounds the reference's metrics and the in-block costs stored as bytes */
        parallel = metrics[i] <= N + 1 && ((edge[i] ^ edge[N + i]) & 3u) == 3u;
    int rc = c->sIn.ensure(1024 + kStreamMaxSteps + 2 * ced::kFpBlock);
    if (rc == CED_OK) rc = c->sOut.ensure(272 + kStreamMaxSteps / 8 + 16);
    if (rc == CED_OK && parallel && !c->sParallel.p) {
        const ced::FpScratch lay = ced::fpScratchLayout(kStreamMaxSteps);
        rc = c->sParallel.ensure(lay.total);
        if (rc == CED_OK)
            CED_CUDA(cudaMemsetAsync(c->sParallel.p + lay.tickets, 0, 16, c->stream));
    }
    const int splitMin = streamSplitMinSegments();
    const bool split = parallel && splitMin > 0 && segmentsIn >= splitMin && !(streamServerEnabled() && !c->fsDisabled) &&
                       cedStreamDecodeSplitTakes(c, segmentsIn);   /* the resident decoder is opt-in and keeps its packets */
    if (rc == CED_OK && split)
        rc = c->sSplit.ensure(cedStreamDecodeSplitScratchBytes(kStreamMaxSteps));
    if (rc == CED_OK && split)
        rc = ensureDoorbell(c);
    if (rc == CED_OK) rc = c->sSurv.ensure(survBytes);
    if (rc == CED_OK) rc = c->sPinIn.ensure(1024 + kStreamMaxSteps);
    if (rc == CED_OK) rc = c->sPinOut.ensure(outBytes);
    if (rc != CED_OK)
        return rc;

    if (parallel && streamServerEnabled() && !c->fsDisabled) {
        /* the resident kernel (frame_server.cuh): no launch, no copy node, no stream synchronise per packet */
        const ced::FpScratch lay = ced::fpScratchLayout(kStreamMaxSteps);
        ced::FpArgs f = {};
        f.cost = c->sParallel.p + lay.cost;
        for (int w = 0; w < ced::kFpWords; w++)
            f.bits[w] = reinterpret_cast<uint32_t *>(c->sParallel.p + lay.bits[w]);
        f.v = reinterpret_cast<int *>(c->sParallel.p + lay.v);
        f.best = reinterpret_cast<uint32_t *>(c->sParallel.p + lay.best);
        f.tickets = reinterpret_cast<unsigned int *>(c->sParallel.p + lay.tickets);
        const size_t nb = (size_t)((total - S - 1) / 8 + 1);
        const int rs = fsDecode(c, f, edge, metrics, segs, segmentsIn, uncoded, nb);
        if (rs == CED_OK)
            return (int)nb;
        if (rs != CED_ERR_UNSUPPORTED)
            return rs;
    } else if (c->fsLaunched) {
        fsStop(c); /* another kind of call: the kernels below should not wait for the idle time-out */
    }

    /* mailbox in : [0,512) edge  [512,768) metrics  [1024,...) segments          (one H2D)
     * mailbox out: [0,256) metrics [256,272) renormCounter [272,...) decoded bytes (one D2H) */
    memcpy(c->sPinIn.p, edge, (size_t)2 * N);
    memcpy(c->sPinIn.p + 512, metrics, (size_t)N);
    if (segmentsIn)
        memcpy(c->sPinIn.p + 1024, segs, (size_t)segmentsIn);
    const int zc = streamZeroCopyMask();
    const bool zcIn = (zc & 2) != 0, zcOut = (zc & 4) != 0;
    if (!zcIn && !parallel)
        CED_CUDA(cudaMemcpyAsync(c->sIn.p, c->sPinIn.p, inBytes, cudaMemcpyHostToDevice, c->stream));
    if (last && it0 > 0) /* chunked packet: bring the earlier decisions back */
        CED_CUDA(cudaMemcpyAsync(c->sSurv.p, surv, (size_t)it0 * W * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                 c->stream));
    const size_t decodedBytes = last ? (size_t)((total - S - 1) / 8 + 1) : 0;
    bool doorbell = false;
    ced::StreamArgs a;
    a.K = K;
    a.n = n;
    a.N = N;
    a.W = W;
    a.iteration = it0;
    a.renormCounter = *renormCounter;
    a.segmentsIn = segmentsIn;
    a.last = last;
    uint8_t *mbIn = (zcIn && !parallel) ? c->sPinIn.p : c->sIn.p, *mbOut = zcOut ? c->sPinOut.p : c->sOut.p;
    a.edge = mbIn;
    a.metricsIn = mbIn + 512;
    a.metrics = mbOut;
    a.segs = mbIn + 1024;
    a.surv = c->sSurv.p;
    a.stateOut = reinterpret_cast<uint32_t *>(mbOut + 256);
    a.out = mbOut + 272;
    if (parallel) {
        const ced::FpScratch lay = ced::fpScratchLayout(kStreamMaxSteps);
        ced::FpArgs f;
        f.T = segmentsIn;
        f.nBlocks = (segmentsIn + ced::kFpBlock - 1) / ced::kFpBlock;
        f.edge = a.edge;
        f.metricsIn = a.metricsIn;
        f.segs = a.segs;
        f.cost = c->sParallel.p + lay.cost;
        for (int w = 0; w < ced::kFpWords; w++)
            f.bits[w] = reinterpret_cast<uint32_t *>(c->sParallel.p + lay.bits[w]);
        f.v = reinterpret_cast<int *>(c->sParallel.p + lay.v);
        f.best = reinterpret_cast<uint32_t *>(c->sParallel.p + lay.best);
        f.tickets = reinterpret_cast<unsigned int *>(c->sParallel.p + lay.tickets);
        f.out = a.out;
        f.stamps = fpStamps();
        f.stampAll = getenv("CED_FP_STAMPS") && atoi(getenv("CED_FP_STAMPS")) > 1;
        if (f.stamps) { /* the device is idle here: every call ends with a synchronise */
            f.stamps[0] = f.stamps[5] = ~0ull;
            f.stamps[1] = f.stamps[6] = 0;
        }
        const int grid = f.nBlocks * 64 / (ced::kFpThreads / 32);
        /* the join kernel writes the packet's bytes to the pinned mailbox and then a count to a doorbell word: the host
         * waits for that word instead of for the stream (CED_STREAM_DOORBELL=0: cudaStreamSynchronize as everywhere else) */
        doorbell = split && zcOut && streamDoorbellEnabled();
        /* the 64 passes over a block all read its segments: those reads stay on the device (one small copy) */
        auto issue = [&]() -> cudaError_t {
            cudaError_t e = cudaSuccess;
            if (split) {
                /* warp_split.cu: blocks of 64 steps from guessed metrics, checked hand-overs; the symbols are read from
                 * the pinned mailbox directly (every block reads its own 160 bytes once) */
                const int rs = cedStreamDecodeSplit(c, edge, metrics, c->sPinIn.p + 1024, segmentsIn, a.out, c->sSplit.p,
                                                    c->sSplit.bytes, c->stream, doorbell ? c->sSplitSeq.p : nullptr,
                                                    doorbell ? reinterpret_cast<volatile unsigned int *>(c->sDoorbell.p) : nullptr);
                if (rs != CED_OK)
                    return cudaErrorNotSupported;
            } else {
            e = cudaMemcpyAsync(c->sIn.p, c->sPinIn.p, inBytes, cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess)
                return e;
            ced::fpBlockKernel<<<grid, ced::kFpThreads, 0, c->stream>>>(f);
            ced::fpSelectKernel<<<grid, ced::kFpThreads, 0, c->stream>>>(f);
            }
            if (!zcOut)
                e = cudaMemcpyAsync(c->sPinOut.p, c->sOut.p, 272 + decodedBytes, cudaMemcpyDeviceToHost, c->stream);
            return e != cudaSuccess ? e : cudaGetLastError();
        };
        if (streamGraphEnabled()) {
            /* per-packet loops call with one packet length: the three submissions become one graph launch */
            /* the graph holds raw pointers into the staging buffers, which ced_stream_encode may have regrown
             * (ensure() frees and reallocates): the cache is keyed on them as well as on the packet length */
            const void *key[5] = {c->sIn.p, c->sOut.p, split ? c->sSplit.p : c->sParallel.p, c->sPinIn.p, c->sPinOut.p};
            /* the split kernels take the labels and the start metrics as kernel arguments: part of the key */
            uint8_t tables[192];
            memcpy(tables, edge, 128);
            memcpy(tables + 128, metrics, 64);
            if (!c->fpGraph || c->fpGraphSegs != segmentsIn || memcmp(key, c->fpGraphKey, sizeof(key)) != 0 ||
                c->fpGraphSplit != split || c->fpGraphDoorbell != doorbell || (split && memcmp(tables, c->fpGraphTables, sizeof(tables)) != 0)) {
                if (c->fpGraph)
                    cudaGraphExecDestroy(c->fpGraph);
                c->fpGraph = nullptr;
                cudaGraph_t g = nullptr;
                CED_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                const cudaError_t e1 = issue();
                const cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
                CED_CUDA(e1);
                CED_CUDA(e2);
                const cudaError_t e3 = cudaGraphInstantiate(&c->fpGraph, g, 0);
                cudaGraphDestroy(g);
                CED_CUDA(e3);
                c->fpGraphSegs = segmentsIn;
                c->fpGraphSplit = split;
                c->fpGraphDoorbell = doorbell;
                memcpy(c->fpGraphTables, tables, sizeof(tables));
                memcpy(c->fpGraphKey, key, sizeof(key));
            }
            CED_CUDA(cudaGraphLaunch(c->fpGraph, c->stream));
        } else {
            CED_CUDA(issue());
        }
        c->launches += 1;
    } else if (N <= 64)
        ced::streamDecodeWarpKernel<<<1, 32, 0, c->stream>>>(a);
    else
        ced::streamDecodeKernel<<<1, std::max(32, N / 2), 0, c->stream>>>(a);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    if (!zcOut && !parallel)
        CED_CUDA(cudaMemcpyAsync(c->sPinOut.p, c->sOut.p, 272 + decodedBytes, cudaMemcpyDeviceToHost, c->stream));
    if (!last && segmentsIn)
        CED_CUDA(cudaMemcpyAsync(c->sPinOut.p + 4096, c->sSurv.p + (size_t)it0 * W,
                                 (size_t)segmentsIn * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (doorbell) {
        const int rw = waitDoorbell(c);
        if (rw != CED_OK)
            return rw;
    } else {
        CED_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (parallel && fpStamps()) {
        static unsigned long long calls = 0;
        const unsigned long long *st = fpStamps();
        if (calls++ % 1000 == 2) { /* every 1000th call, so a tight loop stays tight */
            fprintf(stderr, "fp T=%d: prologue %.1f  chain %.1f (SM clock %.0f MHz) | kernel gap %.1f | walk+write %.1f  us",
                    segmentsIn, (st[3] - st[2]) * 1e-3, (st[4] - st[3]) * 1e-3,
                    (double)(st[11] - st[10]) / (double)(st[4] - st[3]) * 1e3, ((double)st[5] - (double)st[4]) * 1e-3,
                    (st[8] - st[7]) * 1e-3);
            if (st[1]) /* CED_FP_STAMPS=2: grid-wide stamps (atomics on host memory, they slow the kernels down) */
                fprintf(stderr, "  passes %.1f  select %.1f", (st[1] - st[0]) * 1e-3, (st[6] - st[5]) * 1e-3);
            fprintf(stderr, "\n");
        }
    }
    if (last) {
        memcpy(uncoded, c->sPinOut.p + 272, decodedBytes);
        /* src/viterbiDecoderButterflyk1.c:259 -- the caller's reset restores metrics/counters */
        return (int)decodedBytes;
    }
    memcpy(metrics, c->sPinOut.p, (size_t)N);
    memcpy(renormCounter, c->sPinOut.p + 256, sizeof(uint32_t));
    if (segmentsIn)
        memcpy(surv + (size_t)it0 * W, c->sPinOut.p + 4096, (size_t)segmentsIn * W * sizeof(uint32_t));
    *iteration = it0 + (uint32_t)segmentsIn;
    return 0;
}

int ced_stream_encode(int K, int n, const uint32_t *taps, uint32_t *reg, const uint8_t *in, int bytesIn,
                      uint8_t *segs, int last)
{
    if (K < 2 || K > 9 || n < 1 || n > CED_MAX_N || !taps || !reg || bytesIn < 0 || (bytesIn > 0 && !in) || !segs) {
        setError("ced_stream_encode: bad argument");
        return CED_ERR_ARG;
    }
    ced_ctx *c = ced_default_ctx();
    if (!c) {
        setError("ced_stream_encode: no CUDA context (%s)", gLastError);
        return CED_ERR_CUDA;
    }
    const int tail = last ? K - 1 : 0;
    const int T = 8 * bytesIn + tail;
    if (T == 0)
        return 0;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const size_t need = (size_t)std::max(bytesIn, 1) + 16;
    int rc = c->hostIn[0].ensure(need);
    if (rc == CED_OK) rc = c->hostOut[0].ensure((size_t)T + 16);
    if (rc == CED_OK) rc = c->sPinIn.ensure(std::max<size_t>(need, 1024 + kStreamMaxSteps));
    if (rc == CED_OK) rc = c->sPinOut.ensure(std::max<size_t>((size_t)T + 16, 4096));
    if (rc != CED_OK)
        return rc;
    /* One frame per synchronous call is pure latency: the kernel reads the message from and writes the segments to
     * the pinned mailbox directly (pinned memory is device-addressable under unified addressing), which saves the
     * two copy submissions.  CED_STREAM_ZEROCOPY bit 0 = 0 goes through device staging buffers instead. */
    const bool zeroCopy = (streamZeroCopyMask() & 1) != 0;
    if (bytesIn) {
        memcpy(c->sPinIn.p, in, (size_t)bytesIn);
        if (!zeroCopy)
            CED_CUDA(cudaMemcpyAsync(c->hostIn[0].p, c->sPinIn.p, (size_t)bytesIn, cudaMemcpyHostToDevice, c->stream));
    }
    ced::EncTaps t;
    for (int i = 0; i < 8; i++)
        t.tap[i] = i < n ? taps[i] : 0u;
    const bool doorbell = zeroCopy && streamDoorbellEnabled();
    if (doorbell) {
        rc = ensureDoorbell(c);
        if (rc != CED_OK)
            return rc;
        ced::encodeStreamKernel<<<1, ced::kEncThreads, 0, c->stream>>>(c->sPinIn.p, bytesIn, c->sPinOut.p, (size_t)T + 16, tail, K, n, t,
                                                                         *reg, c->sSplitSeq.p,
                                                                         reinterpret_cast<volatile unsigned int *>(c->sDoorbell.p));
    } else {
        ced::encodeBatchKernel<0, 0><<<1, ced::kEncThreads, 0, c->stream>>>(
            zeroCopy ? c->sPinIn.p : c->hostIn[0].p, (size_t)std::max(bytesIn, 1), 1, bytesIn,
            zeroCopy ? c->sPinOut.p : c->hostOut[0].p, (size_t)T + 16, tail, K, n, t, *reg, 1);
    }
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    if (!zeroCopy)
        CED_CUDA(cudaMemcpyAsync(c->sPinOut.p, c->hostOut[0].p, (size_t)T, cudaMemcpyDeviceToHost, c->stream));
    if (doorbell) {
        rc = waitDoorbell(c);
        if (rc != CED_OK)
            return rc;
    } else {
        CED_CUDA(cudaStreamSynchronize(c->stream));
    }
    memcpy(segs, c->sPinOut.p, (size_t)T);
    /* shift-register bookkeeping only (src/convEncode.c:93,122): which input bits are still in the window */
    uint32_t r = *reg;
    for (int i = std::max(0, bytesIn - 2); i < bytesIn; i++)
        r = (r << 8) | in[i];
    *reg = last ? 0u : (r & ((1u << K) - 1u));
    return T;
}

} // extern "C"
