/*
 * softq_decode.cu -- the 3-bit soft-decision path (include/ced_abi.h): cedDecodeBatchSoftQ (behind ced_decode_batch_softq,
 * which ced_abi.cu wraps in its wave pipeline) and ced_quantize_soft.  Kernels in softq_decode.cuh; nothing here computes on the host except the 12 KB cost table of a code.
 */
#include "ced_internal.cuh"
#include "softq_decode.cuh"

#include <vector>

template <class Code>
static int softqTableFor(ced_ctx *c, int slot, cudaStream_t s, const uint4 **out)
{
    if (!c->softqTable[slot].p) {
        std::vector<uint32_t> t((size_t)ced::kSoftQTableUint4 * 4);
        ced::buildSoftQTable<Code>(t.data());
        int rc = c->softqTable[slot].ensure(t.size() * sizeof(uint32_t));
        if (rc != CED_OK)
            return rc;
        CED_CUDA(cudaMemcpyAsync(c->softqTable[slot].p, t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        CED_CUDA(cudaStreamSynchronize(s)); /* once per context and code: later calls on any stream see it */
    }
    *out = c->softqTable[slot].p;
    return CED_OK;
}

int cedDecodeBatchSoftQ(ced_ctx *c, const ced_code_t *code, const uint8_t *dSyms, size_t symStride, int nFrames,
                        int frameBits, uint8_t *dOut, size_t outStride, void *stream, int slot)
{
    if (!c || nFrames < 0 || frameBits <= 0 || (frameBits & 7) || (nFrames > 0 && (!dSyms || !dOut))) {
        setError("ced_decode_batch_softq: bad argument (frameBits must be a positive multiple of 8)");
        return CED_ERR_ARG;
    }
    const CodeId id = classify(code);
    if (id != CodeId::K7_0113_0171 && id != CodeId::K7_0133_0171) {
        setError("ced_decode_batch_softq: K=7 rate-1/2 codes 0113/0171 and 0133/0171 only");
        return CED_ERR_UNSUPPORTED;
    }
    const int T = frameBits + ced::kTailSteps;
    if (symStride < (size_t)T || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch_softq: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    cedStopPacketServer();
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const uint4 *table = nullptr;
    int rc = id == CodeId::K7_0113_0171 ? softqTableFor<Code0113>(c, 0, s, &table) : softqTableFor<Code0133>(c, 1, s, &table);
    if (rc != CED_OK)
        return rc;
    const DecodeWorkingSet ws = decodeWorkingSet((size_t)nFrames, T, c->maxWaveFrames);
    ced_ctx::Work &wk = c->work[slot];
    if (wk.scratch.bytes < ws.scratchBytes || wk.schedState.bytes < ws.stateBytes || wk.schedFlags.bytes < ws.flagBytes) {
        CED_CUDA(cudaDeviceSynchronize());
        rc = wk.scratch.ensure(ws.scratchBytes);
        if (rc == CED_OK) rc = wk.schedState.ensure(ws.stateBytes);
        if (rc == CED_OK) rc = wk.schedFlags.ensure(ws.flagBytes);
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    const bool aligned16 = (reinterpret_cast<uintptr_t>(dSyms) & 15u) == 0 && (symStride & 15u) == 0;
    c->profWaves = 0;
    for (size_t f0 = 0; f0 < (size_t)nFrames; f0 += ws.waveMax) {
        const bool prof = c->profiling && c->profWaves < ced_ctx::kMaxProfWaves;
        const int pw = c->profWaves;
        const int wave = (int)std::min<size_t>(ws.waveMax, (size_t)nFrames - f0);
        const int groups = (wave + 31) / 32;
        const int perSm = std::max(3, std::min({5, c->fwdResident, groups / (4 * c->sms)}));
        const int blocks = std::max(1, std::min(c->sms * perSm, (groups + 3) / 4));
        ced::FwdSched sched;
        sched.counter = reinterpret_cast<unsigned int *>(wk.schedFlags.p);
        sched.done = wk.schedFlags.p + 1;
        sched.state = wk.schedState.p;
        CED_CUDA(cudaMemsetAsync(wk.schedFlags.p, 0, (size_t)(groups + 1) * sizeof(int), s));
        if (prof)
            CED_CUDA(cudaEventRecord(c->prof[pw][0], s));
        const uint8_t *in = dSyms + f0 * symStride;
#define CED_SOFTQ(CODE, AL)                                                                                             \
    ced::k7SoftQForwardKernel<CODE, AL><<<blocks, ced::kFwdThreads, 0, s>>>(in, symStride, wave, T, wk.scratch.p, table,   \
                                                                          c->bm0113.minusOne, sched, 2)
        if (id == CodeId::K7_0113_0171) {
            if (aligned16) CED_SOFTQ(Code0113, true); else CED_SOFTQ(Code0113, false);
        } else {
            if (aligned16) CED_SOFTQ(Code0133, true); else CED_SOFTQ(Code0133, false);
        }
#undef CED_SOFTQ
        if (prof)
            CED_CUDA(cudaEventRecord(c->prof[pw][1], s));
        ced::k7TracebackKernel<ced::Lanes8><<<(wave + ced::kTbThreads - 1) / ced::kTbThreads, ced::kTbThreads, 0, s>>>(
            wk.scratch.p, wave, T, dOut + f0 * outStride, outStride);
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

/* forward pass of one slice of continuous streams (decodeWindowImpl, ced_abi.cu); the caller holds c->mu */
int cedSoftQForwardWindow(ced_ctx *c, const ced_code_t *code, bool aligned16, int blocks, cudaStream_t s, const uint8_t *in,
                          size_t symStride, int wave, int nSegments, uint4 *scratch, ced::FwdSched sched, ced::FwdWindow win)
{
    const CodeId id = classify(code);
    if (id != CodeId::K7_0113_0171 && id != CodeId::K7_0133_0171) {
        setError("ced_decode_window_batch_softq: K=7 rate-1/2 codes 0113/0171 and 0133/0171 only");
        return CED_ERR_UNSUPPORTED;
    }
    const uint4 *table = nullptr;
    const int rc = id == CodeId::K7_0113_0171 ? softqTableFor<Code0113>(c, 0, s, &table) : softqTableFor<Code0133>(c, 1, s, &table);
    if (rc != CED_OK)
        return rc;
#define CED_SOFTQ_WIN(CODE, AL)                                                                                         \
    ced::k7SoftQForwardKernel<CODE, AL, true><<<blocks, ced::kFwdThreads, 0, s>>>(in, symStride, wave, nSegments, scratch,  \
                                                                                table, c->bm0113.minusOne, sched, 2, win)
    if (id == CodeId::K7_0113_0171) {
        if (aligned16) CED_SOFTQ_WIN(Code0113, true); else CED_SOFTQ_WIN(Code0113, false);
    } else {
        if (aligned16) CED_SOFTQ_WIN(Code0133, true); else CED_SOFTQ_WIN(Code0133, false);
    }
#undef CED_SOFTQ_WIN
    return CED_OK;
}

extern "C" {

int ced_quantize_soft(ced_ctx *c, const int8_t *dSoft, size_t softStride, int nFrames, int segsPerFrame, double delta,
                      uint8_t *dSyms, size_t symStride, void *stream)
{
    if (!c || !dSoft || !dSyms || nFrames < 0 || segsPerFrame <= 0 || softStride < (size_t)2 * segsPerFrame ||
        symStride < (size_t)segsPerFrame || !(delta > 0.0)) {
        setError("ced_quantize_soft: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const long long work = (long long)nFrames * segsPerFrame;
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 16);
    ced::quantizeSoftKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(dSoft, softStride, nFrames, segsPerFrame,
                                                                                        dSyms, symStride, (float)(1.0 / delta));
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

} // extern "C"
