/*
 * soft_decode.cu -- ABI entry points of the soft-decision path (include/ced_abi.h): ced_decode_batch_soft,
 * ced_awgn_channel, ced_slice_soft_to_bytes.  Kernels in soft_decode.cuh; nothing here computes on the host.
 */
#include "ced_internal.cuh"
#include "soft_decode.cuh"

extern "C" {

int ced_decode_batch_soft(ced_ctx *c, const ced_code_t *code, const int8_t *dSoft, size_t softStride, int nFrames,
                          int frameBits, uint8_t *dOut, size_t outStride, void *stream)
{
    if (!c || nFrames < 0 || frameBits <= 0 || (frameBits & 7) || (nFrames > 0 && (!dSoft || !dOut))) {
        setError("ced_decode_batch_soft: bad argument (frameBits must be a positive multiple of 8)");
        return CED_ERR_ARG;
    }
    const CodeId id = classify(code);
    if (id != CodeId::K7_0113_0171 && id != CodeId::K7_0133_0171) {
        setError("ced_decode_batch_soft: K=7 rate-1/2 codes 0113/0171 and 0133/0171 only");
        return CED_ERR_UNSUPPORTED;
    }
    const int T = frameBits + ced::kTailSteps;
    if (softStride < (size_t)2 * T || outStride < (size_t)(frameBits / 8)) {
        setError("ced_decode_batch_soft: stride shorter than a frame");
        return CED_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(dSoft) & 15u) || (softStride & 15u)) {
        setError("ced_decode_batch_soft: soft symbol rows must start on 16-byte boundaries");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    DecodeWorkingSet ws = decodeWorkingSet((size_t)nFrames, T, c->maxWaveFrames);
    ws.stateBytes = ws.firstGroups * ced::kSoftStateUint4 * 32 * sizeof(uint4);
    ced_ctx::Work &wk = c->work[0];
    if (wk.scratch.bytes < ws.scratchBytes || wk.schedState.bytes < ws.stateBytes || wk.schedFlags.bytes < ws.flagBytes) {
        CED_CUDA(cudaDeviceSynchronize());
        int rc = wk.scratch.ensure(ws.scratchBytes);
        if (rc == CED_OK) rc = wk.schedState.ensure(ws.stateBytes);
        if (rc == CED_OK) rc = wk.schedFlags.ensure(ws.flagBytes);
        if (rc != CED_OK)
            return rc;
    }
    if (wk.lastStream && wk.lastStream != s)
        CED_CUDA(cudaStreamWaitEvent(s, wk.idle, 0));
    /* persistent grid: the 16-bit kernel holds 32 metric registers, so fewer CTAs are resident than for the
     * hard kernel; never more warps than 32-frame groups */
    static int resident = 0;
    if (!resident) {
        CED_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, ced::k7SoftForwardKernel<Code0113>,
                                                               ced::kFwdThreads, 0));
        resident = std::max(1, resident);
    }
    static const int envPerSm = getenv("CED_SOFT_BLOCKS_PER_SM") ? atoi(getenv("CED_SOFT_BLOCKS_PER_SM")) : 0;
    static const int envCpu = getenv("CED_SOFT_CHUNKS_PER_UNIT") ? atoi(getenv("CED_SOFT_CHUNKS_PER_UNIT")) : 0;
    const int chunksPerUnit = envCpu > 0 ? envCpu : 4; /* 192 steps between hand-offs, as in the hard kernel */
    c->profWaves = 0;
    for (size_t f0 = 0; f0 < (size_t)nFrames; f0 += ws.waveMax) {
        const bool prof = c->profiling && c->profWaves < ced_ctx::kMaxProfWaves;
        const int pw = c->profWaves;
        const int wave = (int)std::min<size_t>(ws.waveMax, (size_t)nFrames - f0);
        const int groups = (wave + 31) / 32;
        const int perSm = envPerSm > 0 ? std::min(envPerSm, resident) : std::max(3, std::min({4, resident, groups / (4 * c->sms)}));
        const int blocks = std::max(1, std::min(c->sms * perSm, (groups + 3) / 4));
        ced::FwdSched sched;
        sched.counter = reinterpret_cast<unsigned int *>(wk.schedFlags.p);
        sched.done = wk.schedFlags.p + 1;
        sched.state = wk.schedState.p;
        CED_CUDA(cudaMemsetAsync(wk.schedFlags.p, 0, (size_t)(groups + 1) * sizeof(int), s));
        if (prof)
            CED_CUDA(cudaEventRecord(c->prof[pw][0], s));
        const int8_t *in = dSoft + f0 * softStride;
        if (id == CodeId::K7_0113_0171)
            ced::k7SoftForwardKernel<Code0113><<<blocks, ced::kFwdThreads, 0, s>>>(in, softStride, wave, T, wk.scratch.p,
                                                                                  c->bm0113.minusOne, sched, chunksPerUnit);
        else
            ced::k7SoftForwardKernel<Code0133><<<blocks, ced::kFwdThreads, 0, s>>>(in, softStride, wave, T, wk.scratch.p,
                                                                                  c->bm0113.minusOne, sched, chunksPerUnit);
        if (prof)
            CED_CUDA(cudaEventRecord(c->prof[pw][1], s));
        ced::k7TracebackKernel<ced::Lanes16><<<(wave + ced::kTbThreads - 1) / ced::kTbThreads, ced::kTbThreads, 0, s>>>(
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

int ced_awgn_channel(ced_ctx *c, const uint8_t *dSegs, size_t segStride, int nFrames, int segsPerFrame, int8_t *dSoft,
                     size_t softStride, double amplitude, double sigma, uint64_t seed, uint64_t firstFrameIndex,
                     uint64_t *dCounters, void *stream)
{
    if (!c || !dSegs || !dSoft || nFrames < 0 || segsPerFrame <= 0 || segStride < (size_t)segsPerFrame ||
        softStride < (size_t)2 * segsPerFrame || (softStride & 1u) || (reinterpret_cast<uintptr_t>(dSoft) & 1u) ||
        !(amplitude > 0.0) || !(sigma >= 0.0)) {
        setError("ced_awgn_channel: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const long long work = (long long)nFrames * segsPerFrame;
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 16);
    ced::awgnChannelKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        dSegs, segStride, nFrames, segsPerFrame, dSoft, softStride, (float)amplitude, (float)sigma, seed, firstFrameIndex,
        reinterpret_cast<unsigned long long *>(dCounters));
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

int ced_slice_soft_to_bytes(ced_ctx *c, const int8_t *dSoft, size_t softStride, int nFrames, int segsPerFrame,
                            uint8_t *dSegs, size_t segStride, void *stream)
{
    if (!c || !dSegs || !dSoft || nFrames < 0 || segsPerFrame <= 0 || segStride < (size_t)segsPerFrame ||
        softStride < (size_t)2 * segsPerFrame || (softStride & 1u) || (reinterpret_cast<uintptr_t>(dSoft) & 1u)) {
        setError("ced_slice_soft_to_bytes: bad argument");
        return CED_ERR_ARG;
    }
    if (nFrames == 0)
        return CED_OK;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    const long long work = (long long)nFrames * segsPerFrame;
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)c->sms * 16);
    ced::sliceSoftToBytesKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        dSoft, softStride, nFrames, segsPerFrame, dSegs, segStride);
    c->launches += 1;
    CED_CUDA(cudaGetLastError());
    return CED_OK;
}

} // extern "C"
