/*
 * multi_gpu.cu -- single-process multi-GPU entry points of include/ced_abi.h: a C caller of
 * the drop-in library shards one host batch over the GPUs of the box without Python, torchrun or MPI.
 *
 *   ced_multi_create                one ced_ctx per device + one persistent worker thread per device, pinned to
 *                                   the CPUs of the device's NUMA node so that its page-locked staging buffers
 *                                   (allocated by that thread on first use) are NUMA-local
 *   ced_{decode,encode}_batch_host_multi   frames [g*N/G, (g+1)*N/G) of the caller's host arrays go to device g
 *                                   through that device's own host pipeline (ced_decode_batch_host); no data-path
 *                                   collective -- frames are independent (src/viterbiDecoderButterflyk1.c:259)
 *   ced_ber_allreduce               the ONE collective of BER mode: ncclAllReduce(sum) of the uint64 counters of
 *                                   every device over a communicator made by ncclCommInitAll (NVLink / NVSwitch).
 *                                   NCCL is opened at first use (dlopen), so the library loads without it.
 *   ced_probe_copy_ceiling          raw page-locked H2D / D2H copy rate of all devices at once: the ceiling the
 *                                   host-buffer (e2e) figures are quoted against
 */
#include "ced_internal.cuh"

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <nccl.h>
#include <sched.h>
#include <string>
#include <thread>

namespace {

/* ---- NCCL, bound at first use ---- */
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*commInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*groupStart)() = nullptr;
    ncclResult_t (*groupEnd)() = nullptr;
    ncclResult_t (*allReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*getErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*getVersion)(int *) = nullptr;
};

NcclApi *ncclApi()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        /* a process that already holds NCCL (torch brings its own libnccl.so.2) gets that copy back */
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle)
                break;
        }
        if (!api.handle)
            return;
        api.commInitAll = reinterpret_cast<decltype(api.commInitAll)>(dlsym(api.handle, "ncclCommInitAll"));
        api.commDestroy = reinterpret_cast<decltype(api.commDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.groupStart = reinterpret_cast<decltype(api.groupStart)>(dlsym(api.handle, "ncclGroupStart"));
        api.groupEnd = reinterpret_cast<decltype(api.groupEnd)>(dlsym(api.handle, "ncclGroupEnd"));
        api.allReduce = reinterpret_cast<decltype(api.allReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.getErrorString = reinterpret_cast<decltype(api.getErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.getVersion = reinterpret_cast<decltype(api.getVersion)>(dlsym(api.handle, "ncclGetVersion"));
        if (!api.commInitAll || !api.commDestroy || !api.groupStart || !api.groupEnd || !api.allReduce) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api.handle ? &api : nullptr;
}

/* CPUs of the NUMA node a device hangs off (sysfs), empty if unknown */
std::vector<int> numaCpusOfDevice(int device)
{
    std::vector<int> cpus;
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return cpus;
    }
    for (char *p = bus; *p; p++)
        *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE *f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1)
            node = -1;
        fclose(f);
    }
    if (node < 0)
        return cpus;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    if (FILE *f = fopen(path, "r")) {
        int a = 0, b = 0;
        while (fscanf(f, "%d", &a) == 1) {
            b = a;
            int ch = fgetc(f);
            if (ch == '-') {
                if (fscanf(f, "%d", &b) != 1)
                    b = a;
                ch = fgetc(f);
            }
            for (int i = a; i <= b; i++)
                cpus.push_back(i);
            if (ch != ',')
                break;
        }
        fclose(f);
    }
    return cpus;
}

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool hasJob = false, done = false, quit = false;
    int rc = CED_OK;
    char err[512] = "";
};

} // namespace

/* Every device gets `lanes` contexts, each with its own worker thread and host pipeline: while one lane drains
 * its pipeline the other one's H2D copies keep the link busy (two callers per device: 57 -> 76 Gbit/s on the
 * 1-GPU box).  Worker w serves device w / lanes; ctx[w] belongs to it. */
struct ced_multi {
    std::vector<int> devices;
    int lanes = 2;
    std::vector<ced_ctx *> ctx;      /* [device * lanes + lane] */
    std::vector<Worker *> workers;   /* one per context */
    std::vector<ncclComm_t> comms;   /* made by the first ced_ber_allreduce, one per device */
    std::mutex mu;
};

namespace {

void workerLoop(ced_multi *m, int i)
{
    Worker *w = m->workers[(size_t)i];
    const int device = m->devices[(size_t)(i / m->lanes)];
    const std::vector<int> cpus = numaCpusOfDevice(device);
    static const bool noPin = getenv("CED_MULTI_NO_PIN") != nullptr;
    if (!cpus.empty() && !noPin) {
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c : cpus)
            if (c < CPU_SETSIZE)
                CPU_SET(c, &set);
        sched_setaffinity(0, sizeof(set), &set); /* best effort: staging this thread allocates is then NUMA-local */
    }
    cudaSetDevice(device);
    for (;;) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lock(w->mu);
            w->cv.wait(lock, [&] { return w->hasJob || w->quit; });
            if (w->quit)
                return;
            job = w->job;
        }
        const int rc = job();
        {
            std::lock_guard<std::mutex> lock(w->mu);
            w->rc = rc;
            if (rc != CED_OK)
                snprintf(w->err, sizeof(w->err), "%s", ced_last_error());
            w->hasJob = false;
            w->done = true;
        }
        w->cv.notify_all();
    }
}

/* run job(i) on every device's worker and wait for all; first failure wins */
int runOnAll(ced_multi *m, const std::function<int(int)> &job)
{
    std::lock_guard<std::mutex> lock(m->mu);
    const int n = (int)m->workers.size();
    for (int i = 0; i < n; i++) {
        Worker *w = m->workers[(size_t)i];
        {
            std::lock_guard<std::mutex> wl(w->mu);
            w->job = [job, i] { return job(i); };
            w->hasJob = true;
            w->done = false;
        }
        w->cv.notify_all();
    }
    int rc = CED_OK;
    for (int i = 0; i < n; i++) {
        Worker *w = m->workers[(size_t)i];
        std::unique_lock<std::mutex> wl(w->mu);
        w->cv.wait(wl, [&] { return w->done; });
        if (w->rc != CED_OK && rc == CED_OK) {
            rc = w->rc;
            setError("device %d: %s", m->devices[(size_t)(i / m->lanes)], w->err);
        }
    }
    return rc;
}

} // namespace

extern "C" {

void ced_shard_range(int nFrames, int nShards, int shard, int *first, int *count)
{
    /* contiguous ranges, the first (nFrames mod nShards) shards one frame longer */
    if (nShards <= 0 || shard < 0 || shard >= nShards || nFrames < 0) {
        if (first) *first = 0;
        if (count) *count = 0;
        return;
    }
    const int base = nFrames / nShards, extra = nFrames % nShards;
    if (first) *first = shard * base + (shard < extra ? shard : extra);
    if (count) *count = base + (shard < extra ? 1 : 0);
}

int ced_multi_create(const int *devices, int nDevices, ced_multi **out)
{
    if (!out || nDevices < 0 || (nDevices > 0 && !devices)) {
        setError("ced_multi_create: bad argument");
        return CED_ERR_ARG;
    }
    *out = nullptr;
    int visible = 0;
    CED_CUDA(cudaGetDeviceCount(&visible));
    ced_multi *m = new ced_multi();
    if (nDevices == 0)
        for (int d = 0; d < visible; d++)
            m->devices.push_back(d);
    else
        m->devices.assign(devices, devices + nDevices);
    const char *envLanes = getenv("CED_MULTI_LANES");
    m->lanes = envLanes ? std::max(1, std::min(atoi(envLanes), 4)) : 2;
    for (int d : m->devices)
        for (int l = 0; l < m->lanes; l++) {
            ced_ctx *c = nullptr;
            const int rc = ced_ctx_create(d, &c);
            if (rc != CED_OK) {
                ced_multi_destroy(m);
                return rc;
            }
            m->ctx.push_back(c);
        }
    for (size_t i = 0; i < m->ctx.size(); i++)
        m->workers.push_back(new Worker());
    for (size_t i = 0; i < m->ctx.size(); i++)
        m->workers[i]->th = std::thread(workerLoop, m, (int)i);
    *out = m;
    return CED_OK;
}

void ced_multi_destroy(ced_multi *m)
{
    if (!m)
        return;
    for (Worker *w : m->workers) {
        {
            std::lock_guard<std::mutex> lock(w->mu);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable())
            w->th.join();
        delete w;
    }
    if (!m->comms.empty())
        if (NcclApi *api = ncclApi())
            for (ncclComm_t c : m->comms)
                api->commDestroy(c);
    for (ced_ctx *c : m->ctx)
        ced_ctx_destroy(c);
    delete m;
}

int ced_multi_device_count(const ced_multi *m)
{
    return m ? (int)m->devices.size() : 0;
}

ced_ctx *ced_multi_ctx(ced_multi *m, int i)
{
    return (m && i >= 0 && i < (int)m->devices.size()) ? m->ctx[(size_t)(i * m->lanes)] : nullptr;
}

int ced_decode_batch_host_multi(ced_multi *m, const ced_code_t *code, const uint8_t *hSegs, size_t segStride,
                                int nFrames, int frameBits, uint8_t *hOut, size_t outStride)
{
    if (!m || m->ctx.empty() || !code || !hSegs || !hOut || nFrames < 0) {
        setError("ced_decode_batch_host_multi: bad argument");
        return CED_ERR_ARG;
    }
    const int G = (int)m->ctx.size(); /* devices x lanes shards, device-major: a device's lanes take neighbours */
    return runOnAll(m, [=](int g) {
        int first = 0, count = 0;
        ced_shard_range(nFrames, G, g, &first, &count);
        if (count == 0)
            return CED_OK;
        return ced_decode_batch_host(m->ctx[(size_t)g], code, hSegs + (size_t)first * segStride, segStride, count, frameBits,
                                     hOut + (size_t)first * outStride, outStride);
    });
}

int ced_encode_batch_host_multi(ced_multi *m, const ced_code_t *code, const uint8_t *hMsg, size_t msgStride, int nFrames,
                                int frameBytes, uint8_t *hSegs, size_t segStride)
{
    if (!m || m->ctx.empty() || !code || !hMsg || !hSegs || nFrames < 0) {
        setError("ced_encode_batch_host_multi: bad argument");
        return CED_ERR_ARG;
    }
    const int G = (int)m->ctx.size();
    return runOnAll(m, [=](int g) {
        int first = 0, count = 0;
        ced_shard_range(nFrames, G, g, &first, &count);
        if (count == 0)
            return CED_OK;
        return ced_encode_batch_host(m->ctx[(size_t)g], code, hMsg + (size_t)first * msgStride, msgStride, count, frameBytes,
                                     hSegs + (size_t)first * segStride, segStride);
    });
}

int ced_ber_allreduce(ced_multi *m, uint64_t *const *dCounters, int count)
{
    if (!m || m->ctx.empty() || !dCounters || count <= 0) {
        setError("ced_ber_allreduce: bad argument");
        return CED_ERR_ARG;
    }
    NcclApi *api = ncclApi();
    if (!api) {
        setError("ced_ber_allreduce: libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "dlopen failed");
        return CED_ERR_UNSUPPORTED;
    }
    std::lock_guard<std::mutex> lock(m->mu);
    const int G = (int)m->devices.size();
    auto devCtx = [&](int g) { return m->ctx[(size_t)(g * m->lanes)]; };
#define CED_NCCL(expr)                                                                                        \
    do {                                                                                                      \
        ncclResult_t r__ = (expr);                                                                            \
        if (r__ != ncclSuccess) {                                                                             \
            setError("%s failed: %s", #expr, api->getErrorString ? api->getErrorString(r__) : "NCCL error");  \
            return CED_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)
    if (m->comms.empty()) {
        m->comms.resize((size_t)G);
        const ncclResult_t r = api->commInitAll(m->comms.data(), G, m->devices.data());
        if (r != ncclSuccess) {
            m->comms.clear();
            setError("ncclCommInitAll failed: %s", api->getErrorString ? api->getErrorString(r) : "NCCL error");
            return CED_ERR_CUDA;
        }
    }
    CED_NCCL(api->groupStart());
    for (int g = 0; g < G; g++) {
        if (!dCounters[g]) {
            api->groupEnd();
            setError("ced_ber_allreduce: NULL counter pointer for device %d", m->devices[(size_t)g]);
            return CED_ERR_ARG;
        }
        CED_NCCL(api->allReduce(dCounters[g], dCounters[g], (size_t)count, ncclUint64, ncclSum, m->comms[(size_t)g],
                                devCtx(g)->stream));
    }
    CED_NCCL(api->groupEnd());
#undef CED_NCCL
    for (int g = 0; g < G; g++) {
        CED_CUDA(cudaSetDevice(m->devices[(size_t)g]));
        CED_CUDA(cudaStreamSynchronize(devCtx(g)->stream));
        devCtx(g)->launches += 1;
    }
    return CED_OK;
}

/* device memory for C callers that keep batches resident (BER mode); plain cudaMalloc / cudaMemcpy on the
 * context's device and stream */
int ced_device_alloc(ced_ctx *c, size_t bytes, void **out)
{
    if (!c || !out || bytes == 0) {
        setError("ced_device_alloc: bad argument");
        return CED_ERR_ARG;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    if (cudaMalloc(out, bytes) != cudaSuccess) {
        cudaGetLastError();
        setError("ced_device_alloc: cudaMalloc(%zu) failed", bytes);
        return CED_ERR_NOMEM;
    }
    CED_CUDA(cudaMemsetAsync(*out, 0, bytes, c->stream));
    return CED_OK;
}

void ced_device_free(ced_ctx *c, void *p)
{
    if (!c || !p)
        return;
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(p);
}

int ced_copy_to_device(ced_ctx *c, void *dDst, const void *hSrc, size_t bytes)
{
    if (!c || !dDst || !hSrc) {
        setError("ced_copy_to_device: bad argument");
        return CED_ERR_ARG;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    CED_CUDA(cudaMemcpyAsync(dDst, hSrc, bytes, cudaMemcpyHostToDevice, c->stream));
    CED_CUDA(cudaStreamSynchronize(c->stream));
    return CED_OK;
}

int ced_copy_to_host(ced_ctx *c, void *hDst, const void *dSrc, size_t bytes)
{
    if (!c || !hDst || !dSrc) {
        setError("ced_copy_to_host: bad argument");
        return CED_ERR_ARG;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    CED_CUDA(cudaMemcpyAsync(hDst, dSrc, bytes, cudaMemcpyDeviceToHost, c->stream));
    CED_CUDA(cudaStreamSynchronize(c->stream));
    return CED_OK;
}

int ced_nccl_version(void)
{
    NcclApi *api = ncclApi();
    int v = 0;
    if (api && api->getVersion)
        api->getVersion(&v);
    return v;
}

/* Raw copy rate between page-locked host memory and one device (no kernels): `bytes` per direction, best of
 * `reps` after one warm-up; both directions separately.  Several callers (threads or ranks) running this at the
 * same time measure the rate the host can sustain for all of them together. */
int ced_probe_copy_ceiling(ced_ctx *c, size_t bytes, int reps, double *h2dBytesPerSecond, double *d2hBytesPerSecond)
{
    if (!c || bytes == 0 || reps <= 0 || !h2dBytesPerSecond || !d2hBytesPerSecond) {
        setError("ced_probe_copy_ceiling: bad argument");
        return CED_ERR_ARG;
    }
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    CED_CUDA(cudaSetDevice(c->device));
    void *host = nullptr, *dev = nullptr;
    CED_CUDA(cudaMallocHost(&host, bytes));
    memset(host, 1, bytes);
    if (cudaMalloc(&dev, bytes) != cudaSuccess) {
        cudaFreeHost(host);
        setError("ced_probe_copy_ceiling: cudaMalloc(%zu) failed", bytes);
        return CED_ERR_NOMEM;
    }
    double best[2] = {0, 0};
    for (int dir = 0; dir < 2; dir++)
        for (int r = 0; r <= reps; r++) {
            const auto t0 = std::chrono::steady_clock::now();
            if (dir == 0)
                cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->h2d);
            else
                cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->d2h);
            cudaStreamSynchronize(dir == 0 ? c->h2d : c->d2h);
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (r > 0)
                best[dir] = std::max(best[dir], (double)bytes / sec);
        }
    cudaFree(dev);
    cudaFreeHost(host);
    CED_CUDA(cudaGetLastError());
    *h2dBytesPerSecond = best[0];
    *d2hBytesPerSecond = best[1];
    return CED_OK;
}

/* the same probe on every device of a ced_multi at once (one worker thread each): aggregate bytes per second */
int ced_multi_probe_copy_ceiling(ced_multi *m, size_t bytesPerDevice, int reps, double *h2dBytesPerSecond,
                                 double *d2hBytesPerSecond)
{
    if (!m || m->ctx.empty() || !h2dBytesPerSecond || !d2hBytesPerSecond) {
        setError("ced_multi_probe_copy_ceiling: bad argument");
        return CED_ERR_ARG;
    }
    const int G = (int)m->ctx.size(), lanes = m->lanes;
    std::vector<double> up((size_t)G, 0.0), down((size_t)G, 0.0);
    const int rc = runOnAll(m, [&](int g) {
        if (g % lanes) /* one copy stream per device saturates its link */
            return CED_OK;
        return ced_probe_copy_ceiling(m->ctx[(size_t)g], bytesPerDevice, reps, &up[(size_t)g], &down[(size_t)g]);
    });
    if (rc != CED_OK)
        return rc;
    *h2dBytesPerSecond = *d2hBytesPerSecond = 0.0;
    for (int g = 0; g < G; g++) {
        *h2dBytesPerSecond += up[(size_t)g];
        *d2hBytesPerSecond += down[(size_t)g];
    }
    return CED_OK;
}

} // extern "C"
