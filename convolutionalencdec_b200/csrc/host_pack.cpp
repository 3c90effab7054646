/*
 * host_pack.cpp -- transfer compression for ced_decode_batch_host (include/ced_abi.h).
 *
 * The reference wire format spends one byte on a 2-bit segment (src/viterbiDecoder.h:154), so a host
 * buffer of symbols crosses PCIe at 4x the necessary size (269 MB for 2^16 x 4096-bit frames: 5 ms at
 * the ~54 GB/s this link delivers, against 1.6 ms of decoding).  Before the H2D copy the host
 * pipeline therefore packs each chunk to the 4-segments-per-byte format of ced_decode_batch_packed
 * with a small pool of host threads (AVX-512 or AVX2 when the CPU has it).  This is data movement only: no
 * encode/decode arithmetic happens on the host.
 */
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace ced_host {

/* byte-per-segment row -> packed row (segment t in bits 2*(t%4).. of byte t/4), low 2 bits of each byte */
static void packRowScalar(const uint8_t *in, int segs, uint8_t *out)
{
    int t = 0, o = 0;
    for (; t + 8 <= segs; t += 8, o += 2) {
        uint64_t w;
        memcpy(&w, in + t, 8);
        w &= 0x0303030303030303ull;
        const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
        out[o] = (uint8_t)(lo | (lo >> 6) | (lo >> 12) | (lo >> 18));
        out[o + 1] = (uint8_t)(hi | (hi >> 6) | (hi >> 12) | (hi >> 18));
    }
    for (; t < segs; t += 4, o++) {
        uint8_t b = 0;
        for (int j = 0; j < 4 && t + j < segs; j++)
            b |= (uint8_t)((in[t + j] & 3u) << (2 * j));
        out[o] = b;
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static void packRowAvx2(const uint8_t *in, int segs, uint8_t *out)
{
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i mul1 = _mm256_set1_epi16(0x0401);      /* s0 + 4*s1 per byte pair            */
    const __m256i mul2 = _mm256_set1_epi32(0x00100001);  /* (s0+4 s1) + 16*(s2+4 s3) per quad   */
    int t = 0, o = 0;
    for (; t + 32 <= segs; t += 32, o += 8) {
        __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(in + t)), three);
        v = _mm256_maddubs_epi16(v, mul1);
        v = _mm256_madd_epi16(v, mul2);                  /* 8 x int32, each one packed byte */
        v = _mm256_packus_epi32(v, v);
        v = _mm256_packus_epi16(v, v);                   /* per 128-bit lane: 4 bytes in the low dword */
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(v, 0), hi = (uint32_t)_mm256_extract_epi32(v, 4);
        memcpy(out + o, &lo, 4);
        memcpy(out + o + 4, &hi, 4);
    }
    if (t < segs)
        packRowScalar(in + t, segs - t, out + o);
}

/* 64 segments -> 16 bytes per round: two multiply-adds and one narrowing move (VPMOVDB) */
__attribute__((target("avx512f,avx512bw"))) static void packRowAvx512(const uint8_t *in, int segs, uint8_t *out)
{
    const __m512i three = _mm512_set1_epi8(3);
    const __m512i mul1 = _mm512_set1_epi16(0x0401);
    const __m512i mul2 = _mm512_set1_epi32(0x00100001);
    int t = 0, o = 0;
    for (; t + 64 <= segs; t += 64, o += 16) {
        __m512i v = _mm512_and_si512(_mm512_loadu_si512(reinterpret_cast<const void *>(in + t)), three);
        v = _mm512_maddubs_epi16(v, mul1);
        v = _mm512_madd_epi16(v, mul2);                  /* 16 x int32, each one packed byte */
        _mm_storeu_si128(reinterpret_cast<__m128i *>(out + o), _mm512_cvtepi32_epi8(v));
    }
    if (t < segs)
        packRowAvx2(in + t, segs - t, out + o);
}
#endif

using PackRowFn = void (*)(const uint8_t *, int, uint8_t *);

static PackRowFn pickPackRow()
{
#if defined(__x86_64__)
    const char *env = getenv("CED_HOST_PACK_ISA");   /* "avx2" keeps the 256-bit loop on an AVX-512 host */
    if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") && !(env && !strcmp(env, "avx2")))
        return packRowAvx512;
    if (__builtin_cpu_supports("avx2"))
        return packRowAvx2;
#endif
    return packRowScalar;
}

/* Minimal fork-join pool: run(fn) calls fn(worker, nWorkers) on every worker incl. the caller. */
class Pool {
public:
    explicit Pool(int n) : n_(n < 1 ? 1 : n)
    {
        for (int i = 1; i < n_; i++)
            threads_.emplace_back([this, i] { loop(i); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &t : threads_)
            t.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int, int)> &fn)
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_ = &fn;
            pending_ = n_ - 1;
            gen_++;
        }
        cv_.notify_all();
        fn(0, n_);
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *fn;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_)
                    return;
                fn = fn_;
            }
            if (fn)
                (*fn)(id, n_);
            {
                std::lock_guard<std::mutex> l(mu_);
                if (--pending_ == 0)
                    done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *fn_ = nullptr;
    int pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct Packer {
    Pool pool;
    PackRowFn fn;
    explicit Packer(int threads) : pool(threads), fn(pickPackRow()) {}
};

Packer *packerCreate(int threads)
{
    return new Packer(threads);
}
void packerDestroy(Packer *p)
{
    delete p;
}
int packerThreads(const Packer *p)
{
    return p->pool.size();
}

/* rows [0, nRows) of `in` (stride inStride, segs valid bytes) -> `out` (stride outStride) */
void packerRun(Packer *p, const uint8_t *in, size_t inStride, int nRows, int segs, uint8_t *out, size_t outStride)
{
    PackRowFn fn = p->fn;
    p->pool.run([=](int w, int nw) {
        const int lo = (int)((long long)nRows * w / nw), hi = (int)((long long)nRows * (w + 1) / nw);
        for (int r = lo; r < hi; r++)
            fn(in + (size_t)r * inStride, segs, out + (size_t)r * outStride);
    });
}

} // namespace ced_host

/* C ABI (include/ced_abi.h): pack byte-per-segment rows on the host, e.g. to feed
 * ced_decode_batch_packed_host or to ship packed symbols over a network. */
extern "C" int ced_host_pack_symbols(const uint8_t *segs, size_t segStride, int nFrames, int segsPerFrame,
                                     uint8_t *packed, size_t packedStride, int threads)
{
    if (!segs || !packed || nFrames < 0 || segsPerFrame <= 0 || segStride < (size_t)segsPerFrame ||
        packedStride < (size_t)(segsPerFrame + 3) / 4)
        return -2; /* CED_ERR_ARG */
    ced_host::Packer p(threads < 1 ? 1 : (threads > 64 ? 64 : threads));
    ced_host::packerRun(&p, segs, segStride, nFrames, segsPerFrame, packed, packedStride);
    return 0;
}
