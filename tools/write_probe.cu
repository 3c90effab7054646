// Measurement only: how fast can this GPU WRITE a large buffer?  The encoder (encode_batch.cuh) is a write stream;
// MEASURED_PEAKS.json's figure is a copy (half reads).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/write_probe tools/write_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) fillKernel(uint4 *p, size_t n16, uint32_t v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 w = make_uint4(v, v + 1, v + 2, v + 3);
    if (MODE == 4) { // 32-byte stores, L2 evict-first
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < n16; i += stride) {
            asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p + 2 * i), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
        }
        return;
    }
    if (MODE == 3) { // 32-byte stores, two uint4 per thread adjacent (256-bit)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < n16; i += stride) {
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(p + 2 * i), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        if (MODE == 0)
            p[i] = w;
        else if (MODE == 1)
            __stcs(p + i, w);
        else if (MODE == 2)
            __stwt(p + i, w);
    }
}

// block-contiguous: each CTA owns a contiguous span (like the encoder's per-frame rows)
__global__ void __launch_bounds__(256) fillSpanKernel(uint4 *p, size_t n16, size_t span16, uint32_t v)
{
    uint4 w = make_uint4(v, v + 1, v + 2, v + 3);
    for (size_t s = blockIdx.x; s * span16 < n16; s += gridDim.x) {
        uint4 *q = p + s * span16;
        for (size_t i = threadIdx.x; i < span16 && s * span16 + i < n16; i += blockDim.x)
            q[i] = w;
    }
}

__global__ void __launch_bounds__(256) copyKernel(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        b[i] = a[i];
}
__global__ void __launch_bounds__(256) readKernel(const uint4 *__restrict__ a, uint4 *out, size_t n16)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 x = a[i];
        acc.x ^= x.x; acc.y ^= x.y; acc.z ^= x.z; acc.w ^= x.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u)
        out[0] = acc;
}

// the encoder's write pattern: one warp per row of `rowBytes` valid bytes at `stride`, lane l writes 32 contiguous
// bytes per 1 KB span as two 16-byte stores (MODE 0) or the warp writes 512 contiguous bytes per store (MODE 1);
// TAIL: the last (rowBytes % 32) bytes of a row go out as byte stores; READ: one LDG.32 per lane per span
template <int MODE, bool TAIL, bool READ>
__global__ void __launch_bounds__(256) rowKernel(uint8_t *p, size_t stride, long long rows, int rowBytes,
                                                 const uint32_t *__restrict__ src, uint32_t v)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += warps) {
        uint8_t *row = p + (size_t)r * stride;
        const int spans = (rowBytes + 16 + 1023) / 1024;
#pragma unroll 4
        for (int sp = 0; sp < spans; sp++) {
            uint32_t x = v;
            if (READ)
                x ^= __ldg(src + (size_t)r * 128 + (sp & 3) * 32 + lane);
            uint4 w = make_uint4(x, x + 1, x + 2, x + 3);
            if (MODE == 0) {
                const int o = 1024 * sp + 32 * lane;
                if (o + 32 <= rowBytes) {
                    *reinterpret_cast<uint4 *>(row + o) = w;
                    *reinterpret_cast<uint4 *>(row + o + 16) = w;
                } else if (TAIL) {
                    for (int b = o; b < rowBytes; b++)
                        row[b] = (uint8_t)x;
                }
            } else if (MODE == 3) { /* ragged end: read the 16-byte piece, merge, write it back whole */
                for (int h = 0; h < 2; h++) {
                    const int o = 1024 * sp + 512 * h + 16 * lane;
                    if (o >= rowBytes)
                        continue;
                    if (o + 16 <= rowBytes)
                        *reinterpret_cast<uint4 *>(row + o) = w;
                    else if (TAIL) {
                        uint4 old = __ldcg(reinterpret_cast<const uint4 *>(row + o));
                        const int cnt = rowBytes - o;
                        uint32_t ow[4] = {old.x, old.y, old.z, old.w}, nw[4] = {w.x, w.y, w.z, w.w};
                        for (int i = 0; i < 4; i++) {
                            const int nb = min(4, max(0, cnt - 4 * i));
                            const uint32_t m = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
                            ow[i] = (ow[i] & ~m) | (nw[i] & m);
                        }
                        *reinterpret_cast<uint4 *>(row + o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                }
            } else if (MODE == 2) {
                const int shift = (int)((reinterpret_cast<uintptr_t>(row) >> 4) & 1);
                for (int h = 0; h < 2; h++) {
                    const int o = 1024 * sp + 512 * h + 16 * lane - 16 * shift;
                    if (o < 0 || o >= rowBytes)
                        continue;
                    if (o + 16 <= rowBytes)
                        *reinterpret_cast<uint4 *>(row + o) = w;
                    else if (TAIL) {
                        int cnt = rowBytes - o, b = o;
                        if (cnt & 8) { *reinterpret_cast<uint2 *>(row + b) = make_uint2(x, x); b += 8; }
                        if (cnt & 4) { *reinterpret_cast<uint32_t *>(row + b) = x; b += 4; }
                        if (cnt & 2) { *reinterpret_cast<uint16_t *>(row + b) = (uint16_t)x; b += 2; }
                        if (cnt & 1) row[b] = (uint8_t)x;
                    }
                }
            } else {
                for (int h = 0; h < 2; h++) {
                    const int o = 1024 * sp + 512 * h + 16 * lane;
                    if (o + 16 <= rowBytes)
                        *reinterpret_cast<uint4 *>(row + o) = w;
                    else if (TAIL)
                        for (int b = o; b < rowBytes; b++)
                            row[b] = (uint8_t)x;
                }
            }
        }
    }
}

// phase 1 of the two-phase scheme: pull a chunk's source words (and each row's ragged last sector) into L2
__global__ void __launch_bounds__(256) warmL2Kernel(const uint4 *__restrict__ src, size_t n16, const uint8_t *p, size_t stride,
                                                    long long rows, int tailOff, uint32_t *sink)
{
    const size_t gstride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (size_t i = tid; i < n16; i += gstride) {
        const uint4 x = __ldcg(src + i);
        acc ^= x.x ^ x.y ^ x.z ^ x.w;
    }
    if (tailOff >= 0)
        for (size_t r = tid; r < (size_t)rows; r += gstride)
            asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p + r * stride + tailOff));
    if (acc == 0x1234567u)
        *sink = acc;
}

template <class F>
static double timeIt(F f, int reps = 10)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++)
        f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps * 1e-3;
}

int main(int argc, char **argv)
{
    for (size_t bytes : {(size_t)4300 << 20}) {
        uint4 *p, *q;
        if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMalloc(&q, bytes) != cudaSuccess) {
            printf("alloc failed\n");
            return 1;
        }
        const size_t n16 = bytes / 16;
        printf("buffer %.0f MB\n", bytes / 1e6);
        for (int grid : {148 * 8, 148 * 64}) {
            printf("  grid %5d:", grid);
            printf(" st %.0f", bytes / timeIt([&] { fillKernel<0><<<grid, 256>>>(p, n16, 1); }) / 1e9);
            printf(" st.cs %.0f", bytes / timeIt([&] { fillKernel<1><<<grid, 256>>>(p, n16, 1); }) / 1e9);
            printf(" st.wt %.0f", bytes / timeIt([&] { fillKernel<2><<<grid, 256>>>(p, n16, 1); }) / 1e9);
            printf(" st.v8 %.0f", bytes / timeIt([&] { fillKernel<3><<<grid, 256>>>(p, n16, 1); }) / 1e9);
            printf(" st.v8.evict_first %.0f", bytes / timeIt([&] { fillKernel<4><<<grid, 256>>>(p, n16, 1); }) / 1e9);
            printf(" span4KB %.0f", bytes / timeIt([&] { fillSpanKernel<<<grid, 256>>>(p, n16, 256, 1); }) / 1e9);
            printf(" span64KB %.0f", bytes / timeIt([&] { fillSpanKernel<<<grid, 256>>>(p, n16, 4096, 1); }) / 1e9);
            printf(" | copy(r+w) %.0f", 2.0 * bytes / timeIt([&] { copyKernel<<<grid, 256>>>(p, q, n16); }) / 1e9);
            printf(" read %.0f GB/s\n", bytes / timeIt([&] { readKernel<<<grid, 256>>>(p, q, n16); }) / 1e9);
        }
        printf("  cudaMemsetAsync %.0f GB/s\n", bytes / timeIt([&] { cudaMemsetAsync(p, 1, bytes); }) / 1e9);
        printf("  cudaMemcpyAsync D2D (r+w) %.0f GB/s\n", 2.0 * bytes / timeIt([&] { cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice); }) / 1e9);
        cudaFree(p);
        cudaFree(q);
    }
    {
        const long long rows = 1 << 20;
        uint8_t *p;
        uint32_t *src;
        cudaMalloc(&p, (size_t)rows * 4128);
        cudaMalloc(&src, (size_t)rows * 512);
        cudaMemset(src, 1, (size_t)rows * 512);
        printf("encoder write pattern, 2^20 rows (GB/s of valid bytes written):\n");
        for (int grid : {148 * 8, 148 * 16, 148 * 32}) {
            for (size_t stride : {(size_t)4096, (size_t)4112, (size_t)4128}) {
                const int full = 4096, withTail = stride == 4096 ? 4096 : 4102;
                const double b = (double)rows;
                printf("  grid %4d stride %zu: RMW tail %.0f, +read %.0f | aligned pieces %.0f, +tail %.0f, +tail+read %.0f | 32B/lane %.0f, +tail %.0f, +tail+read %.0f | 512B/warp %.0f, +tail %.0f, +tail+read %.0f\n", grid, stride,
                       b * withTail / timeIt([&] { rowKernel<3, true, false><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<3, true, true><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * full / timeIt([&] { rowKernel<2, false, false><<<grid, 256>>>(p, stride, rows, full, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<2, true, false><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<2, true, true><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * full / timeIt([&] { rowKernel<0, false, false><<<grid, 256>>>(p, stride, rows, full, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<0, true, false><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<0, true, true><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * full / timeIt([&] { rowKernel<1, false, false><<<grid, 256>>>(p, stride, rows, full, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<1, true, false><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9,
                       b * withTail / timeIt([&] { rowKernel<1, true, true><<<grid, 256>>>(p, stride, rows, withTail, src, 1); }) / 1e9);
            }
        }
        printf("two-phase (warm L2 with a chunk's source + tail sectors, then write the chunk), stride 4112, 4102-byte rows:\n");
        for (long long chunk : {8192LL, 16384LL, 32768LL, 65536LL})
            for (int tails : {0, 1})
                for (int mode : {1, 3}) {
                    const double t = timeIt([&] {
                        for (long long r0 = 0; r0 < rows; r0 += chunk) {
                            warmL2Kernel<<<148 * 4, 256>>>(reinterpret_cast<const uint4 *>(src + r0 * 128), (size_t)chunk * 32, p + r0 * 4112,
                                                           4112, chunk, tails ? 4096 : -1, src);
                            if (mode == 1)
                                rowKernel<1, true, true><<<148 * 32, 256>>>(p + r0 * 4112, 4112, chunk, 4102, src + r0 * 128, 1);
                            else
                                rowKernel<3, true, true><<<148 * 32, 256>>>(p + r0 * 4112, 4112, chunk, 4102, src + r0 * 128, 1);
                        }
                    }, 5);
                    printf("  chunk %6lld rows, tail sectors %s, %s: %.0f GB/s written\n", chunk, tails ? "prefetched" : "not prefetched",
                           mode == 1 ? "partial stores" : "RMW piece", (double)rows * 4102 / t / 1e9);
                }
        cudaFree(p);
        cudaFree(src);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
