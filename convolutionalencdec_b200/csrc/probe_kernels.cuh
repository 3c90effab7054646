/*
 * probe_kernels.cuh -- measurement only: the INT-ALU roofline denominator.
 * SURVEY 8(d) defines INT_peak as the rate of a dependent-free IADD3/LOP3 stream
 * on the ALU pipe; MEASURED_PEAKS.json has no such number, so bench.py measures
 * it on the box with this kernel (mode 0), and separately the rate with IMAD
 * co-issued on the FMA pipe (mode 1) for the notes in DESIGN.md.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ced {

constexpr int kProbeChains = 8;
constexpr int kProbeUnroll = 16;

template <int MODE>
__global__ void __launch_bounds__(256) intProbeKernel(uint32_t *out, int iters, uint32_t b, uint32_t c)
{
    uint32_t a[kProbeChains];
#pragma unroll
    for (int j = 0; j < kProbeChains; j++)
        a[j] = threadIdx.x * 2654435761u + j;
    uint32_t m[kProbeChains];
#pragma unroll
    for (int j = 0; j < kProbeChains; j++)
        m[j] = threadIdx.x + 17u * j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < kProbeUnroll; u++) {
#pragma unroll
            for (int j = 0; j < kProbeChains; j++) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
                if (MODE == 1)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[j]) : "r"(b), "r"(c));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < kProbeChains; j++)
        acc ^= a[j] ^ m[j];
    if (acc == 0x12345678u)
        out[0] = acc;
}

} // namespace ced
