// Dependent-chain latencies (cycles per instruction, one warp alone on an SM) of the integer instructions the
// sequential min-plus chain of frame_parallel.cuh can be built from.  nvcc -arch=sm_100a -o tools/_bin/latency_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int N = 512;
template <int OP>
__global__ void chain(uint32_t seed, uint32_t *out, long long *cycles)
{
    __shared__ uint32_t sm[64];
    sm[threadIdx.x] = threadIdx.x ^ 1;
    sm[threadIdx.x + 32] = threadIdx.x;
    __syncthreads();
    uint32_t x = seed + threadIdx.x, y = seed * 3 + 1, z = 0xFFFFFFFFu;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = __viaddmin_u16x2(x, y, z);
        if (OP == 1) x = __vminu2(x, y + i);
        if (OP == 2) x = min((int)x + (int)y, (int)z + i);           // IADD + VIMNMX
        if (OP == 3) x = __vimin3_s32((int)x, (int)y + i, (int)z - i);
        if (OP == 4) x = __byte_perm(x, y, 0x4140 + (i & 1));
        if (OP == 5) x = __shfl_xor_sync(0xFFFFFFFFu, x, 1);
        if (OP == 6) x = sm[x & 63];
        if (OP == 7) x = __viaddmin_s32((int)x, (int)y, (int)z + i);
        if (OP == 8) x = x + y + i;
        if (OP == 9) { sm[(threadIdx.x + i) & 63] = x; __syncwarp(); x = sm[(threadIdx.x + i + 1) & 63]; }
        if (OP == 10) { asm volatile("bar.sync 1, 32;" ::: "memory"); x += 1; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
template <int OP> void run(const char *name, uint32_t *out, long long *cyc)
{
    chain<OP><<<1, 32>>>(12345, out, cyc);
    chain<OP><<<1, 32>>>(12345, out, cyc);
    cudaDeviceSynchronize();
    printf("%-28s %6.1f cycles per dependent op\n", name, (double)cyc[0] / N);
}
int main()
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 256); cudaMallocManaged(&cyc, 8);
    run<0>("VIADDMNMX.U16x2", out, cyc);
    run<1>("VIMNMX.U16x2 (+IADD)", out, cyc);
    run<2>("IADD + VIMNMX.S32", out, cyc);
    run<3>("VIMNMX3.S32", out, cyc);
    run<4>("PRMT", out, cyc);
    run<5>("SHFL.BFLY", out, cyc);
    run<6>("LDS (dependent address)", out, cyc);
    run<7>("VIADDMNMX.S32", out, cyc);
    run<8>("IADD3", out, cyc);
    run<9>("STS + syncwarp + LDS", out, cyc);
    run<10>("bar.sync (one warp)", out, cyc);
    return 0;
}
