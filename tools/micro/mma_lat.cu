// dependent-chain latency of legacy mma.sync bf16 m16n8k16 on sm_100a, and ILP scaling within one warp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ILP> __global__ void k(float *out, int iters, long long *cycles) {
    float d[ILP][4] = {};
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) mma_bf16(d[j], a, b);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < ILP; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    float *out; cudaMalloc(&out, 1 << 20); long long *cyc; cudaMallocManaged(&cyc, 8);
    const int iters = 4096;
#define RUN(I, W) { k<I><<<1, 32 * W>>>(out, iters, cyc); cudaDeviceSynchronize(); k<I><<<1, 32 * W>>>(out, iters, cyc); cudaDeviceSynchronize(); \
    printf("ILP %d warps %d: %.1f clk per mma per warp-chain step (%.2f clk/mma overall)\n", I, W, (double)*cyc / iters, (double)*cyc / iters / I / W); }
    RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(1, 8) RUN(4, 8)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
