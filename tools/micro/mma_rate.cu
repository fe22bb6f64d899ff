// microbenchmark: legacy mma.sync TF32 / BF16 and FFMA2 issue rates on sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE> __global__ void k(float *out, int iters) {
    float d[8][4] = {};
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    uint64_t acc[8] = {}; uint64_t x = 0x3f8000003f800000ull + threadIdx.x, w = 0x3f0000003f000000ull;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { _Pragma("unroll")
            for (int j = 0; j < 8; ++j) mma_tf32(d[j], a, b); }
        else if (MODE == 1) { _Pragma("unroll")
            for (int j = 0; j < 8; ++j) mma_bf16(d[j], a, b); }
        else { _Pragma("unroll")
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(x), "l"(w)); }
    }
    float s = 0; for (int j = 0; j < 8; ++j) { s += d[j][0] + d[j][1] + d[j][2] + d[j][3] + (float)(acc[j] & 0xff); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, warps * 32>>>(out, iters); else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters); else k<2><<<148, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double ops = (double)iters * 8 * warps * 148;      // warp-instructions
        double macs = mode == 0 ? 1024.0 : (mode == 1 ? 2048.0 : 64.0);
        printf("mode %d (%s) warps/SM %2d: %.3f ms  %.2f warp-instr/clk/SM (at 1.9GHz)  %.1f TMAC/s\n", mode,
               mode == 0 ? "mma tf32 m16n8k8" : mode == 1 ? "mma bf16 m16n8k16" : "ffma2", warps, best,
               ops / 148 / (best * 1e-3 * 1.9e9), ops * macs / (best * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
