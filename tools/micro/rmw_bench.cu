// Microbenchmark: how fast is the epilogue's shared-memory read-modify-write (LDS.32 / FADD / STS.32, 32 lanes = one 128 B row) when the
// SM also runs (a) nothing else, (b) warps polling mbarriers, (c) warps issuing cp.async / TMA?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
__device__ __forceinline__ float lds(u32 a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(u32 a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
// mode bit0: 4 extra warps poll an mbarrier that never completes (try_wait); bit1: same with test_wait; bit2: 3 warps stream cp.async 16 B
// per lane from global; bit3: 3 warps do LDG.128 + STS.128
__global__ void __launch_bounds__(384) k(int mode, const uint4 *__restrict__ src, u32 n_src, long long *out, int iters, int batch) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *acc = (float *)smem;                         // 512 rows x 32
    unsigned char *stage = smem + 65536;                // 32 KB staging for the copy warps
    __shared__ __align__(8) u64 bar;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 512 * 32; i += 384) acc[i] = 0.f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((u32)__cvta_generic_to_shared(&bar))); stop = 0; }
    __syncthreads();
    const u32 acc0 = (u32)__cvta_generic_to_shared(acc) + lane * 4;
    const u32 barA = (u32)__cvta_generic_to_shared(&bar);
    if (warp < 4) {
        u32 rng = warp * 7919u + blockIdx.x * 31u + 1u;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            u32 ra[8]; float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { rng = rng * 1664525u + 1013904223u; ra[j] = acc0 + ((rng >> 9) & 511u) * 128u; }
            if (batch) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = lds(ra[j]);
#pragma unroll
                for (int j = 0; j < 8; ++j) sts(ra[j], a[j] + 1.f);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) { a[j] = lds(ra[j]); sts(ra[j], a[j] + 1.f); }
            }
        }
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
        __syncwarp();
        if (warp == 0 && lane == 0) stop = 1;
    } else if (warp < 8) {
        if (mode & 3) {
            while (!stop) {
                u32 done;
                if (mode & 1) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(barA), "r"(0u) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(barA), "r"(0u) : "memory");
            }
        }
    } else if (warp < 11) {
        if (mode & 12) {
            u32 rng = warp * 104729u + blockIdx.x;
            const u32 st = (u32)__cvta_generic_to_shared(stage) + (warp - 8) * 8192 + lane * 16;
            int slot = 0;
            while (!stop) {
                rng = rng * 1664525u + 1013904223u;
                const uint4 *p = src + (size_t)((rng >> 4) % (n_src / 8)) * 8 + (lane & 7) + (size_t)(lane >> 3) * 0;   // 8 lanes = one 128 B row
                if (mode & 4) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + slot * 512), "l"(p));
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 8;" ::: "memory");
                } else {
                    uint4 v = __ldg(p);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(st + slot * 512), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                }
                slot = (slot + 1) & 15;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
    }
}
int main() {
    const u32 n_src = 8u << 20;                       // 8 M uint4 = 128 MB
    uint4 *src; cudaMalloc(&src, (size_t)n_src * 16); cudaMemset(src, 1, (size_t)n_src * 16);
    long long *out; cudaMalloc(&out, 296 * 4 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    for (int batch = 0; batch < 2; ++batch)
        for (int mode : {0, 1, 2, 4, 8, 5, 9}) {
            k<<<296, 384, 100 * 1024>>>(mode, src, n_src, out, iters, batch);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[296 * 4]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 296 * 4; ++i) s += h[i];
            printf("batch=%d mode=%2d: %.1f clk per 8-pair RMW batch per warp (%.1f clk per pair)\n", batch, mode, s / (296 * 4) / iters, s / (296 * 4) / iters / 8);
        }
    return 0;
}
