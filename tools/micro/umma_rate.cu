// Microbenchmark: issue cost / latency of small tcgen05.mma (kind::f16, M = 128, K = 16) on B200.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
// One thread of one CTA per SM issues `reps` groups of G MMAs and commits after every group; variants:
//   same D (dependent accumulate) vs G different D column ranges; A from shared memory vs A from TMEM; N = 32 .. 256.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;

__device__ __forceinline__ u64 umma_desc(u32 smem_addr) {
    return (u64)((smem_addr >> 4) & 0x3FFFu) | ((u64)(128u >> 4) << 16) | ((u64)(512u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ u32 idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(N >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ss(u32 d, u64 a, u64 b, u32 id, u32 acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(u32 d, u32 a, u64 b, u32 id, u32 acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    for (u32 spins = 0; spins < (1u << 22); ++spins) {
        u32 done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ bool elect_one() {
    u32 pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// mode bit0: A from TMEM; bit1: G distinct accumulators; G = MMAs per group; wait_each: wait for the commit after every group
__global__ void __launch_bounds__(128) k(int N, int G, int mode, int reps, int wait_each, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) u64 bar;
    __shared__ u32 tmem_base;
    for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<u32 *>(smem)[i] = 0x3C003C00u;
    const u32 b = (u32)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        const u32 dst = (u32)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 td = tmem_base;
    if (threadIdx.x == 0) {
        const u32 sa = (u32)__cvta_generic_to_shared(smem), sb = sa + 16384;
        const u32 id = idesc(N);
        u32 parity = 0;
        const long long t0 = clock64();
        long long t_issue = 0;
        for (int r = 0; r < reps; ++r) {
            const long long ta = clock64();
            for (int g = 0; g < G; ++g) {
                const u32 d = td + ((mode & 2) ? (u32)((g * N) % 256) : 0u);
                if (mode & 1) mma_ts(d, td + 256 + (g & 1) * 8, umma_desc(sb + (g & 1) * 256), id, 1u);
                else mma_ss(d, umma_desc(sa + (g & 1) * 256), umma_desc(sb + (g & 1) * 256), id, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
            t_issue += clock64() - ta;
            if (wait_each || r == reps - 1) {
                if (!wait_each) { /* only the final commit completes a phase we wait on: earlier commits flipped it too */ }
                mbar_wait(b, parity);
                parity ^= 1u;
            } else {
                mbar_wait(b, parity);      // keep phase accounting simple: every commit is consumed, but ...
                parity ^= 1u;
            }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

// pipelined variant: never waits inside the loop; commits go to a ring of 8 barriers waited 4 groups later
__global__ void __launch_bounds__(128) kp(int N, int G, int mode, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) u64 bar[8];
    __shared__ u32 tmem_base;
    for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<u32 *>(smem)[i] = 0x3C003C00u;
    const u32 b0 = (u32)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + i * 8)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        const u32 dst = (u32)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 td = tmem_base;
    if (threadIdx.x == 0) {
        const u32 sa = (u32)__cvta_generic_to_shared(smem), sb = sa + 16384;
        const u32 id = idesc(N);
        const long long t0 = clock64();
        long long t_issue = 0;
        for (int r = 0; r < reps; ++r) {
            if (r >= 4) mbar_wait(b0 + ((r - 4) & 7) * 8, ((r - 4) >> 3) & 1);
            const long long ta = clock64();
            const u32 dbuf = td + (u32)((r & 3) * 32) * ((mode & 4) ? 1u : 0u);      // bit2: rotate over 4 accumulator buffers per group
            for (int g = 0; g < G; ++g) {
                const u32 d = dbuf + ((mode & 2) ? (u32)((g * N) % 256) : 0u);
                if (mode & 1) mma_ts(d, td + 256 + (g & 1) * 8, umma_desc(sb + (g & 1) * 256), id, 1u);
                else mma_ss(d, umma_desc(sa + (g & 1) * 256), umma_desc(sb + (g & 1) * 256), id, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + (r & 7) * 8) : "memory");
            t_issue += clock64() - ta;
        }
        for (int r = reps > 4 ? reps - 4 : 0; r < reps; ++r) mbar_wait(b0 + (r & 7) * 8, (r >> 3) & 1);
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

// as kp, but the whole warp runs the loop (uniform control flow, descriptors in uniform registers) and ONE elected lane issues
__global__ void __launch_bounds__(128) ke(int N, int G, int mode, int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) u64 bar[8];
    __shared__ u32 tmem_base;
    for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<u32 *>(smem)[i] = 0x3C003C00u;
    const u32 b0 = (u32)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + i * 8)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        const u32 dst = (u32)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 td = tmem_base;
    if (threadIdx.x < 32) {
        const u32 sa = (u32)__cvta_generic_to_shared(smem), sb = sa + 16384;
        const u32 id = idesc(N);
        const long long t0 = clock64();
        long long t_issue = 0;
        for (int r = 0; r < reps; ++r) {
            if (r >= 4) mbar_wait(b0 + ((r - 4) & 7) * 8, ((r - 4) >> 3) & 1);
            const long long ta = clock64();
            const u32 dbuf = td + (u32)((r & 3) * 32) * ((mode & 4) ? 1u : 0u);
            if (elect_one()) {
                for (int g = 0; g < G; ++g) {
                    const u32 d = dbuf + ((mode & 2) ? (u32)((g * N) % 256) : 0u);
                    if (mode & 1) mma_ts(d, td + 256 + (g & 1) * 8, umma_desc(sb + (g & 1) * 256), id, 1u);
                    else mma_ss(d, umma_desc(sa + (g & 1) * 256), umma_desc(sb + (g & 1) * 256), id, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + (r & 7) * 8) : "memory");
            }
            __syncwarp();
            t_issue += clock64() - ta;
        }
        for (int r = reps > 4 ? reps - 4 : 0; r < reps; ++r) mbar_wait(b0 + (r & 7) * 8, (r >> 3) & 1);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

// pure issue throughput: W warps (each its own accumulator columns) issue `reps` MMAs back to back, ONE commit per warp at the end;
// commit_every > 0: additionally commit (to a scratch barrier nobody waits on) after every commit_every MMAs
__global__ void __launch_bounds__(128) kq(int N, int W, int reps, int commit_every, int ts, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) u64 bar[8];
    __shared__ u32 tmem_base;
    for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<u32 *>(smem)[i] = 0x3C003C00u;
    const u32 b0 = (u32)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + i * 8)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        const u32 dst = (u32)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 td = tmem_base;
    const int w = threadIdx.x >> 5;
    if (w < W) {
        const u32 sa = (u32)__cvta_generic_to_shared(smem), sb = sa + 16384;
        const u32 id = idesc(N);
        const u32 d = td + (u32)(w * 64 % 256);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int r = 0; r < reps; ++r) {
                if (ts) mma_ts(d, td + 256 + (r & 1) * 8, umma_desc(sb + (r & 1) * 256), id, 1u);
                else mma_ss(d, umma_desc(sa + (r & 1) * 256), umma_desc(sb + (r & 1) * 256), id, 1u);
                if (commit_every > 0 && (r % commit_every) == commit_every - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + (4 + w) * 8) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0 + w * 8) : "memory");
        }
        __syncwarp();
        const long long t1 = clock64();
        mbar_wait(b0 + w * 8, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

int main() {
    long long *out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    printf("# issue throughput: back-to-back MMAs, one commit at the end (issue = last MMA issued, done = all complete)\n");
    for (int ts : {0, 1})
        for (int N : {32, 64, 128, 256})
            for (int W : {1, 2, 4})
                for (int ce : {0, 6, 1}) {
                    if (N * 1 > 64 && W > 1) continue;
                    const int reps = 600;
                    kq<<<148, 128, 49152 + 1024>>>(N, W, reps, ce, ts, out);
                    long long h[2]; cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("issue  A=%s N=%3d warps=%d commit_every=%d: issue %.1f clk/MMA, done %.1f clk/MMA (per warp)\n", ts ? "tmem" : "smem", N, W, ce,
                           (double)h[0] / reps, (double)h[1] / reps);
                }
    return 0;
}
int main_old() {
    long long *out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    const int reps = 200;
    printf("# serial: commit + wait after every group of G MMAs (group latency)\n");
    for (int mode = 0; mode < 1; ++mode)
        for (int N : {32})
            for (int G : {1, 6}) {
                if ((mode & 2) && G * N > 256) continue;
                k<<<148, 128, 49152 + 1024>>>(N, G, mode, reps, 1, out);
                long long h[2]; cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("serial A=%s D=%s N=%3d G=%d: %.1f clk/group  (issue %.1f)\n", (mode & 1) ? "tmem" : "smem", (mode & 2) ? "distinct" : "same", N, G,
                       (double)h[0] / reps, (double)h[1] / reps);
            }
    printf("# pipelined: 4 groups in flight\n");
    for (int mode = 0; mode < 8; ++mode)
        for (int N : {32, 64, 128, 256})
            for (int G : {1, 6}) {
                if ((mode & 2) && G * N > 256) continue;
                if ((mode & 4) && N != 32) continue;
                kp<<<148, 128, 49152 + 1024>>>(N, G, mode, reps, out);
                long long h[2]; cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("pipe   A=%s D=%s rot=%d N=%3d G=%d: %.1f clk/group = %.1f clk/MMA (issue %.1f/group)\n", (mode & 1) ? "tmem" : "smem", (mode & 2) ? "distinct" : "same",
                       (mode >> 2) & 1, N, G, (double)h[0] / reps, (double)h[0] / reps / G, (double)h[1] / reps);
            }
    cudaFuncSetAttribute(ke, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    printf("# elected lane in a uniform warp loop, 4 groups in flight\n");
    for (int mode = 0; mode < 8; ++mode)
        for (int N : {32, 64, 128, 256})
            for (int G : {1, 6}) {
                if ((mode & 2) && G * N > 256) continue;
                if ((mode & 4) && N != 32) continue;
                ke<<<148, 128, 49152 + 1024>>>(N, G, mode, reps, out);
                long long h[2]; cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("elect  A=%s D=%s rot=%d N=%3d G=%d: %.1f clk/group = %.1f clk/MMA (issue %.1f/group)\n", (mode & 1) ? "tmem" : "smem", (mode & 2) ? "distinct" : "same",
                       (mode >> 2) & 1, N, G, (double)h[0] / reps, (double)h[0] / reps / G, (double)h[1] / reps);
            }
    return 0;
}
