// Microbenchmark / bring-up: TMA tile::gather4 (4 arbitrary rows of a 2D tensor per instruction, 128B swizzle) feeding a
// tcgen05.mma whose A operand is described with a SWIZZLE_128B K-major shared-memory descriptor.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_gather4 tma_gather4.cu && ./tma_gather4
// X [N rows][64 bf16] (a "split row": 32 hi | 32 lo); the kernel gathers 128 rows, multiplies the hi and the lo half by an
// identity B tile and accumulates: D[m][n] = X[idx[m]][n] + X[idx[m]][32 + n], checked on the host.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64; typedef uint16_t u16;

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
// canonical K-major no-swizzle tile offset (B operand)
__host__ __device__ inline u32 tile_off(int r, int c) { return (u32)((r >> 3) * 512 + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2); }
__device__ __forceinline__ u64 desc_noswz(u32 a) { return (u64)((a >> 4) & 0x3FFFu) | ((u64)(128u >> 4) << 16) | ((u64)(512u >> 4) << 32) | (1ull << 46); }
// K-major SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused (1), layout type 2
__device__ __forceinline__ u64 desc_sw128(u32 a) { return (u64)((a >> 4) & 0x3FFFu) | ((u64)1 << 16) | ((u64)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61); }
constexpr u32 IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmap, const int *__restrict__ idx, unsigned char *__restrict__ raw,
                                         float *__restrict__ dout) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *a_tile = (unsigned char *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);     // 16 KB, 1024-aligned
    unsigned char *b_tile = a_tile + 16384;                                                      // 2 KB identity, canonical no swizzle
    __shared__ __align__(8) u64 bar[2];
    __shared__ u32 tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 bar0 = (u32)__cvta_generic_to_shared(&bar[0]), bar1 = bar0 + 8;
    for (int i = tid; i < 32 * 32; i += 128) {
        const int n = i >> 5, kk = i & 31;
        *reinterpret_cast<u16 *>(b_tile + tile_off(n, kk)) = (n == kk) ? (u16)0x3F80 : (u16)0;      // bf16 1.0
    }
    for (int i = tid; i < 16384 / 4; i += 128) reinterpret_cast<u32 *>(a_tile)[i] = 0xDEADBEEFu;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        const u32 dst = (u32)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 td = tmem_base;
    const u32 a_addr = (u32)__cvta_generic_to_shared(a_tile), b_addr = (u32)__cvta_generic_to_shared(b_tile);
    if (warp == 0) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(16384u) : "memory");
        __syncwarp();
        // lane l gathers rows 4 l .. 4 l + 3 of the tile
        const int r0 = idx[4 * lane], r1 = idx[4 * lane + 1], r2 = idx[4 * lane + 2], r3 = idx[4 * lane + 3];
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(a_addr + lane * 512), "l"(reinterpret_cast<u64>(&tmap)), "r"(bar0), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    }
    mbar_wait(bar0, 0);
    for (int i = tid; i < 16384 / 16; i += 128) reinterpret_cast<uint4 *>(raw)[i] = reinterpret_cast<const uint4 *>(a_tile)[i];
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        if (elect_one()) {
            for (int ks = 0; ks < 4; ++ks) {       // K = 64 = hi (2 steps) | lo (2 steps); B = identity for every step
                const u64 ad = desc_sw128(a_addr + ks * 32);
                const u64 bd = desc_noswz(b_addr + (ks & 1) * 256);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                             ::"r"(td), "l"(ad), "l"(bd), "r"(IDESC), "r"((u32)(ks > 0)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar1) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(bar1, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    u32 d[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
                   "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
                   "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
                   "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                 : "r"(td + ((u32)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) dout[tid * 32 + j] = __uint_as_float(d[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(td) : "memory");
}

static u16 f2bf(float f) { u32 u; memcpy(&u, &f, 4); return (u16)(u >> 16); }

int main() {
    const int N = 2048;
    std::vector<u16> X((size_t)N * 64);
    auto val = [](int r, int c) { return (float)((r * 7 + c * 3) % 97); };
    for (int r = 0; r < N; ++r) for (int c = 0; c < 64; ++c) X[(size_t)r * 64 + c] = f2bf(val(r, c));
    std::vector<int> idx(128);
    for (int m = 0; m < 128; ++m) idx[m] = (m * 601 + 13) % N;
    u16 *dX; int *didx; unsigned char *draw; float *dD;
    cudaMalloc(&dX, X.size() * 2); cudaMalloc(&didx, 128 * 4); cudaMalloc(&draw, 16384); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(didx, idx.data(), 128 * 4, cudaMemcpyHostToDevice);

    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {64, (cuuint64_t)N};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 1; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480 + 1024);
    k<<<1, 128, 20480 + 1024>>>(tmap, didx, draw, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned char> raw(16384); std::vector<float> D(128 * 32);
    cudaMemcpy(raw.data(), draw, 16384, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    // layout check: tile row m, 16-byte chunk j expected at m * 128 + ((j ^ (m & 7)) * 16)
    int bad_layout = 0;
    for (int m = 0; m < 128; ++m) for (int j = 0; j < 8; ++j) {
        const u16 *p = (const u16 *)(raw.data() + m * 128 + ((j ^ (m & 7)) * 16));
        for (int e2 = 0; e2 < 8; ++e2) if (p[e2] != X[(size_t)idx[m] * 64 + j * 8 + e2]) ++bad_layout;
    }
    printf("gather4 + 128B swizzle layout: %s (%d mismatching elements)\n", bad_layout ? "MISMATCH" : "ok", bad_layout);
    if (bad_layout) {
        for (int m = 0; m < 2; ++m) { printf("row %d (src %d):", m, idx[m]); for (int b = 0; b < 64; ++b) { u32 u = (u32)((const u16 *)(raw.data() + m * 128))[b] << 16; float f; memcpy(&f, &u, 4); printf(" %g", f); } printf("\n  expect:"); for (int c = 0; c < 64; ++c) printf(" %g", val(idx[m], c)); printf("\n"); }
    }
    int bad = 0; float worst = 0;
    for (int m = 0; m < 128; ++m) for (int n2 = 0; n2 < 32; ++n2) {
        const float ref = val(idx[m], n2) + val(idx[m], 32 + n2);
        const float err = fabsf(D[m * 32 + n2] - ref);
        if (err > 0) { ++bad; if (err > worst) worst = err; }
    }
    printf("tcgen05.mma with SWIZZLE_128B A descriptor: %s (%d wrong, worst %g)\n", bad ? "MISMATCH" : "ok", bad, worst);
    if (bad) { printf("D[0][0..7] ="); for (int j = 0; j < 8; ++j) printf(" %g", D[j]); printf("   expect"); for (int j = 0; j < 8; ++j) printf(" %g", val(idx[0], j) + val(idx[0], 32 + j)); printf("\n"); }
    return 0;
}
