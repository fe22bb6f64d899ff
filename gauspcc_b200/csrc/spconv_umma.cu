// spconv_umma.cu -- sparse convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Reference: every spnn.Conv3d(C, C, 5) of src/ai_pcc/GausPcgc/network_ue_4stage_conv.py:17-62 (torchsparse 2.1.0
// gather - implicit GEMM - scatter); semantics as restated in SURVEY.md 8(c): y[o] = sum_k W[k]^T x[nbr_k(o)].
// v9  = first, unpipelined version (kept for A/B: variant 70).
// v10 = the production kernel: warp-specialised gather -> tcgen05.mma -> TMEM -> scatter-add pipeline over
//       "split rows" (every activation row stored as 32 x bf16 hi | 32 x bf16 lo, see gpc_rows_split).
#include "common.cuh"
// =====================================================================================================
// v9: tcgen05 / TMEM contraction.  One CTA (4 warps) owns TM consecutive output rows with fp32 accumulators in
// shared memory; the tile's pair stream (sorted by offset, row) is cut into chunks of <= 128 pairs of ONE offset.
// Per chunk: thread i gathers input row i, splits it into bf16 hi / lo and stores both into UMMA canonical
// K-major (no-swizzle) operand tiles; W[k] hi / lo come as a pre-packed canonical image; ONE thread issues
//     D[128 pairs x 32 co] (TMEM, fp32)  =  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi     (6 x tcgen05.mma kind::f16, K = 16)
// and commits to an mbarrier; every warp then reads its 32 TMEM lanes (tcgen05.ld 32x32b.x32: one full output row
// per thread) and adds it into the accumulator row of its pair.  Operands never pass through the register file
// as MMA fragments (v6: 4 KB of W^T fragments per offset change per warp, 12 HMMA + 24 conversions per 8 pairs);
// here each thread does one row gather + one row scatter-add per pair and the tensor core does the rest.
// Chunks are processed in stream order with a CTA barrier between them: fixed accumulation order per output row.
// =====================================================================================================
constexpr int SC9_ACC = 36;
constexpr u32 SC9_IDESC = (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ |
                          ((32u >> 3) << 17) /* N = 32 */ | ((128u >> 4) << 24) /* M = 128 */;   // A, B K-major, dense

// canonical K-major no-swizzle tile: element (row r, col c) of a [rows x 32] bf16 tile lives at
//   (r / 8) * 512 + (c / 8) * 128 + (r % 8) * 16 + (c % 8) * 2      (8 x 16-byte core matrices, LBO = 128, SBO = 512)
__host__ __device__ __forceinline__ u32 umma_off(int r, int c) { return (u32)((r >> 3) * 512 + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2); }

// W [n_kernels*125][32 ci][32 co] fp32 -> Wc [n_kernels*125][2 (hi, lo)][2 KB canonical image of B[n = co][c = ci]]
__global__ void pack_weights_umma_kernel(const float *__restrict__ W, u16 *__restrict__ Wc, int n_kernels) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (i64)n_kernels * GPC_K3 * GPC_C * GPC_C) return;
    const int co = (int)(g & 31), ci = (int)((g >> 5) & 31);
    const i64 k = g >> 10;
    const float w = W[k * 1024 + ci * 32 + co];
    const float w1 = bf16_round(w), w2 = bf16_round(w - w1);
    u16 *dst = Wc + k * 2048;                       // 4 KB per offset = 2048 u16
    dst[umma_off(co, ci) / 2] = (u16)(__float_as_uint(w1) >> 16);
    dst[1024 + umma_off(co, ci) / 2] = (u16)(__float_as_uint(w2) >> 16);
}
extern "C" int gpc_spconv_pack_weights_umma(const float *W, int n_kernels, void *Wc, void *stream) {
    const i64 total = (i64)n_kernels * GPC_K3 * GPC_C * GPC_C;
    pack_weights_umma_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(W, (u16 *)Wc, n_kernels);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

__device__ __forceinline__ u64 umma_desc(u32 smem_addr) {      // K-major, SWIZZLE_NONE, LBO = 128 B, SBO = 512 B
    return (u64)((smem_addr >> 4) & 0x3FFFu) | ((u64)(128u >> 4) << 16) | ((u64)(512u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    // bounded: a mis-programmed MMA / commit must fault the launch (trap), never hang the GPU
    for (u32 spins = 0; spins < (1u << 24); ++spins) {
        u32 done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

template <int TM>
struct Sc9Smem {
    float acc[TM][SC9_ACC];
    __align__(128) unsigned char a_hi[8192];
    __align__(128) unsigned char a_lo[8192];
    __align__(128) unsigned char b[4096];          // hi image then lo image
    __align__(8) u64 mbar;
    u32 tmem_base;
    u32 seg[GPC_K3 + 1];
};

template <int TM>
__global__ void __launch_bounds__(128) spconv_fwd_v9_kernel(const float *__restrict__ x, const uint4 *__restrict__ Wc,
                                                            const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs, i64 n,
                                                            const float *__restrict__ residual, int flags, float *__restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Sc9Smem<TM> &s = *reinterpret_cast<Sc9Smem<TM> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const i64 t = blockIdx.x;
    const i64 r0 = t * TM;
    const int rows = (int)min((i64)TM, n - r0);

    for (int i = tid; i <= GPC_K3; i += 128) s.seg[i] = seg_g[t * (GPC_K3 + 1) + i];
    for (int i = tid; i < TM * SC9_ACC / 4; i += 128) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const u32 bar = (u32)__cvta_generic_to_shared(&s.mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        const u32 dst = (u32)__cvta_generic_to_shared(&s.tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_d = s.tmem_base;
    const u32 a_hi = (u32)__cvta_generic_to_shared(s.a_hi), a_lo = (u32)__cvta_generic_to_shared(s.a_lo);
    const u32 b_hi = (u32)__cvta_generic_to_shared(s.b), b_lo = b_hi + 2048;
    const u32 my_a = (u32)((tid >> 3) * 512 + (tid & 7) * 16);          // row `tid` of the operand tile, column group 0

    u32 parity = 0;
    int cur_k = -1;
    const u32 p_end = s.seg[GPC_K3];
    int k = 0;
    for (u32 p = s.seg[0]; p < p_end;) {
        while (p >= s.seg[k + 1]) ++k;
        const int cnt = (int)min(128u, s.seg[k + 1] - p);
        // ---- gather + split: thread i <-> pair i of the chunk
        u32 my_row = 0xFFFFu;
        if (tid < cnt) {
            const u64 e = __ldg(pairs + p + tid);
            my_row = (u32)(e >> 32) & 0xFFFFu;
            const float4 *src = reinterpret_cast<const float4 *>(x + (i64)(u32)e * GPC_C);
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                 // 8 channels per 16-byte core-matrix row
                const float4 v0 = __ldg(src + 2 * j), v1 = __ldg(src + 2 * j + 1);
                u32 h0, l0, h1, l1, h2, l2, h3, l3;
                split_bf16(v0.x, v0.y, h0, l0); split_bf16(v0.z, v0.w, h1, l1);
                split_bf16(v1.x, v1.y, h2, l2); split_bf16(v1.z, v1.w, h3, l3);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_hi + my_a + j * 128), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_lo + my_a + j * 128), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
            }
        }
        if (k != cur_k) {                                                  // block-uniform: 4 KB canonical image of W[k]
            cur_k = k;
            const uint4 *wsrc = Wc + (size_t)k * 256;
            reinterpret_cast<uint4 *>(s.b)[tid] = __ldg(wsrc + tid);
            reinterpret_cast<uint4 *>(s.b)[tid + 128] = __ldg(wsrc + tid + 128);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_bf16(tmem_d, umma_desc(a_lo + ks * 256), umma_desc(b_hi + ks * 256), SC9_IDESC, ks);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_bf16(tmem_d, umma_desc(a_hi + ks * 256), umma_desc(b_lo + ks * 256), SC9_IDESC, 1u);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) umma_bf16(tmem_d, umma_desc(a_hi + ks * 256), umma_desc(b_hi + ks * 256), SC9_IDESC, 1u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: thread i holds row i of D (32 fp32) -> add into the accumulator row of its pair
        u32 d[32];
        const u32 taddr = tmem_d + ((u32)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
                       "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
                       "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
                       "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (my_row != 0xFFFFu) {
            float4 *a = reinterpret_cast<float4 *>(&s.acc[my_row][0]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 v = a[j];
                v.x += __uint_as_float(d[4 * j]); v.y += __uint_as_float(d[4 * j + 1]);
                v.z += __uint_as_float(d[4 * j + 2]); v.w += __uint_as_float(d[4 * j + 3]);
                a[j] = v;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                       // operand tiles, TMEM and the accumulator rows are free for the next chunk
        p += cnt;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d) : "memory");
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    const int lane = tid & 31;
    for (int r = warp; r < rows; r += 4) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TM>
static int launch_spconv_v9(const float *x, const void *Wc, const u32 *seg, const u64 *pairs, i64 n, const float *residual,
                            int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc9Smem<TM>) + 128;
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v9_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_fwd_v9_kernel<TM><<<(unsigned)tiles, 128, smem, st>>>(x, (const uint4 *)Wc, seg, pairs, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variant 70: Wc from gpc_spconv_pack_weights_umma; pair stream with pad = 1 and tile_rows = TM
extern "C" int gpc_spconv_fwd_v9(const float *x, const void *Wc, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int tile_rows, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 70) {
        if (tile_rows == 128) return launch_spconv_v9<128>(x, Wc, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 256) return launch_spconv_v9<256>(x, Wc, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 512) return launch_spconv_v9<512>(x, Wc, seg, pairs, n, residual, flags, y, st);
    }
    gpc_set_error("unsupported conv v9 variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}


// =====================================================================================================
// "split rows": one activation row = 128 B = 32 x bf16 hi (channels 0..31) | 32 x bf16 lo, x = hi + lo to 16 mantissa
// bits.  Same bytes per row as fp32, but a gathered row drops into the tensor-core operand tile as eight 16-byte
// cp.async copies (no register staging, no conversion instructions in the conv), and the three-term product
// hi.Whi + hi.Wlo + lo.Whi keeps the contraction within 1.5e-4 of fp32 on the probabilities.
// =====================================================================================================
__global__ void rows_split_kernel(const float2 *__restrict__ x, i64 n, u32 *__restrict__ xs) {
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * 16) return;
    const i64 r = g >> 4;
    const int cp = (int)(g & 15);
    const float2 v = x[g];
    u32 hi, lo;
    split_bf16(v.x, v.y, hi, lo);
    xs[r * 32 + cp] = hi;
    xs[r * 32 + 16 + cp] = lo;
}
__global__ void rows_join_kernel(const u32 *__restrict__ xs, i64 n, float2 *__restrict__ x) {
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * 16) return;
    const i64 r = g >> 4;
    const int cp = (int)(g & 15);
    x[g] = join_bf16(xs[r * 32 + cp], xs[r * 32 + 16 + cp]);
}
extern "C" int gpc_rows_split(const float *x, int64_t n, void *xs, void *stream) {
    if (n <= 0) return GPC_OK;
    rows_split_kernel<<<cdiv(n * 16, 256), 256, 0, as_stream(stream)>>>((const float2 *)x, n, (u32 *)xs);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
extern "C" int gpc_rows_join(const void *xs, int64_t n, float *x, void *stream) {
    if (n <= 0) return GPC_OK;
    rows_join_kernel<<<cdiv(n * 16, 256), 256, 0, as_stream(stream)>>>((const u32 *)xs, n, (float2 *)x);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// =====================================================================================================
// v10: warp-specialised tcgen05 pipeline.  One CTA owns TM consecutive output rows = 4 QUARTERS of TM/4 rows; the pair
// stream is the ordinary one built with tile_rows = TM/4 (sorted by sub-tile, offset, row; pad = 1).
//
//   chunk (k, j)      = for each quarter q the pairs [seg[q][k] + 32 j, + 32) of offset k  ->  A rows 32 q .. 32 q + 31
//                       (M = 128 = 4 quarters x 32 lanes, all of ONE offset k, so one W[k] serves the whole MMA)
//   gather warps 4..7 : warp q, lane l copies its pair's split row into the canonical K-major operand tiles of a
//                       ring stage (8 x cp.async 16 B) + 1/128 of W[k]'s 4 KB image; cp.async.wait_group lagging D
//                       chunks behind, fence.proxy.async, arrive on full[stage]
//   MMA warp 8, lane 0: wait full[stage] and a free TMEM buffer; 6 x tcgen05.mma (lo.Whi, hi.Wlo, hi.Whi; K = 2 x 16)
//                       -> D[128 x 32] fp32 in TMEM; tcgen05.commit frees the stage and publishes the buffer
//   epilogue warps 0..3: warp q reads TMEM lanes 32 q .. 32 q + 31 (tcgen05.ld 32x32b.x32: one 32-channel row per
//                       thread) and adds each row into the fp32 accumulator row of its pair in shared memory.
// Quarter q's accumulator rows are touched by warp q only and its chunks arrive in offset order, so the summation
// order per output row is fixed (encoder and decoder CDFs stay bit-identical) with no CTA barrier in the loop.
// =====================================================================================================
constexpr int SC10_NB = 4;            // TMEM accumulator buffers of 32 columns
constexpr int SC10_STAGE = 20480;     // A hi 8 KB | A lo 8 KB | W[k] hi 2 KB | W[k] lo 2 KB
constexpr int SC10_THREADS = 288;
constexpr int SC10_ENT = 8;           // pair-entry ring (chunks), filled by cp.async D + 1 chunks ahead
constexpr int SC10_RID = 16;          // row-id ring (chunks): gather warps -> epilogue warps; >= S + SC10_NB

template <int TM, int S>
struct Sc10Smem {
    float acc[TM][GPC_C];                         // 16-byte chunk j of row r lives at chunk j ^ (r & 7): conflict-free RMW
    __align__(128) unsigned char stage[S][SC10_STAGE];
    u64 ent[4][SC10_ENT][32];
    u16 rid[SC10_RID][128];
    u32 seg[4][GPC_K3 + 3];
    u32 cstart[GPC_K3 + 3];
    u16 tab[GPC_K3 * (TM / 128) + 4];             // chunk -> k | j << 8
    __align__(8) u64 full[S];
    u64 empty[S];
    u64 dfull[SC10_NB];
    u64 dempty[SC10_NB];
    u32 tmem_base;
};

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(u32 dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async8(u32 dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src));
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// one lane of a converged warp; keeps the surrounding control flow warp-uniform so that the tcgen05 operands (descriptors, TMEM
// addresses) live in uniform registers -- issuing from `if (lane == 0)` makes ptxas wrap every UTCHMMA in an election loop
__device__ __forceinline__ bool elect_one() {
    u32 pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(u32 bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// role profile (PROF instantiation only; tools/conv_ab.py --prof): cycles summed over the chunks of CTAs, lane 0 of one warp per role
//  [0] gather: wait empty   [1] gather: issue copies   [2] gather: wait_group + fence + arrive
//  [3] mma: wait full       [4] mma: wait dempty       [5] mma: issue + commit
//  [6] epi: wait dfull      [7] epi: tcgen05.ld        [8] epi: scatter-add     [9] chunks   [10] CTA total   [11] setup  [12] write-out
__device__ unsigned long long g_sc10_prof[32];
extern "C" int gpc_debug_conv_profile(unsigned long long *out_h, int reset) {
    GPC_CUDA_CHECK(cudaDeviceSynchronize());
    if (out_h) GPC_CUDA_CHECK(cudaMemcpyFromSymbol(out_h, g_sc10_prof, sizeof(unsigned long long) * 32));
    if (reset) { unsigned long long z[32] = {0}; GPC_CUDA_CHECK(cudaMemcpyToSymbol(g_sc10_prof, z, sizeof(z))); }
    return GPC_OK;
}
#define SC10_T(var) do { if (PROF) var = clock64(); } while (0)
// fine-grained marks of the gather loop ([16 + i], lane 0 of warp 4): time since the previous mark
#define SC10_MARK(i) do { if (PROF && lane == 0 && warp == 4) { const long long _n = clock64(); fine[i] += _n - t_mark; t_mark = _n; } } while (0)
#define SC10_ACCUM(i, a, b) do { if (PROF && lane == 0) acc_t[i] += (b) - (a); } while (0)

template <int TM, int S, int D, bool PROF>
__global__ void __launch_bounds__(SC10_THREADS, 1)
spconv_fwd_v10_kernel(const unsigned char *__restrict__ xs, const unsigned char *__restrict__ Wc, const u32 *__restrict__ seg_g,
                      const u64 *__restrict__ pairs, i64 n, const void *__restrict__ residual, int flags,
                      float *__restrict__ y, u32 *__restrict__ ys) {
    static_assert(D >= 1 && D < S, "signal lag must leave a free stage");
    static_assert(D + 1 < SC10_ENT && S + SC10_NB <= SC10_RID, "ring depths");
    constexpr int QR = TM / 4;                                   // rows per quarter == tile_rows of the pair stream
    constexpr int P = D + 1;                                     // pair entries are fetched P chunks ahead
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Sc10Smem<TM, S> &s = *reinterpret_cast<Sc10Smem<TM, S> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = blockIdx.x;
    const i64 n_sub = (n + QR - 1) / QR;
    long long acc_t[3] = {0, 0, 0}, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t_begin = 0, t_setup = 0;
    SC10_T(t_begin);

    // ---- setup: segment table, chunk table, barriers, TMEM
    for (int i = tid; i < 4 * (GPC_K3 + 1); i += SC10_THREADS) {
        const int q = i / (GPC_K3 + 1), k = i - q * (GPC_K3 + 1);
        const i64 st = t * 4 + q;
        s.seg[q][k] = st < n_sub ? seg_g[st * (GPC_K3 + 1) + k] : 0u;
    }
    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init((u32)__cvta_generic_to_shared(&s.full[i]), 128);
            mbar_init((u32)__cvta_generic_to_shared(&s.empty[i]), 1);
        }
        for (int i = 0; i < SC10_NB; ++i) {
            mbar_init((u32)__cvta_generic_to_shared(&s.dfull[i]), 1);
            mbar_init((u32)__cvta_generic_to_shared(&s.dempty[i]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        const u32 dst = (u32)__cvta_generic_to_shared(&s.tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid < GPC_K3) {
        u32 m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) m = max(m, s.seg[q][tid + 1] - s.seg[q][tid]);
        s.cstart[tid] = (m + 31) >> 5;
    }
    __syncthreads();
    if (warp == 0) {                                             // exclusive scan of the 125 chunk counts
        u32 v[4], sum = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int k = lane * 4 + i; v[i] = k < GPC_K3 ? s.cstart[k] : 0u; sum += v[i]; }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
        u32 run = incl - sum;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int k = lane * 4 + i; if (k <= GPC_K3) s.cstart[k] = run; run += v[i]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid < GPC_K3) {
        const u32 b = s.cstart[tid], e = s.cstart[tid + 1];
        for (u32 c = b; c < e; ++c) s.tab[c] = (u16)(tid | ((c - b) << 8));
    }
    __syncthreads();
    const u32 n_chunks = s.cstart[GPC_K3];
    const u32 tmem_d = s.tmem_base;
    const u32 stage0 = (u32)__cvta_generic_to_shared(&s.stage[0][0]);
    const u32 full0 = (u32)__cvta_generic_to_shared(&s.full[0]), empty0 = (u32)__cvta_generic_to_shared(&s.empty[0]);
    const u32 dfull0 = (u32)__cvta_generic_to_shared(&s.dfull[0]), dempty0 = (u32)__cvta_generic_to_shared(&s.dempty[0]);
    SC10_T(t_setup);

    if (warp >= 4 && warp < 8) {
        // =================================================================== gather warps
        const int q = warp - 4;
        const int pt = q * 32 + lane;                            // A row of this thread == its 1/128 share of W[k]
        const u32 a_off = (u32)((pt >> 3) * 512 + (pt & 7) * 16);
        const u32 ent0 = (u32)__cvta_generic_to_shared(&s.ent[q][0][lane]);
        // stream index of this lane's pair in chunk c, or ~0 (the lane idles in that chunk)
        auto pair_idx = [&](u32 c) -> u32 {
            const u32 kj = s.tab[c], k = kj & 0xFFu, j = kj >> 8;
            const u32 idx = s.seg[q][k] + 32u * j + (u32)lane;
            return idx < s.seg[q][k + 1] ? idx : 0xFFFFFFFFu;
        };
        for (u32 c = 0; c < (u32)P && c < n_chunks; ++c) {       // entries of the first P chunks
            const u32 idx = pair_idx(c);
            if (idx != 0xFFFFFFFFu) cp_async8(ent0 + (c % SC10_ENT) * 256, pairs + idx);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        cp_async_wait<0>();
        u32 st_i = 0, st_ph = 0;                                 // stage index / use parity of chunk c
        u32 sg_i = 0;                                            // stage index of chunk c - D
        for (u32 c = 0; c < n_chunks; ++c) {
            if (c + P < n_chunks) {                              // entry of chunk c + P rides in this chunk's copy group
                const u32 idx = pair_idx(c + P);
                if (idx != 0xFFFFFFFFu) cp_async8(ent0 + ((c + P) % SC10_ENT) * 256, pairs + idx);
            }
            const u32 my = pair_idx(c);
            SC10_T(t0);
            if (c >= (u32)S) mbar_wait(empty0 + st_i * 8, st_ph ^ 1u);
            SC10_T(t1);
            const u32 st = stage0 + st_i * SC10_STAGE;
            u32 row = 0xFFFFu;
            if (my != 0xFFFFFFFFu) {
                const u64 e = s.ent[q][c % SC10_ENT][lane];
                row = (u32)(e >> 32) & 0xFFFFu;
                const unsigned char *src = xs + (size_t)(u32)e * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cp_async16(st + a_off + j * 128, src + 16 * j);
                    cp_async16(st + 8192 + a_off + j * 128, src + 64 + 16 * j);
                }
            }
            s.rid[c % SC10_RID][pt] = (u16)row;                  // published to the epilogue warps through full -> dfull
            const unsigned char *wsrc = Wc + (size_t)(s.tab[c] & 0xFFu) * 4096 + pt * 16;
            cp_async16(st + 16384 + pt * 16, wsrc);
            cp_async16(st + 18432 + pt * 16, wsrc + 2048);
            asm volatile("cp.async.commit_group;" ::: "memory");
            SC10_T(t2);
            if (c >= (u32)D) {
                cp_async_wait<D>();                              // chunk c - D has landed (and the entries of chunk c + 1)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(full0 + sg_i * 8);
                if (++sg_i == (u32)S) sg_i = 0;
            } else {
                cp_async_wait<D>();
            }
            SC10_T(t3);
            SC10_ACCUM(0, t0, t1); SC10_ACCUM(1, t1, t2); SC10_ACCUM(2, t2, t3);
            if (++st_i == (u32)S) { st_i = 0; st_ph ^= 1u; }
        }
        if (PROF && lane == 0 && warp == 4) {
            atomicAdd(&g_sc10_prof[0], (unsigned long long)acc_t[0]); atomicAdd(&g_sc10_prof[1], (unsigned long long)acc_t[1]);
            atomicAdd(&g_sc10_prof[2], (unsigned long long)acc_t[2]);
        }
        cp_async_wait<0>();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (u32 c = n_chunks > (u32)D ? n_chunks - D : 0u; c < n_chunks; ++c) {
            mbar_arrive(full0 + sg_i * 8);
            if (++sg_i == (u32)S) sg_i = 0;
        }
    } else if (warp == 8) {
        // =================================================================== MMA issue (one thread)
        if (lane == 0) {
            u32 st_i = 0, st_ph = 0;
            for (u32 c = 0; c < n_chunks; ++c) {
                const u32 b = c & (SC10_NB - 1), v = c / SC10_NB;
                SC10_T(t0);
                mbar_wait(full0 + st_i * 8, st_ph);
                SC10_T(t1);
                if (v > 0) mbar_wait(dempty0 + b * 8, (v & 1u) ^ 1u);
                SC10_T(t2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 a_hi = stage0 + st_i * SC10_STAGE, a_lo = a_hi + 8192, b_hi = a_hi + 16384, b_lo = a_hi + 18432;
                const u32 d = tmem_d + b * 32;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16(d, umma_desc(a_lo + ks * 256), umma_desc(b_hi + ks * 256), SC9_IDESC, ks);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16(d, umma_desc(a_hi + ks * 256), umma_desc(b_lo + ks * 256), SC9_IDESC, 1u);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16(d, umma_desc(a_hi + ks * 256), umma_desc(b_hi + ks * 256), SC9_IDESC, 1u);
                umma_commit(empty0 + st_i * 8);                  // the stage is free once these MMAs have read it
                umma_commit(dfull0 + b * 8);                     // ... and the accumulator buffer is complete
                SC10_T(t3);
                SC10_ACCUM(0, t0, t1); SC10_ACCUM(1, t1, t2); SC10_ACCUM(2, t2, t3);
                if (++st_i == (u32)S) { st_i = 0; st_ph ^= 1u; }
            }
            if (PROF) {
                atomicAdd(&g_sc10_prof[3], (unsigned long long)acc_t[0]); atomicAdd(&g_sc10_prof[4], (unsigned long long)acc_t[1]);
                atomicAdd(&g_sc10_prof[5], (unsigned long long)acc_t[2]); atomicAdd(&g_sc10_prof[9], (unsigned long long)n_chunks);
            }
        }
        __syncwarp();
    } else {
        // =================================================================== epilogue warps (quarter q = warp)
        const int q = warp;
        float4 *accq = reinterpret_cast<float4 *>(&s.acc[q * QR][0]);
        for (int i = lane; i < QR * 8; i += 32) accq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        for (u32 c = 0; c < n_chunks; ++c) {
            const u32 b = c & (SC10_NB - 1), v = c / SC10_NB;
            SC10_T(t0);
            mbar_wait(dfull0 + b * 8, v & 1u);
            SC10_T(t1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const u32 r0 = s.rid[c % SC10_RID][q * 32 + lane];
            const bool any = __any_sync(0xFFFFFFFFu, r0 != 0xFFFFu);
            u32 d[32];
            if (any) {
                const u32 taddr = tmem_d + ((u32)(q * 32) << 16) + b * 32;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                               "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]),
                               "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]),
                               "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty0 + b * 8);         // the values are in registers: hand the buffer back
            SC10_T(t2);
            if (r0 != 0xFFFFu) {
                float4 *a = accq + r0 * 8;
                const u32 sw = r0 & 7u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 w = a[j ^ sw];
                    w.x += __uint_as_float(d[4 * j]); w.y += __uint_as_float(d[4 * j + 1]);
                    w.z += __uint_as_float(d[4 * j + 2]); w.w += __uint_as_float(d[4 * j + 3]);
                    a[j ^ sw] = w;
                }
            }
            SC10_T(t3);
            SC10_ACCUM(0, t0, t1); SC10_ACCUM(1, t1, t2); SC10_ACCUM(2, t2, t3);
        }
        __syncwarp();
        SC10_T(t0);
        if (PROF && lane == 0 && warp == 0) {
            atomicAdd(&g_sc10_prof[6], (unsigned long long)acc_t[0]); atomicAdd(&g_sc10_prof[7], (unsigned long long)acc_t[1]);
            atomicAdd(&g_sc10_prof[8], (unsigned long long)acc_t[2]);
        }
        // ---- write-out of the quarter: (+ residual) (ReLU) -> fp32 rows and / or split rows, one 128 B store per row
        const bool relu = (flags & GPC_CONV_RELU) != 0, res_split = (flags & GPC_CONV_RES_SPLIT) != 0;
        const i64 g0 = t * TM + (i64)q * QR;
        const int rows = (int)max((i64)0, min((i64)QR, n - g0));
        const int cp = lane & 15;
        for (int r = 0; r < rows; ++r) {
            const float4 a4 = accq[r * 8 + ((cp >> 1) ^ (r & 7))];
            float2 a = (cp & 1) ? make_float2(a4.z, a4.w) : make_float2(a4.x, a4.y);
            const i64 g = g0 + r;
            if (residual) {
                float2 rv;
                if (res_split) {
                    const u32 *rs = reinterpret_cast<const u32 *>(residual) + g * 32;
                    rv = join_bf16(__ldg(rs + cp), __ldg(rs + 16 + cp));
                } else {
                    rv = __ldg(reinterpret_cast<const float2 *>(residual) + g * 16 + cp);
                }
                a.x += rv.x; a.y += rv.y;
            }
            if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); }
            if (y && lane < 16) reinterpret_cast<float2 *>(y)[g * 16 + cp] = a;
            if (ys) {
                u32 hi, lo;
                split_bf16(a.x, a.y, hi, lo);
                ys[g * 32 + lane] = lane < 16 ? hi : lo;
            }
        }
        if (PROF && lane == 0 && warp == 0) atomicAdd(&g_sc10_prof[12], (unsigned long long)(clock64() - t0));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
    }
    if (PROF && tid == 0) {
        atomicAdd(&g_sc10_prof[10], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_sc10_prof[11], (unsigned long long)(t_setup - t_begin));
    }
}

template <int TM, int S, int D, bool PROF = false>
static int launch_spconv_v10(const void *xs, const void *Wc, const u32 *seg, const u64 *pairs, i64 n, const void *residual,
                             int flags, float *y, void *ys, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc10Smem<TM, S>) + 128;
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v10_kernel<TM, S, D, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_fwd_v10_kernel<TM, S, D, PROF><<<(unsigned)tiles, SC10_THREADS, smem, st>>>(
        (const unsigned char *)xs, (const unsigned char *)Wc, seg, pairs, n, residual, flags, y, (u32 *)ys);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// =====================================================================================================
// v11: v10 with the gathered operand kept OUT of shared memory.  v10's role profile (tools/conv_ab.py, variant 89) showed
// all three roles waiting on the shared-memory pipe: per chunk the gather wrote 16 KB of operand tiles, the six MMAs re-read
// 30 KB of them (N = 32 is too thin to amortise an A read) and the scatter-add moved another 20-40 KB.  Here the pairs' rows
// are the MMA's A operand in TENSOR MEMORY: a gather thread loads its pair's split row into registers (8 x LDG.128, NBUF
// chunks in flight per thread), tcgen05.st's the hi and lo halves into its own TMEM lane (16 + 16 columns) and the MMA
// runs tcgen05.mma [D], [A in TMEM], W[k] descriptor.  Shared memory then only carries W[k] (4 KB per chunk) and the fp32
// accumulators.  Roles, chunk schedule, barriers and the fixed summation order are v10's.
// TMEM columns: [0, 128) = 4 accumulator buffers D, [128, 128 + 32 S) = S operand stages (hi 16 | lo 16 columns).
// =====================================================================================================
constexpr int SC11_ENT = 16;
#define GPC_CONV_X_ROTATE 0x100
#define GPC_CONV_X_SKIPW 0x200          // pair-entry ring (chunks)
template <int TM, int S, int NBUF>
struct Sc11Smem {
    float acc[TM][GPC_C];                         // 16-byte chunk j of row r lives at chunk j ^ (r & 7)
    __align__(128) unsigned char w[S][4096];      // W[k] hi 2 KB | lo 2 KB, canonical K-major images
    __align__(128) unsigned char stg[4][NBUF][4096];   // per gather warp: 32 gathered split rows, 16-byte chunk j of row r at j ^ (r & 7)
    u64 ent[4][SC11_ENT][32];
    u16 rid[SC10_RID][128];
    u32 seg[4][GPC_K3 + 3];
    u32 cstart[GPC_K3 + 3];
    u16 tab[GPC_K3 * (TM / 128) + 4];
    __align__(8) u64 full[S];
    u64 empty[S];
    u64 dfull[SC10_NB];
    u64 dempty[SC10_NB];
    u32 tmem_base;
};

__device__ __forceinline__ void umma_bf16_ts(u32 tmem_d, u32 tmem_a, u64 bdesc, u32 idesc, u32 accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(u32 taddr, const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w),
                   "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w) : "memory");
}

template <int TM, int S, int NBUF, bool PROF>
__global__ void __launch_bounds__(SC10_THREADS, 1)
spconv_fwd_v11_kernel(const unsigned char *__restrict__ xs, const unsigned char *__restrict__ Wc, const u32 *__restrict__ seg_g,
                      const u64 *__restrict__ pairs, i64 n, const void *__restrict__ residual, int flags,
                      float *__restrict__ y, u32 *__restrict__ ys) {
    constexpr int QR = TM / 4;
    constexpr int P = 2 * NBUF - 1;              // pair entries are fetched P chunks ahead (group accounting below)
    static_assert(P < SC11_ENT && S + SC10_NB <= SC10_RID && NBUF <= S, "ring depths");
    static_assert(128 + 32 * S <= 512, "TMEM columns");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Sc11Smem<TM, S, NBUF> &s = *reinterpret_cast<Sc11Smem<TM, S, NBUF> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = blockIdx.x;
    const i64 n_sub = (n + QR - 1) / QR;
    long long acc_t[3] = {0, 0, 0}, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t_begin = 0, t_setup = 0, t_mark = 0;
    long long fine[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    SC10_T(t_begin);

    for (int i = tid; i < 4 * (GPC_K3 + 1); i += SC10_THREADS) {
        const int q = i / (GPC_K3 + 1), k = i - q * (GPC_K3 + 1);
        const i64 st = t * 4 + q;
        s.seg[q][k] = st < n_sub ? seg_g[st * (GPC_K3 + 1) + k] : 0u;
    }
    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init((u32)__cvta_generic_to_shared(&s.full[i]), 128);
            mbar_init((u32)__cvta_generic_to_shared(&s.empty[i]), 1);
        }
        for (int i = 0; i < SC10_NB; ++i) {
            mbar_init((u32)__cvta_generic_to_shared(&s.dfull[i]), 1);
            mbar_init((u32)__cvta_generic_to_shared(&s.dempty[i]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        const u32 dst = (u32)__cvta_generic_to_shared(&s.tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    // offsets are visited in the order k = (i + rot) % 125: a fixed function of the tile (so every row still has ONE summation
    // order, identical in encoder and decoder) that keeps the 148 CTAs from all fetching the same W[k] image at the same time
    const int rot = (flags & GPC_CONV_X_ROTATE) ? (int)((t * 37) % GPC_K3) : 0;
    if (tid < GPC_K3) {
        const int k = (tid + rot) % GPC_K3;
        u32 m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) m = max(m, s.seg[q][k + 1] - s.seg[q][k]);
        s.cstart[tid] = (m + 31) >> 5;
    }
    __syncthreads();
    if (warp == 0) {
        u32 v[4], sum = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int k = lane * 4 + i; v[i] = k < GPC_K3 ? s.cstart[k] : 0u; sum += v[i]; }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
        u32 run = incl - sum;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int k = lane * 4 + i; if (k <= GPC_K3) s.cstart[k] = run; run += v[i]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid < GPC_K3) {
        const u32 b = s.cstart[tid], e = s.cstart[tid + 1];
        for (u32 c = b; c < e; ++c) s.tab[c] = (u16)(((tid + rot) % GPC_K3) | ((c - b) << 8));
    }
    __syncthreads();
    const u32 n_chunks = s.cstart[GPC_K3];
    const u32 tmem_d = s.tmem_base;
    const u32 w0 = (u32)__cvta_generic_to_shared(&s.w[0][0]);
    const u32 full0 = (u32)__cvta_generic_to_shared(&s.full[0]), empty0 = (u32)__cvta_generic_to_shared(&s.empty[0]);
    const u32 dfull0 = (u32)__cvta_generic_to_shared(&s.dfull[0]), dempty0 = (u32)__cvta_generic_to_shared(&s.dempty[0]);
    SC10_T(t_setup);

    if (warp >= 4 && warp < 8) {
        // =================================================================== gather warps
        // Eight lanes copy one 128 B row (coalesced: a warp instruction touches 4 rows, not 32), cp.async, into a per-warp staging
        // slot; NBUF - 1 chunks later lane l reads row l back (conflict-free through the chunk swizzle) and stores it to TMEM.
        const int q = warp - 4;
        const int pt = q * 32 + lane;
        const int g8 = lane >> 3, j8 = lane & 7;
        const u32 ent0 = (u32)__cvta_generic_to_shared(&s.ent[q][0][lane]);
        const u32 stg0 = (u32)__cvta_generic_to_shared(&s.stg[q][0][0]);
        const u32 lane_base = tmem_d + ((u32)(q * 32) << 16) + 128u;       // this warp's TMEM lanes, first operand column
        auto pair_idx = [&](u32 c) -> u32 {
            const u32 kj = s.tab[c], k = kj & 0xFFu, j = kj >> 8;
            const u32 idx = s.seg[q][k] + 32u * j + (u32)lane;
            return idx < s.seg[q][k + 1] ? idx : 0xFFFFFFFFu;
        };
        auto issue_w = [&](u32 c) {
            const u32 si = c % (u32)S;
            if (c >= (u32)S) mbar_wait(empty0 + si * 8, ((c / (u32)S) & 1u) ^ 1u);
            if ((flags & GPC_CONV_X_SKIPW) && c >= (u32)S) return;          // timing experiment only (wrong results)
            const unsigned char *wsrc = Wc + (size_t)(s.tab[c] & 0xFFu) * 4096 + pt * 16;
            cp_async16(w0 + si * 4096 + pt * 16, wsrc);
            cp_async16(w0 + si * 4096 + 2048 + pt * 16, wsrc + 2048);
        };
        auto issue_rows = [&](u32 c) {
            const u32 kj = s.tab[c], k = kj & 0xFFu, j = kj >> 8;
            const u32 first = s.seg[q][k] + 32u * j, cnt = min(32u, s.seg[q][k + 1] - min(s.seg[q][k + 1], first));
            const u64 *ent = &s.ent[q][c % SC11_ENT][0];
            const u32 dst = stg0 + (c % (u32)NBUF) * 4096;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u32 r = 4u * i + g8;
                if (r < cnt) cp_async16(dst + r * 128 + ((j8 ^ (r & 7)) << 4), xs + (size_t)(u32)ent[r] * 128 + j8 * 16);
            }
            s.rid[c % SC10_RID][pt] = (u32)lane < cnt ? (u16)((u32)(ent[lane] >> 32) & 0xFFFFu) : (u16)0xFFFFu;
        };
        // ---- prologue: entries of chunks 0..P-1, then W and rows of chunks 0..NBUF-2
        for (u32 c = 0; c < (u32)P && c < n_chunks; ++c) {
            const u32 idx = pair_idx(c);
            if (idx != 0xFFFFFFFFu) cp_async8(ent0 + (c % SC11_ENT) * 256, pairs + idx);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NBUF - 1; ++u) {
            if ((u32)u < n_chunks) { issue_w(u); issue_rows(u); }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // ---- steady state.  Group accounting: every iteration commits ONE group = {entry of chunk c + P, W and rows of chunk
        // c + NBUF - 1}; cp.async.wait_group<NBUF-1> in iteration c therefore covers W and rows of chunk c (committed NBUF - 1
        // iterations ago) and the entry of chunk c - (NBUF-1) + P = c + NBUF, which iteration c + 1 dereferences.
        // All shared-memory table reads of an iteration are issued together at its top (plain loads, no asm in between), so their
        // latencies overlap instead of forming a load -> address -> cp.async chain per copy.
        for (u32 c = 0; c < n_chunks; ++c) {
            SC10_T(t0);
            if (PROF) t_mark = t0;
            __syncwarp();                                        // the slot re-filled below was read by other lanes last iteration
            SC10_MARK(0);
            const u32 cr = c + NBUF - 1, ce = c + P;             // chunk whose rows / W are issued now, chunk whose entries are fetched now
            const bool do_rows = cr < n_chunks, do_ent = ce < n_chunks;
            const u32 kj_r = s.tab[do_rows ? cr : c], kj_e = s.tab[do_ent ? ce : c];
            const u32 k_r = kj_r & 0xFFu, k_e = kj_e & 0xFFu;
            const u32 beg_r = s.seg[q][k_r] + 32u * (kj_r >> 8), end_r = s.seg[q][k_r + 1];
            const u32 idx_e = s.seg[q][k_e] + 32u * (kj_e >> 8) + (u32)lane, end_e = s.seg[q][k_e + 1];
            const u32 cnt = do_rows ? min(32u, end_r - min(end_r, beg_r)) : 0u;
            const u64 *ent = &s.ent[q][cr % SC11_ENT][0];
            u64 e8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e8[i] = ent[4 * i + g8];
            const u64 e_own = ent[lane];
            SC10_MARK(1);
            if (do_rows) {
                const u32 si = cr % (u32)S;
                if (cr >= (u32)S) mbar_wait(empty0 + si * 8, ((cr / (u32)S) & 1u) ^ 1u);
                SC10_MARK(2);
                const unsigned char *wsrc = Wc + (size_t)k_r * 4096 + pt * 16;
                cp_async16(w0 + si * 4096 + pt * 16, wsrc);
                cp_async16(w0 + si * 4096 + 2048 + pt * 16, wsrc + 2048);
                const u32 dst = stg0 + (cr % (u32)NBUF) * 4096;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const u32 r = 4u * i + g8;
                    if (r < cnt) cp_async16(dst + r * 128 + ((j8 ^ (r & 7)) << 4), xs + (size_t)(u32)e8[i] * 128 + j8 * 16);
                }
                s.rid[cr % SC10_RID][pt] = (u32)lane < cnt ? (u16)((u32)(e_own >> 32) & 0xFFFFu) : (u16)0xFFFFu;
            }
            if (do_ent && idx_e < end_e) cp_async8(ent0 + (ce % SC11_ENT) * 256, pairs + idx_e);
            SC10_MARK(3);
            asm volatile("cp.async.commit_group;" ::: "memory");
            SC10_T(t1);
            cp_async_wait<NBUF - 1>();
            SC10_MARK(4);
            __syncwarp();                                        // rows / entries copied by the other lanes of the warp are visible
            SC10_MARK(5);
            SC10_T(t2);
            // ---- publish chunk c: staging row `lane` -> registers -> this thread's TMEM lane (idle lanes store stale bytes: their D rows are ignored)
            const u32 src = stg0 + (c % (u32)NBUF) * 4096 + (u32)lane * 128;
            uint4 v[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[jj].x), "=r"(v[jj].y), "=r"(v[jj].z), "=r"(v[jj].w)
                             : "r"(src + (u32)((jj ^ (lane & 7)) << 4)));
            const u32 si = c % (u32)S;
            const u32 ta = lane_base + si * 32;
            tmem_st16(ta, v[0], v[1], v[2], v[3]);
            SC10_MARK(6);
            tmem_st16(ta + 16, v[4], v[5], v[6], v[7]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            SC10_MARK(7);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            SC10_MARK(8);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(full0 + si * 8);
            SC10_MARK(9);
            SC10_T(t3);
            SC10_ACCUM(0, t0, t1); SC10_ACCUM(1, t1, t2); SC10_ACCUM(2, t2, t3);
        }
        cp_async_wait<0>();
        if (PROF && lane == 0 && warp == 4) {
            atomicAdd(&g_sc10_prof[0], (unsigned long long)acc_t[0]); atomicAdd(&g_sc10_prof[1], (unsigned long long)acc_t[1]);
            atomicAdd(&g_sc10_prof[2], (unsigned long long)acc_t[2]);
            for (int i = 0; i < 10; ++i) atomicAdd(&g_sc10_prof[16 + i], (unsigned long long)fine[i]);
        }
    } else if (warp == 8) {
        // =================================================================== MMA issue (whole warp loops, one elected lane issues)
        {
            u32 st_i = 0, st_ph = 0;
            for (u32 c = 0; c < n_chunks; ++c) {
                const u32 b = c & (SC10_NB - 1), v = c / SC10_NB;
                SC10_T(t0);
                mbar_wait(full0 + st_i * 8, st_ph);
                SC10_T(t1);
                if (v > 0) mbar_wait(dempty0 + b * 8, (v & 1u) ^ 1u);
                SC10_T(t2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 a_hi = tmem_d + 128u + st_i * 32, a_lo = a_hi + 16;
                const u32 b_hi = w0 + st_i * 4096, b_lo = b_hi + 2048;
                const u32 d = tmem_d + b * 32;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_bf16_ts(d, a_lo + ks * 8, umma_desc(b_hi + ks * 256), SC9_IDESC, ks);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_bf16_ts(d, a_hi + ks * 8, umma_desc(b_lo + ks * 256), SC9_IDESC, 1u);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) umma_bf16_ts(d, a_hi + ks * 8, umma_desc(b_hi + ks * 256), SC9_IDESC, 1u);
                    umma_commit(empty0 + st_i * 8);
                    umma_commit(dfull0 + b * 8);
                }
                __syncwarp();
                SC10_T(t3);
                SC10_ACCUM(0, t0, t1); SC10_ACCUM(1, t1, t2); SC10_ACCUM(2, t2, t3);
                if (++st_i == (u32)S) { st_i = 0; st_ph ^= 1u; }
            }
            if (PROF && lane == 0) {
                atomicAdd(&g_sc10_prof[3], (unsigned long long)acc_t[0]); atomicAdd(&g_sc10_prof[4], (unsigned long long)acc_t[1]);
                atomicAdd(&g_sc10_prof[5], (unsigned long long)acc_t[2]); atomicAdd(&g_sc10_prof[9], (unsigned long long)n_chunks);
            }
        }
    } else {
        // =================================================================== epilogue warps (quarter q = warp)
        const int q = warp;
        float4 *accq = reinterpret_cast<float4 *>(&s.acc[q * QR][0]);
        for (int i = lane; i < QR * 8; i += 32) accq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        for (u32 c = 0; c < n_chunks; ++c) {
            const u32 b = c & (SC10_NB - 1), v = c / SC10_NB;
            SC10_T(t0);
            mbar_wait(dfull0 + b * 8, v & 1u);
            SC10_T(t1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const u32 r0 = s.rid[c % SC10_RID][q * 32 + lane];
            const bool any = __any_sync(0xFFFFFFFFu, r0 != 0xFFFFu);
            u32 d[32];
            if (any) {
                const u32 taddr = tmem_d + ((u32)(q * 32) << 16) + b * 32;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                               "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]),
                               "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]),
                               "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty0 + b * 8);
            SC10_T(t2);
            if (r0 != 0xFFFFu) {
                float4 *a = accq + r0 * 8;
                const u32 sw = r0 & 7u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 w = a[j ^ sw];
                    w.x += __uint_as_float(d[4 * j]); w.y += __uint_as_float(d[4 * j + 1]);
                    w.z += __uint_as_float(d[4 * j + 2]); w.w += __uint_as_float(d[4 * j + 3]);
                    a[j ^ sw] = w;
                }
            }
            SC10_T(t3);
            SC10_ACCUM(0, t0, t1); SC10_ACCUM(1, t1, t2); SC10_ACCUM(2, t2, t3);
        }
        __syncwarp();
        SC10_T(t0);
        if (PROF && lane == 0 && warp == 0) {
            atomicAdd(&g_sc10_prof[6], (unsigned long long)acc_t[0]); atomicAdd(&g_sc10_prof[7], (unsigned long long)acc_t[1]);
            atomicAdd(&g_sc10_prof[8], (unsigned long long)acc_t[2]);
        }
        const bool relu = (flags & GPC_CONV_RELU) != 0, res_split = (flags & GPC_CONV_RES_SPLIT) != 0;
        const i64 g0 = t * TM + (i64)q * QR;
        const int rows = (int)max((i64)0, min((i64)QR, n - g0));
        const int cp = lane & 15;
        for (int r = 0; r < rows; ++r) {
            const float4 a4 = accq[r * 8 + ((cp >> 1) ^ (r & 7))];
            float2 a = (cp & 1) ? make_float2(a4.z, a4.w) : make_float2(a4.x, a4.y);
            const i64 g = g0 + r;
            if (residual) {
                float2 rv;
                if (res_split) {
                    const u32 *rs = reinterpret_cast<const u32 *>(residual) + g * 32;
                    rv = join_bf16(__ldg(rs + cp), __ldg(rs + 16 + cp));
                } else {
                    rv = __ldg(reinterpret_cast<const float2 *>(residual) + g * 16 + cp);
                }
                a.x += rv.x; a.y += rv.y;
            }
            if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); }
            if (y && lane < 16) reinterpret_cast<float2 *>(y)[g * 16 + cp] = a;
            if (ys) {
                u32 hi, lo;
                split_bf16(a.x, a.y, hi, lo);
                ys[g * 32 + lane] = lane < 16 ? hi : lo;
            }
        }
        if (PROF && lane == 0 && warp == 0) atomicAdd(&g_sc10_prof[12], (unsigned long long)(clock64() - t0));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_d) : "memory");
    }
    if (PROF && tid == 0) {
        atomicAdd(&g_sc10_prof[10], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_sc10_prof[11], (unsigned long long)(t_setup - t_begin));
    }
}

template <int TM, int S, int NBUF, bool PROF = false>
static int launch_spconv_v11(const void *xs, const void *Wc, const u32 *seg, const u64 *pairs, i64 n, const void *residual,
                             int flags, float *y, void *ys, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc11Smem<TM, S, NBUF>) + 128;
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v11_kernel<TM, S, NBUF, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_fwd_v11_kernel<TM, S, NBUF, PROF><<<(unsigned)tiles, SC10_THREADS, smem, st>>>(
        (const unsigned char *)xs, (const unsigned char *)Wc, seg, pairs, n, residual, flags, y, (u32 *)ys);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variant 80: pair stream with pad = 1 and tile_rows = cta_rows / 4; xs = split rows; y (fp32) and / or ys (split) output
extern "C" int gpc_spconv_fwd_v10(const void *xs, const void *Wc, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                  int cta_rows, const void *residual, int flags, float *y, void *ys, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(y || ys, GPC_EINVAL, "no output requested");
    GPC_REQUIRE(xs != ys && xs != (const void *)y, GPC_EINVAL, "conv is out of place (rows are gathered from xs while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 80) {
        if (cta_rows == 256) return launch_spconv_v10<256, 8, 6>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 512) return launch_spconv_v10<512, 7, 5>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_v10<1024, 4, 2>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    }
    if (variant == 91 || variant == 92) {
        flags |= variant == 91 ? GPC_CONV_X_ROTATE : GPC_CONV_X_SKIPW;
        if (cta_rows == 512) return launch_spconv_v11<512, 8, 5>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_v11<1024, 6, 3>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    }
    if (variant == 90) {             // v11: gathered operand through tensor memory
        if (cta_rows == 512) return launch_spconv_v11<512, 8, 5>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_v11<1024, 6, 3>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    }
    if (variant == 99) {             // v11 role profile
        if (cta_rows == 512) return launch_spconv_v11<512, 8, 5, true>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_v11<1024, 6, 3, true>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    }
    if (variant == 89) {             // role profile (gpc_debug_conv_profile)
        if (cta_rows == 512) return launch_spconv_v10<512, 7, 5, true>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_v10<1024, 4, 2, true>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    }
    gpc_set_error("unsupported conv v10 variant %d / cta_rows %d", variant, cta_rows);
    return GPC_EINVAL;
}
