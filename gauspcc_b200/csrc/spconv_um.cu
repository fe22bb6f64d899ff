// spconv_um.cu -- the sparse convolution of the big octree levels: TMA row gather -> tcgen05.mma -> per-row fp32 sums.
//
// Reference: every spnn.Conv3d(C, C, 5) of src/ai_pcc/GausPcgc/network_ue_4stage_conv.py:17-62 (torchsparse 2.1.0 gather -
// implicit GEMM - scatter); semantics as restated in SURVEY.md 8(c):  y[o] = act( sum_k W[k]^T x[nbr_k(o)] (+ residual[o]) ).
//
// Formulation (TRANSPOSED: the pairs are the N dimension of the MMA, so a chunk is as long as the tile has pairs of one offset,
// in steps of 16, and nothing is padded to 128):
//     D^T[co][p] = sum_ci W[k]^T[co][ci] . x[nbr(p)][ci]          for the pairs p of ONE offset k inside one tile of TM output rows
//   A operand  W[k]^T, bf16 hi | lo, in TENSOR MEMORY (32 columns), the 32 output channels replicated into all four lane
//              quadrants (M = 128): every epilogue warp finds the whole product in the quadrant it may read
//   B operand  the gathered rows, K-major with 128 B swizzle, in shared memory: written by TMA tile::gather4 (four arbitrary rows
//              per instruction, one lane each; padding entries are out-of-bounds rows = zero fill), never touched by a warp
//   D          fp32 in tensor memory, N columns; six MMAs per chunk (K = 2 x 16 for Whi.xhi, Whi.xlo, Wlo.xhi)
//   epilogue   warp e owns TM/4 consecutive output rows (fp32 sums in shared memory).  The pairs of a chunk are sorted by output
//              row, so the warp's pairs are ONE column range of D: tcgen05.ld gives lane = channel, register = pair, and a pair is
//              added to its row with one conflict-free 128 B read-modify-write.  Offsets arrive in ascending order and a row is
//              touched by one warp only: one fixed summation order per row (encoder and decoder CDFs stay bit-identical).
// Activations are "split rows" (32 x bf16 hi | 32 x bf16 lo per row, spconv_fmt.cu): a gathered row is an operand row as it
// stands; x = hi + lo to 16 mantissa bits and the three-term product keeps the contraction within ~1.5e-4 of fp32 on the
// probabilities.
//
// Warp roles (12 warps, two CTAs per SM): 0 = MMA issue (+ TMEM allocation), 1-3 = TMA producers, 4-7 = weights -> TMEM (one warp
// per lane quadrant), 8-11 = epilogue.  All hand-offs are mbarriers (plus one 4-warp barrier per chunk inside the epilogue);
// after the last chunk all twelve warps write the tile out.
#include <cuda.h>
#include "umma.cuh"

constexpr int UM_THREADS = 12 * 32;
constexpr int UM_IS = 8;              // index ring: the row indices of a chunk are requested UM_IS chunks before they are used

template <int TM, int NMAX, int GS, int DS, int WS>
struct UmSmem {
    float acc[TM + 4][GPC_C];                          // rows TM .. TM + 3: one dummy row per epilogue warp (tail lanes of a batch)
    __align__(1024) unsigned char g[GS][NMAX * 128];   // gathered rows of a chunk (128 B swizzle atoms of 8 rows)
    __align__(16) u16 rid[GS + DS][NMAX];              // output row (within the tile) of every pair of a chunk; 0xFFFF = padding
    __align__(16) uint4 idx[3][2 * UM_IS][NMAX / 8];   // per producer warp: the input rows of ITS quads (4 rows = 16 B) of the next chunks
    u32 seg[GPC_K3 + 3];
    u32 cstart[GPC_K3 + 3];
    u32 tab[GPC_K3 * (TM / NMAX) + 8];                 // chunk -> k | j << 8 | len << 12 | last chunk of the offset << 20
    __align__(8) u64 full_g[GS];
    u64 empty_g[GS];
    u64 full_d[DS];
    u64 empty_d[DS];
    u64 full_w[WS];
    u64 empty_w[WS];
    u32 tmem_base;
};

// K-major, SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused, descriptor version 1, layout type 2
__device__ __forceinline__ u64 um_desc_sw128(u32 smem_addr) {
    return (u64)((smem_addr >> 4) & 0x3FFFu) | ((u64)1 << 16) | ((u64)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void um_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void um_gather4(u32 dst, const CUtensorMap *tmap, u32 bar, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<u64>(tmap)), "r"(bar), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void um_bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void um_tmem_st32(u32 taddr, const uint4 (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(v[0].x), "r"(v[0].y), "r"(v[0].z), "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w),
                   "r"(v[2].x), "r"(v[2].y), "r"(v[2].z), "r"(v[2].w), "r"(v[3].x), "r"(v[3].y), "r"(v[3].z), "r"(v[3].w),
                   "r"(v[4].x), "r"(v[4].y), "r"(v[4].z), "r"(v[4].w), "r"(v[5].x), "r"(v[5].y), "r"(v[5].z), "r"(v[5].w),
                   "r"(v[6].x), "r"(v[6].y), "r"(v[6].z), "r"(v[6].w), "r"(v[7].x), "r"(v[7].y), "r"(v[7].z), "r"(v[7].w) : "memory");
}
// shared-memory accumulator accesses of the epilogue: volatile asm keeps the written order (all loads of a batch, then all stores)
// without a "memory" clobber that would serialise every read-modify-write behind the previous one
__device__ __forceinline__ float um_lds(u32 addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void um_sts(u32 addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

// role profile (PROF instantiation only; tools/um_check.py): cycles summed over chunks, lane 0 of one warp per role
//  [0] producer: wait empty_g   [1] producer: issue (includes the wait for the chunk's row indices)
//  [2] mma: wait full_w   [3] mma: wait full_g   [4] mma: wait empty_d   [5] mma: issue + commit
//  [6] weights: wait empty_w   [7] weights: store + arrive
//  [8] epilogue: wait full_d   [9] epilogue: column range   [10] epilogue: tcgen05.ld + read-modify-write
//  [11] chunks   [12] CTA total   [13] setup   [14] write-out   [15] CTAs
__device__ unsigned long long g_um_prof[16];
extern "C" int gpc_debug_conv_um_profile(unsigned long long *out_h, int reset) {
    GPC_CUDA_CHECK(cudaDeviceSynchronize());
    if (out_h) GPC_CUDA_CHECK(cudaMemcpyFromSymbol(out_h, g_um_prof, sizeof(unsigned long long) * 16));
    if (reset) { unsigned long long z[16] = {0}; GPC_CUDA_CHECK(cudaMemcpyToSymbol(g_um_prof, z, sizeof(z))); }
    return GPC_OK;
}
#define UM_T(var) do { if (PROF) var = clock64(); } while (0)
#define UM_ACC(i, a, b) do { if (PROF) pacc[i] += (b) - (a); } while (0)
#define UM_FLUSH(slot, i) do { if (PROF && lane == 0) atomicAdd(&g_um_prof[slot], (unsigned long long)pacc[i]); } while (0)

template <int TM, int NMAX, int GS, int DS, int WS, int TCOLS, int MINB, bool TMAG, bool PROF>
__global__ void __launch_bounds__(UM_THREADS, MINB)
spconv_um_kernel(const __grid_constant__ CUtensorMap tmap, const unsigned char *__restrict__ xs, const uint4 *__restrict__ Wp, const u32 *__restrict__ seg_g,
                 const u32 *__restrict__ pair_nbr, const u16 *__restrict__ pair_row, i64 n, i64 tile0,
                 const void *__restrict__ residual, int flags, float *__restrict__ y, u32 *__restrict__ ys) {
    constexpr int RS = GS + DS;                  // row-id ring: a chunk's ids live from its gather to the start of its epilogue
    constexpr int CPW = NMAX / 4;                // columns of a full chunk per epilogue warp
    constexpr int NP = 3;                        // producer warps (1, 2, 3)
    constexpr u32 A_COL = DS * NMAX;             // first weight column of tensor memory
    static_assert(A_COL + WS * 32 <= TCOLS, "TMEM columns");
    static_assert(NMAX == 64 || NMAX == 128, "chunk length");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    typedef UmSmem<TM, NMAX, GS, DS, WS> Smem;
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    if (((u32)__cvta_generic_to_shared(smem_raw) & 1023u) != 0u) __trap();      // 128 B swizzle atoms are 1024 B aligned
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = tile0 + blockIdx.x;
    long long pacc[4] = {0, 0, 0, 0}, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t_begin = 0, t_setup = 0;
    UM_T(t_begin);

    // ---- setup: segment table, chunk table, barriers, tensor memory, zeroed sums
    for (int i = tid; i <= GPC_K3; i += UM_THREADS) s.seg[i] = seg_g[t * (GPC_K3 + 1) + i];
    for (int i = tid; i < TM * GPC_C / 4; i += UM_THREADS) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        for (int i = 0; i < GS; ++i) { mbar_init((u32)__cvta_generic_to_shared(&s.full_g[i]), TMAG ? NP : NP * 32); mbar_init((u32)__cvta_generic_to_shared(&s.empty_g[i]), 1); }
        for (int i = 0; i < DS; ++i) { mbar_init((u32)__cvta_generic_to_shared(&s.full_d[i]), 1); mbar_init((u32)__cvta_generic_to_shared(&s.empty_d[i]), 4); }
        for (int i = 0; i < WS; ++i) { mbar_init((u32)__cvta_generic_to_shared(&s.full_w[i]), 4); mbar_init((u32)__cvta_generic_to_shared(&s.empty_w[i]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        const u32 dst = (u32)__cvta_generic_to_shared(&s.tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "n"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid < GPC_K3) s.cstart[tid] = (s.seg[tid + 1] - s.seg[tid] + NMAX - 1) / NMAX;
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
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    if (tid < GPC_K3) {                                          // chunk -> k | j << 8 | len << 12 | last << 20
        const u32 b = s.cstart[tid], e = s.cstart[tid + 1], tot = s.seg[tid + 1] - s.seg[tid];
        for (u32 c = b; c < e; ++c) {
            const u32 j = c - b, len = min((u32)NMAX, tot - j * NMAX);
            s.tab[c] = (u32)tid | (j << 8) | (len << 12) | ((c + 1 == e ? 1u : 0u) << 20);
        }
    }
    __syncthreads();
    const u32 n_chunks = s.cstart[GPC_K3];
    const u32 tmem = s.tmem_base;
    const u32 full_g0 = (u32)__cvta_generic_to_shared(&s.full_g[0]), empty_g0 = (u32)__cvta_generic_to_shared(&s.empty_g[0]);
    const u32 full_d0 = (u32)__cvta_generic_to_shared(&s.full_d[0]), empty_d0 = (u32)__cvta_generic_to_shared(&s.empty_d[0]);
    const u32 full_w0 = (u32)__cvta_generic_to_shared(&s.full_w[0]), empty_w0 = (u32)__cvta_generic_to_shared(&s.empty_w[0]);
    const u32 g0 = (u32)__cvta_generic_to_shared(&s.g[0][0]);
    UM_T(t_setup);

    if (warp == 0) {
        // =================================================================== MMA issue (whole warp loops, one elected lane issues)
        constexpr u32 IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);      // D f32, A = B = bf16 K-major, M = 128
        u32 wi = 0xFFFFFFFFu, ws = 0;
        for (u32 c = 0; c < n_chunks; ++c) {
            const u32 e = s.tab[c], len = (e >> 12) & 0xFFu;
            UM_T(t0);
            if (((e >> 8) & 0xFu) == 0) {                        // first chunk of an offset: its weights
                ++wi;
                ws = wi % (u32)WS;
                mbar_wait(full_w0 + ws * 8, (wi / (u32)WS) & 1u);
            }
            const u32 gs = c % (u32)GS, ds = c % (u32)DS, fd = c / (u32)DS;
            UM_T(t1);
            mbar_wait(full_g0 + gs * 8, (c / (u32)GS) & 1u);
            UM_T(t2);
            if (fd) mbar_wait(empty_d0 + ds * 8, (fd & 1u) ^ 1u);
            UM_T(t3);
            if (!TMAG) fence_async_smem();                       // rows written by cp.async (generic proxy) -> visible to the tensor core
            tmem_fence_after();
            if (elect_one()) {
                const u32 idesc = IDESC | ((len >> 3) << 17);
                const u32 d = tmem + ds * NMAX, a = tmem + A_COL + ws * 32, b = g0 + gs * (NMAX * 128);
                umma_bf16_ts(d, a, um_desc_sw128(b), idesc, 0u);                 // Whi . xhi   (K 0..15)
                umma_bf16_ts(d, a + 8, um_desc_sw128(b + 32), idesc, 1u);        //             (K 16..31)
                umma_bf16_ts(d, a, um_desc_sw128(b + 64), idesc, 1u);            // Whi . xlo
                umma_bf16_ts(d, a + 8, um_desc_sw128(b + 96), idesc, 1u);
                umma_bf16_ts(d, a + 16, um_desc_sw128(b), idesc, 1u);            // Wlo . xhi
                umma_bf16_ts(d, a + 24, um_desc_sw128(b + 32), idesc, 1u);
                umma_commit(empty_g0 + gs * 8);
                umma_commit(full_d0 + ds * 8);
                if ((e >> 20) & 1u) umma_commit(empty_w0 + ws * 8);              // last chunk of the offset
            }
            __syncwarp();
            UM_T(t4);
            UM_ACC(0, t0, t1); UM_ACC(1, t1, t2); UM_ACC(2, t2, t3); UM_ACC(3, t3, t4);
        }
        UM_FLUSH(2, 0); UM_FLUSH(3, 1); UM_FLUSH(4, 2); UM_FLUSH(5, 3);
        if (PROF && lane == 0) { atomicAdd(&g_um_prof[11], (unsigned long long)n_chunks); atomicAdd(&g_um_prof[15], 1ull); }
    } else if (warp < 4) {
        // =================================================================== producers: the rows of a chunk -> shared memory
        // Quad q of a chunk (rows 4q .. 4q + 3) belongs to warp 1 + q % 3.  The four row indices of a quad are one 16-byte word of the
        // pair stream: every warp copies ITS words of chunk c + UM_IS into its own ring with one cp.async per chunk, in the commit
        // group of chunk c's rows, so they have landed long before they are needed and no producer waits for another.
        //   cp.async (default): one instruction = four rows (eight lanes x 16 B each); 16-byte piece j of row r goes to piece
        //     j ^ (r & 7) (the 128 B swizzle the B descriptor expects); padding rows are not copied (their D columns are never added
        //     to a row).  A chunk is handed to the MMA warp two chunks later (cp.async.wait_group), behind a proxy fence.
        //   TMA tile::gather4 (flag GPC_CONV_TMA_GATHER): one lane = one quad; measured ~47 clk per UTMALDG issue (lane by lane
        //     from uniform registers), 2-3x slower than the cp.async path end to end (profiles/r02_conv_um.md).
        const u32 p = (u32)warp - 1u;
        const u32 rid0 = (u32)__cvta_generic_to_shared(&s.rid[0][0]);
        const u32 ring0 = (u32)__cvta_generic_to_shared(&s.idx[p][0][0]);
        constexpr u32 IR = 2 * UM_IS;
        auto prefetch = [&](u32 c) {                             // my index words of chunk c -> ring slot c % IR (lane t: quad p + 3 t)
            if (c < n_chunks) {
                const u32 e = s.tab[c], k = e & 0xFFu, len = (e >> 12) & 0xFFu;
                const u32 q = p + (u32)NP * (u32)lane;
                if (4u * q < len)
                    cp_async16(ring0 + ((c % IR) * (NMAX / 8) + (u32)lane) * 16u,
                               reinterpret_cast<const uint4 *>(pair_nbr + s.seg[k] + (u32)NMAX * ((e >> 8) & 0xFu)) + q);
            }
        };
        for (u32 c = 0; c < (u32)UM_IS; ++c) prefetch(c);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        const u32 rl = (u32)lane >> 3, j8 = (u32)lane & 7u;
        for (u32 c = 0; c < n_chunks + 2; ++c) {
            UM_T(t0);
            if (c < n_chunks) {
                const u32 e = s.tab[c], k = e & 0xFFu, len = (e >> 12) & 0xFFu;
                const u32 start = s.seg[k] + (u32)NMAX * ((e >> 8) & 0xFu);
                const u32 quads = len >> 2;
                const u32 gs = c % (u32)GS, f = c / (u32)GS;
                const u32 slot = g0 + gs * (NMAX * 128);
                if (f) mbar_wait(empty_g0 + gs * 8, (f & 1u) ^ 1u);      // the MMAs of chunk c - GS have read the slot
                UM_T(t1);
                if (TMAG) {
                    const u32 q = p + (u32)NP * (u32)lane;
                    const u32 mine = quads > p ? (quads - p + (u32)NP - 1u) / (u32)NP : 0u;
                    if (lane == 0) um_expect_tx(full_g0 + gs * 8, mine * 512u + (p == 0 ? len * 2u : 0u));
                    __syncwarp();
                    if (q < quads) {
                        const uint4 ix = s.idx[p][c % IR][lane];
                        um_gather4(slot + q * 512, &tmap, full_g0 + gs * 8, (int)ix.x, (int)ix.y, (int)ix.z, (int)ix.w);
                    }
                    if (p == 0 && lane == 0) um_bulk_g2s(rid0 + (c % (u32)RS) * (NMAX * 2), pair_row + start, len * 2u, full_g0 + gs * 8);
                } else {
                    const u32 *ring = reinterpret_cast<const u32 *>(&s.idx[p][c % IR][0]);
                    u32 t = 0;
                    for (u32 q = p; q < quads; q += (u32)NP, ++t) {
                        const u32 nb = ring[4u * t + rl];
                        const u32 r = 4u * q + rl;
                        if (nb != 0xFFFFFFFFu) cp_async16(slot + r * 128u + ((j8 ^ (r & 7u)) << 4), xs + (size_t)nb * 128 + j8 * 16);
                    }
                    if (p == 0 && 8u * (u32)lane < len)                   // the chunk's output rows, for the epilogue
                        cp_async16(rid0 + (c % (u32)RS) * (NMAX * 2) + (u32)lane * 16u, pair_row + start + 8u * (u32)lane);
                }
                prefetch(c + (u32)UM_IS);
            }
            cp_async_commit();
            UM_T(t2);
            if (!TMAG && c >= 2) {                               // chunk c - 2: my copies have landed -> visible to the tensor core -> arrive
                cp_async_wait<2>();
                fence_async_smem();
                mbar_arrive(full_g0 + ((c - 2) % (u32)GS) * 8);
                __syncwarp();                                    // index words copied by the other lanes are visible
            } else if (TMAG) {
                cp_async_wait<2>();                              // index words only
                __syncwarp();
            }
            UM_T(t3);
            UM_ACC(0, t0, t1); UM_ACC(1, t1, t2); UM_ACC(2, t2, t3);
        }
        cp_async_wait<0>();
        if (warp == 1) { UM_FLUSH(0, 0); UM_FLUSH(1, 1); UM_FLUSH(6, 2); }
    } else if (warp < 8) {
        // =================================================================== weights: W[k]^T hi | lo -> this warp's lane quadrant
        const u32 q = (u32)warp & 3u;
        const u32 ta = tmem + ((q * 32u) << 16) + A_COL;
        int k = 0;
        while (k < GPC_K3 && s.seg[k + 1] == s.seg[k]) ++k;
        uint4 vn[8];
        if (k < GPC_K3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vn[i] = __ldg(Wp + ((size_t)k * 32 + lane) * 8 + i);
        }
        for (u32 wi = 0; k < GPC_K3; ++wi) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = vn[i];
            ++k;
            while (k < GPC_K3 && s.seg[k + 1] == s.seg[k]) ++k;
            if (k < GPC_K3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) vn[i] = __ldg(Wp + ((size_t)k * 32 + lane) * 8 + i);
            }
            const u32 ws = wi % (u32)WS, f = wi / (u32)WS;
            UM_T(t0);
            if (f) mbar_wait(empty_w0 + ws * 8, (f & 1u) ^ 1u);
            UM_T(t1);
            tmem_fence_after();
            um_tmem_st32(ta + ws * 32, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tmem_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_w0 + ws * 8);
            UM_T(t2);
            UM_ACC(0, t0, t1); UM_ACC(1, t1, t2);
        }
        (void)t0; (void)t1;
    } else {
        // =================================================================== epilogue: warp e adds the columns [e CPW, (e + 1) CPW) of every
        // chunk to their rows.  Within a chunk the rows are distinct; a barrier between chunks keeps the offsets of a row in order.
        const u32 e = (u32)warp - 8u;                            // == warp & 3: the lane quadrant this warp may read
        const u32 acc0 = (u32)__cvta_generic_to_shared(&s.acc[0][0]) + (u32)lane * 4u;
        const u32 dummy = (u32)TM + e;
        const u32 td = tmem + ((e * 32u) << 16) + e * CPW;
        constexpr int NG = CPW / 16;                             // 16-column groups of a full chunk per warp
        // load(c): D columns and row ids of chunk c -> registers (asynchronous); done(c): wait for them and hand the D buffer back
        auto load = [&](u32 c, u32 (&d)[CPW], uint4 (&r)[2 * NG], u32 &len) {
            const u32 ds = c % (u32)DS;
            len = (s.tab[c] >> 12) & 0xFFu;
            mbar_wait(full_d0 + ds * 8, (c / (u32)DS) & 1u);
            tmem_fence_after();
            const uint4 *rid = reinterpret_cast<const uint4 *>(&s.rid[c % (u32)RS][e * CPW]);
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) {
                if (e * CPW + gq * 16 < len) {                   // len is a multiple of 16: a group of 16 columns is inside or outside
                    tmem_ld16(td + ds * NMAX + gq * 16, *reinterpret_cast<u32(*)[16]>(&d[gq * 16]));
                    r[2 * gq] = rid[2 * gq];
                    r[2 * gq + 1] = rid[2 * gq + 1];
                }
            }
        };
        auto done = [&](u32 c) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_d0 + (c % (u32)DS) * 8);     // values and row ids are in registers: hand the buffer back
        };
        auto rmw = [&](const u32 (&d)[CPW], const uint4 (&r)[2 * NG], u32 len) {
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) {
                if (e * CPW + gq * 16 < len) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint4 rr = r[2 * gq + h];
                        const u32 w[4] = {rr.x, rr.y, rr.z, rr.w};
                        u32 ra[8];
                        float a[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const u32 row = (jj & 1) ? (w[jj >> 1] >> 16) : (w[jj >> 1] & 0xFFFFu);
                            ra[jj] = acc0 + min(row, dummy) * 128u;                   // padding (0xFFFF) -> dummy row
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) a[jj] = um_lds(ra[jj]);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) um_sts(ra[jj], a[jj] + __uint_as_float(d[gq * 16 + 8 * h + jj]));
                    }
                }
            }
        };
        u32 dA[CPW], dB[CPW], lenA = 0, lenB = 0;
        uint4 rA[2 * NG], rB[2 * NG];
        if (n_chunks) { load(0, dA, rA, lenA); done(0); }
        for (u32 c = 0; c < n_chunks; c += 2) {
            // chunk c (set A) is in registers; request chunk c + 1 (set B) before adding chunk c to its rows, and so on
            UM_T(t0);
            if (c + 1 < n_chunks) load(c + 1, dB, rB, lenB);
            UM_T(t1);
            asm volatile("bar.sync 1, 128;" ::: "memory");       // every epilogue warp is done with the previous chunk
            rmw(dA, rA, lenA);
            UM_T(t2);
            if (c + 1 < n_chunks) {
                done(c + 1);
                if (c + 2 < n_chunks) load(c + 2, dA, rA, lenA);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                rmw(dB, rB, lenB);
                if (c + 2 < n_chunks) done(c + 2);
            }
            UM_T(t3);
            UM_ACC(0, t0, t1); UM_ACC(1, t1, t2); UM_ACC(2, t2, t3);
        }
        if (warp == 8) { UM_FLUSH(8, 0); UM_FLUSH(9, 1); UM_FLUSH(10, 2); }
        asm volatile("" ::: "memory");
    }
    tmem_fence_before();
    __syncthreads();
    UM_T(t0);
    if (warp == 0) {
        tmem_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS) : "memory");
    }
    // ---- write-out by all warps: (+ residual) (ReLU) -> fp32 rows and / or split rows, one 128 B store per row
    {
        const bool relu = (flags & GPC_CONV_RELU) != 0, res_split = (flags & GPC_CONV_RES_SPLIT) != 0;
        const i64 gbase = t * TM;
        const int rows = (int)max((i64)0, min((i64)TM, n - gbase));
        constexpr int NW = UM_THREADS / 32, UR = 4;
        for (int rb = warp; rb < rows; rb += NW * UR) {
            float v[UR];
            u32 wh[UR], wl[UR];
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const int r = rb + u * NW;
                v[u] = 0.f; wh[u] = 0u; wl[u] = 0u;
                if (r < rows) {
                    v[u] = s.acc[r][lane];
                    const i64 gr = gbase + r;
                    if (residual) {
                        if (res_split) {
                            const u32 *rs = reinterpret_cast<const u32 *>(residual) + gr * 32 + (lane >> 1);
                            wh[u] = __ldg(rs); wl[u] = __ldg(rs + 16);
                        } else {
                            wh[u] = __float_as_uint(__ldg(reinterpret_cast<const float *>(residual) + gr * 32 + lane));
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const int r = rb + u * NW;
                if (r < rows) {                                  // warp-uniform
                    const i64 gr = gbase + r;
                    float o = v[u];
                    if (residual) {
                        if (res_split) o += (lane & 1) ? __uint_as_float(wh[u] & 0xFFFF0000u) + __uint_as_float(wl[u] & 0xFFFF0000u)
                                                       : __uint_as_float(wh[u] << 16) + __uint_as_float(wl[u] << 16);
                        else o += __uint_as_float(wh[u]);
                    }
                    if (relu) o = fmaxf(o, 0.f);
                    if (y) y[gr * 32 + lane] = o;
                    if (ys) {
                        const float vh = bf16_round(o), vl = bf16_round(o - vh);
                        const u32 mine = (__float_as_uint(vh) >> 16) | (__float_as_uint(vl) & 0xFFFF0000u);      // hi in the low half, lo in the high half
                        const u32 other = __shfl_xor_sync(0xFFFFFFFFu, mine, 1);
                        // even lane 2j: hi word j = (hi[2j], hi[2j+1]); odd lane 2j+1: lo word 16 + j = (lo[2j], lo[2j+1])
                        const u32 word = (lane & 1) ? ((other >> 16) | (mine & 0xFFFF0000u)) : ((mine & 0xFFFFu) | (other << 16));
                        ys[gr * 32 + ((lane & 1) << 4) + (lane >> 1)] = word;
                    }
                }
            }
        }
    }
    if (PROF && tid == 0) {
        atomicAdd(&g_um_prof[14], (unsigned long long)(clock64() - t0));
        atomicAdd(&g_um_prof[12], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_um_prof[13], (unsigned long long)(t_setup - t_begin));
    }
}

// ---- host side
typedef CUresult (*um_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static um_encode_fn um_encoder() {
    static um_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (um_encode_fn)p;
    }
    return fn;
}

template <int TM, int NMAX, int GS, int DS, int WS, int TCOLS, int MINB, bool TMAG, bool PROF = false>
static int launch_spconv_um(const CUtensorMap &tmap, const void *xs, const void *Wp, const u32 *seg, const u32 *pair_nbr, const u16 *pair_row, i64 n,
                            i64 tile0, i64 tiles, const void *residual, int flags, float *y, void *ys, cudaStream_t st) {
    static bool configured = false;
    typedef UmSmem<TM, NMAX, GS, DS, WS> Smem;
    constexpr size_t smem = sizeof(Smem) + 1024;
    static_assert(smem * MINB <= 232448 - 1024 * MINB, "shared memory per SM");
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_um_kernel<TM, NMAX, GS, DS, WS, TCOLS, MINB, TMAG, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    spconv_um_kernel<TM, NMAX, GS, DS, WS, TCOLS, MINB, TMAG, PROF><<<(unsigned)tiles, UM_THREADS, smem, st>>>(
        tmap, (const unsigned char *)xs, (const uint4 *)Wp, seg, pair_nbr, pair_row, n, tile0, residual, flags, y, (u32 *)ys);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// xs = split rows [n][128 B]; Wp = this conv's slice of gpc_spconv_pack_weights_um; seg / pair_nbr / pair_row = the pair stream built
// with tile_rows (512 or 1024), pad = 16 and padding entries 0xFFFFFFFF / 0xFFFF.  Output rows [row0, row1) (whole tiles; row1 <= 0
// or >= n: to the end).  y (fp32 rows) and / or ys (split rows).
extern "C" int gpc_spconv_fwd_um(const void *xs, const void *Wp, const uint32_t *seg, const uint32_t *pair_nbr,
                                 const uint16_t *pair_row, int64_t n, int tile_rows, const void *residual, int flags, float *y,
                                 void *ys, int64_t row0, int64_t row1, void *stream) {
    const bool prof = (flags & GPC_CONV_PROFILE) != 0, tmag = (flags & GPC_CONV_TMA_GATHER) != 0;
    flags &= ~(GPC_CONV_PROFILE | GPC_CONV_TMA_GATHER);
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(y || ys, GPC_EINVAL, "no output requested");
    GPC_REQUIRE(xs != ys && xs != (const void *)y, GPC_EINVAL, "conv is out of place (rows are gathered from xs while y is written)");
    GPC_REQUIRE(tile_rows == 512 || tile_rows == 1024, GPC_EINVAL, "tile_rows must be 512 or 1024");
    if (row1 <= 0 || row1 > n) row1 = n;
    GPC_REQUIRE(row0 >= 0 && row0 % tile_rows == 0 && (row1 % tile_rows == 0 || row1 == n), GPC_EINVAL, "row range must cover whole tiles");
    if (row0 >= row1) return GPC_OK;
    um_encode_fn enc = um_encoder();
    GPC_REQUIRE(enc != nullptr, GPC_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {64, (cuuint64_t)n};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(xs), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { gpc_set_error("cuTensorMapEncodeTiled failed: %d", (int)rc); return GPC_ECUDA; }
    cudaStream_t st = as_stream(stream);
    const i64 tile0 = row0 / tile_rows, tiles = (row1 - row0 + tile_rows - 1) / tile_rows;
#define UM_GO(TMr, NMx, GSn, DSn, WSn, TC, MB) \
    do { \
        if (tmag) return prof ? launch_spconv_um<TMr, NMx, GSn, DSn, WSn, TC, MB, true, true>(tmap, xs, Wp, seg, pair_nbr, pair_row, n, tile0, tiles, residual, flags, y, ys, st) \
                              : launch_spconv_um<TMr, NMx, GSn, DSn, WSn, TC, MB, true, false>(tmap, xs, Wp, seg, pair_nbr, pair_row, n, tile0, tiles, residual, flags, y, ys, st); \
        return prof ? launch_spconv_um<TMr, NMx, GSn, DSn, WSn, TC, MB, false, true>(tmap, xs, Wp, seg, pair_nbr, pair_row, n, tile0, tiles, residual, flags, y, ys, st) \
                    : launch_spconv_um<TMr, NMx, GSn, DSn, WSn, TC, MB, false, false>(tmap, xs, Wp, seg, pair_nbr, pair_row, n, tile0, tiles, residual, flags, y, ys, st); \
    } while (0)
    if (tile_rows == 512) UM_GO(512, 64, 4, 3, 2, 256, 2);
    UM_GO(1024, 128, 4, 3, 4, 512, 1);
#undef UM_GO
}

// W [n_kernels*125][32 ci][32 co] fp32 -> Wp [n_kernels*125][32 co][32 words]: words 0..15 = bf16x2 (hi(W[2j][co]), hi(W[2j+1][co])),
// words 16..31 = the lo halves: one TMEM lane (128 B) per output channel, K pairs packed as the A operand wants them
__global__ void pack_weights_um_kernel(const float *__restrict__ W, u32 *__restrict__ Wp, i64 total) {
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (k, co, j)
    if (g >= total) return;
    const int j = (int)(g & 15), co = (int)((g >> 4) & 31);
    const i64 k = g >> 9;
    const float w0 = W[k * 1024 + (2 * j) * 32 + co], w1 = W[k * 1024 + (2 * j + 1) * 32 + co];
    const float h0 = bf16_round(w0), h1 = bf16_round(w1);
    const float l0 = bf16_round(w0 - h0), l1 = bf16_round(w1 - h1);
    u32 *dst = Wp + (k * 32 + co) * 32;
    dst[j] = (__float_as_uint(h0) >> 16) | (__float_as_uint(h1) & 0xFFFF0000u);
    dst[16 + j] = (__float_as_uint(l0) >> 16) | (__float_as_uint(l1) & 0xFFFF0000u);
}
extern "C" int gpc_spconv_pack_weights_um(const float *W, int n_kernels, void *Wp, void *stream) {
    const i64 total = (i64)n_kernels * GPC_K3 * 32 * 16;
    pack_weights_um_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(W, (u32 *)Wp, total);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
