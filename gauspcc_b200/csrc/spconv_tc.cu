// spconv_tc.cu -- the production sparse convolution: gather -> tcgen05.mma (operands in tensor memory) -> scatter-add.
//
// Reference: every spnn.Conv3d(C, C, 5) of src/ai_pcc/GausPcgc/network_ue_4stage_conv.py:17-62 (torchsparse 2.1.0 gather -
// implicit GEMM - scatter); semantics as restated in SURVEY.md 8(c):  y[o] = act( sum_k W[k]^T x[nbr_k(o)] (+ residual[o]) ).
//
// Data formats
//   activations  "split rows": 128 B per row = 32 x bf16 hi | 32 x bf16 lo, x = hi + lo (gpc_rows_split); the contraction is
//                hi.Whi + hi.Wlo + lo.Whi with fp32 accumulation: 1.5e-4 from fp32 on the probabilities
//   weights      gpc_spconv_pack_weights_umma: per offset k a 4 KB image = canonical K-major tiles of W[k]^T hi | lo
//   kernel map   the ordinary pair stream (kmap.cu) built with tile_rows = TM / 4 and pad = 1
//
// One CTA owns TM consecutive output rows = 4 QUARTERS of TM/4 rows; fp32 accumulators of the tile live in shared memory.
//   chunk (k, j)   for each quarter q the pairs [seg[q][k] + 32 j, + 32) of offset k -> MMA rows 32 q .. 32 q + 31
//                  (M = 128 = 4 quarters x 32 lanes, all of ONE offset k, so one W[k] serves the whole MMA)
//   gather warps   8 warps = 2 sets (even / odd chunks) x 4 quarters.  Eight lanes copy one 128 B row (cp.async, coalesced)
//                  into the warp's staging ring; NBUF - 1 own chunks later lane l reads row l back and tcgen05.st's its
//                  hi / lo halves into ITS lane of tensor memory (16 + 16 columns of an operand stage); the set also copies
//                  W[k]'s 4 KB image into the stage's shared-memory slot and arrives on full[stage]
//   MMA warp       one elected lane: 6 x tcgen05.mma  D[128 x 32] (+)= A[tmem] . W[k][smem]   (lo.Whi, hi.Wlo, hi.Whi; K = 2 x 16),
//                  tcgen05.commit -> empty[stage] (operands may be overwritten) and dfull[buffer] (accumulator complete)
//   epilogue warps 8 warps = 4 quarters x 2 channel halves: tcgen05.ld of the warp's 32 lanes x 16 columns, then every lane
//                  adds its pair's 16 channels into the accumulator row of that pair
// Quarter q's accumulator rows are only touched by the two epilogue warps of q (disjoint channels), and chunks arrive in offset
// order, so every output row has ONE summation order: encoder and decoder CDFs stay bit-identical, with no CTA barrier in the loop.
// Why the operand goes through tensor memory, and the other measurements behind this shape: profiles/r01_conv_tcgen05.md.
#include "umma.cuh"

constexpr int TC_NB = 4;              // TMEM accumulator buffers (32 columns each): columns [0, 128)
constexpr int TC_RID = 16;            // row-id ring (chunks): gather -> epilogue; >= S + TC_NB
constexpr int TC_THREADS = 17 * 32;   // warps 0-7 epilogue, 8-15 gather, 16 MMA issue (+ TMEM allocation)

template <int TM, int S, int NBUF, int ENT>
struct TcSmem {
    float acc[TM][GPC_C];                              // 16-byte chunk j of row r lives at chunk j ^ (r & 7): conflict-free across consecutive rows
    __align__(128) unsigned char w[S][4096];           // W[k] hi 2 KB | lo 2 KB of the stage's chunk
    __align__(128) unsigned char stg[8][NBUF][4096];   // per gather warp: 32 gathered rows, 16-byte chunk j of row r at j ^ (r & 7)
    u64 ent[8][ENT][32];                               // per gather warp: pair entries (nbr | row << 32 | k << 48) of its next chunks
    u16 rid[TC_RID][128];                              // accumulator row of each MMA row of a chunk (0xFFFF: idle lane)
    u32 seg[4][GPC_K3 + 3];
    u32 cstart[GPC_K3 + 3];
    u16 tab[GPC_K3 * ((TM + 127) / 128) + 4];          // chunk -> k | j << 8
    __align__(8) u64 full[S];
    u64 empty[S];
    u64 dfull[TC_NB];
    u64 dempty[TC_NB];
    u32 tmem_base;
};

// role profile (PROF instantiations only; tools/conv_ab.py): cycles per role summed over chunks, one lane per role
//  [0] gather: table loads + wait empty + issue copies   [1] gather: wait_group   [2] gather: staging -> TMEM + arrive
//  [3] mma: wait full       [4] mma: wait dempty       [5] mma: issue + commit
//  [6] epi: wait dfull      [7] epi: tcgen05.ld        [8] epi: scatter-add     [9] chunks   [10] CTA total   [11] setup  [12] write-out
__device__ unsigned long long g_tc_prof[16];
extern "C" int gpc_debug_conv_tc_profile(unsigned long long *out_h, int reset) {
    GPC_CUDA_CHECK(cudaDeviceSynchronize());
    if (out_h) GPC_CUDA_CHECK(cudaMemcpyFromSymbol(out_h, g_tc_prof, sizeof(unsigned long long) * 16));
    if (reset) { unsigned long long z[16] = {0}; GPC_CUDA_CHECK(cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z))); }
    return GPC_OK;
}
#define TC_T(var) do { if (PROF) var = clock64(); } while (0)
#define TC_ACCUM(i, a, b) do { if (PROF && lane == 0) acc_t[i] += (b) - (a); } while (0)

template <int TM, int S, int NBUF, int ENT, bool PROF>
__global__ void __launch_bounds__(TC_THREADS, 1)
spconv_tc_kernel(const unsigned char *__restrict__ xs, const unsigned char *__restrict__ Wc, const u32 *__restrict__ seg_g,
                 const u64 *__restrict__ pairs, i64 n, const void *__restrict__ residual, int flags,
                 float *__restrict__ y, u32 *__restrict__ ys) {
    constexpr int QR = TM / 4;                   // rows per quarter == tile_rows of the pair stream
    constexpr int P = 2 * NBUF - 1;              // a gather warp fetches pair entries P of its own chunks ahead (group accounting below)
    static_assert(P < ENT && S + TC_NB <= TC_RID && 2 * NBUF <= S && (S % 2) == 0, "ring depths");
    static_assert(128 + 32 * S <= 512, "TMEM columns");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem<TM, S, NBUF, ENT> &s = *reinterpret_cast<TcSmem<TM, S, NBUF, ENT> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = blockIdx.x;
    const i64 n_sub = (n + QR - 1) / QR;
    long long acc_t[3] = {0, 0, 0}, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t_begin = 0, t_setup = 0;
    TC_T(t_begin);

    // ---- setup: segment table, chunk table, barriers, TMEM
    for (int i = tid; i < 4 * (GPC_K3 + 1); i += TC_THREADS) {
        const int q = i / (GPC_K3 + 1), k = i - q * (GPC_K3 + 1);
        const i64 st = t * 4 + q;
        s.seg[q][k] = st < n_sub ? seg_g[st * (GPC_K3 + 1) + k] : 0u;
    }
    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init((u32)__cvta_generic_to_shared(&s.full[i]), 128);
            mbar_init((u32)__cvta_generic_to_shared(&s.empty[i]), 1);
        }
        for (int i = 0; i < TC_NB; ++i) {
            mbar_init((u32)__cvta_generic_to_shared(&s.dfull[i]), 1);
            mbar_init((u32)__cvta_generic_to_shared(&s.dempty[i]), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        const u32 dst = (u32)__cvta_generic_to_shared(&s.tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
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
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    if (tid < GPC_K3) {
        const u32 b = s.cstart[tid], e = s.cstart[tid + 1];
        for (u32 c = b; c < e; ++c) s.tab[c] = (u16)(tid | ((c - b) << 8));
    }
    __syncthreads();
    const u32 n_chunks = s.cstart[GPC_K3];
    const u32 tmem_d = s.tmem_base;
    const u32 w0 = (u32)__cvta_generic_to_shared(&s.w[0][0]);
    const u32 full0 = (u32)__cvta_generic_to_shared(&s.full[0]), empty0 = (u32)__cvta_generic_to_shared(&s.empty[0]);
    const u32 dfull0 = (u32)__cvta_generic_to_shared(&s.dfull[0]), dempty0 = (u32)__cvta_generic_to_shared(&s.dempty[0]);
    TC_T(t_setup);

    if (warp >= 8 && warp < 16) {
        // =================================================================== gather warps
        const int gw = warp - 8, q = gw & 3;
        const u32 set = (u32)gw >> 2;                            // this warp's chunks: c = set, set + 2, set + 4, ...
        const int pt = q * 32 + lane;                            // MMA row of this lane == its 1/128 share of W[k]
        const int g8 = lane >> 3, j8 = lane & 7;                 // copy role: row 4 i + g8 of the chunk, 16-byte piece j8
        const u32 ent0 = (u32)__cvta_generic_to_shared(&s.ent[gw][0][lane]);
        const u32 stg0 = (u32)__cvta_generic_to_shared(&s.stg[gw][0][0]);
        const u32 lane_base = tmem_d + ((u32)(q * 32) << 16) + 128u;       // this warp's TMEM lanes, first operand column
        // Lookaheads are counted in OWN chunks (m = c >> 1).  Every loop iteration commits ONE cp.async group = {entry of own chunk
        // m + P, W and rows of own chunk m + NBUF - 1}; cp.async.wait_group<NBUF-1> in iteration m therefore covers W and rows of
        // own chunk m (committed NBUF - 1 iterations ago) and the entry of own chunk m + NBUF, which iteration m + 1 dereferences.
        auto fetch_entry = [&](u32 c) {                          // this lane's pair entry of chunk c -> entry ring
            const u32 kj = s.tab[c], k = kj & 0xFFu;
            const u32 idx = s.seg[q][k] + 32u * (kj >> 8) + (u32)lane;
            if (idx < s.seg[q][k + 1]) cp_async8(ent0 + ((c >> 1) % ENT) * 256, pairs + idx);
        };
        auto issue_chunk = [&](u32 c) {                          // W[k] and the gathered rows of chunk c -> stage slot / staging ring
            const u32 kj = s.tab[c], k = kj & 0xFFu;
            const u32 beg = s.seg[q][k] + 32u * (kj >> 8), end = s.seg[q][k + 1];
            const u32 cnt = min(32u, end - min(end, beg));
            const u64 *ent = &s.ent[gw][(c >> 1) % ENT][0];
            u64 e8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e8[i] = ent[4 * i + g8];
            const u64 e_own = ent[lane];
            const u32 si = c % (u32)S;
            if (c >= (u32)S) mbar_wait(empty0 + si * 8, ((c / (u32)S) & 1u) ^ 1u);      // MMAs of chunk c - S are done with the stage
            const unsigned char *wsrc = Wc + (size_t)k * 4096 + pt * 16;
            cp_async16(w0 + si * 4096 + pt * 16, wsrc);
            cp_async16(w0 + si * 4096 + 2048 + pt * 16, wsrc + 2048);
            s.rid[c % TC_RID][pt] = (u32)lane < cnt ? (u16)((u32)(e_own >> 32) & 0xFFFFu) : (u16)0xFFFFu;
            const u32 dst = stg0 + ((c >> 1) % (u32)NBUF) * 4096;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u32 r = 4u * i + g8;
                if (r < cnt) cp_async16(dst + r * 128 + ((j8 ^ (r & 7)) << 4), xs + (size_t)(u32)e8[i] * 128 + j8 * 16);
            }
        };
        for (u32 m = 0; m < (u32)P; ++m) { const u32 c = set + 2 * m; if (c < n_chunks) fetch_entry(c); }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int m = 0; m < NBUF - 1; ++m) {
            const u32 c = set + 2 * m;
            if (c < n_chunks) issue_chunk(c);
            cp_async_commit();
        }
        for (u32 c = set; c < n_chunks; c += 2) {
            TC_T(t0);
            __syncwarp();                                        // the staging slot re-filled below was read by other lanes last iteration
            if (c + 2 * P < n_chunks) fetch_entry(c + 2 * P);
            if (c + 2 * (NBUF - 1) < n_chunks) issue_chunk(c + 2 * (NBUF - 1));
            cp_async_commit();
            TC_T(t1);
            cp_async_wait<NBUF - 1>();
            __syncwarp();                                        // rows / entries copied by the other lanes of the warp are visible
            TC_T(t2);
            // ---- publish chunk c: staging row `lane` -> registers -> this thread's TMEM lane (idle lanes store stale bytes: their D rows are ignored)
            const u32 src = stg0 + ((c >> 1) % (u32)NBUF) * 4096 + (u32)lane * 128;
            uint4 v[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[jj].x), "=r"(v[jj].y), "=r"(v[jj].z), "=r"(v[jj].w)
                             : "r"(src + (u32)((jj ^ (lane & 7)) << 4)));
            const u32 si = c % (u32)S;
            const u32 ta = lane_base + si * 32;
            tmem_st16(ta, v[0], v[1], v[2], v[3]);
            tmem_st16(ta + 16, v[4], v[5], v[6], v[7]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_async_smem();                                  // W[k] (generic-proxy cp.async writes) -> visible to the tensor core
            tmem_fence_before();
            mbar_arrive(full0 + si * 8);
            TC_T(t3);
            TC_ACCUM(0, t0, t1); TC_ACCUM(1, t1, t2); TC_ACCUM(2, t2, t3);
        }
        cp_async_wait<0>();
        if (PROF && lane == 0 && warp == 8) {
            atomicAdd(&g_tc_prof[0], (unsigned long long)acc_t[0]); atomicAdd(&g_tc_prof[1], (unsigned long long)acc_t[1]);
            atomicAdd(&g_tc_prof[2], (unsigned long long)acc_t[2]);
        }
    } else if (warp == 16) {
        // =================================================================== MMA issue (whole warp loops, one elected lane issues)
        u32 st_i = 0, st_ph = 0;
        for (u32 c = 0; c < n_chunks; ++c) {
            const u32 b = c & (TC_NB - 1), v = c / TC_NB;
            TC_T(t0);
            mbar_wait(full0 + st_i * 8, st_ph);
            TC_T(t1);
            if (v > 0) mbar_wait(dempty0 + b * 8, (v & 1u) ^ 1u);
            TC_T(t2);
            tmem_fence_after();
            const u32 a_hi = tmem_d + 128u + st_i * 32, a_lo = a_hi + 16;
            const u32 b_hi = w0 + st_i * 4096, b_lo = b_hi + 2048;
            const u32 d = tmem_d + b * 32;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16_ts(d, a_lo + ks * 8, umma_smem_desc(b_hi + ks * 256), UMMA_IDESC_BF16_M128_N32, ks);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16_ts(d, a_hi + ks * 8, umma_smem_desc(b_lo + ks * 256), UMMA_IDESC_BF16_M128_N32, 1u);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) umma_bf16_ts(d, a_hi + ks * 8, umma_smem_desc(b_hi + ks * 256), UMMA_IDESC_BF16_M128_N32, 1u);
                umma_commit(empty0 + st_i * 8);
                umma_commit(dfull0 + b * 8);
            }
            __syncwarp();
            TC_T(t3);
            TC_ACCUM(0, t0, t1); TC_ACCUM(1, t1, t2); TC_ACCUM(2, t2, t3);
            if (++st_i == (u32)S) { st_i = 0; st_ph ^= 1u; }
        }
        if (PROF && lane == 0) {
            atomicAdd(&g_tc_prof[3], (unsigned long long)acc_t[0]); atomicAdd(&g_tc_prof[4], (unsigned long long)acc_t[1]);
            atomicAdd(&g_tc_prof[5], (unsigned long long)acc_t[2]); atomicAdd(&g_tc_prof[9], (unsigned long long)n_chunks);
        }
    } else {
        // =================================================================== epilogue warps: quarter q, channels 16 h .. 16 h + 15
        const int q = warp & 3, h = warp >> 2;
        float4 *accq = reinterpret_cast<float4 *>(&s.acc[q * QR][0]);
        for (int i = h * 32 + lane; i < QR * 8; i += 64) accq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");             // the two warps of the quarter
        for (u32 c = 0; c < n_chunks; ++c) {
            const u32 b = c & (TC_NB - 1), v = c / TC_NB;
            TC_T(t0);
            mbar_wait(dfull0 + b * 8, v & 1u);
            TC_T(t1);
            tmem_fence_after();
            const u32 r0 = s.rid[c % TC_RID][q * 32 + lane];
            const bool any = __any_sync(0xFFFFFFFFu, r0 != 0xFFFFu);
            u32 d[16];
            if (any) {
                tmem_ld16(tmem_d + ((u32)(q * 32) << 16) + b * 32 + h * 16, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            tmem_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dempty0 + b * 8);         // the values are in registers: hand the buffer back
            TC_T(t2);
            if (r0 != 0xFFFFu) {
                float4 *a = accq + r0 * 8;
                const u32 sw = r0 & 7u;
                // one 16-byte read-modify-write at a time: batching the four loads ahead of the stores was measured 8 % SLOWER here
                // (the opposite of the mma.sync kernel): the interleaved form spreads the conflicting wavefronts out
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u32 jj = (u32)(4 * h + j) ^ sw;
                    float4 w = a[jj];
                    w.x += __uint_as_float(d[4 * j]); w.y += __uint_as_float(d[4 * j + 1]);
                    w.z += __uint_as_float(d[4 * j + 2]); w.w += __uint_as_float(d[4 * j + 3]);
                    a[jj] = w;
                }
            }
            TC_T(t3);
            TC_ACCUM(0, t0, t1); TC_ACCUM(1, t1, t2); TC_ACCUM(2, t2, t3);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");             // both channel halves of the quarter are complete
        TC_T(t0);
        if (PROF && lane == 0 && warp == 0) {
            atomicAdd(&g_tc_prof[6], (unsigned long long)acc_t[0]); atomicAdd(&g_tc_prof[7], (unsigned long long)acc_t[1]);
            atomicAdd(&g_tc_prof[8], (unsigned long long)acc_t[2]);
        }
        // ---- write-out of the quarter (rows r = h, h + 2, ...): (+ residual) (ReLU) -> fp32 rows and / or split rows, one 128 B store per row
        const bool relu = (flags & GPC_CONV_RELU) != 0, res_split = (flags & GPC_CONV_RES_SPLIT) != 0;
        const i64 g0 = t * TM + (i64)q * QR;
        const int rows = (int)max((i64)0, min((i64)QR, n - g0));
        const int cp = lane & 15;
        for (int r = h; r < rows; r += 2) {
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
        if (PROF && lane == 0 && warp == 0) atomicAdd(&g_tc_prof[12], (unsigned long long)(clock64() - t0));
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 16) {
        tmem_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_d) : "memory");
    }
    if (PROF && tid == 0) {
        atomicAdd(&g_tc_prof[10], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_tc_prof[11], (unsigned long long)(t_setup - t_begin));
    }
}

template <int TM, int S, int NBUF, int ENT, bool PROF = false>
static int launch_spconv_tc(const void *xs, const void *Wc, const u32 *seg, const u64 *pairs, i64 n, const void *residual,
                            int flags, float *y, void *ys, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(TcSmem<TM, S, NBUF, ENT>) + 128;
    static_assert(sizeof(TcSmem<TM, S, NBUF, ENT>) + 128 <= 232448, "shared memory per CTA");
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel<TM, S, NBUF, ENT, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_tc_kernel<TM, S, NBUF, ENT, PROF><<<(unsigned)tiles, TC_THREADS, smem, st>>>(
        (const unsigned char *)xs, (const unsigned char *)Wc, seg, pairs, n, residual, flags, y, (u32 *)ys);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// xs = split rows; pair stream with pad = 1 and tile_rows = cta_rows / 4; Wc from gpc_spconv_pack_weights_umma;
// y (fp32 rows) and / or ys (split rows) output.  profile != 0: the instrumented instantiation (gpc_debug_conv_tc_profile).
extern "C" int gpc_spconv_fwd_tc(const void *xs, const void *Wc, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int cta_rows, const void *residual, int flags, float *y, void *ys, int profile, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(y || ys, GPC_EINVAL, "no output requested");
    GPC_REQUIRE(xs != ys && xs != (const void *)y, GPC_EINVAL, "conv is out of place (rows are gathered from xs while y is written)");
    cudaStream_t st = as_stream(stream);
    if (!profile) {
        if (cta_rows == 640) return launch_spconv_tc<640, 6, 3, 8>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 768) return launch_spconv_tc<768, 4, 2, 4>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 512) return launch_spconv_tc<512, 8, 3, 8>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_tc<1024, 4, 2, 4>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    } else {
        if (cta_rows == 512) return launch_spconv_tc<512, 8, 3, 8, true>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
        if (cta_rows == 1024) return launch_spconv_tc<1024, 4, 2, 4, true>(xs, Wc, seg, pairs, n, residual, flags, y, ys, st);
    }
    gpc_set_error("unsupported tcgen05 conv cta_rows %d (512, 640, 768, 1024)", cta_rows);
    return GPC_EINVAL;
}
