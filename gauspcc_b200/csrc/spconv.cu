// spconv.cu -- submanifold sparse convolution forward, K=5, C=32 -> 32, bias-less (a-7, a-10, a-12).
//
// Replaces torchsparse's implicit-GEMM conv behind spnn.Conv3d(32,32,5)
// (src/ai_pcc/GausPcgc/kit/nn.py:14-16, network_ue_4stage_conv.py:18-61):
//     y[o,:] = act( sum_{k : nbr_k(o) exists} x[nbr_k(o),:] . W[k]  (+ residual[o,:]) )
//
// Output-stationary: one CTA owns `TM` consecutive output rows with fp32 accumulators in shared
// memory and walks the populated offsets k in ascending order (fixed accumulation order => the
// encoder and the decoder produce bit-identical features).  For each k the CTA's pair list
// (kmap.cu) gives (output row, input row); W[k] is staged through a double-buffered 4 KB shared
// tile, the input rows are gathered with coalesced 128 B loads.
//
// v1 contraction: fp32 FFMA, lane = output channel, input row broadcast from shared memory.
#include "common.cuh"

constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_P = 4;                       // pairs in flight per warp step

template <int TM>
struct ScSmem {
    float acc[TM][GPC_C];
    float w[2][GPC_C][GPC_C];
    float xs[SC_WARPS][SC_P][GPC_C];
    u32 seg[GPC_K3 + 1];
    int klist[GPC_K3];
    int nk;
};

template <int TM>
__global__ void __launch_bounds__(SC_THREADS) spconv_fwd_kernel(const float *__restrict__ x, const float *__restrict__ W,
                                                                 const u32 *__restrict__ seg, const u32 *__restrict__ pair_nbr,
                                                                 const u16 *__restrict__ pair_row, i64 n,
                                                                 const float *__restrict__ residual, int flags,
                                                                 float *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScSmem<TM> &s = *reinterpret_cast<ScSmem<TM> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = blockIdx.x;
    const i64 r0 = t * TM;
    const int rows = (int)min((i64)TM, n - r0);

    for (int i = tid; i <= GPC_K3; i += SC_THREADS) s.seg[i] = seg[t * (GPC_K3 + 1) + i];
    for (int i = tid; i < TM * GPC_C; i += SC_THREADS) (&s.acc[0][0])[i] = 0.f;
    __syncthreads();
    if (tid == 0) {
        int c = 0;
        for (int k = 0; k < GPC_K3; ++k) if (s.seg[k + 1] != s.seg[k]) s.klist[c++] = k;
        s.nk = c;
    }
    __syncthreads();
    const int nk = s.nk;

    float4 wnext = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nk > 0) wnext = __ldg(reinterpret_cast<const float4 *>(W + (i64)s.klist[0] * (GPC_C * GPC_C)) + tid);

    for (int it = 0; it < nk; ++it) {
        const int buf = it & 1;
        const int k = s.klist[it];
        reinterpret_cast<float4 *>(&s.w[buf][0][0])[tid] = wnext;
        __syncthreads();                       // W[k] visible; all adds of the previous offset are done
        if (it + 1 < nk) wnext = __ldg(reinterpret_cast<const float4 *>(W + (i64)s.klist[it + 1] * (GPC_C * GPC_C)) + tid);

        float w[GPC_C];
#pragma unroll
        for (int ci = 0; ci < GPC_C; ++ci) w[ci] = s.w[buf][ci][lane];

        const u32 seg_b = s.seg[k], seg_e = s.seg[k + 1];
        for (u32 p = seg_b + warp * SC_P; p < seg_e; p += SC_WARPS * SC_P) {
            const int np = (int)min((u32)SC_P, seg_e - p);
            u32 my_nbr = 0, my_row = 0;
            if (lane < np) { my_nbr = pair_nbr[p + lane]; my_row = pair_row[p + lane]; }
#pragma unroll
            for (int j = 0; j < SC_P; ++j) {
                const u32 nb = __shfl_sync(0xFFFFFFFFu, my_nbr, j);
                if (j < np) s.xs[warp][j][lane] = __ldg(x + (i64)nb * GPC_C + lane);
            }
            __syncwarp();
            float a[SC_P];
#pragma unroll
            for (int j = 0; j < SC_P; ++j) a[j] = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < GPC_C / 4; ++c4) {
#pragma unroll
                for (int j = 0; j < SC_P; ++j) {
                    const float4 xv = reinterpret_cast<const float4 *>(&s.xs[warp][j][0])[c4];
                    a[j] = fmaf(xv.x, w[4 * c4 + 0], a[j]);
                    a[j] = fmaf(xv.y, w[4 * c4 + 1], a[j]);
                    a[j] = fmaf(xv.z, w[4 * c4 + 2], a[j]);
                    a[j] = fmaf(xv.w, w[4 * c4 + 3], a[j]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < SC_P; ++j) {
                const u32 r = __shfl_sync(0xFFFFFFFFu, my_row, j);
                if (j < np) s.acc[r][lane] += a[j];     // row r appears once per offset: no other warp touches it now
            }
        }
    }
    __syncthreads();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = warp; r < rows; r += SC_WARPS) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TM>
static int launch_spconv(const float *x, const float *W, const u32 *seg, const u32 *pair_nbr, const u16 *pair_row, i64 n,
                         const float *residual, int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(ScSmem<TM>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_fwd_kernel<TM><<<(unsigned)tiles, SC_THREADS, smem, st>>>(x, W, seg, pair_nbr, pair_row, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

static int spconv_fwd_v1(const float *x, const float *W, const uint32_t *seg, const uint32_t *pair_nbr,
                              const uint16_t *pair_row, int64_t n, int tile_rows, const float *residual, int flags,
                              float *y, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    switch (tile_rows) {
        case 128: return launch_spconv<128>(x, W, seg, pair_nbr, pair_row, n, residual, flags, y, st);
        case 256: return launch_spconv<256>(x, W, seg, pair_nbr, pair_row, n, residual, flags, y, st);
        case 512: return launch_spconv<512>(x, W, seg, pair_nbr, pair_row, n, residual, flags, y, st);
        default: gpc_set_error("unsupported tile_rows %d (128, 256, 512)", tile_rows); return GPC_EINVAL;
    }
}


// =====================================================================================================
// v2: software-pipelined gather (cp.async ring), packed-pair FFMA2 contraction.
//
// The tile's pair list is walked as a stream of chunks (<= CH pairs of ONE offset k).  Chunk c+D-1 is
// gathered with 16-byte cp.async copies into a D-stage shared ring while chunk c is contracted, so the
// L2/HBM latency of the row gathers is hidden behind math instead of being paid once per offset.  W[k]
// rides the same pipeline through its own (D+1)-slot ring.  Pair indices are prefetched into registers
// one chunk further ahead.  Contraction: lane = output channel, two input channels per FFMA2
// (fma.rn.f32x2): weights are pre-packed as float2 (W[k][2i][co], W[k][2i+1][co]).
// One __syncthreads per chunk orders the accumulator updates (a row appears once per offset), so the
// accumulation order is a fixed function of the kernel map: encoder and decoder agree bit for bit.
// =====================================================================================================
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void pack_weights_kernel(const float *__restrict__ W, float2 *__restrict__ Wp, int n_kernels) {
    // W [n_kernels*125][32 ci][32 co] -> Wp [n_kernels*125][16][32 co] of (W[2i][co], W[2i+1][co])
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (i64)n_kernels * GPC_K3 * 16 * GPC_C) return;
    const int co = (int)(g & 31), i = (int)((g >> 5) & 15);
    const i64 k = g >> 9;
    const float *src = W + k * (GPC_C * GPC_C);
    Wp[g] = make_float2(src[(2 * i) * GPC_C + co], src[(2 * i + 1) * GPC_C + co]);
}
extern "C" int gpc_spconv_pack_weights(const float *W, int n_kernels, float *Wp, void *stream) {
    const i64 total = (i64)n_kernels * GPC_K3 * 16 * GPC_C;
    pack_weights_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(W, (float2 *)Wp, n_kernels);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

template <int TM, int NW, int CH, int D>
struct Sc2Smem {
    float acc[TM][GPC_C];
    float xs[D][CH][GPC_C];
    float2 ws[D + 1][16][GPC_C];
    u32 rows[D][CH];
    u32 seg[GPC_K3 + 1];
};

template <int CH>
struct ChunkIt {
    int k, segord;
    u32 p, end;
    __device__ __forceinline__ void init(const u32 *seg) { k = 0; segord = -1; p = seg[0]; end = seg[GPC_K3]; }
    // next chunk: [cp, cp+cnt) of offset ck; cord = ordinal of its (non-empty) segment; first = starts the segment
    __device__ __forceinline__ bool next(const u32 *seg, u32 &cp, int &cnt, int &ck, int &cord, bool &first) {
        if (p >= end) return false;
        while (p >= seg[k + 1]) ++k;
        first = (p == seg[k]);
        if (first) ++segord;
        cnt = (int)min((u32)CH, seg[k + 1] - p);
        cp = p; ck = k; cord = segord;
        p += cnt;
        return true;
    }
};

template <int TM, int NW, int CH, int D>
__global__ void __launch_bounds__(NW * 32) spconv_fwd_v2_kernel(const float *__restrict__ x, const float2 *__restrict__ Wp,
                                                                const u32 *__restrict__ seg_g, const u32 *__restrict__ pair_nbr,
                                                                const u16 *__restrict__ pair_row, i64 n,
                                                                const float *__restrict__ residual, int flags,
                                                                float *__restrict__ y) {
    constexpr int THREADS = NW * 32;
    constexpr int PT = CH * 8 / THREADS;            // 16-byte pieces per thread per chunk
    constexpr int WPT = 256 / THREADS;              // 16-byte pieces of W[k] (4 KB) per thread
    constexpr int WD = D + 1;
    constexpr int P = 4;
    static_assert(PT >= 1 && CH * 8 % THREADS == 0 && 256 % THREADS == 0, "bad config");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sc2Smem<TM, NW, CH, D> &s = *reinterpret_cast<Sc2Smem<TM, NW, CH, D> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = blockIdx.x;
    const i64 r0 = t * TM;
    const int rows = (int)min((i64)TM, n - r0);

    for (int i = tid; i <= GPC_K3; i += THREADS) s.seg[i] = seg_g[t * (GPC_K3 + 1) + i];
    for (int i = tid; i < TM * GPC_C / 4; i += THREADS) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    ChunkIt<CH> it_idx, it_iss, it_cmp;
    it_idx.init(s.seg); it_iss.init(s.seg); it_cmp.init(s.seg);

    // registers holding the indices of the next chunk to be issued
    u32 nb[PT];
    u32 rowreg = 0;
    auto idx_load = [&]() {
        u32 cp; int cnt, ck, cord; bool first;
        if (!it_idx.next(s.seg, cp, cnt, ck, cord, first)) return;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const int slot = (tid + i * THREADS) >> 3;
            nb[i] = slot < cnt ? __ldg(pair_nbr + cp + slot) : 0u;
        }
        if (tid < CH) rowreg = tid < cnt ? (u32)__ldg(pair_row + cp + tid) : 0u;
    };
    int iss_count = 0;
    auto issue = [&]() {
        u32 cp; int cnt, ck, cord; bool first;
        if (it_iss.next(s.seg, cp, cnt, ck, cord, first)) {
            const int stage = iss_count % D;
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                const int q = tid + i * THREADS;
                const int slot = q >> 3, piece = q & 7;
                if (slot < cnt) cp_async16(&s.xs[stage][slot][piece * 4], x + (i64)nb[i] * GPC_C + piece * 4);
            }
            if (tid < CH) s.rows[stage][tid] = rowreg;
            if (first) {
                const float2 *src = Wp + (i64)ck * (16 * GPC_C);
                float2 *dst = &s.ws[cord % WD][0][0];
#pragma unroll
                for (int i = 0; i < WPT; ++i) {
                    const int q = tid + i * THREADS;
                    cp_async16(dst + q * 2, src + q * 2);
                }
            }
            idx_load();
        }
        ++iss_count;
        cp_async_commit();
    };

    idx_load();
#pragma unroll
    for (int st = 0; st < D - 1; ++st) issue();

    u64 w2[16];
    int cur_ord = -1;
    for (int c = 0;; ++c) {
        cp_async_wait<D - 2>();
        __syncthreads();
        issue();
        u32 cp; int cnt, ck, cord; bool first;
        if (!it_cmp.next(s.seg, cp, cnt, ck, cord, first)) break;
        const int stage = c % D;
        if (cord != cur_ord) {
            cur_ord = cord;
            const u64 *wsrc = reinterpret_cast<const u64 *>(&s.ws[cord % WD][0][0]);
#pragma unroll
            for (int i = 0; i < 16; ++i) w2[i] = wsrc[i * GPC_C + lane];
        }
        for (int j0 = warp * P; j0 < cnt; j0 += NW * P) {
            u64 a2[P];
#pragma unroll
            for (int j = 0; j < P; ++j) a2[j] = 0ull;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    // rows past cnt hold stale data of an earlier chunk: computed and discarded
                    const ulonglong2 xv = *reinterpret_cast<const ulonglong2 *>(&s.xs[stage][(j0 + j) % CH][q * 4]);
                    a2[j] = ffma2(xv.x, w2[2 * q], a2[j]);
                    a2[j] = ffma2(xv.y, w2[2 * q + 1], a2[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < P; ++j) {
                if (j0 + j < cnt) {
                    const float2 f = *reinterpret_cast<const float2 *>(&a2[j]);
                    const u32 r = s.rows[stage][j0 + j];
                    s.acc[r][lane] += f.x + f.y;
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = warp; r < rows; r += NW) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TM, int NW, int CH, int D>
static int launch_spconv_v2(const float *x, const float *Wp, const u32 *seg, const u32 *pair_nbr, const u16 *pair_row, i64 n,
                            const float *residual, int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc2Smem<TM, NW, CH, D>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v2_kernel<TM, NW, CH, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_fwd_v2_kernel<TM, NW, CH, D><<<(unsigned)tiles, NW * 32, smem, st>>>(x, (const float2 *)Wp, seg, pair_nbr, pair_row, n,
                                                                                 residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}


// =====================================================================================================
// v3: fixed-size gather blocks, warp-owned accumulator rows, W through L1.
//
// The tile's pair stream (u64 entries: nbr | row<<32 | k<<48, sorted by (k,row)) is cut into blocks of G
// pairs REGARDLESS of offset boundaries.  Block c+D-1's rows are gathered by cp.async while block c is
// contracted; the pair entries themselves ride a deeper ring (DM >= 2D-1 blocks ahead), so no global load
// sits on the loop's critical path.  Output row r of the tile is owned by warp r % NW: a warp contracts
// exactly the pairs of its rows, in stream order, so accumulation needs no block barrier per offset and
// its order is a fixed function of the kernel map (encoder == decoder, bit for bit).  One barrier per
// G pairs (data arrival + ring reuse).  W[k] (4 KB, packed float2) is read straight into registers
// through L1: the NW warps of a CTA need the same W[k] at the same time, so it is one L2 fetch per
// (CTA, offset) like a shared-memory stage, without a ring to manage.
// =====================================================================================================
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
}

template <int TM, int NW, int G, int D, int DM>
struct Sc3Smem {
    float acc[TM][GPC_C];
    float xs[D][G][GPC_C];
    u64 meta[DM][G];
};

template <int TM, int NW, int G, int D, int DM>
__global__ void __launch_bounds__(NW * 32) spconv_fwd_v3_kernel(const float *__restrict__ x, const float2 *__restrict__ Wp,
                                                                const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs,
                                                                i64 n, const float *__restrict__ residual, int flags,
                                                                float *__restrict__ y) {
    constexpr int THREADS = NW * 32;
    constexpr int PT = G * 8 / THREADS;
    constexpr u64 INVALID = 0xFFFFFFFFFFFFFFFFull;
    static_assert(G % 32 == 0 && G * 8 % THREADS == 0 && DM >= 2 * D - 1 && (NW & (NW - 1)) == 0 && THREADS >= G, "bad config");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sc3Smem<TM, NW, G, D, DM> &s = *reinterpret_cast<Sc3Smem<TM, NW, G, D, DM> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 t = blockIdx.x;
    const i64 r0 = t * TM;
    const int rows = (int)min((i64)TM, n - r0);
    const u32 p_begin = seg_g[t * (GPC_K3 + 1)], p_end = seg_g[t * (GPC_K3 + 1) + GPC_K3];
    const int nblk = (int)((p_end - p_begin + G - 1) / G);

    for (int i = tid; i < TM * GPC_C / 4; i += THREADS) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto issue_meta = [&](int b) {
        if (b < nblk && tid < G) {
            const u32 p = p_begin + (u32)b * G + tid;
            if (p < p_end) cp_async8(&s.meta[b % DM][tid], pairs + p);
            else s.meta[b % DM][tid] = INVALID;
        }
    };
    auto issue_rows = [&](int b) {
        if (b < nblk) {
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                const int q = tid + i * THREADS;
                const int slot = q >> 3, piece = q & 7;
                const u64 m = s.meta[b % DM][slot];
                if (m != INVALID) cp_async16(&s.xs[b % D][slot][piece * 4], x + (i64)(u32)m * GPC_C + piece * 4);
            }
        }
    };
    // prologue: entries of blocks 0..DM-2, then rows of blocks 0..D-2
    for (int b = 0; b < DM - 1; ++b) issue_meta(b);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int b = 0; b < D - 1; ++b) { issue_rows(b); cp_async_commit(); }

    u64 w2[16];
    u32 cur_k = 0xFFFFFFFFu;
    const u64 *Wg = reinterpret_cast<const u64 *>(Wp);
    for (int c = 0; c < nblk; ++c) {
        cp_async_wait<D - 2>();
        __syncthreads();
        issue_meta(c + DM - 1);
        issue_rows(c + D - 1);
        cp_async_commit();

        const u64 *mb = s.meta[c % DM];
        const float(*xb)[GPC_C] = s.xs[c % D];
#pragma unroll 1
        for (int h = 0; h < G / 32; ++h) {
            const u64 m = mb[h * 32 + lane];
            const u32 row = (u32)(m >> 32) & 0xFFFFu, k = (u32)(m >> 48);
            const bool mine = (m != INVALID) && ((row & (NW - 1)) == (u32)warp);
            u32 bal = __ballot_sync(0xFFFFFFFFu, mine);
            while (bal) {
                // up to 4 of this warp's pairs that share one offset
                const int l0 = __ffs(bal) - 1;
                bal &= bal - 1;
                const u32 k0 = __shfl_sync(0xFFFFFFFFu, k, l0);
                int l1 = l0, l2 = l0, l3 = l0, np = 1;
                if (bal) { const int cnd = __ffs(bal) - 1; if (__shfl_sync(0xFFFFFFFFu, k, cnd) == k0) { l1 = cnd; bal &= bal - 1; np = 2; } }
                if (np == 2 && bal) { const int cnd = __ffs(bal) - 1; if (__shfl_sync(0xFFFFFFFFu, k, cnd) == k0) { l2 = cnd; bal &= bal - 1; np = 3; } }
                if (np == 3 && bal) { const int cnd = __ffs(bal) - 1; if (__shfl_sync(0xFFFFFFFFu, k, cnd) == k0) { l3 = cnd; bal &= bal - 1; np = 4; } }
                if (k0 != cur_k) {
                    cur_k = k0;
                    const u64 *wsrc = Wg + (size_t)k0 * (16 * GPC_C) + lane;
#pragma unroll
                    for (int i = 0; i < 16; ++i) w2[i] = __ldg(wsrc + i * GPC_C);
                }
                const float *x0 = xb[h * 32 + l0], *x1 = xb[h * 32 + l1], *x2 = xb[h * 32 + l2], *x3 = xb[h * 32 + l3];
                u64 a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const ulonglong2 v0 = reinterpret_cast<const ulonglong2 *>(x0)[q];
                    const ulonglong2 v1 = reinterpret_cast<const ulonglong2 *>(x1)[q];
                    const ulonglong2 v2 = reinterpret_cast<const ulonglong2 *>(x2)[q];
                    const ulonglong2 v3 = reinterpret_cast<const ulonglong2 *>(x3)[q];
                    a0 = ffma2(v0.x, w2[2 * q], a0); a1 = ffma2(v1.x, w2[2 * q], a1);
                    a2 = ffma2(v2.x, w2[2 * q], a2); a3 = ffma2(v3.x, w2[2 * q], a3);
                    a0 = ffma2(v0.y, w2[2 * q + 1], a0); a1 = ffma2(v1.y, w2[2 * q + 1], a1);
                    a2 = ffma2(v2.y, w2[2 * q + 1], a2); a3 = ffma2(v3.y, w2[2 * q + 1], a3);
                }
                const u32 rw0 = __shfl_sync(0xFFFFFFFFu, row, l0), rw1 = __shfl_sync(0xFFFFFFFFu, row, l1);
                const u32 rw2 = __shfl_sync(0xFFFFFFFFu, row, l2), rw3 = __shfl_sync(0xFFFFFFFFu, row, l3);
                { const float2 f = *reinterpret_cast<const float2 *>(&a0); s.acc[rw0][lane] += f.x + f.y; }
                if (np > 1) { const float2 f = *reinterpret_cast<const float2 *>(&a1); s.acc[rw1][lane] += f.x + f.y; }
                if (np > 2) { const float2 f = *reinterpret_cast<const float2 *>(&a2); s.acc[rw2][lane] += f.x + f.y; }
                if (np > 3) { const float2 f = *reinterpret_cast<const float2 *>(&a3); s.acc[rw3][lane] += f.x + f.y; }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = warp; r < rows; r += NW) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TM, int NW, int G, int D, int DM>
static int launch_spconv_v3(const float *x, const float *Wp, const u32 *seg, const u64 *pairs, i64 n, const float *residual,
                            int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc3Smem<TM, NW, G, D, DM>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v3_kernel<TM, NW, G, D, DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TM - 1) / TM;
    spconv_fwd_v3_kernel<TM, NW, G, D, DM><<<(unsigned)tiles, NW * 32, smem, st>>>(x, (const float2 *)Wp, seg, pairs, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

extern "C" int gpc_spconv_fwd_v3(const float *x, const float *Wp, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int tile_rows, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
#define GPC_V3(TM_, NW_, G_, D_, DM_) return launch_spconv_v3<TM_, NW_, G_, D_, DM_>(x, Wp, seg, pairs, n, residual, flags, y, st)
    if (variant == 10) {
        if (tile_rows == 64) GPC_V3(64, 8, 64, 4, 8);
        if (tile_rows == 128) GPC_V3(128, 8, 64, 4, 8);
        if (tile_rows == 256) GPC_V3(256, 8, 64, 4, 8);
        if (tile_rows == 512) GPC_V3(512, 8, 64, 4, 8);
    } else if (variant == 11) {
        if (tile_rows == 128) GPC_V3(128, 4, 64, 4, 8);
        if (tile_rows == 256) GPC_V3(256, 4, 64, 4, 8);
        if (tile_rows == 512) GPC_V3(512, 4, 64, 4, 8);
    } else if (variant == 12) {
        if (tile_rows == 128) GPC_V3(128, 8, 128, 3, 6);
        if (tile_rows == 256) GPC_V3(256, 8, 128, 3, 6);
        if (tile_rows == 512) GPC_V3(512, 8, 128, 3, 6);
    } else if (variant == 13) {
        if (tile_rows == 256) GPC_V3(256, 16, 128, 4, 8);
        if (tile_rows == 512) GPC_V3(512, 16, 128, 4, 8);
    }
#undef GPC_V3
    gpc_set_error("unsupported conv v3 variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}


// =====================================================================================================
// v4: warp-private sub-tiles, split-bf16 tensor-core contraction (mma.sync m16n8k16), fp32 accumulate.
//
// Why not FFMA: with lane = output channel every pair needs its 32 inputs delivered to all 32 lanes
// (4 KB of register-file traffic per pair = 32 clk on the 128 B/clk shared/L1 data path; ncu: l1tex 89 %,
// FMA pipe 37 %).  Why not one tcgen05 tile per offset: the pairs of one (row tile, offset) are few
// (5-90), M=128 tiles would be mostly padding and N=32 makes the MMA shared-memory bound; and a single
// TF32 pass misses the 1e-3 parity bound (measured 7.5e-3), see DESIGN.md.  mma.sync works on 16 pairs
// at a time, takes its operands from registers that are loaded ONCE per pair (A) / once per
// (sub-tile, offset) (B) and accumulates in fp32.
//
// fp32-faithful products from bf16 tensor cores: x = x1 + x2, w = w1 + w2 (bf16 each, 16 mantissa bits
// together); x.w ~= x1.w1 + x1.w2 + x2.w1 (dropped x2.w2 ~ 2^-16 relative; oracle-measured 1.5e-4 max
// on probabilities vs the 1e-3 bound).  3 terms x 2 k-steps x 4 n-tiles = 24 mma per 16 pairs.
//
// One warp (= one CTA) owns TW consecutive output rows: its own pair stream (u64 entries sorted by
// (offset,row)), its own fp32 accumulators in shared memory, its own cp.async ring -> no block
// barriers at all, and a fixed accumulation order (encoder == decoder bit for bit).
// =====================================================================================================
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const u32 (&a)[4], u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// W [n_kernels*125][32 ci][32 co] fp32 -> Wb [n_kernels*125][2 (hi,lo)][8 q][32 co] of 4 x bf16 (channels 4q..4q+3)
__global__ void pack_weights_bf16_kernel(const float *__restrict__ W, u64 *__restrict__ Wb, int n_kernels) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (i64)n_kernels * GPC_K3 * 8 * GPC_C) return;
    const int n = (int)(g & 31), q = (int)((g >> 5) & 7);
    const i64 k = g >> 8;
    const float *src = W + k * (GPC_C * GPC_C);
    u64 hi = 0, lo = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float w = src[(4 * q + c) * GPC_C + n];
        const float w1 = bf16_round(w);
        const float w2 = bf16_round(w - w1);
        hi |= (u64)(__float_as_uint(w1) >> 16) << (16 * c);
        lo |= (u64)(__float_as_uint(w2) >> 16) << (16 * c);
    }
    Wb[(k * 2 + 0) * 256 + q * 32 + n] = hi;
    Wb[(k * 2 + 1) * 256 + q * 32 + n] = lo;
}
extern "C" int gpc_spconv_pack_weights_bf16(const float *W, int n_kernels, void *Wb, void *stream) {
    const i64 total = (i64)n_kernels * GPC_K3 * 8 * GPC_C;
    pack_weights_bf16_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(W, (u64 *)Wb, n_kernels);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

constexpr int SC4_XS = 36;      // padded row stride (floats) of the gather ring: conflict-free A-fragment loads
constexpr int SC4_ACC = 40;     // padded row stride (floats) of the accumulators

template <int TW, int D, int MW>
struct Sc4Smem {
    float acc[TW][SC4_ACC];
    float xs[D][16][SC4_XS];
    u64 meta[MW];
    u32 seg[GPC_K3 + 1];
};

struct GroupIt {                 // walks one sub-tile's stream in groups of <= 16 pairs of one offset
    int k;
    u32 p, end;
    __device__ __forceinline__ void init(const u32 *seg) { k = 0; p = seg[0]; end = seg[GPC_K3]; }
    __device__ __forceinline__ bool next(const u32 *seg, u32 &gp, int &cnt, int &gk) {
        if (p >= end) return false;
        while (p >= seg[k + 1]) ++k;
        cnt = (int)min(16u, seg[k + 1] - p);
        gp = p; gk = k;
        p += cnt;
        return true;
    }
};

template <int TW, int D, int MW>
__global__ void __launch_bounds__(32) spconv_fwd_v4_kernel(const float *__restrict__ x, const u64 *__restrict__ Wb,
                                                           const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs, i64 n,
                                                           const float *__restrict__ residual, int flags, float *__restrict__ y) {
    static_assert((MW & (MW - 1)) == 0 && MW >= 32 * (2 * D + 2), "meta window too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sc4Smem<TW, D, MW> &s = *reinterpret_cast<Sc4Smem<TW, D, MW> *>(smem_raw);
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    const i64 st = blockIdx.x;
    const i64 r0 = st * TW;
    const int rows = (int)min((i64)TW, n - r0);

    for (int i = lane; i <= GPC_K3; i += 32) s.seg[i] = seg_g[st * (GPC_K3 + 1) + i];
    for (int i = lane; i < TW * SC4_ACC / 4; i += 32) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const u32 p_begin = s.seg[0];
    const u32 total = s.seg[GPC_K3] - p_begin;

    u32 fetched = 0;
    auto refill = [&](u32 dead_before) {         // entries [dead_before, fetched) are live; keep the window as full as possible
        while (fetched < total && fetched + 32 - dead_before <= (u32)MW) {
            const u32 idx = fetched + lane;
            if (idx < total) cp_async8(&s.meta[idx & (MW - 1)], pairs + p_begin + idx);
            fetched += 32;
        }
    };
    refill(0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    GroupIt it_iss, it_cmp;
    it_iss.init(s.seg);
    it_cmp.init(s.seg);
    int n_issued = 0;
    auto issue = [&]() {
        u32 gp; int cnt, gk;
        if (it_iss.next(s.seg, gp, cnt, gk)) {
            const u32 i0 = gp - p_begin;
            float(*dst)[SC4_XS] = s.xs[n_issued % D];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int slot = (lane >> 3) + 4 * i, piece = lane & 7;
                if (slot < cnt) {
                    const u32 nb = (u32)s.meta[(i0 + slot) & (MW - 1)];
                    cp_async16(&dst[slot][piece * 4], x + (i64)nb * GPC_C + piece * 4);
                }
            }
        }
        ++n_issued;
    };
    for (int i = 0; i < D - 1; ++i) { issue(); cp_async_commit(); }

    u32 b1r[2][4][2], b2r[2][4][2];       // [u][j][b0/b1]: hi and lo halves of W[k] as mma B fragments
    int cur_k = -1;
    for (int c = 0;; ++c) {
        u32 gp; int cnt, gk;
        const bool have = it_cmp.next(s.seg, gp, cnt, gk);
        if (!have) break;
        const u32 i0 = gp - p_begin;
        refill(i0);
        issue();
        cp_async_commit();
        cp_async_wait<D - 1>();
        __syncwarp();

        if (gk != cur_k) {
            cur_k = gk;
            const u64 *wsrc = Wb + (size_t)gk * 512;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u64 h = __ldg(wsrc + (2 * t + u) * 32 + 8 * j + g);
                    const u64 l = __ldg(wsrc + 256 + (2 * t + u) * 32 + 8 * j + g);
                    b1r[u][j][0] = (u32)h; b1r[u][j][1] = (u32)(h >> 32);
                    b2r[u][j][0] = (u32)l; b2r[u][j][1] = (u32)(l >> 32);
                }
        }
        // A fragments: rows g and g+8 of the group, channels 8t..8t+7, split into bf16 hi / lo
        const float(*xb)[SC4_XS] = s.xs[c % D];
        float xa[8], xc[8];
        *reinterpret_cast<float4 *>(&xa[0]) = *reinterpret_cast<const float4 *>(&xb[g][8 * t]);
        *reinterpret_cast<float4 *>(&xa[4]) = *reinterpret_cast<const float4 *>(&xb[g][8 * t + 4]);
        *reinterpret_cast<float4 *>(&xc[0]) = *reinterpret_cast<const float4 *>(&xb[g + 8][8 * t]);
        *reinterpret_cast<float4 *>(&xc[4]) = *reinterpret_cast<const float4 *>(&xb[g + 8][8 * t + 4]);
        u32 a1[2][4], a2[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {           // h = 0: channels 4u+0,1 (logical cols 2t,2t+1); h = 1: 4u+2,3 (cols 2t+8,2t+9)
                const float p0 = xa[4 * u + 2 * h], p1 = xa[4 * u + 2 * h + 1];
                const float q0 = xc[4 * u + 2 * h], q1 = xc[4 * u + 2 * h + 1];
                const float p0h = bf16_round(p0), p1h = bf16_round(p1), q0h = bf16_round(q0), q1h = bf16_round(q1);
                a1[u][2 * h] = (__float_as_uint(p0h) >> 16) | (__float_as_uint(p1h) & 0xFFFF0000u);
                a1[u][2 * h + 1] = (__float_as_uint(q0h) >> 16) | (__float_as_uint(q1h) & 0xFFFF0000u);
                a2[u][2 * h] = pack_bf16x2(p0 - p0h, p1 - p1h);
                a2[u][2 * h + 1] = pack_bf16x2(q0 - q0h, q1 - q1h);
            }
        }
        float d[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f; }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                mma_bf16_16816(d[j], a2[u], b1r[u][j][0], b1r[u][j][1]);      // x2.w1
                mma_bf16_16816(d[j], a1[u], b2r[u][j][0], b2r[u][j][1]);      // x1.w2
                mma_bf16_16816(d[j], a1[u], b1r[u][j][0], b1r[u][j][1]);      // x1.w1
            }
        // scatter-add into this warp's accumulator rows (fragment rows g / g+8 <-> pairs g / g+8 of the group)
        if (g < cnt) {
            const u32 r = (u32)(s.meta[(i0 + g) & (MW - 1)] >> 32) & 0xFFFFu;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 *dst = reinterpret_cast<float2 *>(&s.acc[r][8 * j + 2 * t]);
                float2 v = *dst; v.x += d[j][0]; v.y += d[j][1]; *dst = v;
            }
        }
        if (g + 8 < cnt) {
            const u32 r = (u32)(s.meta[(i0 + g + 8) & (MW - 1)] >> 32) & 0xFFFFu;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 *dst = reinterpret_cast<float2 *>(&s.acc[r][8 * j + 2 * t]);
                float2 v = *dst; v.x += d[j][2]; v.y += d[j][3]; *dst = v;
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = 0; r < rows; ++r) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TW, int D, int MW>
static int launch_spconv_v4(const float *x, const void *Wb, const u32 *seg, const u64 *pairs, i64 n, const float *residual,
                            int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc4Smem<TW, D, MW>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v4_kernel<TW, D, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TW - 1) / TW;
    spconv_fwd_v4_kernel<TW, D, MW><<<(unsigned)tiles, 32, smem, st>>>(x, (const u64 *)Wb, seg, pairs, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variant 20: Wb from gpc_spconv_pack_weights_bf16; tile_rows = rows per warp (64 / 128 / 256)
extern "C" int gpc_spconv_fwd_v4(const float *x, const void *Wb, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int tile_rows, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 20) {
        if (tile_rows == 64) return launch_spconv_v4<64, 4, 512>(x, Wb, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v4<128, 4, 512>(x, Wb, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 256) return launch_spconv_v4<256, 4, 512>(x, Wb, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 21) {
        if (tile_rows == 64) return launch_spconv_v4<64, 8, 1024>(x, Wb, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v4<128, 8, 1024>(x, Wb, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 256) return launch_spconv_v4<256, 8, 1024>(x, Wb, seg, pairs, n, residual, flags, y, st);
    }
    gpc_set_error("unsupported conv v4 variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}


// =====================================================================================================
// v5: as v4 but with the roles of the MMA operands swapped: W^T[k] (32 co x 32 ci) is the A operand (held
// in registers per offset, loaded with 8 x LDG.128 from a fragment-ordered pack), the gathered rows are the
// B operand (n = 8 pairs per MMA tile).  Padding granularity drops from 16 to 8 pairs per (sub-tile,
// offset) and the D fragment gives each lane (co, pair) scalars whose smem read-modify-write is one
// conflict-free wavefront for consecutive rows.  Leaner integer path: hardware cvt.rn.bf16x2 for the
// hi/lo split, a compacted list of non-empty offsets, L1 prefetch of the next offset's W.
// =====================================================================================================
// Wa [n_kernels*125][2 part][2 mt][2 u][32 lane] uint4 = (a0,a1,a2,a3) of mma.m16n8k16 for A = W^T:
//   lane (g,t): a0 = (co 16mt+g,   ch 8t+4u+{0,1}), a1 = (co 16mt+g+8, same ch),
//               a2 = (co 16mt+g,   ch 8t+4u+{2,3}), a3 = (co 16mt+g+8, same ch);  part 0 = bf16 hi, 1 = bf16 lo
__global__ void pack_weights_frag_kernel(const float *__restrict__ W, uint4 *__restrict__ Wa, int n_kernels) {
    i64 gi = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (i64)n_kernels * GPC_K3 * 8 * 32) return;
    const int lane = (int)(gi & 31), u = (int)((gi >> 5) & 1), mt = (int)((gi >> 6) & 1), part = (int)((gi >> 7) & 1);
    const i64 k = gi >> 8;
    const int g = lane >> 2, t = lane & 3;
    const float *src = W + k * (GPC_C * GPC_C);
    u32 out[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int co = 16 * mt + g + 8 * (r & 1);
        const int ch = 8 * t + 4 * u + 2 * (r >> 1);
        u32 v = 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float w = src[(ch + e) * GPC_C + co];
            const float w1 = bf16_round(w);
            const float pv = part == 0 ? w1 : bf16_round(w - w1);
            v |= (__float_as_uint(pv) >> 16) << (16 * e);
        }
        out[r] = v;
    }
    Wa[gi] = make_uint4(out[0], out[1], out[2], out[3]);
}
extern "C" int gpc_spconv_pack_weights_frag(const float *W, int n_kernels, void *Wa, void *stream) {
    const i64 total = (i64)n_kernels * GPC_K3 * 8 * 32;
    pack_weights_frag_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(W, (uint4 *)Wa, n_kernels);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

constexpr int SC5_XS = 36;
constexpr int SC5_ACC = 36;

template <int TW, int D, int MW>
struct Sc5Smem {
    float acc[TW][SC5_ACC];
    float xs[D][16][SC5_XS];
    u64 meta[MW];
    u32 sk[GPC_K3 + 3];          // compacted non-empty offsets: k
    u32 sb[GPC_K3 + 3];          // their stream begin (relative); sb[nseg] = total
};

__device__ __forceinline__ void mma_bf16_a4(float (&d)[4], const uint4 &a, u32 b0, u32 b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

template <int TW, int D, int MW>
__global__ void __launch_bounds__(32) spconv_fwd_v5_kernel(const float *__restrict__ x, const uint4 *__restrict__ Wa,
                                                           const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs, i64 n,
                                                           const float *__restrict__ residual, int flags, float *__restrict__ y) {
    static_assert((MW & (MW - 1)) == 0 && MW >= 32 * (D + 3), "meta window too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sc5Smem<TW, D, MW> &s = *reinterpret_cast<Sc5Smem<TW, D, MW> *>(smem_raw);
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    const i64 st = blockIdx.x;
    const i64 r0 = st * TW;
    const int rows = (int)min((i64)TW, n - r0);
    const u32 *seg = seg_g + st * (GPC_K3 + 1);
    const u32 p_begin = __ldg(seg);

    // compact the non-empty offsets of this sub-tile
    int nseg = 0;
    for (int base = 0; base < GPC_K3; base += 32) {
        const int k = base + lane;
        u32 b = 0, e = 0;
        if (k < GPC_K3) { b = __ldg(seg + k); e = __ldg(seg + k + 1); }
        const bool ne = e > b;
        const u32 bal = __ballot_sync(0xFFFFFFFFu, ne);
        if (ne) { const int idx = nseg + __popc(bal & ((1u << lane) - 1u)); s.sk[idx] = (u32)k; s.sb[idx] = b - p_begin; }
        nseg += __popc(bal);
    }
    const u32 total = __ldg(seg + GPC_K3) - p_begin;
    if (lane == 0) s.sb[nseg] = total;
    for (int i = lane; i < TW * SC5_ACC / 4; i += 32) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    u32 fetched = 0;
    auto refill = [&](u32 dead_before) {
        while (fetched < total && fetched + 32 - dead_before <= (u32)MW) {
            const u32 idx = fetched + lane;
            if (idx < total) cp_async8(&s.meta[idx & (MW - 1)], pairs + p_begin + idx);
            fetched += 32;
        }
    };
    refill(0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    // issue-side and compute-side cursors over units (<= 16 pairs of one offset)
    int si_i = 0; u32 p_i = 0; int n_issued = 0;
    auto issue = [&]() {
        if (p_i < total) {
            while (p_i >= s.sb[si_i + 1]) ++si_i;
            const int cnt = (int)min(16u, s.sb[si_i + 1] - p_i);
            if (p_i == s.sb[si_i]) {      // first unit of an offset: pull its W (4 KB = 32 lines) towards L1
                const char *wl = reinterpret_cast<const char *>(Wa) + (size_t)s.sk[si_i] * 4096 + lane * 128;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(wl));
            }
            float(*dst)[SC5_XS] = s.xs[n_issued % D];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int slot = (lane >> 3) + 4 * i, piece = lane & 7;
                if (slot < cnt) {
                    const u32 nb = (u32)s.meta[(p_i + slot) & (MW - 1)];
                    cp_async16(&dst[slot][piece * 4], x + (i64)nb * GPC_C + piece * 4);
                }
            }
            p_i += cnt;
        }
        ++n_issued;
    };
    for (int i = 0; i < D - 1; ++i) { issue(); cp_async_commit(); }

    uint4 w1[2][2], w2[2][2];             // [mt][u] A fragments of W^T[k]: bf16 hi / lo
    int si_c = 0, cur_si = -1;
    u32 p_c = 0;
    for (int c = 0; p_c < total; ++c) {
        while (p_c >= s.sb[si_c + 1]) ++si_c;
        const int cnt = (int)min(16u, s.sb[si_c + 1] - p_c);
        refill(p_c);
        issue();
        cp_async_commit();
        if (si_c != cur_si) {
            cur_si = si_c;
            const uint4 *wsrc = Wa + (size_t)s.sk[si_c] * 256 + lane;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    w1[mt][u] = __ldg(wsrc + (0 * 4 + mt * 2 + u) * 32);
                    w2[mt][u] = __ldg(wsrc + (1 * 4 + mt * 2 + u) * 32);
                }
        }
        cp_async_wait<D - 1>();
        __syncwarp();

        const float(*xb)[SC5_XS] = s.xs[c % D];
        const int ntile = cnt > 8 ? 2 : 1;
#pragma unroll 1
        for (int nt = 0; nt < ntile; ++nt) {
            // B fragments: pair (8nt + g), channels 8t..8t+7
            const float4 xa = *reinterpret_cast<const float4 *>(&xb[8 * nt + g][8 * t]);
            const float4 xc = *reinterpret_cast<const float4 *>(&xb[8 * nt + g][8 * t + 4]);
            u32 x1[2][2], x2[2][2];       // [u][b0/b1] hi / lo
            split_bf16(xa.x, xa.y, x1[0][0], x2[0][0]);
            split_bf16(xa.z, xa.w, x1[0][1], x2[0][1]);
            split_bf16(xc.x, xc.y, x1[1][0], x2[1][0]);
            split_bf16(xc.z, xc.w, x1[1][1], x2[1][1]);
            float d[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f; }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16_a4(d[mt], w1[mt][u], x2[u][0], x2[u][1]);       // w1.x2
                    mma_bf16_a4(d[mt], w2[mt][u], x1[u][0], x1[u][1]);       // w2.x1
                    mma_bf16_a4(d[mt], w1[mt][u], x1[u][0], x1[u][1]);       // w1.x1
                }
            // D^T fragment: d[mt][0] = (co 16mt+g, pair 2t), [1] = (co, pair 2t+1), [2] = (co+8, pair 2t), [3] = (co+8, pair 2t+1)
            const int pa = 8 * nt + 2 * t;
            if (pa < cnt) {
                const u32 r = (u32)(s.meta[(p_c + pa) & (MW - 1)] >> 32) & 0xFFFFu;
                float *a = &s.acc[r][g];
                a[0] += d[0][0]; a[8] += d[0][2]; a[16] += d[1][0]; a[24] += d[1][2];
            }
            if (pa + 1 < cnt) {
                const u32 r = (u32)(s.meta[(p_c + pa + 1) & (MW - 1)] >> 32) & 0xFFFFu;
                float *a = &s.acc[r][g];
                a[0] += d[0][1]; a[8] += d[0][3]; a[16] += d[1][1]; a[24] += d[1][3];
            }
        }
        p_c += cnt;
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = 0; r < rows; ++r) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TW, int D, int MW>
static int launch_spconv_v5(const float *x, const void *Wa, const u32 *seg, const u64 *pairs, i64 n, const float *residual,
                            int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc5Smem<TW, D, MW>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v5_kernel<TW, D, MW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TW - 1) / TW;
    spconv_fwd_v5_kernel<TW, D, MW><<<(unsigned)tiles, 32, smem, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variant 30/31: Wa from gpc_spconv_pack_weights_frag; tile_rows = rows per warp (32 / 64 / 128)
extern "C" int gpc_spconv_fwd_v5(const float *x, const void *Wa, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int tile_rows, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 30) {
        if (tile_rows == 32) return launch_spconv_v5<32, 3, 256>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 64) return launch_spconv_v5<64, 3, 256>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v5<128, 3, 256>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 31) {
        if (tile_rows == 64) return launch_spconv_v5<64, 5, 256>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v5<128, 5, 256>(x, Wa, seg, pairs, n, residual, flags, y, st);
    }
    gpc_set_error("unsupported conv v5 variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}


// =====================================================================================================
// v6: v5's contraction, re-plumbed for instruction count and latency.
//   * the pair stream is padded per (sub-tile, offset) to whole 8-entry MMA tiles (kmap pad = 8), so the
//     kernel walks a flat list of tiles: no offset iterators, no window bookkeeping; a tile's offset is the
//     k field of its first entry;
//   * explicit software pipeline inside the warp (in-order issue): while the 12 HMMAs of tile c run, the raw
//     inputs / row ids / W fragments of tile c+1 are already being loaded and are converted before the
//     read-modify-write of tile c -- no instruction waits on a load issued in the same iteration;
//   * tile c+D's rows are gathered by cp.async (2 x 16 B per lane), its entry was prefetched to a register
//     one iteration earlier.
// =====================================================================================================
constexpr int SC6_XS = 36;
constexpr int SC6_ACC = 36;

template <int TW, int D>
struct Sc6Smem {
    float acc[TW][SC6_ACC];
    float xs[D][8][SC6_XS];
    u32 rowk[D][12];              // [0..7] row of each pair in the tile (0xFFFF = padding), [8] = offset k
};

// G > 1 ("split offsets"): the CTA has G warps on the SAME TW rows; warp w contracts the offsets [125 w / G, 125 (w+1) / G) into its
// own accumulators and the partial sums are added in warp order at the end (still one fixed summation order per row).  The coarse
// octree levels have a few hundred to a few thousand rows with 35-70 neighbours each: one warp per 8 rows walked ~100 dependent
// 8-pair tiles and every such launch took ~70 us whatever its size (launch list, profiles/r01_step_breakdown.md).
template <int TW, int D, int G>
__global__ void __launch_bounds__(32 * G) spconv_fwd_v6_kernel(const float *__restrict__ x, const uint4 *__restrict__ Wa,
                                                               const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs, i64 n,
                                                               const float *__restrict__ residual, int flags, float *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = G > 1 ? (int)(threadIdx.x >> 5) : 0;
    Sc6Smem<TW, D> &s = reinterpret_cast<Sc6Smem<TW, D> *>(smem_raw)[wid];
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const i64 st = blockIdx.x;
    const i64 r0 = st * TW;
    const int rows = (int)min((i64)TW, n - r0);
    const int kb = GPC_K3 * wid / G, ke = GPC_K3 * (wid + 1) / G;
    const u32 p_begin = __ldg(seg_g + st * (GPC_K3 + 1) + kb);
    const int ntiles = (int)((__ldg(seg_g + st * (GPC_K3 + 1) + ke) - p_begin) >> 3);
    const u64 *tile_base = pairs + p_begin;

    for (int i = lane; i < TW * SC6_ACC / 4; i += 32) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int er = lane & 7, epg = lane >> 3;          // this lane copies row `er` of a tile, 32-byte piece group `epg`
    auto load_entry = [&](int tile) -> u64 {
        return tile < ntiles ? __ldg(tile_base + (i64)tile * 8 + er) : 0xFFFFFFFFFFFFFFFFull;
    };
    auto issue_tile = [&](int tile, u64 e) {             // gather rows of `tile` into ring slot tile % D
        if (tile < ntiles) {
            const int slot = tile % D;
            const u32 nb = (u32)e;
            if (nb != 0xFFFFFFFFu) {
                const float *src = x + (i64)nb * GPC_C + epg * 8;
                cp_async16(&s.xs[slot][er][epg * 8], src);
                cp_async16(&s.xs[slot][er][epg * 8 + 4], src + 4);
            }
            if (lane < 8) s.rowk[slot][lane] = (u32)(e >> 32) & 0xFFFFu;
            if (lane == 0) s.rowk[slot][8] = (u32)(e >> 48);
        }
    };

    // prologue: tiles 0..D-1 in flight, entry of tile D in a register
    u64 e_next = load_entry(0);
#pragma unroll 1
    for (int i = 0; i < D; ++i) {
        const u64 e = e_next;
        e_next = load_entry(i + 1);
        issue_tile(i, e);
        cp_async_commit();
    }
    cp_async_wait<D - 1>();                              // tile 0 landed
    __syncwarp();

    uint4 w1[2][2], w2[2][2];                            // A fragments of W^T[k] (bf16 hi / lo), [mt][u]
    u32 xf1[2][2], xf2[2][2];                            // B fragments of the current tile (bf16 hi / lo), [u][b0/b1]
    u32 row_a = 0xFFFFu, row_b = 0xFFFFu;                // accumulator rows of pairs 2t, 2t+1 of the current tile
    u32 k_cur = 0xFFFFFFFFu;
    if (ntiles > 0) {
        const float4 xa = *reinterpret_cast<const float4 *>(&s.xs[0][g][8 * t]);
        const float4 xc = *reinterpret_cast<const float4 *>(&s.xs[0][g][8 * t + 4]);
        row_a = s.rowk[0][2 * t]; row_b = s.rowk[0][2 * t + 1];
        k_cur = s.rowk[0][8];
        const uint4 *wsrc = Wa + (size_t)k_cur * 256 + lane;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 2; ++u) { w1[mt][u] = __ldg(wsrc + (mt * 2 + u) * 32); w2[mt][u] = __ldg(wsrc + (4 + mt * 2 + u) * 32); }
        split_bf16(xa.x, xa.y, xf1[0][0], xf2[0][0]);
        split_bf16(xa.z, xa.w, xf1[0][1], xf2[0][1]);
        split_bf16(xc.x, xc.y, xf1[1][0], xf2[1][0]);
        split_bf16(xc.z, xc.w, xf1[1][1], xf2[1][1]);
    }

#pragma unroll 1
    for (int c = 0; c < ntiles; ++c) {
        // A. tiles <= c+1 have landed
        cp_async_wait<D - 2>();
        __syncwarp();
        // B. refill the ring slot tile c just vacated (its data lives in registers since the previous iteration)
        {
            const u64 e = e_next;
            e_next = load_entry(c + D + 1);
            issue_tile(c + D, e);
            cp_async_commit();
        }
        // C. loads for tile c+1 (raw rows, accumulator rows, offset, W if it changes)
        const bool more = c + 1 < ntiles;
        const int ns = (c + 1) % D;
        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xc = xa;
        u32 nrow_a = 0xFFFFu, nrow_b = 0xFFFFu, k_next = k_cur;
        if (more) {
            xa = *reinterpret_cast<const float4 *>(&s.xs[ns][g][8 * t]);
            xc = *reinterpret_cast<const float4 *>(&s.xs[ns][g][8 * t + 4]);
            nrow_a = s.rowk[ns][2 * t]; nrow_b = s.rowk[ns][2 * t + 1];
            k_next = s.rowk[ns][8];
        }
        uint4 n1[2][2], n2[2][2];
        const bool newk = k_next != k_cur;               // warp-uniform
        if (newk) {
            const uint4 *wsrc = Wa + (size_t)k_next * 256 + lane;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 2; ++u) { n1[mt][u] = __ldg(wsrc + (mt * 2 + u) * 32); n2[mt][u] = __ldg(wsrc + (4 + mt * 2 + u) * 32); }
        }
        // D. contraction of tile c
        float d[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f; }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_bf16_a4(d[mt], w1[mt][u], xf2[u][0], xf2[u][1]);
                mma_bf16_a4(d[mt], w2[mt][u], xf1[u][0], xf1[u][1]);
                mma_bf16_a4(d[mt], w1[mt][u], xf1[u][0], xf1[u][1]);
            }
        // E. convert tile c+1's inputs while the HMMAs drain
        u32 y1[2][2], y2[2][2];
        split_bf16(xa.x, xa.y, y1[0][0], y2[0][0]);
        split_bf16(xa.z, xa.w, y1[0][1], y2[0][1]);
        split_bf16(xc.x, xc.y, y1[1][0], y2[1][0]);
        split_bf16(xc.z, xc.w, y1[1][1], y2[1][1]);
        // F. scatter-add tile c: d[mt][0] = (co 16mt+g, pair 2t), [1] = (co, pair 2t+1), [2]/[3] = co+8.  The two pairs of a lane
        // belong to ONE offset, hence to different rows: all eight loads are issued before the first store (written as two
        // separate read-modify-writes the compiler has to assume they alias and serialises load - add - store - load - add - store)
        {
            const bool va = row_a != 0xFFFFu, vb = row_b != 0xFFFFu;
            float *pa = &s.acc[va ? row_a : 0u][g], *pb = &s.acc[vb ? row_b : 0u][g];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            if (va) { a0 = pa[0]; a1 = pa[8]; a2 = pa[16]; a3 = pa[24]; }
            if (vb) { b0 = pb[0]; b1 = pb[8]; b2 = pb[16]; b3 = pb[24]; }
            a0 += d[0][0]; a1 += d[0][2]; a2 += d[1][0]; a3 += d[1][2];
            b0 += d[0][1]; b1 += d[0][3]; b2 += d[1][1]; b3 += d[1][3];
            if (va) { pa[0] = a0; pa[8] = a1; pa[16] = a2; pa[24] = a3; }
            if (vb) { pb[0] = b0; pb[8] = b1; pb[16] = b2; pb[24] = b3; }
        }
        // G. rotate
#pragma unroll
        for (int u = 0; u < 2; ++u) { xf1[u][0] = y1[u][0]; xf1[u][1] = y1[u][1]; xf2[u][0] = y2[u][0]; xf2[u][1] = y2[u][1]; }
        row_a = nrow_a; row_b = nrow_b;
        if (newk) {
            k_cur = k_next;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 2; ++u) { w1[mt][u] = n1[mt][u]; w2[mt][u] = n2[mt][u]; }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
    if (G > 1) __syncthreads();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = wid; r < rows; r += G) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < G; ++w) v += reinterpret_cast<Sc6Smem<TW, D> *>(smem_raw)[w].acc[r][lane];      // warp order: offsets ascending
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

// ---- v6d: v6 without the shared-memory gather ring.  Lane (g, t) of an 8-pair tile needs exactly the 32 contiguous bytes
// [8t, 8t+8) of pair g's row, so the four lanes of a quad read one 128 B row straight into the B fragments (2 x LDG.128 per lane):
// no cp.async, no staging, no LDS, no wait_group / syncwarp per tile.  Entries come 32 at a time (4 tiles, one coalesced 256 B load)
// and are handed out by shuffles; the rows of tile T + 3 are requested before tile T is multiplied (four register slots indexed
// statically by the unrolled loop), W^T[k] of tile T + 1 while tile T is multiplied.  Accumulators and their fixed order as in v6.
template <int TW>
__global__ void __launch_bounds__(32) spconv_fwd_v6d_kernel(const float *__restrict__ x, const uint4 *__restrict__ Wa,
                                                            const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs, i64 n,
                                                            const float *__restrict__ residual, int flags, float *__restrict__ y,
                                                            i64 tile0) {
    __shared__ float acc[TW][SC6_ACC];
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    const i64 st = tile0 + blockIdx.x;                   // tiles [tile0, tile0 + gridDim.x) of the level
    const i64 r0 = st * TW;
    const int rows = (int)min((i64)TW, n - r0);
    const u32 p_begin = __ldg(seg_g + st * (GPC_K3 + 1));
    const int ntiles = (int)((__ldg(seg_g + st * (GPC_K3 + 1) + GPC_K3) - p_begin) >> 3);
    const u64 *tile_base = pairs + p_begin;
    for (int i = lane; i < TW * SC6_ACC / 4; i += 32) reinterpret_cast<float4 *>(&acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    auto load_bulk = [&](int k) -> u64 {                 // lane l: entry (tile 4k + l / 8, pair l % 8)
        const int e = k * 32 + lane;
        return e < ntiles * 8 ? __ldg(tile_base + e) : ~0ull;
    };
    u64 eb[2] = {load_bulk(0), load_bulk(1)};   // bulks 2j / 2j + 1: statically indexed (a register hand-over would wait for every load in flight)
    float4 xa[4], xc[4];
    u32 rwa[4], rwb[4], kof[4];
    // an entry is handed out in two steps one tile apart, so that nothing waits for a shuffle: pick (shuffles into p_*), then
    // fetch (row loads from the picked neighbour index, bookkeeping into the tile's slot)
    u32 p_nb = 0xFFFFFFFFu, p_ra = 0xFFFFu, p_rb = 0xFFFFu, p_kk = 0;
    auto pick = [&](u64 ebv, int tl) {
        p_nb = __shfl_sync(0xFFFFFFFFu, (u32)ebv, tl * 8 + g);
        const u32 hi = (u32)(ebv >> 32);
        p_ra = __shfl_sync(0xFFFFFFFFu, hi, tl * 8 + 2 * t) & 0xFFFFu;             // padding entries: 0xFFFF
        p_rb = __shfl_sync(0xFFFFFFFFu, hi, tl * 8 + 2 * t + 1) & 0xFFFFu;
        p_kk = __shfl_sync(0xFFFFFFFFu, hi, tl * 8) >> 16;                         // entry 0 of a tile is always a real pair
    };
    auto fetch = [&](float4 &a, float4 &c, u32 &ra, u32 &rb, u32 &kk) {
        ra = p_ra; rb = p_rb; kk = p_kk;
        a = make_float4(0.f, 0.f, 0.f, 0.f); c = a;
        if (p_nb != 0xFFFFFFFFu) {
            const float4 *src = reinterpret_cast<const float4 *>(x + (i64)p_nb * GPC_C + 8 * t);
            a = __ldg(src); c = __ldg(src + 1);
        }
    };
    auto load_w = [&](uint4 (&a1)[2][2], uint4 (&a2)[2][2], u32 k) {
        const uint4 *wsrc = Wa + (size_t)k * 256 + lane;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 2; ++u) { a1[mt][u] = __ldg(wsrc + (mt * 2 + u) * 32); a2[mt][u] = __ldg(wsrc + (4 + mt * 2 + u) * 32); }
    };
    uint4 w1[2][2], w2[2][2];
    u32 k_cur = 0xFFFFFFFFu;
    if (ntiles > 0) {
        pick(eb[0], 0); fetch(xa[0], xc[0], rwa[0], rwb[0], kof[0]);
        if (ntiles > 1) { pick(eb[0], 1); fetch(xa[1], xc[1], rwa[1], rwb[1], kof[1]); }
        if (ntiles > 2) { pick(eb[0], 2); fetch(xa[2], xc[2], rwa[2], rwb[2], kof[2]); }
        pick(eb[0], 3);
        k_cur = kof[0];
        load_w(w1, w2, k_cur);
    }
#pragma unroll 1
    for (int kb = 0; 8 * kb < ntiles; ++kb) {
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) {
            const int u = u8 & 3, h = u8 >> 2;
            const int T = 8 * kb + u8;
            if (T >= ntiles) break;
            // tile T - 1's slot is free: its fragments were converted and its rows / offset consumed in the previous step
            if (T + 3 < ntiles) fetch(xa[(u + 3) & 3], xc[(u + 3) & 3], rwa[(u + 3) & 3], rwb[(u + 3) & 3], kof[(u + 3) & 3]);
            // bulk eb[h] (tiles T .. T + 3) was picked completely during the previous four steps: fetch the bulk after the next one
            if (u == 0) eb[h] = load_bulk(2 * kb + h + 2);
            pick(eb[h ^ 1], u);                          // tile T + 4
            // W of the next tile if its offset differs (warp-uniform): in flight during this tile's MMAs.  (Requesting it as soon as
            // the offset shows up in the 3-tile look-ahead window was measured 20 % SLOWER: the fragments then live across iterations.)
            const u32 k_next = T + 1 < ntiles ? kof[(u + 1) & 3] : k_cur;
            const bool newk = k_next != k_cur;
            uint4 n1[2][2], n2[2][2];
            if (newk) load_w(n1, n2, k_next);
            u32 xf1[2][2], xf2[2][2];
            split_bf16(xa[u].x, xa[u].y, xf1[0][0], xf2[0][0]);
            split_bf16(xa[u].z, xa[u].w, xf1[0][1], xf2[0][1]);
            split_bf16(xc[u].x, xc[u].y, xf1[1][0], xf2[1][0]);
            split_bf16(xc[u].z, xc[u].w, xf1[1][1], xf2[1][1]);
            float d[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) { d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f; }
#pragma unroll
            for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16_a4(d[mt], w1[mt][uu], xf2[uu][0], xf2[uu][1]);
                    mma_bf16_a4(d[mt], w2[mt][uu], xf1[uu][0], xf1[uu][1]);
                    mma_bf16_a4(d[mt], w1[mt][uu], xf1[uu][0], xf1[uu][1]);
                }
            {   // scatter-add: the two pairs of a lane are different rows; all loads before the first store
                const u32 row_a = rwa[u], row_b = rwb[u];
                const bool va = row_a != 0xFFFFu, vb = row_b != 0xFFFFu;
                float *pa = &acc[va ? row_a : 0u][g], *pb = &acc[vb ? row_b : 0u][g];
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
                if (va) { a0 = pa[0]; a1 = pa[8]; a2 = pa[16]; a3 = pa[24]; }
                if (vb) { b0 = pb[0]; b1 = pb[8]; b2 = pb[16]; b3 = pb[24]; }
                a0 += d[0][0]; a1 += d[0][2]; a2 += d[1][0]; a3 += d[1][2];
                b0 += d[0][1]; b1 += d[0][3]; b2 += d[1][1]; b3 += d[1][3];
                if (va) { pa[0] = a0; pa[8] = a1; pa[16] = a2; pa[24] = a3; }
                if (vb) { pb[0] = b0; pb[8] = b1; pb[16] = b2; pb[24] = b3; }
            }
            __syncwarp();                                // the next tile may touch the same accumulator rows from other lanes
            if (newk) {
                k_cur = k_next;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int uu = 0; uu < 2; ++uu) { w1[mt][uu] = n1[mt][uu]; w2[mt][uu] = n2[mt][uu]; }
            }
        }
    }
    __syncwarp();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = 0; r < rows; ++r) {
        float v = acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TW, int D, int G = 1>
static int launch_spconv_v6(const float *x, const void *Wa, const u32 *seg, const u64 *pairs, i64 n, const float *residual,
                            int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc6Smem<TW, D>) * G;
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v6_kernel<TW, D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TW - 1) / TW;
    spconv_fwd_v6_kernel<TW, D, G><<<(unsigned)tiles, 32 * G, smem, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variants 40/41: Wa from gpc_spconv_pack_weights_frag; pair stream built with pad = 8
extern "C" int gpc_spconv_fwd_v6(const float *x, const void *Wa, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int tile_rows, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 40) {
        if (tile_rows == 32) return launch_spconv_v6<32, 6>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 64) return launch_spconv_v6<64, 6>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v6<128, 6>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 256) return launch_spconv_v6<256, 6>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 41) {
        if (tile_rows == 64) return launch_spconv_v6<64, 12>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v6<128, 12>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 42) {
        if (tile_rows == 8) return launch_spconv_v6<8, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 16) return launch_spconv_v6<16, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 32) return launch_spconv_v6<32, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 64) return launch_spconv_v6<64, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v6<128, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 43) {
        if (tile_rows == 64) return launch_spconv_v6<64, 3>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 44) {          // split offsets over 4 warps (coarse levels)
        if (tile_rows == 8) return launch_spconv_v6<8, 4, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 16) return launch_spconv_v6<16, 4, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 32) return launch_spconv_v6<32, 4, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 64) return launch_spconv_v6<64, 4, 4>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 46) {          // split offsets over 8 warps
        if (tile_rows == 8) return launch_spconv_v6<8, 4, 8>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 16) return launch_spconv_v6<16, 4, 8>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 32) return launch_spconv_v6<32, 4, 8>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 48) {          // v6d: rows straight into the MMA fragments (no shared-memory gather ring)
        const i64 tiles = (n + tile_rows - 1) / tile_rows;
        if (tile_rows == 64) { spconv_fwd_v6d_kernel<64><<<(unsigned)tiles, 32, 0, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y, 0); GPC_LAUNCH_CHECK(); return GPC_OK; }
        if (tile_rows == 32) { spconv_fwd_v6d_kernel<32><<<(unsigned)tiles, 32, 0, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y, 0); GPC_LAUNCH_CHECK(); return GPC_OK; }
        if (tile_rows == 128) { spconv_fwd_v6d_kernel<128><<<(unsigned)tiles, 32, 0, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y, 0); GPC_LAUNCH_CHECK(); return GPC_OK; }
    } else if (variant == 47) {          // split offsets over 16 warps
        if (tile_rows == 8) return launch_spconv_v6<8, 4, 16>(x, Wa, seg, pairs, n, residual, flags, y, st);
    } else if (variant == 45) {          // split offsets over 2 warps
        if (tile_rows == 16) return launch_spconv_v6<16, 4, 2>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 32) return launch_spconv_v6<32, 4, 2>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 64) return launch_spconv_v6<64, 4, 2>(x, Wa, seg, pairs, n, residual, flags, y, st);
    }
    gpc_set_error("unsupported conv v6 variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}

// v6d (variant 48) for the output rows [row0, row1) only: row0 a multiple of tile_rows, row1 a multiple of tile_rows or n.  x, y,
// residual and the pair stream are the level's.  Used by the decoder's stage wavefront (codec.py).
extern "C" int gpc_spconv_fwd_v6_rows(const float *x, const void *Wa, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                      int tile_rows, const float *residual, int flags, float *y, int variant, int64_t row0,
                                      int64_t row1, void *stream) {
    if (n <= 0 || row1 <= row0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    GPC_REQUIRE(variant == 48 && (tile_rows == 128 || tile_rows == 64 || tile_rows == 32), GPC_EINVAL, "row ranges: v6d (variant 48) only");
    GPC_REQUIRE(row0 >= 0 && row1 <= n && row0 % tile_rows == 0 && (row1 % tile_rows == 0 || row1 == n), GPC_EINVAL,
                "row range must be made of whole tiles");
    cudaStream_t st = as_stream(stream);
    const i64 tile0 = row0 / tile_rows, tiles = (row1 - row0 + tile_rows - 1) / tile_rows;
    if (tile_rows == 128) spconv_fwd_v6d_kernel<128><<<(unsigned)tiles, 32, 0, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y, tile0);
    else if (tile_rows == 64) spconv_fwd_v6d_kernel<64><<<(unsigned)tiles, 32, 0, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y, tile0);
    else spconv_fwd_v6d_kernel<32><<<(unsigned)tiles, 32, 0, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y, tile0);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}


// =====================================================================================================
// v7: output-stationary, ROW-TIED tiles: the fp32 accumulators never leave registers.
//
// One warp owns 64 consecutive output rows = 8 groups of 8 rows.  D^T[32 co x 8 rows] of every group is an
// mma accumulator fragment (2 m-tiles x 4 regs), 64 registers for the sub-tile, live for the whole kernel.
// For each offset k present in the sub-tile (kmap "rt8" header byte = touched groups) W^T[k] is loaded once
// (8 x LDG.128, prefetched one offset ahead) and every touched group gets one MMA tile whose B operand holds
// the neighbour rows of its 8 output rows (zero where absent).  No scatter, no shared-memory accumulators,
// no read-modify-write: ncu on v6 showed the L1/shared data path at 71 % with the RMW (31 of 83 wavefronts per
// tile) and the per-tile W reload (32) as its main consumers.  Row-tied tiles are ~2x more numerous than
// compacted ones (fill 0.3-0.45) but the tensor pipe was only 25 % busy.
// Accumulation order per output element: offsets ascending, fixed -> encoder == decoder bit for bit.
// =====================================================================================================
constexpr int SC7_XS = 36;

constexpr int SC7_DE = 32;     // entry ring depth (tiles); entries are fetched SC7_EA tiles ahead of their gather
constexpr int SC7_EA = 16;

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src));
}

template <int D>
struct Sc7Smem {
    float xs[D][8][SC7_XS];
    u32 ent[SC7_DE][8];
    u32 vmask[D + 2];
    u32 klist[128];
};

template <int D>
__global__ void __launch_bounds__(32) spconv_fwd_v7_kernel(const float *__restrict__ x, const uint4 *__restrict__ Wa,
                                                           const u32 *__restrict__ toff, const u8 *__restrict__ hdr,
                                                           const u32 *__restrict__ tiles, i64 n,
                                                           const float *__restrict__ residual, int flags, float *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sc7Smem<D> &s = *reinterpret_cast<Sc7Smem<D> *>(smem_raw);
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    const i64 st = blockIdx.x;
    const i64 r0 = st * 64;

    // present offsets of this sub-tile: k | touched-group mask << 8
    int nk = 0;
#pragma unroll
    for (int base = 0; base < 128; base += 32) {
        const u32 m = hdr[st * 128 + base + lane];
        const u32 bal = __ballot_sync(0xFFFFFFFFu, m != 0);
        if (m) s.klist[nk + __popc(bal & ((1u << lane) - 1u))] = (u32)(base + lane) | (m << 8);
        nk += __popc(bal);
    }
    const u32 t_begin = __ldg(toff + st * (GPC_K3 + 1));
    const int T = (int)(__ldg(toff + st * (GPC_K3 + 1) + GPC_K3) - t_begin);
    const u32 *tb = tiles + (i64)t_begin * 8;
    __syncwarp();

    const int er = lane & 7, epg = lane >> 3;
    // the 8 input-row ids of a tile ride a shared ring, fetched SC7_EA tiles before the tile's rows are gathered, so
    // no global load sits between consecutive tiles (ncu on the first v7: ~1300 clk per tile = one exposed L2 trip)
    auto fetch_entries = [&](int tile) {
        if (tile < T && lane < 8) cp_async4(&s.ent[tile % SC7_DE][lane], tb + (i64)tile * 8 + lane);
    };
    auto issue_tile = [&](int tile) {
        if (tile < T) {
            const int slot = tile % D;
            const u32 nb = s.ent[tile % SC7_DE][er];
            const bool valid = nb != 0xFFFFFFFFu;
            if (valid) {
                const float *src = x + (i64)nb * GPC_C + epg * 8;
                cp_async16(&s.xs[slot][er][epg * 8], src);
                cp_async16(&s.xs[slot][er][epg * 8 + 4], src + 4);
            }
            const u32 vm = __ballot_sync(0xFFFFFFFFu, valid && lane < 8);
            if (lane == 0) s.vmask[slot] = vm;
        }
    };
#pragma unroll 1
    for (int i = 0; i < D + SC7_EA; ++i) fetch_entries(i);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < D; ++i) { issue_tile(i); cp_async_commit(); }
    cp_async_wait<D - 1>();
    __syncwarp();

    uint4 w1[2][2], w2[2][2];
    u32 xf1[2][2], xf2[2][2];
    float d[8][2][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { d[j][mt][0] = d[j][mt][1] = d[j][mt][2] = d[j][mt][3] = 0.f; }

    auto load_w = [&](u32 k, uint4 (&a1)[2][2], uint4 (&a2)[2][2]) {
        const uint4 *wsrc = Wa + (size_t)k * 256 + lane;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 2; ++u) { a1[mt][u] = __ldg(wsrc + (mt * 2 + u) * 32); a2[mt][u] = __ldg(wsrc + (4 + mt * 2 + u) * 32); }
    };
    auto load_frag = [&](int slot, u32 (&f1)[2][2], u32 (&f2)[2][2]) {
        const bool valid = (s.vmask[slot] >> g) & 1u;
        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xc = xa;
        if (valid) {
            xa = *reinterpret_cast<const float4 *>(&s.xs[slot][g][8 * t]);
            xc = *reinterpret_cast<const float4 *>(&s.xs[slot][g][8 * t + 4]);
        }
        split_bf16(xa.x, xa.y, f1[0][0], f2[0][0]);
        split_bf16(xa.z, xa.w, f1[0][1], f2[0][1]);
        split_bf16(xc.x, xc.y, f1[1][0], f2[1][0]);
        split_bf16(xc.z, xc.w, f1[1][1], f2[1][1]);
    };
    if (T > 0) { load_w(s.klist[0] & 0xFFu, w1, w2); load_frag(0, xf1, xf2); }

    int c = 0;
#pragma unroll 1
    for (int ki = 0; ki < nk; ++ki) {
        const u32 mask = s.klist[ki] >> 8;
        uint4 n1[2][2], n2[2][2];
        const bool more_k = ki + 1 < nk;
        if (more_k) load_w(s.klist[ki + 1] & 0xFFu, n1, n2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if ((mask >> j) & 1u) {                     // warp-uniform
                cp_async_wait<D - 2>();                 // tiles <= c+1 landed
                __syncwarp();
                fetch_entries(c + D + SC7_EA);
                issue_tile(c + D);
                cp_async_commit();
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_bf16_a4(d[j][mt], w1[mt][u], xf2[u][0], xf2[u][1]);
                        mma_bf16_a4(d[j][mt], w2[mt][u], xf1[u][0], xf1[u][1]);
                        mma_bf16_a4(d[j][mt], w1[mt][u], xf1[u][0], xf1[u][1]);
                    }
                if (c + 1 < T) load_frag((c + 1) % D, xf1, xf2);
                ++c;
            }
        }
        if (more_k) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 2; ++u) { w1[mt][u] = n1[mt][u]; w2[mt][u] = n2[mt][u]; }
        }
    }
    cp_async_wait<0>();

    // epilogue: d[j][mt][0] = (co 16mt+g, row 8j+2t), [1] = (co, row 8j+2t+1), [2]/[3] = co+8
    const bool relu = (flags & GPC_CONV_RELU) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const i64 row = r0 + 8 * j + 2 * t + e;
            if (row < n) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int co = 16 * mt + g + 8 * h;
                        float v = d[j][mt][2 * h + e];
                        if (residual) v += __ldg(residual + row * GPC_C + co);
                        if (relu) v = fmaxf(v, 0.f);
                        y[row * GPC_C + co] = v;
                    }
            }
        }
}

template <int D>
static int launch_spconv_v7(const float *x, const void *Wa, const u32 *toff, const u8 *hdr, const u32 *tiles, i64 n,
                            const float *residual, int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc7Smem<D>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v7_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 nst = (n + 63) / 64;
    spconv_fwd_v7_kernel<D><<<(unsigned)nst, 32, smem, st>>>(x, (const uint4 *)Wa, toff, hdr, tiles, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variants 50 (D=6) / 51 (D=10): Wa from gpc_spconv_pack_weights_frag; kernel map from gpc_kmap_rt8_*
extern "C" int gpc_spconv_fwd_v7(const float *x, const void *Wa, const uint32_t *toff, const uint8_t *hdr, const uint32_t *tiles,
                                 int64_t n, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 50) return launch_spconv_v7<6>(x, Wa, toff, hdr, tiles, n, residual, flags, y, st);
    if (variant == 51) return launch_spconv_v7<10>(x, Wa, toff, hdr, tiles, n, residual, flags, y, st);
    gpc_set_error("unsupported conv v7 variant %d", variant);
    return GPC_EINVAL;
}


// =====================================================================================================
// v8: v6 (compacted 8-pair tiles, shared-memory accumulators, software pipeline) with the per-tile
// instruction count cut: ncu showed v6/v7 bound by the warp's own dependent instruction chain
// (stall_wait ~50 %, 2-3 warps per scheduler, 183 / 130 instructions per tile), not by a memory pipe.
//   * ring depth is a power of two, slots advance by mask; all shared addresses are 32-bit, computed once;
//   * the loop is unrolled by two with ping-pong register sets for W^T[k] and the B fragments: no
//     register-to-register rotation (v6 spent 40 moves per tile on it); W for the next tile is always
//     (re)loaded through L1 -- 8 LDG.128, no branch;
//   * pair entries ride a shared ring fetched 16 tiles ahead (no global load between tiles).
// =====================================================================================================
constexpr int SC8_D = 8;        // gather ring depth (tiles)
constexpr int SC8_DE = 32;      // entry ring depth (tiles)
constexpr int SC8_EA = 16;      // entries fetched this many tiles ahead of the gather
constexpr int SC8_XS = 36;
constexpr int SC8_ACC = 36;

template <int TW>
struct Sc8Smem {
    float acc[TW][SC8_ACC];
    float xs[SC8_D][8][SC8_XS];
    u64 ent[SC8_DE][8];
};

__device__ __forceinline__ void cp_async16_s(u32 smem_addr, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async8_s(u32 smem_addr, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gmem_src));
}
__device__ __forceinline__ float4 lds128(u32 a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ u64 lds64(u32 a) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}

template <int TW>
__global__ void __launch_bounds__(32) spconv_fwd_v8_kernel(const float *__restrict__ x, const uint4 *__restrict__ Wa,
                                                           const u32 *__restrict__ seg_g, const u64 *__restrict__ pairs, i64 n,
                                                           const float *__restrict__ residual, int flags, float *__restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Sc8Smem<TW> &s = *reinterpret_cast<Sc8Smem<TW> *>(smem_raw);
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    const i64 st = blockIdx.x;
    const i64 r0 = st * TW;
    const int rows = (int)min((i64)TW, n - r0);
    const u32 p_begin = __ldg(seg_g + st * (GPC_K3 + 1));
    const int T = (int)((__ldg(seg_g + st * (GPC_K3 + 1) + GPC_K3) - p_begin) >> 3);
    const u64 *tb = pairs + p_begin;

    for (int i = lane; i < TW * SC8_ACC / 4; i += 32) reinterpret_cast<float4 *>(&s.acc[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const u32 xs_base = (u32)__cvta_generic_to_shared(&s.xs[0][0][0]);
    const u32 ent_base = (u32)__cvta_generic_to_shared(&s.ent[0][0]);
    const int er = lane & 7, epg = lane >> 3;
    const u32 my_cp_off = (u32)(er * SC8_XS + epg * 8) * 4u;        // where this lane copies to inside a ring slot
    const u32 my_ld_off = (u32)(g * SC8_XS + 8 * t) * 4u;           // where this lane reads its B fragment from
    constexpr u32 SLOT_BYTES = 8 * SC8_XS * 4;

    auto fetch_entries = [&](int tile) {
        if (tile < T && lane < 8) cp_async8_s(ent_base + (u32)((tile & (SC8_DE - 1)) * 8 + lane) * 8u, tb + (i64)tile * 8 + lane);
    };
    auto issue_tile = [&](int tile) {
        if (tile < T) {
            const u64 e = lds64(ent_base + (u32)((tile & (SC8_DE - 1)) * 8 + er) * 8u);
            const u32 nb = (u32)e;
            if (nb != 0xFFFFFFFFu) {
                const float *src = x + (i64)nb * GPC_C + epg * 8;
                const u32 dst = xs_base + (u32)(tile & (SC8_D - 1)) * SLOT_BYTES + my_cp_off;
                cp_async16_s(dst, src);
                cp_async16_s(dst + 16, src + 4);
            }
        }
    };
    for (int i = 0; i < SC8_D + SC8_EA; ++i) fetch_entries(i);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int i = 0; i < SC8_D; ++i) { issue_tile(i); cp_async_commit(); }
    cp_async_wait<SC8_D - 1>();
    __syncwarp();

    struct Frag { u32 f1[2][2], f2[2][2]; u32 row_a, row_b; };
    struct Wt { uint4 a1[2][2], a2[2][2]; };
    auto load_w = [&](u32 k, Wt &w) {
        const uint4 *wsrc = Wa + (size_t)k * 256 + lane;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 2; ++u) { w.a1[mt][u] = __ldg(wsrc + (mt * 2 + u) * 32); w.a2[mt][u] = __ldg(wsrc + (4 + mt * 2 + u) * 32); }
    };
    // raw loads of tile `tile` (must have landed): B-fragment inputs, accumulator rows of pairs 2t / 2t+1, offset
    auto load_raw = [&](int tile, float4 &xa, float4 &xc, u32 &ra, u32 &rb, u32 &k) {
        const u32 a = xs_base + (u32)(tile & (SC8_D - 1)) * SLOT_BYTES + my_ld_off;
        xa = lds128(a);
        xc = lds128(a + 16);
        const u32 eb = ent_base + (u32)((tile & (SC8_DE - 1)) * 8) * 8u;
        const u64 e0 = lds64(eb), ea = lds64(eb + (u32)(2 * t) * 8u), ebb = lds64(eb + (u32)(2 * t + 1) * 8u);
        k = (u32)(e0 >> 48);
        ra = (u32)(ea >> 32) & 0xFFFFu;
        rb = (u32)(ebb >> 32) & 0xFFFFu;
    };
    auto convert = [&](const float4 &xa, const float4 &xc, Frag &f) {
        split_bf16(xa.x, xa.y, f.f1[0][0], f.f2[0][0]);
        split_bf16(xa.z, xa.w, f.f1[0][1], f.f2[0][1]);
        split_bf16(xc.x, xc.y, f.f1[1][0], f.f2[1][0]);
        split_bf16(xc.z, xc.w, f.f1[1][1], f.f2[1][1]);
    };
    const u32 acc_base = (u32)__cvta_generic_to_shared(&s.acc[0][0]) + (u32)g * 4u;
    // one pipeline step: contract tile c (cur), prepare tile c+1 (nxt)
    u32 k_cur = 0xFFFFFFFFu;
    auto step = [&](int c, const Wt &wc, const Frag &fc, Wt &wn, Frag &fn) {
        cp_async_wait<SC8_D - 2>();                 // tiles <= c+1 landed
        __syncwarp();
        fetch_entries(c + SC8_D + SC8_EA);
        issue_tile(c + SC8_D);
        cp_async_commit();
        float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xc = xa;
        u32 kn = 0;
        fn.row_a = fn.row_b = 0xFFFFu;
        if (c + 1 < T) {
            load_raw(c + 1, xa, xc, fn.row_a, fn.row_b, kn);
            if (kn != k_cur) load_w(kn, wn);        // warp-uniform
            else wn = wc;
            k_cur = kn;
        }
        float d[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f; }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_bf16_a4(d[mt], wc.a1[mt][u], fc.f2[u][0], fc.f2[u][1]);
                mma_bf16_a4(d[mt], wc.a2[mt][u], fc.f1[u][0], fc.f1[u][1]);
                mma_bf16_a4(d[mt], wc.a1[mt][u], fc.f1[u][0], fc.f1[u][1]);
            }
        convert(xa, xc, fn);
        if (fc.row_a != 0xFFFFu) {
            float *a = reinterpret_cast<float *>(smem_raw) + 0;   // (address below is in the shared window)
            (void)a;
            const u32 ad = acc_base + fc.row_a * (u32)(SC8_ACC * 4);
            float v0, v1, v2, v3;
            asm volatile("ld.shared.f32 %0, [%4]; ld.shared.f32 %1, [%4+32]; ld.shared.f32 %2, [%4+64]; ld.shared.f32 %3, [%4+96];"
                         : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(ad));
            v0 += d[0][0]; v1 += d[0][2]; v2 += d[1][0]; v3 += d[1][2];
            asm volatile("st.shared.f32 [%0], %1; st.shared.f32 [%0+32], %2; st.shared.f32 [%0+64], %3; st.shared.f32 [%0+96], %4;"
                         ::"r"(ad), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
        }
        if (fc.row_b != 0xFFFFu) {
            const u32 ad = acc_base + fc.row_b * (u32)(SC8_ACC * 4);
            float v0, v1, v2, v3;
            asm volatile("ld.shared.f32 %0, [%4]; ld.shared.f32 %1, [%4+32]; ld.shared.f32 %2, [%4+64]; ld.shared.f32 %3, [%4+96];"
                         : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(ad));
            v0 += d[0][1]; v1 += d[0][3]; v2 += d[1][1]; v3 += d[1][3];
            asm volatile("st.shared.f32 [%0], %1; st.shared.f32 [%0+32], %2; st.shared.f32 [%0+64], %3; st.shared.f32 [%0+96], %4;"
                         ::"r"(ad), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
        }
    };

    Wt wA, wB;
    Frag fA, fB;
    if (T > 0) {
        float4 xa, xc;
        u32 k0;
        load_raw(0, xa, xc, fA.row_a, fA.row_b, k0);
        load_w(k0, wA);
        k_cur = k0;
        convert(xa, xc, fA);
    }
    int c = 0;
#pragma unroll 1
    for (; c + 1 < T; c += 2) {
        step(c, wA, fA, wB, fB);
        step(c + 1, wB, fB, wA, fA);
    }
    if (c < T) step(c, wA, fA, wB, fB);
    cp_async_wait<0>();
    __syncwarp();
    const bool relu = (flags & GPC_CONV_RELU) != 0;
    for (int r = 0; r < rows; ++r) {
        float v = s.acc[r][lane];
        if (residual) v += __ldg(residual + (r0 + r) * GPC_C + lane);
        if (relu) v = fmaxf(v, 0.f);
        y[(r0 + r) * GPC_C + lane] = v;
    }
}

template <int TW>
static int launch_spconv_v8(const float *x, const void *Wa, const u32 *seg, const u64 *pairs, i64 n, const float *residual,
                            int flags, float *y, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(Sc8Smem<TW>);
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(spconv_fwd_v8_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const i64 tiles = (n + TW - 1) / TW;
    spconv_fwd_v8_kernel<TW><<<(unsigned)tiles, 32, smem, st>>>(x, (const uint4 *)Wa, seg, pairs, n, residual, flags, y);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// variant 60: Wa from gpc_spconv_pack_weights_frag; pair stream built with pad = 8
extern "C" int gpc_spconv_fwd_v8(const float *x, const void *Wa, const uint32_t *seg, const uint64_t *pairs, int64_t n,
                                 int tile_rows, const float *residual, int flags, float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 60) {
        if (tile_rows == 32) return launch_spconv_v8<32>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 64) return launch_spconv_v8<64>(x, Wa, seg, pairs, n, residual, flags, y, st);
        if (tile_rows == 128) return launch_spconv_v8<128>(x, Wa, seg, pairs, n, residual, flags, y, st);
    }
    gpc_set_error("unsupported conv v8 variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}


// variant: 0 = v1 (unpipelined, W in reference layout), 1.. = v2 configurations (W packed by gpc_spconv_pack_weights)
extern "C" int gpc_spconv_fwd(const float *x, const float *W, const uint32_t *seg, const uint32_t *pair_nbr,
                              const uint16_t *pair_row, int64_t n, int tile_rows, const float *residual, int flags,
                              float *y, int variant, void *stream) {
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(x != y, GPC_EINVAL, "conv is out of place (rows are gathered from x while y is written)");
    cudaStream_t st = as_stream(stream);
    if (variant == 0) return spconv_fwd_v1(x, W, seg, pair_nbr, pair_row, n, tile_rows, residual, flags, y, stream);
#define GPC_V2(TM_, NW_, CH_, D_) return launch_spconv_v2<TM_, NW_, CH_, D_>(x, W, seg, pair_nbr, pair_row, n, residual, flags, y, st)
    if (variant == 1) {
        if (tile_rows == 256) GPC_V2(256, 4, 32, 4);
        if (tile_rows == 512) GPC_V2(512, 4, 32, 4);
        if (tile_rows == 128) GPC_V2(128, 4, 32, 4);
    } else if (variant == 2) {
        if (tile_rows == 256) GPC_V2(256, 8, 64, 3);
        if (tile_rows == 512) GPC_V2(512, 8, 64, 3);
        if (tile_rows == 1024) GPC_V2(1024, 8, 64, 3);
    } else if (variant == 3) {
        if (tile_rows == 256) GPC_V2(256, 8, 32, 4);
        if (tile_rows == 512) GPC_V2(512, 8, 32, 4);
    }
#undef GPC_V2
    gpc_set_error("unsupported conv variant %d / tile_rows %d", variant, tile_rows);
    return GPC_EINVAL;
}
