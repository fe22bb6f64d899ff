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

extern "C" int gpc_spconv_fwd(const float *x, const float *W, const uint32_t *seg, const uint32_t *pair_nbr,
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
