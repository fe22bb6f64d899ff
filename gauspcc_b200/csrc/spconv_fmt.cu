// spconv_fmt.cu -- operand formats of the tcgen05 sparse conv (spconv_tc.cu): split activation rows and W[k] operand images.
#include "umma.cuh"

// =====================================================================================================
// "split rows": one activation row = 128 B = 32 x bf16 hi (channels 0..31) | 32 x bf16 lo, x = hi + lo to 16 mantissa
// bits.  Same bytes per row as fp32, but a gathered row is a tensor-core operand row as it stands (no conversion in the conv),
// and the three-term product hi.Whi + hi.Wlo + lo.Whi keeps the contraction within 1.5e-4 of fp32 on the probabilities
// (DESIGN.md 5).
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
// weight images: per offset k, W[k]^T (32 co x 32 ci) as two canonical K-major (no swizzle) bf16 tiles, hi then lo, 2 KB each:
// the B operand of tcgen05.mma, copied to shared memory as it stands
// =====================================================================================================
// W [n_kernels*125][32 ci][32 co] fp32 -> Wc [n_kernels*125][2 (hi, lo)][2 KB canonical image of B[n = co][c = ci]]
__global__ void pack_weights_umma_kernel(const float *__restrict__ W, u16 *__restrict__ Wc, int n_kernels) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (i64)n_kernels * GPC_K3 * GPC_C * GPC_C) return;
    const int co = (int)(g & 31), ci = (int)((g >> 5) & 31);
    const i64 k = g >> 10;
    const float w = W[k * 1024 + ci * 32 + co];
    const float w1 = bf16_round(w), w2 = bf16_round(w - w1);
    u16 *dst = Wc + k * 2048;                       // 4 KB per offset = 2048 u16
    dst[umma_tile_off(co, ci) / 2] = (u16)(__float_as_uint(w1) >> 16);
    dst[1024 + umma_tile_off(co, ci) / 2] = (u16)(__float_as_uint(w2) >> 16);
}
extern "C" int gpc_spconv_pack_weights_umma(const float *W, int n_kernels, void *Wc, void *stream) {
    const i64 total = (i64)n_kernels * GPC_K3 * GPC_C * GPC_C;
    pack_weights_umma_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(W, (u16 *)Wc, n_kernels);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

