// spconv_fmt.cu -- the activation format of the tcgen05 sparse conv (spconv_um.cu): split rows.
#include "common.cuh"

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


