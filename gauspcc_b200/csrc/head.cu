// head.cu -- embeddings (a-6, a-9, a-12 context) and the fused prediction head (a-12, a-13).
//
// Reference: src/ai_pcc/GausPcgc/network_ue_4stage_conv.py:15 (prior_embedding), kit/nn.py:101-117
// (TargetEmbedding), network_ue_4stage_conv.py:65-94 (pred_head_s{i}, pred_head_s{i}_emb) and the
// probability -> CDF -> int16 step of pcc_utils.py:146-171 + kit/op.py:50-79.  In the reference the
// head is 2 cuBLAS GEMMs + softmax + cat + cumsum + clamp + mul + round + cast + add per stage; here
// it is one kernel: Linear-ReLU-Linear-softmax-cumsum-clamp-scale-round-(+k)-uint16.
#include "common.cuh"

// ---------------------------------------------------------------- embeddings (float4 lanes: 8 threads per 128 B row)
// Every producer of conv inputs can write fp32 rows, split rows (32 x bf16 hi | 32 x bf16 lo, the operand format of the tcgen05 conv,
// spconv_um.cu) or both: channels 4c .. 4c + 3 of a row are hi words 2c, 2c + 1 and lo words 16 + 2c, 17 + 2c.
__device__ __forceinline__ void store_quad(float4 *out, u32 *out_split, i64 row, int c, float4 v) {
    if (out) out[row * 8 + c] = v;
    if (out_split) {
        u32 h0, l0, h1, l1;
        split_bf16(v.x, v.y, h0, l0);
        split_bf16(v.z, v.w, h1, l1);
        *reinterpret_cast<uint2 *>(out_split + row * 32 + 2 * c) = make_uint2(h0, h1);
        *reinterpret_cast<uint2 *>(out_split + row * 32 + 16 + 2 * c) = make_uint2(l0, l1);
    }
}
__global__ void embed_rows_kernel(const u8 *__restrict__ idx, i64 n, const float4 *__restrict__ table, float4 *__restrict__ out,
                                  u32 *__restrict__ out_split) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 o = g >> 3;
    if (o >= n) return;
    const int c = (int)(g & 7);
    store_quad(out, out_split, o, c, __ldg(table + (i64)idx[o] * 8 + c));
}
extern "C" int gpc_embed_rows(const uint8_t *idx, int64_t n, const float *table, float *out, void *out_split, void *stream) {
    if (n <= 0) return GPC_OK;
    embed_rows_kernel<<<cdiv(n * 8, 256), 256, 0, as_stream(stream)>>>(idx, n, (const float4 *)table, (float4 *)out, (u32 *)out_split);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

__global__ void gather_parent_add_octant_kernel(const float4 *__restrict__ feat, const u32 *__restrict__ parent,
                                                const u64 *__restrict__ child_keys, i64 n, const float4 *__restrict__ temb,
                                                float4 *__restrict__ out, u32 *__restrict__ out_split) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 j = g >> 3;
    if (j >= n) return;
    const int c = (int)(g & 7);
    const float4 a = __ldg(feat + (i64)parent[j] * 8 + c);
    const float4 e = __ldg(temb + (i64)key_octant(child_keys[j]) * 8 + c);
    store_quad(out, out_split, j, c, make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w));
}
extern "C" int gpc_gather_parent_add_octant(const float *feat, const uint32_t *parent, const uint64_t *child_keys,
                                            int64_t n_child, const float *temb, float *out, void *out_split, void *stream) {
    if (n_child <= 0) return GPC_OK;
    gather_parent_add_octant_kernel<<<cdiv(n_child * 8, 256), 256, 0, as_stream(stream)>>>(
        (const float4 *)feat, parent, child_keys, n_child, (const float4 *)temb, (float4 *)out, (u32 *)out_split);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

__global__ void add_ctx_embed_kernel(const float4 *__restrict__ u, const u8 *__restrict__ occ, int shift,
                                     const float4 *__restrict__ emb, i64 n, float4 *__restrict__ out, u32 *__restrict__ out_split) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 o = g >> 3;
    if (o >= n) return;
    const int c = (int)(g & 7);
    const float4 a = u[o * 8 + c];
    const float4 e = __ldg(emb + (i64)(occ[o] >> shift) * 8 + c);
    store_quad(out, out_split, o, c, make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w));
}
extern "C" int gpc_add_ctx_embed(const float *u, const uint8_t *occ, int shift, const float *emb, int64_t n, float *out,
                                 void *out_split, void *stream) {
    if (n <= 0) return GPC_OK;
    add_ctx_embed_kernel<<<cdiv(n * 8, 256), 256, 0, as_stream(stream)>>>((const float4 *)u, occ, shift, (const float4 *)emb, n,
                                                                          (float4 *)out, (u32 *)out_split);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// ---------------------------------------------------------------- symbol split / merge (pcc_utils.py:112-115, 369)
__global__ void split_symbol_kernel(const u8 *__restrict__ occ, i64 n, int shift, int mask, u8 *__restrict__ sym) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sym[i] = (u8)((occ[i] >> shift) & mask);
}
__global__ void merge_symbol_kernel(u8 *__restrict__ occ, i64 n, int shift, const u8 *__restrict__ sym) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) occ[i] = (u8)(occ[i] | (sym[i] << shift));
}
extern "C" int gpc_split_symbol(const uint8_t *occ, int64_t n, int shift, int mask, uint8_t *sym, void *stream) {
    if (n <= 0) return GPC_OK;
    split_symbol_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(occ, n, shift, mask, sym);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
extern "C" int gpc_merge_symbol(uint8_t *occ, int64_t n, int shift, const uint8_t *sym, void *stream) {
    if (n <= 0) return GPC_OK;
    merge_symbol_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(occ, n, shift, sym);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// ---------------------------------------------------------------- fused head
constexpr int HD_ROWS = 128;     // rows per CTA == threads per CTA (one thread per row)

template <int A>
__global__ void __launch_bounds__(HD_ROWS) head_cdf_kernel(const float *__restrict__ f, i64 n, const float *__restrict__ W1,
                                                           const float *__restrict__ b1, const float *__restrict__ W2,
                                                           const float *__restrict__ b2, u16 *__restrict__ cdf,
                                                           float *__restrict__ prob, const u8 *__restrict__ occ, int shift,
                                                           u32 *__restrict__ lohi) {
    constexpr int Lp = A + 1;
    __shared__ float fs[HD_ROWS][GPC_C + 1];
    __shared__ __align__(16) float w1s[GPC_C][GPC_C];
    __shared__ __align__(16) float w2s[A][GPC_C];
    __shared__ float b1s[GPC_C], b2s[A];
    __shared__ __align__(16) u16 cs[HD_ROWS * Lp + 8];
    const int tid = threadIdx.x;
    const i64 r0 = (i64)blockIdx.x * HD_ROWS;
    const int rows = (int)min((i64)HD_ROWS, n - r0);
    for (int i = tid; i < GPC_C * GPC_C; i += HD_ROWS) (&w1s[0][0])[i] = __ldg(W1 + i);
    for (int i = tid; i < A * GPC_C; i += HD_ROWS) (&w2s[0][0])[i] = __ldg(W2 + i);
    if (tid < GPC_C) b1s[tid] = __ldg(b1 + tid);
    if (tid < A) b2s[tid] = __ldg(b2 + tid);
    for (int i = tid; i < rows * GPC_C; i += HD_ROWS) fs[i >> 5][i & 31] = f[r0 * GPC_C + i];     // coalesced
    __syncthreads();
    if (tid < rows) {
        float x[GPC_C], h[GPC_C];
#pragma unroll
        for (int i = 0; i < GPC_C; ++i) x[i] = fs[tid][i];
#pragma unroll
        for (int j = 0; j < GPC_C; ++j) {
            float s = b1s[j];
#pragma unroll
            for (int i4 = 0; i4 < GPC_C / 4; ++i4) {
                const float4 wv = reinterpret_cast<const float4 *>(&w1s[j][0])[i4];
                s = fmaf(x[4 * i4 + 0], wv.x, s); s = fmaf(x[4 * i4 + 1], wv.y, s);
                s = fmaf(x[4 * i4 + 2], wv.z, s); s = fmaf(x[4 * i4 + 3], wv.w, s);
            }
            h[j] = fmaxf(s, 0.f);
        }
        float lg[A];
        float mx = -INFINITY;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            float s = b2s[a];
#pragma unroll
            for (int i4 = 0; i4 < GPC_C / 4; ++i4) {
                const float4 wv = reinterpret_cast<const float4 *>(&w2s[a][0])[i4];
                s = fmaf(h[4 * i4 + 0], wv.x, s); s = fmaf(h[4 * i4 + 1], wv.y, s);
                s = fmaf(h[4 * i4 + 2], wv.z, s); s = fmaf(h[4 * i4 + 3], wv.w, s);
            }
            lg[a] = s;
            mx = fmaxf(mx, s);
        }
        float sum = 0.f;
#pragma unroll
        for (int a = 0; a < A; ++a) { lg[a] = expf(lg[a] - mx); sum += lg[a]; }
        const float scale = 65536.0f - (float)A;       // 2^16 - (Lp - 1), kit/op.py:67-70
        float c = 0.f;
        cs[tid * Lp] = 0;
        const int my_sym = lohi ? (int)((occ[r0 + tid] >> shift) & (A - 1)) : -1;      // a-11 symbol split fused
        u32 c_lo = 0, c_hi = 0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float p = lg[a] / sum;
            if (prob) prob[(r0 + tid) * A + a] = p;
            c += p;                                     // sequential fp32 cumsum
            const float cc = fminf(fmaxf(c, 0.f), 1.f);
            const u16 q = (u16)((u32)__float2int_rn(cc * scale) + (u32)(a + 1));          // int16 wrap == uint16 bits
            cs[tid * Lp + a + 1] = q;
            if (a + 1 == my_sym) c_lo = q;
            if (a == my_sym && a != A - 1) c_hi = q;    // top symbol: c_hi stays 0 == 0x10000
        }
        if (lohi) lohi[r0 + tid] = c_lo | (c_hi << 16);
    }
    __syncthreads();
    if (cdf) {
        u16 *dst = cdf + r0 * Lp;
        for (int i = tid; i < rows * Lp; i += HD_ROWS) dst[i] = cs[i];
    }
}

static int head_launch(const float *f, int64_t n, const float *W1, const float *b1, const float *W2, const float *b2, int A,
                       uint16_t *cdf, float *prob, const uint8_t *occ, int shift, uint32_t *lohi, void *stream) {
    if (n <= 0) return GPC_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned grid = cdiv(n, HD_ROWS);
    switch (A) {
        case 2: head_cdf_kernel<2><<<grid, HD_ROWS, 0, st>>>(f, n, W1, b1, W2, b2, cdf, prob, occ, shift, lohi); break;
        case 4: head_cdf_kernel<4><<<grid, HD_ROWS, 0, st>>>(f, n, W1, b1, W2, b2, cdf, prob, occ, shift, lohi); break;
        case 16: head_cdf_kernel<16><<<grid, HD_ROWS, 0, st>>>(f, n, W1, b1, W2, b2, cdf, prob, occ, shift, lohi); break;
        default: gpc_set_error("unsupported alphabet %d (2, 4, 16)", A); return GPC_EINVAL;
    }
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
extern "C" int gpc_head_cdf(const float *f, int64_t n, const float *W1, const float *b1, const float *W2, const float *b2,
                            int A, uint16_t *cdf, float *prob, void *stream) {
    return head_launch(f, n, W1, b1, W2, b2, A, cdf, prob, nullptr, 0, nullptr, stream);
}
// encoder side: the symbol of this stage is (occ >> shift) & (A-1); writes lohi[o] = c_low | c_high << 16 (c_high == 0
// means 0x10000) for gpc_ac_encode_lohi_h; cdf / prob may be NULL
extern "C" int gpc_head_cdf_sym(const float *f, int64_t n, const float *W1, const float *b1, const float *W2, const float *b2,
                                int A, const uint8_t *occ, int shift, uint32_t *lohi, uint16_t *cdf, float *prob, void *stream) {
    GPC_REQUIRE(occ && lohi, GPC_EINVAL, "occ and lohi are required");
    return head_launch(f, n, W1, b1, W2, b2, A, cdf, prob, occ, shift, lohi, stream);
}
