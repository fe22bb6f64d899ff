// pyramid.cu -- octree occupancy pyramid: FOG (down) and FCG (up), a-4 / a-8 / a-16.
//
// Reference: src/ai_pcc/GausPcgc/kit/nn.py:38-55 (FOG: parent = floor(c/2), occupancy = sum of
// 2^((x%2)+2(y%2)+4(z%2)) over children) and :77-98 (FCG: children 2c+(i&1,(i>>1)&1,(i>>2)&1) where
// bit i is set), followed in the reference by op.sort_CF (pcc_utils.py:96,105,307).
//
// Every level lives in HBM as keys[n] ascending == (z,y,x) order.  Going DOWN needs one radix sort of
// the parent keys (siblings are not adjacent in (z,y,x) order) and a run-length OR.  Going UP needs
// no sort at all: the (z,y,x) rank of every child follows from prefix sums over the sorted parents
// (per-(bz,by) child counts, row heads, slab heads), so children are written directly in order.
#include "common.cuh"

// ---------------------------------------------------------------- unique of sorted keys
struct HeadFlagLoad {
    const u64 *k;
    __device__ u32 operator()(i64 i) const { return (i == 0 || k[i] != k[i - 1]) ? 1u : 0u; }
};
__global__ void unique_compact_kernel(const u64 *__restrict__ keys, i64 n, const u32 *__restrict__ pos,
                                      u64 *__restrict__ out, u32 *__restrict__ n_out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == n) { *n_out = pos[n]; return; }
    if (i > n) return;
    if (i == 0 || keys[i] != keys[i - 1]) out[pos[i]] = keys[i];
}
extern "C" size_t gpc_pyramid_workspace_bytes(int64_t n) {
    const i64 m = n > 0 ? n : 1;
    return 2 * align_up((size_t)m * 8, 256) + 2 * align_up((size_t)m * 4, 256) + align_up((size_t)(m + 1) * 4, 256) +
           align_up(scan_workspace_bytes<u32>(m), 256) + gpc_sort_workspace_bytes(m) + 1024;
}
extern "C" int gpc_unique_sorted(const uint64_t *keys, int64_t n, uint64_t *out_keys, uint32_t *n_out, void *ws,
                                 size_t ws_bytes, void *stream) {
    cudaStream_t st = as_stream(stream);
    GPC_REQUIRE(ws_bytes >= gpc_pyramid_workspace_bytes(n), GPC_ENOSPC, "workspace too small");
    if (n <= 0) { GPC_CUDA_CHECK(cudaMemsetAsync(n_out, 0, 4, st)); return GPC_OK; }
    u32 *pos = (u32 *)ws;
    void *scan_ws = (char *)ws + align_up((size_t)(n + 1) * 4, 256);
    HeadFlagLoad hl{keys};
    int rc = device_exclusive_scan<u32, HeadFlagLoad>(hl, n, pos, scan_ws, st);
    if (rc) return rc;
    unique_compact_kernel<<<cdiv(n + 1, 256), 256, 0, st>>>(keys, n, pos, out_keys, n_out);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// ---------------------------------------------------------------- pyramid down (FOG)
__global__ void parent_code_kernel(const u64 *__restrict__ ck, i64 n, u64 *__restrict__ pk, u32 *__restrict__ code) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = ck[i];
    pk[i] = key_parent(k);
    code[i] = 1u << key_octant(k);
}
__global__ void parent_reduce_kernel(const u64 *__restrict__ spk, const u32 *__restrict__ scode, i64 n,
                                     const u32 *__restrict__ pos, u64 *__restrict__ out_keys, u8 *__restrict__ out_occ,
                                     u32 *__restrict__ n_out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == n) { *n_out = pos[n]; return; }
    if (i > n) return;
    const u64 k = spk[i];
    if (i != 0 && spk[i - 1] == k) return;
    u32 occ = 0;
    for (i64 j = i; j < n && j < i + 8 && spk[j] == k; ++j) occ |= scode[j];   // unique children: OR == the reference's sum
    out_keys[pos[i]] = k;
    out_occ[pos[i]] = (u8)occ;
}
extern "C" int gpc_pyramid_down(const uint64_t *child_keys, int64_t n, gpc_key_xform parent_xf, uint64_t *parent_keys,
                                uint8_t *parent_occ, uint32_t *n_parent, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = as_stream(stream);
    GPC_REQUIRE(ws_bytes >= gpc_pyramid_workspace_bytes(n), GPC_ENOSPC, "workspace too small");
    if (n <= 0) { GPC_CUDA_CHECK(cudaMemsetAsync(n_parent, 0, 4, st)); return GPC_OK; }
    char *b = (char *)ws;
    size_t off = 0;
    u64 *pk = (u64 *)(b + off); off += align_up((size_t)n * 8, 256);
    u64 *spk = (u64 *)(b + off); off += align_up((size_t)n * 8, 256);
    u32 *code = (u32 *)(b + off); off += align_up((size_t)n * 4, 256);
    u32 *scode = (u32 *)(b + off); off += align_up((size_t)n * 4, 256);
    u32 *pos = (u32 *)(b + off); off += align_up((size_t)(n + 1) * 4, 256);
    void *scan_ws = b + off; off += align_up(scan_workspace_bytes<u32>(n), 256);
    void *sort_ws = b + off;
    parent_code_kernel<<<cdiv(n, 256), 256, 0, st>>>(child_keys, n, pk, code);
    GPC_LAUNCH_CHECK();
    int rc = gpc_sort_pairs(pk, code, spk, scode, n, parent_xf, sort_ws, ws_bytes - off, stream);
    if (rc) return rc;
    HeadFlagLoad hl{spk};
    rc = device_exclusive_scan<u32, HeadFlagLoad>(hl, n, pos, scan_ws, st);
    if (rc) return rc;
    parent_reduce_kernel<<<cdiv(n + 1, 256), 256, 0, st>>>(spk, scode, n, pos, parent_keys, parent_occ, n_parent);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// ---------------------------------------------------------------- expand children in (z,y,x) order (FCG + sort_CF)
// Per parent i (sorted): c[q] = number of children with (bz,by) = (q>>1, q&1) -- 0..2 (bx = 0/1);
// rh / sh = 1 when i starts a new (z,y) row / z slab.  A = exclusive prefix sums of that 6-tuple.
struct V6 {
    u32 c[4], rh, sh;
    __host__ __device__ V6() {}
    __host__ __device__ explicit V6(int) { c[0] = c[1] = c[2] = c[3] = rh = sh = 0; }
};
__device__ __forceinline__ V6 operator+(const V6 &a, const V6 &b) {
    V6 r;
    r.c[0] = a.c[0] + b.c[0]; r.c[1] = a.c[1] + b.c[1]; r.c[2] = a.c[2] + b.c[2]; r.c[3] = a.c[3] + b.c[3];
    r.rh = a.rh + b.rh; r.sh = a.sh + b.sh;
    return r;
}
struct ExpandLoad {
    const u64 *k;
    const u8 *occ;
    __device__ V6 operator()(i64 i) const {
        V6 v;
        const u32 o = occ[i];
        v.c[0] = __popc(o & 3u); v.c[1] = __popc((o >> 2) & 3u); v.c[2] = __popc((o >> 4) & 3u); v.c[3] = __popc((o >> 6) & 3u);
        const u64 key = k[i];
        const u64 prev = i > 0 ? k[i - 1] : ~key;
        v.rh = (key >> 21) != (prev >> 21);
        v.sh = (key >> 42) != (prev >> 42);
        return v;
    }
};
__global__ void expand_starts_kernel(const u64 *__restrict__ k, i64 n, const V6 *__restrict__ A,
                                     u32 *__restrict__ row_start, u32 *__restrict__ slab_start) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { row_start[A[n].rh] = (u32)n; slab_start[A[n].sh] = (u32)n; return; }
    const u64 key = k[i];
    const u64 prev = i > 0 ? k[i - 1] : ~key;
    if ((key >> 21) != (prev >> 21)) row_start[A[i].rh] = (u32)i;
    if ((key >> 42) != (prev >> 42)) slab_start[A[i].sh] = (u32)i;
}
__global__ void expand_children_kernel(const u64 *__restrict__ k, const u8 *__restrict__ occ, i64 n, const V6 *__restrict__ A,
                                       const u32 *__restrict__ row_start, const u32 *__restrict__ slab_start,
                                       u64 *__restrict__ child_keys, u32 *__restrict__ child_parent, i64 n_child) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 key = k[i];
    const u64 prev = i > 0 ? k[i - 1] : ~key;
    const V6 a = A[i];
    const u32 rho = a.rh - (((key >> 21) != (prev >> 21)) ? 0u : 1u);
    const u32 sig = a.sh - (((key >> 42) != (prev >> 42)) ? 0u : 1u);
    const u32 rs = row_start[rho], re = row_start[rho + 1], ss = slab_start[sig], se = slab_start[sig + 1];
    const V6 ars = A[rs], are = A[re], ass = A[ss], ase = A[se];
    const u32 T = ass.c[0] + ass.c[1] + ass.c[2] + ass.c[3];
    const u32 Z0 = (ase.c[0] - ass.c[0]) + (ase.c[1] - ass.c[1]);
    const u32 o = occ[i];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int bz = q >> 1, by = q & 1;
        const u32 m = (o >> (2 * q)) & 3u;
        if (!m) continue;
        u32 base = T + (bz ? Z0 : 0u) + (ars.c[2 * bz] - ass.c[2 * bz]) + (ars.c[2 * bz + 1] - ass.c[2 * bz + 1]) +
                   (by ? (are.c[2 * bz] - ars.c[2 * bz]) : 0u) + (a.c[q] - ars.c[q]);
        if (m & 1u) {
            if (base < n_child) { child_keys[base] = key_child(key, (u32)(2 * by + 4 * bz)); child_parent[base] = (u32)i; }
            ++base;
        }
        if (m & 2u) {
            if (base < n_child) { child_keys[base] = key_child(key, (u32)(1 + 2 * by + 4 * bz)); child_parent[base] = (u32)i; }
        }
    }
}
extern "C" size_t gpc_expand_workspace_bytes(int64_t n) {
    const i64 m = n > 0 ? n : 1;
    return align_up((size_t)(m + 1) * sizeof(V6), 256) + 2 * align_up((size_t)(m + 2) * 4, 256) +
           align_up(scan_workspace_bytes<V6>(m), 256) + 1024;
}
extern "C" int gpc_expand_children(const uint64_t *parent_keys, const uint8_t *parent_occ, int64_t n, int64_t n_child,
                                   uint64_t *child_keys, uint32_t *child_parent, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n <= 0 || n_child <= 0) return GPC_OK;
    GPC_REQUIRE(ws && ws_bytes >= gpc_expand_workspace_bytes(n), GPC_ENOSPC, "workspace too small");
    char *b = (char *)ws;
    size_t off = 0;
    V6 *A = (V6 *)(b + off); off += align_up((size_t)(n + 1) * sizeof(V6), 256);
    u32 *row_start = (u32 *)(b + off); off += align_up((size_t)(n + 2) * 4, 256);
    u32 *slab_start = (u32 *)(b + off); off += align_up((size_t)(n + 2) * 4, 256);
    void *scan_ws = b + off;
    ExpandLoad el{parent_keys, parent_occ};
    int rc = device_exclusive_scan<V6, ExpandLoad>(el, n, A, scan_ws, st);
    if (rc) return rc;
    expand_starts_kernel<<<cdiv(n + 1, 256), 256, 0, st>>>(parent_keys, n, A, row_start, slab_start);
    GPC_LAUNCH_CHECK();
    expand_children_kernel<<<cdiv(n, 256), 256, 0, st>>>(parent_keys, parent_occ, n, A, row_start, slab_start,
                                                         child_keys, child_parent, n_child);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// ---------------------------------------------------------------- decode epilogue: leaves, parent-major (pcc_utils.py:375-379)
struct PopcLoad {
    const u8 *occ;
    __device__ u32 operator()(i64 i) const { return (u32)__popc((u32)occ[i]); }
};
__global__ void expand_leaves_kernel(const u64 *__restrict__ k, const u8 *__restrict__ occ, i64 n, const u32 *__restrict__ pos,
                                     float scale, float *__restrict__ xyz, i64 n_child) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 key = k[i];
    u32 o = occ[i];
    i64 p = pos[i];
    while (o) {
        const u32 oct = __ffs(o) - 1;
        o &= o - 1;
        if (p < n_child) {
            const u64 c = key_child(key, oct);
            xyz[3 * p] = (float)key_x(c) * scale; xyz[3 * p + 1] = (float)key_y(c) * scale; xyz[3 * p + 2] = (float)key_z(c) * scale;
        }
        ++p;
    }
}
extern "C" int gpc_expand_leaves_f32(const uint64_t *parent_keys, const uint8_t *parent_occ, int64_t n, int64_t n_child,
                                     float scale, float *xyz, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n <= 0 || n_child <= 0) return GPC_OK;
    GPC_REQUIRE(ws && ws_bytes >= gpc_expand_workspace_bytes(n), GPC_ENOSPC, "workspace too small");
    u32 *pos = (u32 *)ws;
    void *scan_ws = (char *)ws + align_up((size_t)(n + 1) * 4, 256);
    PopcLoad pl{parent_occ};
    int rc = device_exclusive_scan<u32, PopcLoad>(pl, n, pos, scan_ws, st);
    if (rc) return rc;
    expand_leaves_kernel<<<cdiv(n, 256), 256, 0, st>>>(parent_keys, parent_occ, n, pos, scale, xyz, n_child);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
