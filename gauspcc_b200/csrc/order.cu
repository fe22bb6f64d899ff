// order.cu -- voxel keys, stable LSD radix sort and calculate_morton_order (a-1, a-2, a-5).
//
// Reference: src/gs_compress/HAC/utils/pcc_utils.py:12-22 (calculate_morton_order: x -= min;
// key = x + y*M + z*M^2; argsort) and src/ai_pcc/GausPcgc/kit/op.py:17-30 (sort_CF: four stable
// torch.sort passes => lexicographic (z,y,x)).  Both orders are "ascending (z,y,x)"; here one
// 8-bit-digit LSD radix sort over a compacted key does it: ceil((bx+by+bz)/8) passes, each
// pass = per-tile digit histogram -> device scan -> stable scatter (warp match_any ranking).
#include "common.cuh"

// ---------------------------------------------------------------- pack / unpack canonical keys
template <typename T>
__global__ void pack_keys_kernel(const T *__restrict__ xyz, i64 n, u64 *__restrict__ keys, i32 *status) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    i32 c[3];
    i32 flag = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T v = xyz[3 * i + a];
        float r;
        if (sizeof(T) == 4 && ((T)0.5 != (T)0)) {            // float input
            r = rintf((float)v);
            if (r != (float)v) flag |= 1;
        } else {
            r = (float)v;                                    // int32: exact below 2^24, range-checked next
            if ((i64)v > GPC_COORD_MAX || (i64)v < -GPC_COORD_MAX) flag |= 2;
        }
        if (!(fabsf(r) <= (float)GPC_COORD_MAX)) { flag |= 2; r = 0.f; }
        c[a] = (i32)r;
    }
    keys[i] = key_pack(c[0], c[1], c[2]);
    if (flag) atomicOr(status, flag);
}

__global__ void unpack_keys_i32_kernel(const u64 *__restrict__ keys, i64 n, i32 *__restrict__ xyz) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = keys[i];
    xyz[3 * i] = key_x(k); xyz[3 * i + 1] = key_y(k); xyz[3 * i + 2] = key_z(k);
}

__global__ void unpack_keys_f32_kernel(const u64 *__restrict__ keys, i64 n, float scale, float *__restrict__ xyz) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = keys[i];
    xyz[3 * i] = (float)key_x(k) * scale; xyz[3 * i + 1] = (float)key_y(k) * scale; xyz[3 * i + 2] = (float)key_z(k) * scale;
}

extern "C" int gpc_pack_keys_f32(const float *xyz, int64_t n, uint64_t *keys, int32_t *status, void *stream) {
    if (n <= 0) return GPC_OK;
    pack_keys_kernel<float><<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(xyz, n, keys, status);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
extern "C" int gpc_pack_keys_i32(const int32_t *xyz, int64_t n, uint64_t *keys, int32_t *status, void *stream) {
    if (n <= 0) return GPC_OK;
    pack_keys_kernel<i32><<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(xyz, n, keys, status);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
extern "C" int gpc_unpack_keys_i32(const uint64_t *keys, int64_t n, int32_t *xyz, void *stream) {
    if (n <= 0) return GPC_OK;
    unpack_keys_i32_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(keys, n, xyz);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
extern "C" int gpc_unpack_keys_f32(const uint64_t *keys, int64_t n, float scale, float *xyz, void *stream) {
    if (n <= 0) return GPC_OK;
    unpack_keys_f32_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(keys, n, scale, xyz);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// ---------------------------------------------------------------- field min/max of canonical keys
__global__ void minmax_init_kernel(u32 *mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) mm[threadIdx.x] = 0u;
}
__global__ void key_minmax_kernel(const u64 *__restrict__ keys, i64 n, u32 *mm) {
    u32 mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0, 0, 0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        u64 k = keys[i];
        u32 f[3] = {(u32)(k & GPC_FIELD_MASK), (u32)((k >> 21) & GPC_FIELD_MASK), (u32)((k >> 42) & GPC_FIELD_MASK)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { mn[a] = min(mn[a], f[a]); mx[a] = max(mx[a], f[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(0xFFFFFFFFu, mn[a]);
        mx[a] = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&mm[a], mn[a]); atomicMax(&mm[3 + a], mx[a]); }
    }
}
extern "C" int gpc_key_minmax(const uint64_t *keys, int64_t n, uint32_t *minmax, void *stream) {
    cudaStream_t st = as_stream(stream);
    minmax_init_kernel<<<1, 32, 0, st>>>(minmax);
    GPC_LAUNCH_CHECK();
    if (n > 0) {
        unsigned grid = min(cdiv(n, 256), 148u * 8u);
        key_minmax_kernel<<<grid, 256, 0, st>>>(keys, n, minmax);
        GPC_LAUNCH_CHECK();
    }
    return GPC_OK;
}

static inline u32 bits_for(u32 extent) {   // bits needed to hold values 0..extent
    u32 b = 0;
    while (b < 32 && (extent >> b) != 0) ++b;
    return b;
}
extern "C" int gpc_make_xform_h(const uint32_t *mm, gpc_key_xform *xf) {
    GPC_REQUIRE(mm && xf, GPC_EINVAL, "null argument");
    if (mm[0] > mm[3]) {    // empty set
        xf->minx = xf->miny = xf->minz = 0; xf->sy = xf->sz = 0; xf->total_bits = 0;
        return GPC_OK;
    }
    u32 bx = bits_for(mm[3] - mm[0]), by = bits_for(mm[4] - mm[1]), bz = bits_for(mm[5] - mm[2]);
    xf->minx = mm[0]; xf->miny = mm[1]; xf->minz = mm[2];
    xf->sy = bx; xf->sz = bx + by; xf->total_bits = bx + by + bz;
    return GPC_OK;
}

// ---------------------------------------------------------------- radix sort ("onesweep": one read + one write of the records per pass)
// LSD, 8-bit digits over the compacted key.  ONE up-front kernel reads the keys once and builds the global digit histogram of every
// pass; each pass is then a single kernel: a tile ranks its 4096 records (warp match_any, no atomics on the order), publishes its
// per-digit counts, and finds where its digits start in the output by DECOUPLED LOOK-BACK over the tiles before it (Merrill &
// Garland's single-pass scan, as used by Adinets & Merrill's onesweep sort) instead of a separate per-tile histogram kernel plus a
// device-wide scan per pass.  Traffic per record: 8 B once, then 12 B in + 12 B out per pass (the 3-kernel version read the keys twice
// per pass).  Tiles take their index from an atomic counter, so a tile only ever waits for tiles that are already running.
constexpr int RS_THREADS = 512;       // 2 CTAs of 16 warps per SM (64 registers per thread): the pass is latency-bound, occupancy is what it needs
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 8;
constexpr int RS_TILE = RS_THREADS * RS_IPT;
constexpr int RS_MAX_PASSES = 8;
constexpr u32 RS_FLAG_AGG = 1u << 30, RS_FLAG_INCL = 2u << 30, RS_FLAG_MASK = 3u << 30, RS_VAL_MASK = ~RS_FLAG_MASK;

template <bool XF>
__device__ __forceinline__ u32 rs_digit(u64 key, const gpc_key_xform &xf, int shift) {
    u64 c = XF ? key_compact(key, xf) : key;
    return (u32)(c >> shift) & 0xFFu;
}

// ghist[p][d] = number of keys whose p-th digit is d (all passes from one read of the keys)
template <bool XF>
__global__ void __launch_bounds__(256) rs_global_hist_kernel(const u64 *__restrict__ keys, i64 n, gpc_key_xform xf, int passes,
                                                             u32 *__restrict__ ghist) {
    __shared__ u32 h[RS_MAX_PASSES][256];
    for (int i = threadIdx.x; i < RS_MAX_PASSES * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (i64 i0 = (i64)blockIdx.x * 256 + (threadIdx.x & ~31); i0 < n; i0 += (i64)gridDim.x * 256) {      // warp-uniform trip count
        const i64 i = i0 + lane;
        const bool valid = i < n;
        const u64 c = valid ? (XF ? key_compact(keys[i], xf) : keys[i]) : 0ull;
        const u32 nvalid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
        for (int p = 0; p < passes; ++p) {
            const u32 d = (u32)(c >> (8 * p)) & 0xFFu;
            const u32 d0 = __shfl_sync(0xFFFFFFFFu, d, 0);
            // sorted or clustered inputs: the high digits of a warp's 32 keys are usually all equal -> one add instead of a 32-way conflict
            if (__all_sync(0xFFFFFFFFu, !valid || d == d0)) { if (lane == 0) atomicAdd(&h[p][d0], nvalid); }
            else if (valid) atomicAdd(&h[p][d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += 256) {
        const u32 v = (&h[0][0])[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}
// in place: ghist[p][d] -> number of keys whose p-th digit is < d
__global__ void __launch_bounds__(256) rs_scan_hist_kernel(u32 *__restrict__ ghist) {
    __shared__ u32 wsum[8];
    u32 *g = ghist + blockIdx.x * 256;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 v = g[tid];
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 off = 0;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    g[tid] = off + incl - v;
}

// Stable scatter of one tile.  The tile is first re-ordered by digit in shared memory, so that the global writes are contiguous runs
// per digit (a 4096-key tile over 256 digits = 128-byte runs) instead of one 8 + 4 byte record per thread at a random address.
struct RsSmem {
    u64 keys[RS_TILE];
    u32 vals[RS_TILE];
    u32 whist[RS_WARPS][256];
    u32 dstart[256];
    u32 gbase[256];
    u32 tile;
};

template <bool XF, bool IOTA>
__global__ void __launch_bounds__(RS_THREADS, 2) rs_onesweep_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
                                                                 u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, i64 n,
                                                                 gpc_key_xform xf, int shift, const u32 *__restrict__ ghist_scanned,
                                                                 u32 *__restrict__ status /* [tiles][256], zeroed */,
                                                                 u32 *__restrict__ tile_counter) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem &s = *reinterpret_cast<RsSmem *>(rs_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s.tile = atomicAdd(tile_counter, 1u);          // tiles start in index order: look-back never waits on a tile that has not started
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&s.whist[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s.tile;
    const i64 tbase = (i64)tile * RS_TILE;
    const i64 wbase = tbase + (i64)warp * (32 * RS_IPT);
    u64 k[RS_IPT];
    u32 v[RS_IPT], rk[RS_IPT];
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
        i64 idx = wbase + r * 32 + lane;
        bool valid = idx < n;
        k[r] = valid ? keys_in[idx] : 0ull;
        v[r] = IOTA ? (u32)idx : (valid ? vals_in[idx] : 0u);
    }
    const u32 lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
        bool valid = (wbase + r * 32 + lane) < n;
        u32 d = valid ? rs_digit<XF>(k[r], xf, shift) : (0x100u | (u32)lane);
        u32 peers = __match_any_sync(0xFFFFFFFFu, d);
        u32 lower = __popc(peers & lt_mask);
        u32 prev = valid ? s.whist[warp][d] : 0u;
        __syncwarp();
        if (valid && lower == 0) s.whist[warp][d] = prev + __popc(peers);
        __syncwarp();
        rk[r] = prev + lower;
    }
    __syncthreads();
    // thread d < 256: count of digit d in this tile (exclusive scan across warps), publish it, look back, tile-local start of each digit
    const int d = tid;
    u32 off = 0, before = 0, incl = 0;
    if (tid < 256) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { u32 t = s.whist[w][d]; s.whist[w][d] = off; off += t; }
        volatile u32 *st = status + (size_t)tile * 256 + d;
        if (tile > 0) *st = RS_FLAG_AGG | off;
        for (i64 t = (i64)tile - 1; t >= 0; --t) {   // records of digit d in the tiles before this one
            volatile const u32 *ps = status + (size_t)t * 256 + d;
            u32 w = *ps, spins = 0;
            while ((w & RS_FLAG_MASK) == 0) {        // bounded: a broken protocol must fault the launch, never hang the GPU
                if (++spins > (1u << 26)) __trap();
                w = *ps;
            }
            before += w & RS_VAL_MASK;
            if ((w & RS_FLAG_MASK) == RS_FLAG_INCL) break;
        }
        *st = RS_FLAG_INCL | (before + off);
        incl = off;                                  // block-wide exclusive scan over the 256 digit counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        s.gbase[d] = incl;                           // temporarily: inclusive within the warp's 32 digits
    }
    __syncthreads();
    u32 start = 0;
    if (tid < 256) {
        u32 warp_off = 0;
        for (int w = 0; w < warp; ++w) warp_off += s.gbase[w * 32 + 31];
        start = warp_off + incl - off;
    }
    __syncthreads();
    if (tid < 256) {
        s.dstart[d] = start;
        s.gbase[d] = ghist_scanned[d] + before - start;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
        if ((wbase + r * 32 + lane) < n) {
            const u32 d = rs_digit<XF>(k[r], xf, shift);
            const u32 lp = s.dstart[d] + s.whist[warp][d] + rk[r];
            s.keys[lp] = k[r];
            s.vals[lp] = v[r];
        }
    }
    __syncthreads();
    const int count = (int)min((i64)RS_TILE, n - tbase);
    for (int p = tid; p < count; p += RS_THREADS) {
        const u64 key = s.keys[p];
        const u32 pos = s.gbase[rs_digit<XF>(key, xf, shift)] + (u32)p;
        keys_out[pos] = key;
        vals_out[pos] = s.vals[p];
    }
}

struct SortWs {
    u32 *ghist;          // [RS_MAX_PASSES][256]
    u32 *counters;       // [RS_MAX_PASSES]
    u32 *status;         // [passes][tiles][256]
    u64 *alt_keys;
    u32 *alt_vals;
    size_t zero_bytes;   // ghist + counters + status are contiguous and zeroed by one memset
    size_t total;
};
static SortWs sort_ws_layout(void *ws, i64 n) {
    SortWs L;
    const i64 nblocks = (n + RS_TILE - 1) / RS_TILE;
    size_t off = 0;
    char *b = (char *)ws;
    L.ghist = (u32 *)(b + off); off += RS_MAX_PASSES * 256 * 4;
    L.counters = (u32 *)(b + off); off += 256;
    L.status = (u32 *)(b + off); off += align_up((size_t)RS_MAX_PASSES * (size_t)(nblocks > 0 ? nblocks : 1) * 256 * 4, 256);
    L.zero_bytes = off;
    L.alt_keys = (u64 *)(b + off); off += align_up((size_t)(n > 0 ? n : 1) * 8, 256);
    L.alt_vals = (u32 *)(b + off); off += align_up((size_t)(n > 0 ? n : 1) * 4, 256);
    L.total = off;
    return L;
}
extern "C" size_t gpc_sort_workspace_bytes(int64_t n) { return sort_ws_layout(nullptr, n).total; }

template <bool XF>
static int sort_pairs_impl(const u64 *keys_in, const u32 *vals_in, u64 *keys_out, u32 *vals_out, i64 n,
                           gpc_key_xform xf, int total_bits, void *ws, size_t ws_bytes, cudaStream_t st) {
    GPC_REQUIRE(vals_out != nullptr && keys_out != nullptr, GPC_EINVAL, "keys_out and vals_out are required");
    GPC_REQUIRE(keys_out != keys_in, GPC_EINVAL, "sort is out of place");
    GPC_REQUIRE(n < (1ll << 30), GPC_EINVAL, "n must be < 2^30");
    if (n <= 0) return GPC_OK;
    SortWs L = sort_ws_layout(ws, n);
    GPC_REQUIRE(ws && ws_bytes >= L.total, GPC_ENOSPC, "sort workspace too small");
    static bool configured = false;
    if (!configured) {
        GPC_CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<XF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        GPC_CUDA_CHECK(cudaFuncSetAttribute(rs_onesweep_kernel<XF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
        configured = true;
    }
    const int P = (total_bits + 7) / 8;
    GPC_REQUIRE(P <= RS_MAX_PASSES, GPC_EINVAL, "key wider than 64 bits");
    const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
    const int passes = P == 0 ? 1 : P;     // all keys equal: one pass over a zero digit = stable identity
    GPC_CUDA_CHECK(cudaMemsetAsync(L.ghist, 0, (size_t)((char *)L.status - (char *)L.ghist) + (size_t)passes * nblocks * 256 * 4, st));
    rs_global_hist_kernel<XF><<<min(cdiv(n, 256 * 16), 148u * 8u), 256, 0, st>>>(keys_in, n, xf, passes, L.ghist);
    GPC_LAUNCH_CHECK();
    rs_scan_hist_kernel<<<passes, 256, 0, st>>>(L.ghist);
    GPC_LAUNCH_CHECK();
    const u64 *src_k = keys_in;
    const u32 *src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        // destinations alternate so that the LAST pass lands in (keys_out, vals_out); the input is never written
        const bool to_out = ((passes - p) & 1) != 0;
        u64 *dst_k = to_out ? keys_out : L.alt_keys;
        u32 *dst_v = to_out ? vals_out : L.alt_vals;
        const int shift = 8 * p;
        u32 *status = L.status + (size_t)p * nblocks * 256;
        if (p == 0 && src_v == nullptr)
            rs_onesweep_kernel<XF, true><<<nblocks, RS_THREADS, sizeof(RsSmem), st>>>(src_k, nullptr, dst_k, dst_v, n, xf, shift, L.ghist + p * 256,
                                                                                     status, L.counters + p);
        else
            rs_onesweep_kernel<XF, false><<<nblocks, RS_THREADS, sizeof(RsSmem), st>>>(src_k, src_v, dst_k, dst_v, n, xf, shift, L.ghist + p * 256,
                                                                                      status, L.counters + p);
        GPC_LAUNCH_CHECK();
        src_k = dst_k; src_v = dst_v;
    }
    return GPC_OK;
}

extern "C" int gpc_sort_pairs(uint64_t *keys_in, uint32_t *vals_in, uint64_t *keys_out, uint32_t *vals_out,
                              int64_t n, gpc_key_xform xf, void *ws, size_t ws_bytes, void *stream) {
    return sort_pairs_impl<true>(keys_in, vals_in, keys_out, vals_out, n, xf, (int)xf.total_bits, ws, ws_bytes, as_stream(stream));
}

// ---------------------------------------------------------------- calculate_morton_order
// per-axis min/max of the raw input (double holds float32 and int32 exactly)
template <typename T>
__global__ void axis_minmax_partial_kernel(const T *__restrict__ xyz, i64 n, double *__restrict__ partial) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double v = (double)xyz[3 * i + a];
            mn[a] = fmin(mn[a], v); mx[a] = fmax(mx[a], v);
        }
    }
    __shared__ double s[6][256];
#pragma unroll
    for (int a = 0; a < 3; ++a) { s[a][threadIdx.x] = mn[a]; s[3 + a][threadIdx.x] = mx[a]; }
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                s[a][threadIdx.x] = fmin(s[a][threadIdx.x], s[a][threadIdx.x + off]);
                s[3 + a][threadIdx.x] = fmax(s[3 + a][threadIdx.x], s[3 + a][threadIdx.x + off]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 6) partial[(i64)blockIdx.x * 6 + threadIdx.x] = s[threadIdx.x][0];
}
__global__ void axis_minmax_final_kernel(const double *__restrict__ partial, int nparts, double *__restrict__ out) {
    const int a = threadIdx.x;
    if (a >= 6) return;
    double v = partial[a];
    for (int p = 1; p < nparts; ++p) v = a < 3 ? fmin(v, partial[p * 6 + a]) : fmax(v, partial[p * 6 + a]);
    out[a] = v;
}

// key = x' + y'*2^bx + z'*2^(bx+by) with x' = trunc(x - min_x): the reference's
// `x - torch.min(x)` in the tensor dtype followed by `.astype(np.int64)` (pcc_utils.py:18-19);
// any M > extent gives the same ranking as the reference's M = max+1.
template <typename T>
__global__ void pack_compact_kernel(const T *__restrict__ xyz, i64 n, const double *__restrict__ mm, int sy, int sz,
                                    u64 *__restrict__ keys) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 f[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T v = xyz[3 * i + a];
        T d = v - (T)mm[a];
        f[a] = (u64)(i64)d;                 // float: truncation toward zero, like astype(int64)
    }
    keys[i] = (f[2] << sz) | (f[1] << sy) | f[0];
}
__global__ void widen_idx_kernel(const u32 *__restrict__ in, i64 n, i64 *__restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (i64)in[i];
}

extern "C" size_t gpc_lexorder_workspace_bytes(int64_t n) {
    const i64 m = n > 0 ? n : 1;
    return gpc_sort_workspace_bytes(n) + 2 * align_up((size_t)m * 8, 256) + align_up((size_t)m * 4, 256) + 4096 * 6 * 8 + 1024;
}

extern "C" int gpc_lexorder_zyx(const void *xyz, int is_f32, int64_t n, int64_t *out_idx, void *ws, size_t ws_bytes,
                                void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n <= 0) return GPC_OK;
    GPC_REQUIRE(xyz && out_idx && ws, GPC_EINVAL, "null argument");
    GPC_REQUIRE(ws_bytes >= gpc_lexorder_workspace_bytes(n), GPC_ENOSPC, "lexorder workspace too small");
    GPC_REQUIRE(n < (1ll << 32), GPC_EINVAL, "n must be < 2^32");
    char *b = (char *)ws;
    size_t off = 0;
    double *partial = (double *)(b + off); off += 4096 * 6 * 8;
    double *mm_d = (double *)(b + off); off += 1024;
    u64 *keys = (u64 *)(b + off); off += align_up((size_t)n * 8, 256);
    u64 *keys_sorted = (u64 *)(b + off); off += align_up((size_t)n * 8, 256);
    u32 *idx32 = (u32 *)(b + off); off += align_up((size_t)n * 4, 256);
    void *sort_ws = b + off;
    const int nparts = (int)min((i64)4096, (n + 255) / 256);
    if (is_f32) axis_minmax_partial_kernel<float><<<nparts, 256, 0, st>>>((const float *)xyz, n, partial);
    else        axis_minmax_partial_kernel<i32><<<nparts, 256, 0, st>>>((const i32 *)xyz, n, partial);
    GPC_LAUNCH_CHECK();
    axis_minmax_final_kernel<<<1, 32, 0, st>>>(partial, nparts, mm_d);
    GPC_LAUNCH_CHECK();
    double mm[6];
    GPC_CUDA_CHECK(cudaMemcpyAsync(mm, mm_d, sizeof(mm), cudaMemcpyDeviceToHost, st));
    GPC_CUDA_CHECK(cudaStreamSynchronize(st));
    u32 bits[3];
    for (int a = 0; a < 3; ++a) {
        double ext = mm[3 + a] - mm[a];
        GPC_REQUIRE(ext == ext && ext < 2097152.0, GPC_ERANGE, "extent per axis must be < 2^21 (the reference's int64 key overflows beyond it)");
        bits[a] = bits_for((u32)ext);
    }
    const int sy = (int)bits[0], sz = (int)(bits[0] + bits[1]), total = (int)(bits[0] + bits[1] + bits[2]);
    if (is_f32) pack_compact_kernel<float><<<cdiv(n, 256), 256, 0, st>>>((const float *)xyz, n, mm_d, sy, sz, keys);
    else        pack_compact_kernel<i32><<<cdiv(n, 256), 256, 0, st>>>((const i32 *)xyz, n, mm_d, sy, sz, keys);
    GPC_LAUNCH_CHECK();
    gpc_key_xform xf = {0, 0, 0, 0, 0, (u32)total};
    int rc = sort_pairs_impl<false>(keys, nullptr, keys_sorted, idx32, n, xf, total, sort_ws, ws_bytes - off, st);
    if (rc) return rc;
    widen_idx_kernel<<<cdiv(n, 256), 256, 0, st>>>(idx32, n, out_idx);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
