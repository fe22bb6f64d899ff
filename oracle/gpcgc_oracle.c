/*
 * gpcgc_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * A plain-C, CPU, fp32 restatement of the GausPcgc anchor-geometry codec that
 * /root/reference/src/gs_compress/HAC/utils/pcc_utils.py drives.  Nothing in the
 * product package (gauspcc_b200/) may import, link or execute this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / CPU baseline.
 *
 * Pinning status: the reference ships no tests or golden vectors and its two
 * arithmetic dependencies (torchsparse==2.1.0, torchac==0.9.3, requirements.txt:6,8)
 * are not vendored.  This oracle is pinned against
 *   - the reference's own kit/op.py and pcc_utils.calculate_morton_order
 *     (imported unchanged, tests/golden/make_golden.py), and
 *   - a run of the reference's own compress_point_cloud/decompress_point_cloud,
 *     Network, FOG, FCG, TargetEmbedding code over pure-torch stand-ins for the
 *     two missing third-party packages (tests/golden/make_golden.py).
 * The sparse-conv offset ordering and the range coder are restated from the
 * published algorithms (see DESIGN.md "oracle").
 *
 * All functions are extern "C"-style, operate on caller-allocated buffers and
 * return 0 on success.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef int32_t i32;
typedef int64_t i64;
typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

/* ------------------------------------------------------------------------- */
/* calculate_morton_order -- pcc_utils.py:12-22                               */
/*   x -= min over rows (per axis); M = x.max()+1 (global);                   */
/*   key = x + y*M + z*M^2 (int64); argsort(key).  Ties broken by index       */
/*   (the reference's quicksort leaves ties unspecified).                     */
/* ------------------------------------------------------------------------- */
typedef struct { i64 key; i64 idx; } keyidx_t;

static int cmp_keyidx(const void *a, const void *b) {
    const keyidx_t *p = (const keyidx_t *)a, *q = (const keyidx_t *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    if (p->idx != q->idx) return p->idx < q->idx ? -1 : 1;
    return 0;
}

int orc_lexorder(const i64 *xyz, i64 n, i64 *out_idx) {
    if (n == 0) return 0;
    i64 mn[3] = { xyz[0], xyz[1], xyz[2] };
    for (i64 i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
            if (xyz[3 * i + a] < mn[a]) mn[a] = xyz[3 * i + a];
    i64 mx = 0;
    for (i64 i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            i64 v = xyz[3 * i + a] - mn[a];
            if (v > mx) mx = v;
        }
    i64 M = mx + 1;
    keyidx_t *ki = (keyidx_t *)malloc(sizeof(keyidx_t) * (size_t)n);
    if (!ki) return -1;
    for (i64 i = 0; i < n; ++i) {
        i64 x = xyz[3 * i] - mn[0], y = xyz[3 * i + 1] - mn[1], z = xyz[3 * i + 2] - mn[2];
        ki[i].key = x + y * M + z * M * M;      /* np.power(M, arange(3)) dot */
        ki[i].idx = i;
    }
    qsort(ki, (size_t)n, sizeof(keyidx_t), cmp_keyidx);
    for (i64 i = 0; i < n; ++i) out_idx[i] = ki[i].idx;
    free(ki);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* sort_CF row order -- kit/op.py:17-30: sort by x, then stably by y, z, b    */
/* => lexicographic (z, y, x) (batch is always 0 on this path).               */
/* ------------------------------------------------------------------------- */
typedef struct { i32 x, y, z; i64 idx; } cidx_t;

static int cmp_zyx(const void *a, const void *b) {
    const cidx_t *p = (const cidx_t *)a, *q = (const cidx_t *)b;
    if (p->z != q->z) return p->z < q->z ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->idx != q->idx) return p->idx < q->idx ? -1 : 1;
    return 0;
}

/* permutation that sorts coords[n,3] into (z,y,x) order, stable */
int orc_sort_zyx(const i32 *coords, i64 n, i64 *perm) {
    cidx_t *c = (cidx_t *)malloc(sizeof(cidx_t) * (size_t)(n ? n : 1));
    if (!c) return -1;
    for (i64 i = 0; i < n; ++i) {
        c[i].x = coords[3 * i]; c[i].y = coords[3 * i + 1]; c[i].z = coords[3 * i + 2]; c[i].idx = i;
    }
    qsort(c, (size_t)n, sizeof(cidx_t), cmp_zyx);
    for (i64 i = 0; i < n; ++i) perm[i] = c[i].idx;
    free(c);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* FOG -- kit/nn.py:38-55 (pos(): code = 2^((x%2)+2(y%2)+4(z%2)), Python %    */
/* => non-negative remainder; all-ones k=2,s=2 conv => parent = floor(c/2),   */
/* occupancy = sum of child codes).  Input voxels must be unique (duplicates   */
/* are merged by the caller, see oracle.py).  Output sorted (z,y,x).           */
/* ------------------------------------------------------------------------- */
int orc_fog(const i32 *coords, i64 n, i32 *pcoords, u8 *pocc, i64 *n_out) {
    cidx_t *c = (cidx_t *)malloc(sizeof(cidx_t) * (size_t)(n ? n : 1));
    if (!c) return -1;
    for (i64 i = 0; i < n; ++i) {
        i32 x = coords[3 * i], y = coords[3 * i + 1], z = coords[3 * i + 2];
        c[i].x = x >> 1; c[i].y = y >> 1; c[i].z = z >> 1;     /* floor(c/2), arithmetic shift */
        c[i].idx = (x & 1) + 2 * (y & 1) + 4 * (z & 1);        /* exponent of the child code */
    }
    qsort(c, (size_t)n, sizeof(cidx_t), cmp_zyx);
    i64 m = 0;
    for (i64 i = 0; i < n; ++i) {
        if (i == 0 || c[i].x != c[i - 1].x || c[i].y != c[i - 1].y || c[i].z != c[i - 1].z) {
            pcoords[3 * m] = c[i].x; pcoords[3 * m + 1] = c[i].y; pcoords[3 * m + 2] = c[i].z;
            pocc[m] = 0;
            ++m;
        }
        pocc[m - 1] = (u8)(pocc[m - 1] + (1u << c[i].idx));
    }
    *n_out = m;
    free(c);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* FCG -- kit/nn.py:64-98: for each parent, children 2c + (i&1,(i>>1)&1,       */
/* (i>>2)&1) for i = 0..7 kept where bit i of the occupancy is set            */
/* (parent-major, i ascending), then sort_CF (pcc_utils.py:105,307).          */
/* Returns children in (z,y,x) order and the parent row of each child.        */
/* ------------------------------------------------------------------------- */
int orc_fcg(const i32 *pcoords, const u8 *pocc, i64 n, i32 *ccoords, i64 *cparent, i64 *n_out) {
    i64 m = 0;
    for (i64 i = 0; i < n; ++i) m += __builtin_popcount(pocc[i]);
    cidx_t *c = (cidx_t *)malloc(sizeof(cidx_t) * (size_t)(m ? m : 1));
    i64 *par = (i64 *)malloc(sizeof(i64) * (size_t)(m ? m : 1));
    if (!c || !par) return -1;
    i64 k = 0;
    for (i64 i = 0; i < n; ++i)
        for (int b = 0; b < 8; ++b)
            if ((pocc[i] >> b) & 1) {
                c[k].x = pcoords[3 * i] * 2 + (b & 1);
                c[k].y = pcoords[3 * i + 1] * 2 + ((b >> 1) & 1);
                c[k].z = pcoords[3 * i + 2] * 2 + ((b >> 2) & 1);
                c[k].idx = k; par[k] = i; ++k;
            }
    qsort(c, (size_t)m, sizeof(cidx_t), cmp_zyx);
    for (i64 j = 0; j < m; ++j) {
        ccoords[3 * j] = c[j].x; ccoords[3 * j + 1] = c[j].y; ccoords[3 * j + 2] = c[j].z;
        cparent[j] = par[c[j].idx];
    }
    *n_out = m;
    free(c); free(par);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Kernel map of a stride-1 odd-K submanifold conv (spnn.Conv3d(C,C,K),       */
/* network_ue_4stage_conv.py:18-61).  coords must be sorted (z,y,x), unique.  */
/* map[o*K^3 + idx(d)] = row of (c_o + d) or -1,                              */
/* idx(d) = ((dz+r)*K + (dy+r))*K + (dx+r), r = K/2 (x fastest).             */
/* ------------------------------------------------------------------------- */
static i64 find_row(const i32 *coords, i64 n, i32 x, i32 y, i32 z) {
    i64 lo = 0, hi = n;
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        const i32 *c = coords + 3 * mid;
        int lt;
        if (c[2] != z) lt = c[2] < z; else if (c[1] != y) lt = c[1] < y; else lt = c[0] < x;
        if (lt) lo = mid + 1; else hi = mid;
    }
    if (lo < n && coords[3 * lo] == x && coords[3 * lo + 1] == y && coords[3 * lo + 2] == z) return lo;
    return -1;
}

int orc_kmap(const i32 *coords, i64 n, int K, i32 *map) {
    const int r = K / 2, K3 = K * K * K;
    #pragma omp parallel for schedule(static)
    for (i64 o = 0; o < n; ++o) {
        const i32 *c = coords + 3 * o;
        for (int dz = -r; dz <= r; ++dz)
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx) {
                    int k = ((dz + r) * K + (dy + r)) * K + (dx + r);
                    map[o * K3 + k] = (dx == 0 && dy == 0 && dz == 0)
                        ? (i32)o : (i32)find_row(coords, n, c[0] + dx, c[1] + dy, c[2] + dz);
                }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Sparse conv forward, bias-less: y[o,:] = sum_k x[map[o,k],:] . W[k]        */
/* W is [K^3, Cin, Cout] (spnn.Conv3d.kernel layout, kit/nn.py:14-15);        */
/* fp32 accumulate in fixed order (k ascending, ci ascending).                */
/* ------------------------------------------------------------------------- */
int orc_conv(const float *x, const float *W, const i32 *map, i64 n, int K3, int Cin, int Cout, float *y) {
    #pragma omp parallel for schedule(dynamic, 256)
    for (i64 o = 0; o < n; ++o) {
        float acc[64];
        for (int co = 0; co < Cout; ++co) acc[co] = 0.f;
        for (int k = 0; k < K3; ++k) {
            i32 j = map[o * K3 + k];
            if (j < 0) continue;
            const float *xr = x + (i64)j * Cin;
            const float *Wk = W + (i64)k * Cin * Cout;
            for (int ci = 0; ci < Cin; ++ci) {
                const float xv = xr[ci];
                const float *w = Wk + ci * Cout;
                for (int co = 0; co < Cout; ++co) acc[co] += xv * w[co];
            }
        }
        for (int co = 0; co < Cout; ++co) y[o * Cout + co] = acc[co];
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* pred_head_s{i}: Linear(C,C) - ReLU - Linear(C,A) - Softmax                 */
/* (network_ue_4stage_conv.py:65-94).  nn.Linear: y = x W^T + b, W [out,in].  */
/* ------------------------------------------------------------------------- */
int orc_head(const float *f, i64 n, int C, const float *W1, const float *b1,
             const float *W2, const float *b2, int A, float *prob) {
    #pragma omp parallel for schedule(static)
    for (i64 o = 0; o < n; ++o) {
        float h[64], lg[16];
        const float *fr = f + o * C;
        for (int j = 0; j < C; ++j) {
            float s = b1[j];
            for (int i = 0; i < C; ++i) s += fr[i] * W1[j * C + i];
            h[j] = s > 0.f ? s : 0.f;
        }
        float mx = -INFINITY;
        for (int a = 0; a < A; ++a) {
            float s = b2[a];
            for (int i = 0; i < C; ++i) s += h[i] * W2[a * C + i];
            lg[a] = s;
            if (s > mx) mx = s;
        }
        float sum = 0.f;
        for (int a = 0; a < A; ++a) { lg[a] = expf(lg[a] - mx); sum += lg[a]; }
        for (int a = 0; a < A; ++a) prob[o * A + a] = lg[a] / sum;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* prob -> CDF -> int16 -- pcc_utils.py:146-148 and kit/op.py:50-79:          */
/*   cdf = clamp(cat(0, cumsum(p)), 0, 1); cdf*(2^16-(Lp-1)); round (half     */
/*   even); cast int16 (wraps mod 2^16); += arange(Lp).  Lp = A+1.            */
/* ------------------------------------------------------------------------- */
int orc_cdf_u16(const float *prob, i64 n, int A, u16 *cdf) {
    const int Lp = A + 1;
    const float scale = 65536.0f - (float)(Lp - 1);
    for (i64 o = 0; o < n; ++o) {
        float c = 0.f;
        for (int k = 0; k < Lp; ++k) {
            if (k > 0) c += prob[o * A + (k - 1)];
            float cc = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
            i32 q = (i32)nearbyintf(cc * scale);           /* round-half-even in the default FP env */
            cdf[o * Lp + k] = (u16)((u32)q + (u32)k);       /* int16 wrap == uint16 bit pattern */
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Range coder (torchac 0.9.3 encode_int16_normalized_cdf /                   */
/* decode_int16_normalized_cdf, call sites pcc_utils.py:174-177,322-366).     */
/* Restated from the in-tree twin of the same algorithm:                      */
/* HAC/submodules/arithmetic.zip!arithmetic/arithmetic_kernel.cu:58-91        */
/* (bit sink), :114-162 (encode), :237-287 (bit source, binsearch),           */
/* :310-355 (decode); with integer CDF rows                                   */
/*   c_low = cdf[s], c_high = (s == Lp-2) ? 0x10000 : cdf[s+1].               */
/* ------------------------------------------------------------------------- */
typedef struct { u8 *out; i64 cap; i64 len; u8 cache; u8 count; int overflow; } bitsink_t;

static inline void sink_append(bitsink_t *s, int bit) {
    s->cache = (u8)((s->cache << 1) | (bit & 1));
    if (++s->count == 8) {
        if (s->len < s->cap) s->out[s->len] = s->cache; else s->overflow = 1;
        s->len++; s->count = 0; s->cache = 0;
    }
}
static inline void sink_bit_and_pending(bitsink_t *s, int bit, u64 *pending) {
    sink_append(s, bit);
    while (*pending > 0) { sink_append(s, !bit); --*pending; }
}

int orc_ac_encode(const u16 *cdf, const int16_t *sym, i64 n, int Lp, u8 *out, i64 cap, i64 *out_len) {
    bitsink_t s = { out, cap, 0, 0, 0, 0 };
    u32 low = 0, high = 0xFFFFFFFFu;
    u64 pending = 0;
    const int max_symbol = Lp - 2;
    for (i64 i = 0; i < n; ++i) {
        const int si = sym[i];
        if (si < 0 || si > max_symbol) return -2;
        const u64 span = (u64)high - (u64)low + 1;
        const u32 c_low = cdf[i * Lp + si];
        const u32 c_high = si == max_symbol ? 0x10000u : cdf[i * Lp + si + 1];
        high = (low - 1) + (u32)((span * (u64)c_high) >> 16);
        low  = low + (u32)((span * (u64)c_low) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                sink_bit_and_pending(&s, 0, &pending);
                low <<= 1; high <<= 1; high |= 1;
            } else if (low >= 0x80000000u) {
                sink_bit_and_pending(&s, 1, &pending);
                low <<= 1; high <<= 1; high |= 1;
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                pending++;
                low <<= 1; low &= 0x7FFFFFFFu;
                high <<= 1; high |= 0x80000001u;
            } else break;
        }
    }
    pending += 1;
    if (low < 0x40000000u) sink_bit_and_pending(&s, 0, &pending);
    else                   sink_bit_and_pending(&s, 1, &pending);
    if (s.count > 0) for (int i = s.count; i < 8; ++i) sink_append(&s, 0);
    *out_len = s.len;
    return s.overflow ? -3 : 0;
}

typedef struct { const u8 *in; i64 len; i64 ptr; u8 cache; u8 bits; } bitsrc_t;

static inline void src_get(bitsrc_t *s, u32 *value) {
    if (s->bits == 0) {
        if (s->ptr == s->len) { *value <<= 1; return; }
        s->cache = s->in[s->ptr++]; s->bits = 8;
    }
    *value <<= 1;
    *value |= (u32)((s->cache >> (s->bits - 1)) & 1);
    s->bits--;
}

int orc_ac_decode(const u16 *cdf, const u8 *in, i64 in_len, i64 n, int Lp, int16_t *sym) {
    bitsrc_t s = { in, in_len, 0, 0, 0 };
    u32 low = 0, high = 0xFFFFFFFFu, value = 0;
    const int max_symbol = Lp - 2;
    for (int i = 0; i < 32; ++i) src_get(&s, &value);
    for (i64 i = 0; i < n; ++i) {
        const u64 span = (u64)high - (u64)low + 1;
        const u16 count = (u16)((((u64)value - (u64)low + 1) * 0x10000ull - 1) / span);
        int left = 0, right = max_symbol + 1;
        const u16 *row = cdf + i * Lp;
        int si = -1;
        while (left + 1 < right) {
            const int m = (left + right) / 2;
            const u16 v = row[m];
            if (v < count) left = m; else if (v > count) right = m; else { si = m; break; }
        }
        if (si < 0) si = left;
        sym[i] = (int16_t)si;
        const u32 c_low = row[si];
        const u32 c_high = si == max_symbol ? 0x10000u : row[si + 1];
        high = (low - 1) + (u32)((span * (u64)c_high) >> 16);
        low  = low + (u32)((span * (u64)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1; high <<= 1; high |= 1;
                src_get(&s, &value);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low <<= 1; low &= 0x7FFFFFFFu;
                high <<= 1; high |= 0x80000001u;
                value -= 0x40000000u;
                src_get(&s, &value);
            } else break;
        }
    }
    return 0;
}
