// ac_host.cu -- host range coder, bit-compatible with torchac 0.9.3's
// encode_int16_normalized_cdf / decode_int16_normalized_cdf as called at
// src/gs_compress/HAC/utils/pcc_utils.py:174-177 and :322,336,351,366 (a-14).
//
// 32-bit low/high coder, 16-bit precision, pending-bit carry resolution, MSB-first bit packing; the
// in-tree statement of the same algorithm is HAC/submodules/arithmetic.zip!arithmetic/
// arithmetic_kernel.cu:58-91,114-162 (encode) and :237-287,310-355 (decode).  A CDF row has Lp uint16
// entries; the last one is never read (the top of the last symbol is 0x10000).  This is product code
// (the CPU oracle under oracle/ has its own, separate restatement).
#include <string.h>
#include "common.cuh"
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace {

struct BitWriter {
    u8 *out;
    i64 cap, len;
    u64 acc;        // bits not yet flushed, right-aligned
    int nacc;
    bool overflow;
    inline void put(u32 bit) {
        acc = (acc << 1) | bit;
        if (++nacc == 8) {
            if (len < cap) out[len] = (u8)acc; else overflow = true;
            ++len; acc = 0; nacc = 0;
        }
    }
    inline void put_with_pending(u32 bit, u64 &pending) {
        put(bit);
        const u32 inv = bit ^ 1u;
        for (; pending > 0; --pending) put(inv);
    }
    inline void finish() { while (nacc != 0) put(0); }
};

struct BitReader {
    const u8 *in;
    i64 len, pos;
    u32 cur;
    int left;
    inline u32 get() {
        if (left == 0) {
            if (pos >= len) return 0;           // zeros past the end of the stream
            cur = in[pos++]; left = 8;
        }
        --left;
        return (cur >> left) & 1u;
    }
};

}  // namespace

extern "C" int gpc_ac_encode_h(const uint16_t *cdf, const uint8_t *sym, int64_t n, int Lp, uint8_t *out, int64_t cap,
                               int64_t *out_len) {
    GPC_REQUIRE(Lp >= 3 && out && out_len, GPC_EINVAL, "bad argument");
    BitWriter bw{out, cap, 0, 0, 0, false};
    u32 low = 0, high = 0xFFFFFFFFu;
    u64 pending = 0;
    const u32 top_sym = (u32)(Lp - 2);
    for (i64 i = 0; i < n; ++i) {
        const u32 s = sym[i];
        if (s > top_sym) { gpc_set_error("symbol %u out of range at row %lld", s, (long long)i); return GPC_EDATA; }
        const u16 *row = cdf + i * Lp;
        const u64 span = (u64)high - (u64)low + 1ull;
        const u64 c_lo = row[s];
        const u64 c_hi = s == top_sym ? 0x10000ull : (u64)row[s + 1];
        high = (low - 1u) + (u32)((span * c_hi) >> 16);
        low = low + (u32)((span * c_lo) >> 16);
        for (;;) {
            if (high < 0x80000000u) {
                bw.put_with_pending(0, pending);
            } else if (low >= 0x80000000u) {
                bw.put_with_pending(1, pending);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                ++pending;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                continue;
            } else {
                break;
            }
            low <<= 1;
            high = (high << 1) | 1u;
        }
    }
    ++pending;
    bw.put_with_pending(low < 0x40000000u ? 0u : 1u, pending);
    bw.finish();
    *out_len = bw.len;
    if (bw.overflow) { gpc_set_error("range coder output buffer too small"); return GPC_ENOSPC; }
    return GPC_OK;
}

// ---- fast paths -------------------------------------------------------------------------------------------------
// Same arithmetic as above, restructured for throughput: the renormalisation shifts out ALL leading bits on which
// low and high agree in one step (count-leading-zeros of low ^ high) instead of one per loop trip, bits are pulled
// from a 64-bit reservoir, and the decoder's 64-bit division is a double-precision divide with an exact +-1 fix-up.
namespace {

struct BitReservoir {
    const u8 *in;
    i64 len, pos;
    u64 buf;          // MSB-aligned, `have` valid bits (zeros past the end of the stream)
    int have;
    inline void refill() {
        if (pos + 8 <= len) {                    // bulk: splice 8 big-endian bytes behind the valid bits, keep whole bytes only
            u64 w;
            memcpy(&w, in + pos, 8);
            w = __builtin_bswap64(w);
            buf |= have ? (w >> have) : w;
            const int adv = (63 - have) >> 3;
            pos += adv;
            have += adv * 8;
        } else {
            while (have <= 56) {
                const u64 byte = pos < len ? in[pos] : 0;
                ++pos;
                buf |= byte << (56 - have);
                have += 8;
            }
        }
    }
    inline u32 take(int s) {       // 0 <= s <= 32; caller keeps have >= 32 by refilling once per symbol
        const u32 r = (u32)((buf >> 32) >> (32 - s));
        buf <<= s;
        have -= s;
        return r;
    }
};

struct FastWriter {
    u8 *out;
    i64 cap, len;
    u64 acc;          // `n` pending output bits, right-aligned
    int n;
    bool overflow;
    inline void put_bits(u32 v, int s) {      // s <= 32
        acc = (acc << s) | v;
        n += s;
        while (n >= 8) {
            n -= 8;
            if (len < cap) out[len] = (u8)(acc >> n); else overflow = true;
            ++len;
        }
    }
    inline void put_run(u32 bit, u64 count) {  // `count` copies of `bit`
        const u32 pat = bit ? 0xFFFFFFFFu : 0u;
        while (count >= 32) { put_bits(pat, 32); count -= 32; }
        if (count) put_bits(pat >> (32 - (int)count), (int)count);
    }
};

}  // namespace

// Decoder, division-free: torchac picks the symbol by count = floor(num / span) and a binary search for the largest m
// with cdf[m] <= count.  floor(num/span) >= cdf[m]  <=>  cdf[m] * span <= num, so the symbol is the number of entries
// cdf[1..Lp-2] whose product with span is <= num -- multiplications (independent, SIMD for the 16-ary stage) replace the
// 64-bit division that heads the serial dependency chain of every symbol, and the products are the very terms the
// interval update needs ((span * c) >> 16).  Renormalisation shifts out the whole common prefix of low/high at once.
#if defined(__x86_64__)
#include <immintrin.h>
#include <cstdlib>
__attribute__((target("avx2"))) static inline int sym16_avx2(const u16 *row, u64 span, u64 num) {
    // prod[m] = row[m] * span for m = 0..15 (row[0] == 0), span = (span-1) + 1 with span-1 < 2^32
    const __m256i sm1 = _mm256_set1_epi64x((long long)(span - 1));
    const __m256i nv = _mm256_set1_epi64x((long long)num);
    int s = 0;
#pragma GCC unroll 4
    for (int v = 0; v < 4; ++v) {
        const __m128i r16 = _mm_loadl_epi64((const __m128i *)(row + 4 * v));          // 4 x u16
        const __m256i r64 = _mm256_cvtepu16_epi64(r16);
        const __m256i p = _mm256_add_epi64(_mm256_mul_epu32(r64, sm1), r64);
        const __m256i gt = _mm256_cmpgt_epi64(p, nv);                                   // product > num  (values < 2^50: signed ok)
        s += 4 - __builtin_popcount((unsigned)_mm256_movemask_pd(_mm256_castsi256_pd(gt)));
    }
    return s - 1;                                                                       // entry 0 (== 0) always counts
}
// AVX-512: the 16 entries as two vectors of eight 64-bit products, compares straight into mask registers
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx2"))) static inline int sym16_avx512(const u16 *row, u64 span, u64 num) {
    const __m512i sm1 = _mm512_set1_epi64((long long)(span - 1));
    const __m512i nv = _mm512_set1_epi64((long long)num);
    const __m256i r16 = _mm256_loadu_si256((const __m256i *)row);                       // 16 x u16 (entry 16 is not read)
    const __m512i lo = _mm512_cvtepu16_epi64(_mm256_castsi256_si128(r16));
    const __m512i hi = _mm512_cvtepu16_epi64(_mm256_extracti128_si256(r16, 1));
    const __m512i plo = _mm512_add_epi64(_mm512_mul_epu32(lo, sm1), lo);
    const __m512i phi = _mm512_add_epi64(_mm512_mul_epu32(hi, sm1), hi);
    const unsigned le = (unsigned)_mm512_cmple_epu64_mask(plo, nv) | ((unsigned)_mm512_cmple_epu64_mask(phi, nv) << 8);
    return __builtin_popcount(le) - 1;                                                   // entry 0 (== 0) always counts
}
#endif

// One decoder body, instantiated for the alphabet (binary / general) and for the host ISA: FAST = AVX2 + BMI2 + LZCNT (any x86-64-v3
// host: single-instruction lzcnt that is defined at 0, shlx / shrx shifts without the flags dependency, the SIMD symbol search).
// Renormalisation: torchac's loop first shifts out the common prefix of low / high (sh bits), then, while low = 01.. and high = 10..
// ("underflow": the interval straddles the midpoint), drops the second bit of both; once the top bits differ they keep differing, so
// the loop is: prefix shift, then u underflow shifts with u = min(leading ones of low << 1, leading zeros of high << 1), done in one step:
//   low' = (low << u) & 0x7FFFFFFF,  high' = (high << u) | 0x80000000 | (2^u - 1),
//   value' = ((value << u) - 2^31 (2^u - 1)) | bits = ((value << u) ^ (u ? 0x80000000 : 0)) | bits      (mod 2^32)
// On the 4- and 16-ary stages the underflow test is taken about as often as not: as a branch it was the decoder's main misprediction
// (-20 % per symbol without it); on the binary stages (0.5 bit per symbol, the interval rarely moves) the branch predicts well and the
// loop form stays.
#if defined(__x86_64__)
#define GPC_AC_FAST_TARGET __attribute__((target("avx2,bmi,bmi2,lzcnt")))
#else
#define GPC_AC_FAST_TARGET
#endif

// Decoder state between calls: one stream may be decoded in several pieces (gpc_ac_decode_begin_h / _more_h), e.g. while
// later CDF rows are still being computed.
struct AcDecState {
    BitReservoir br;
    u32 low, high, value;
};
static inline void ac_decode_start(AcDecState &st, const u8 *in, i64 in_len) {
    st.br = BitReservoir{in, in_len, 0, 0, 0};
    st.br.refill();
    st.low = 0; st.high = 0xFFFFFFFFu; st.value = st.br.take(32);
    st.br.refill();
}

template <int LP /* 3, 5, 17 or 0 = any */, int ISA /* 0 = base, 1 = AVX2 + BMI2 + LZCNT, 2 = + AVX-512 */>
static inline __attribute__((always_inline)) void ac_decode_body(AcDecState &state, const u16 *cdf, i64 n, int Lp_rt, u8 *sym) {
    constexpr bool FAST = ISA >= 1;
    constexpr bool BINARY = LP == 3;
    const int Lp = LP ? LP : Lp_rt;
    BitReservoir br = state.br;                                                 // locals: kept in registers across the loop
    u32 low = state.low, high = state.high, value = state.value;
    const int top_sym = Lp - 2;
    auto clz32 = [](u32 x) -> int {
#if defined(__x86_64__)
        if (FAST) return (int)__builtin_ia32_lzcnt_u32(x);
#endif
        return x ? __builtin_clz(x) : 32;
    };
    for (i64 i = 0; i < n; ++i) {
        const u16 *row = cdf + i * Lp;
        const u64 span = (u64)high - (u64)low + 1ull;
        const u64 num = ((((u64)value - (u64)low) + 1ull) << 16) - 1ull;        // < 2^49
        int s;
        u64 p_lo, p_hi;
        if (BINARY) {
            const u64 p1 = (u64)row[1] * span;
            s = p1 <= num;
            p_lo = s ? p1 : 0;
            p_hi = s ? (span << 16) : p1;
        } else {
#if defined(__x86_64__)
            if (ISA == 2 && LP == 17) {
                s = sym16_avx512(row, span, num);
            } else if (FAST && LP == 17) {
                s = sym16_avx2(row, span, num);
            } else
#endif
            if (LP == 5) {                                                      // three scalar products: no vector set-up for so few
                s = ((u64)row[1] * span <= num) + ((u64)row[2] * span <= num) + ((u64)row[3] * span <= num);
            } else {
                s = 0;
                for (int m = 1; m <= top_sym; ++m) s += (u64)row[m] * span <= num;
            }
            p_lo = (u64)row[s] * span;                                          // (re)computed: cheaper than reloading a vector store
            p_hi = s == top_sym ? (span << 16) : (u64)row[s + 1] * span;       // top of the last symbol = 0x10000
        }
        sym[i] = (u8)s;
        high = (low - 1u) + (u32)(p_hi >> 16);
        low = low + (u32)(p_lo >> 16);
        const int sh = clz32(low ^ high);                                       // common prefix, 0..32 bits
        if (br.have < 40) br.refill();
        low = (u32)((u64)low << sh);
        high = (u32)(((u64)high << sh) | ((1ull << sh) - 1ull));
        if (!BINARY) {
            // prefix shift and underflow shifts with ONE read of sh + u bits: the bits of the two steps are consecutive in the stream and
            // the underflow flip touches bit 31 only, so value'' = ((value << (sh + u)) | bits) ^ (u ? 2^31 : 0)   (mod 2^32)
            const int ul = clz32(~(low << 1)), uh = clz32(high << 1);
            int u = ul < uh ? ul : uh;
            u = u > 31 ? 31 : u;                                                // keeps the shifts defined (low = 01..1, high = 10..0 for 31 bits cannot both hold)
            const int t = sh + u;
            low = (low << u) & 0x7FFFFFFFu;
            high = (high << u) | 0x80000000u | ((1u << u) - 1u);
            if (__builtin_expect(t <= 32, 1)) {
                value = ((u32)((u64)value << t) | br.take(t)) ^ (u ? 0x80000000u : 0u);
            } else {                                                            // > 32 bits at once: the two reads of the plain form
                value = (u32)((u64)value << sh) | br.take(sh);
                if (br.have < 40) br.refill();
                value = ((value << u) ^ 0x80000000u) | br.take(u);
            }
            continue;
        }
        value = (u32)((u64)value << sh) | br.take(sh);
        if (BINARY) {
            while (low >= 0x40000000u && high < 0xC0000000u) {                  // rare here
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                value = (value << 1) | br.take(1);
            }
        } else {
            const int ul = clz32(~(low << 1)), uh = clz32(high << 1);
            int u = ul < uh ? ul : uh;
            u = u > 31 ? 31 : u;                                                // keeps the shifts defined (low = 01..1, high = 10..0 for 31 bits cannot both hold)
            if (br.have < 40) br.refill();
            low = (low << u) & 0x7FFFFFFFu;
            high = (high << u) | 0x80000000u | ((1u << u) - 1u);
            value = ((value << u) ^ (u ? 0x80000000u : 0u)) | br.take(u);
        }
    }
    state.br = br;
    state.low = low; state.high = high; state.value = value;
}
GPC_AC_FAST_TARGET static void ac_decode_fast(AcDecState &st, const u16 *cdf, i64 n, int Lp, u8 *sym) {
    if (Lp == 3) ac_decode_body<3, 1>(st, cdf, n, Lp, sym);
    else if (Lp == 5) ac_decode_body<5, 1>(st, cdf, n, Lp, sym);
    else if (Lp == 17) ac_decode_body<17, 1>(st, cdf, n, Lp, sym);
    else ac_decode_body<0, 1>(st, cdf, n, Lp, sym);
}
#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx2,bmi,bmi2,lzcnt"))) static void ac_decode_512(AcDecState &st, const u16 *cdf, i64 n, u8 *sym) {
    ac_decode_body<17, 2>(st, cdf, n, 17, sym);
}
#endif
static void ac_decode_base(AcDecState &st, const u16 *cdf, i64 n, int Lp, u8 *sym) {
    if (Lp == 3) ac_decode_body<3, 0>(st, cdf, n, Lp, sym);
    else if (Lp == 5) ac_decode_body<5, 0>(st, cdf, n, Lp, sym);
    else ac_decode_body<0, 0>(st, cdf, n, Lp, sym);
}
static void ac_decode_dispatch(AcDecState &st, const u16 *cdf, i64 n, int Lp, u8 *sym) {
#if defined(__x86_64__)
    static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    static const bool wide = fast && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                             __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") && !getenv("GPC_AC_NO_AVX512");
    if (wide && Lp == 17) { ac_decode_512(st, cdf, n, sym); return; }
    if (fast) { ac_decode_fast(st, cdf, n, Lp, sym); return; }
#endif
    ac_decode_base(st, cdf, n, Lp, sym);
}

extern "C" int gpc_ac_decode_h(const uint16_t *cdf, const uint8_t *in, int64_t in_len, int64_t n, int Lp, uint8_t *sym) {
    GPC_REQUIRE(Lp >= 3 && Lp <= 17 && sym, GPC_EINVAL, "bad argument");
    AcDecState st;
    ac_decode_start(st, in, in_len);
    ac_decode_dispatch(st, cdf, n, Lp, sym);
    return GPC_OK;
}

// The same decoder in pieces: `state` is caller-owned memory of gpc_ac_decode_state_bytes() bytes; begin binds it to a stream
// (which must stay alive), every `more` call decodes the next n symbols with the next n CDF rows.  Any split of the rows gives the
// symbols gpc_ac_decode_h gives (tests/test_cabi_cpu.py).
extern "C" int64_t gpc_ac_decode_state_bytes(void) { return (int64_t)sizeof(AcDecState); }
extern "C" int gpc_ac_decode_begin_h(void *state, const uint8_t *in, int64_t in_len) {
    GPC_REQUIRE(state && (in || in_len == 0) && in_len >= 0, GPC_EINVAL, "bad argument");
    ac_decode_start(*(AcDecState *)state, in, in_len);
    return GPC_OK;
}
extern "C" int gpc_ac_decode_more_h(void *state, const uint16_t *cdf, int64_t n, int Lp, uint8_t *sym) {
    GPC_REQUIRE(state && Lp >= 3 && Lp <= 17 && (n == 0 || (cdf && sym)) && n >= 0, GPC_EINVAL, "bad argument");
    ac_decode_dispatch(*(AcDecState *)state, cdf, n, Lp, sym);
    return GPC_OK;
}

// Encoder fed with (c_low | c_high << 16) per symbol, c_high == 0 meaning 0x10000 (top symbol): what the fused head
// kernel writes when it is given the symbol (gpc_head_cdf_sym) -- 4 bytes per row cross PCIe instead of 2*(A+1)+1.
extern "C" int gpc_ac_encode_lohi_h(const uint32_t *lohi, int64_t n, uint8_t *out, int64_t cap, int64_t *out_len) {
    GPC_REQUIRE(out && out_len, GPC_EINVAL, "bad argument");
    FastWriter fw{out, cap, 0, 0, 0, false};
    u32 low = 0, high = 0xFFFFFFFFu;
    u64 pending = 0;
    for (i64 i = 0; i < n; ++i) {
        const u32 e = lohi[i];
        const u64 c_lo = e & 0xFFFFu;
        const u64 c_hi = (e >> 16) ? (u64)(e >> 16) : 0x10000ull;
        const u64 span = (u64)high - (u64)low + 1ull;
        high = (low - 1u) + (u32)((span * c_hi) >> 16);
        low = low + (u32)((span * c_lo) >> 16);
        for (;;) {
            const u32 diff = low ^ high;
            if (!(diff & 0x80000000u)) {
                const int sh = diff ? __builtin_clz(diff) : 32;
                const u32 first = low >> 31;
                fw.put_bits(first, 1);
                if (pending) { fw.put_run(first ^ 1u, pending); pending = 0; }
                if (sh > 1) fw.put_bits(sh == 32 ? (low & 0x7FFFFFFFu) : ((low << 1) >> (32 - (sh - 1))), sh - 1);
                if (sh == 32) { low = 0; high = 0xFFFFFFFFu; }
                else { low <<= sh; high = (high << sh) | ((1u << sh) - 1u); }
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                ++pending;
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
            } else {
                break;
            }
        }
    }
    ++pending;
    const u32 last = low < 0x40000000u ? 0u : 1u;
    fw.put_bits(last, 1);
    fw.put_run(last ^ 1u, pending);
    if (fw.n) fw.put_bits(0, 8 - fw.n);
    *out_len = fw.len;
    if (fw.overflow) { gpc_set_error("range coder output buffer too small"); return GPC_ENOSPC; }
    return GPC_OK;
}
