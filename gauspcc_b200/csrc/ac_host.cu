// ac_host.cu -- host range coder, bit-compatible with torchac 0.9.3's
// encode_int16_normalized_cdf / decode_int16_normalized_cdf as called at
// src/gs_compress/HAC/utils/pcc_utils.py:174-177 and :322,336,351,366 (a-14).
//
// 32-bit low/high coder, 16-bit precision, pending-bit carry resolution, MSB-first bit packing; the
// in-tree statement of the same algorithm is HAC/submodules/arithmetic.zip!arithmetic/
// arithmetic_kernel.cu:58-91,114-162 (encode) and :237-287,310-355 (decode).  A CDF row has Lp uint16
// entries; the last one is never read (the top of the last symbol is 0x10000).  This is product code
// (the CPU oracle under oracle/ has its own, separate restatement).
#include "common.cuh"

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

extern "C" int gpc_ac_decode_h(const uint16_t *cdf, const uint8_t *in, int64_t in_len, int64_t n, int Lp, uint8_t *sym) {
    GPC_REQUIRE(Lp >= 3 && sym, GPC_EINVAL, "bad argument");
    BitReader br{in, in_len, 0, 0, 0};
    u32 low = 0, high = 0xFFFFFFFFu, value = 0;
    for (int i = 0; i < 32; ++i) value = (value << 1) | br.get();
    const int top_sym = Lp - 2;
    for (i64 i = 0; i < n; ++i) {
        const u16 *row = cdf + i * Lp;
        const u64 span = (u64)high - (u64)low + 1ull;
        const u16 count = (u16)(((((u64)value - (u64)low + 1ull) << 16) - 1ull) / span);
        int s;
        if (Lp == 3) {
            s = row[1] <= count;
        } else {
            int lo = 0, hi = top_sym + 1;
            s = -1;
            while (lo + 1 < hi) {
                const int mid = (lo + hi) >> 1;
                const u16 v = row[mid];
                if (v < count) lo = mid; else if (v > count) hi = mid; else { s = mid; break; }
            }
            if (s < 0) s = lo;
        }
        sym[i] = (u8)s;
        const u64 c_lo = row[s];
        const u64 c_hi = s == top_sym ? 0x10000ull : (u64)row[s + 1];
        high = (low - 1u) + (u32)((span * c_hi) >> 16);
        low = low + (u32)((span * c_lo) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1;
                high = (high << 1) | 1u;
                value = (value << 1) | br.get();
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                value = (value << 1) | br.get();
            } else {
                break;
            }
        }
    }
    return GPC_OK;
}
