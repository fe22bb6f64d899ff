"""CPU parity oracle for the GausPcgc anchor-geometry codec -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (gauspcc_b200/) never does.

The driver below restates, line by line, the reference's
    compress_point_cloud    src/gs_compress/HAC/utils/pcc_utils.py:73-203
    decompress_point_cloud  src/gs_compress/HAC/utils/pcc_utils.py:271-381
on top of the C primitives in gpcgc_oracle.c (each cites the reference lines it follows).
Third-party arithmetic that is not vendored in the reference (torchsparse 2.1.0 sparse conv,
torchac 0.9.3 range coder) is restated from its published algorithm; see DESIGN.md "Oracle" for
how it is pinned (tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgpcgc_oracle.so")
_lib = None

STAGE_ALPHABETS = (2, 2, 4, 16)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "gpcgc_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _chk(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed rc={rc}")


# ----------------------------------------------------------------------------- primitives
def lexorder(xyz: np.ndarray) -> np.ndarray:
    """calculate_morton_order (pcc_utils.py:12-22) -> int64 permutation."""
    assert xyz.ndim == 2 and xyz.shape[1] == 3, f'Input data must be a 3D point cloud, but got {xyz.shape}.'
    x = np.ascontiguousarray(xyz).astype(np.int64)
    out = np.empty(x.shape[0], dtype=np.int64)
    _chk(lib().orc_lexorder(_p(x), C.c_int64(x.shape[0]), _p(out)), "lexorder")
    return out


def sort_zyx_perm(coords: np.ndarray) -> np.ndarray:
    c = np.ascontiguousarray(coords, dtype=np.int32)
    perm = np.empty(c.shape[0], dtype=np.int64)
    _chk(lib().orc_sort_zyx(_p(c), C.c_int64(c.shape[0]), _p(perm)), "sort_zyx")
    return perm


def fog(coords: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """kit/nn.py:38-55.  coords int32 [n,3] unique -> (parents [m,3] sorted zyx, occupancy u8 [m])."""
    c = np.ascontiguousarray(coords, dtype=np.int32)
    n = c.shape[0]
    pc = np.empty((max(n, 1), 3), dtype=np.int32)
    po = np.empty(max(n, 1), dtype=np.uint8)
    m = C.c_int64(0)
    _chk(lib().orc_fog(_p(c), C.c_int64(n), _p(pc), _p(po), C.byref(m)), "fog")
    return pc[: m.value].copy(), po[: m.value].copy()


def fcg(pcoords: np.ndarray, pocc: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """kit/nn.py:77-98 + sort_CF.  -> (children [m,3] sorted zyx, parent row of each child [m])."""
    pc = np.ascontiguousarray(pcoords, dtype=np.int32)
    po = np.ascontiguousarray(pocc, dtype=np.uint8)
    m_exp = int(np.unpackbits(po).sum())
    cc = np.empty((max(m_exp, 1), 3), dtype=np.int32)
    par = np.empty(max(m_exp, 1), dtype=np.int64)
    m = C.c_int64(0)
    _chk(lib().orc_fcg(_p(pc), _p(po), C.c_int64(pc.shape[0]), _p(cc), _p(par), C.byref(m)), "fcg")
    assert m.value == m_exp
    return cc[:m_exp].copy(), par[:m_exp].copy()


def fcg_parent_major(pcoords: np.ndarray, pocc: np.ndarray) -> np.ndarray:
    """kit/nn.py:86-92 as called at pcc_utils.py:375: children in parent-major order (parent rows as
    given, octant index i = 0..7 ascending), NOT re-sorted -- this is the row order of the decoded cloud."""
    pc = np.asarray(pcoords, dtype=np.int32)
    bits = ((np.asarray(pocc, dtype=np.uint8)[:, None] >> np.arange(8)) & 1).astype(bool)      # [n,8]
    base = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.int32)   # kit/nn.py:64-73
    allc = pc[:, None, :] * 2 + base[None, :, :]
    return allc[bits]


def kmap(coords: np.ndarray, K: int = 5) -> np.ndarray:
    """Dense kernel map int32 [n, K^3] (x-fastest offset index), -1 = absent.  coords sorted zyx."""
    c = np.ascontiguousarray(coords, dtype=np.int32)
    out = np.empty((c.shape[0], K ** 3), dtype=np.int32)
    _chk(lib().orc_kmap(_p(c), C.c_int64(c.shape[0]), C.c_int(K), _p(out)), "kmap")
    return out


def conv(x: np.ndarray, W: np.ndarray, km: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    n, K3 = km.shape
    assert W.shape[0] == K3 and W.shape[1] == x.shape[1]
    y = np.empty((n, W.shape[2]), dtype=np.float32)
    _chk(lib().orc_conv(_p(x), _p(W), _p(km), C.c_int64(n), C.c_int(K3), C.c_int(W.shape[1]),
                        C.c_int(W.shape[2]), _p(y)), "conv")
    return y


def head(f: np.ndarray, W1, b1, W2, b2) -> np.ndarray:
    f = np.ascontiguousarray(f, dtype=np.float32)
    A = W2.shape[0]
    prob = np.empty((f.shape[0], A), dtype=np.float32)
    _chk(lib().orc_head(_p(f), C.c_int64(f.shape[0]), C.c_int(f.shape[1]),
                        _p(np.ascontiguousarray(W1, dtype=np.float32)), _p(np.ascontiguousarray(b1, dtype=np.float32)),
                        _p(np.ascontiguousarray(W2, dtype=np.float32)), _p(np.ascontiguousarray(b2, dtype=np.float32)),
                        C.c_int(A), _p(prob)), "head")
    return prob


def cdf_u16(prob: np.ndarray) -> np.ndarray:
    prob = np.ascontiguousarray(prob, dtype=np.float32)
    n, A = prob.shape
    out = np.empty((n, A + 1), dtype=np.uint16)
    _chk(lib().orc_cdf_u16(_p(prob), C.c_int64(n), C.c_int(A), _p(out)), "cdf")
    return out


def ac_encode(cdf: np.ndarray, sym: np.ndarray) -> bytes:
    cdf = np.ascontiguousarray(cdf, dtype=np.uint16)
    sym = np.ascontiguousarray(sym, dtype=np.int16)
    n, Lp = cdf.shape
    cap = 4 * n + 64
    out = np.empty(cap, dtype=np.uint8)
    ln = C.c_int64(0)
    _chk(lib().orc_ac_encode(_p(cdf), _p(sym), C.c_int64(n), C.c_int(Lp), _p(out), C.c_int64(cap), C.byref(ln)),
         "ac_encode")
    return out[: ln.value].tobytes()


def ac_decode(cdf: np.ndarray, stream: bytes) -> np.ndarray:
    cdf = np.ascontiguousarray(cdf, dtype=np.uint16)
    n, Lp = cdf.shape
    buf = np.frombuffer(stream, dtype=np.uint8) if len(stream) else np.zeros(1, dtype=np.uint8)
    sym = np.empty(n, dtype=np.int16)
    _chk(lib().orc_ac_decode(_p(cdf), _p(buf), C.c_int64(len(stream)), C.c_int64(n), C.c_int(Lp), _p(sym)),
         "ac_decode")
    return sym


# ----------------------------------------------------------------------------- HAC attribute coder
# HAC/submodules/arithmetic.zip!arithmetic/arithmetic_kernel.cu restated on top of the range coder above: the reference codes
# chunk c (symbols [c * chunk_size, (c + 1) * chunk_size)) with its own coder (:94-163) whose integer CDF entries are
# c = rn(cdf_float * (2^16 - (Lp - 1))) + index (:121-122; binsearch :264-287 compares the same values as uint16).
def attr_cdf(mean, scale, Q, min_value: int, max_value: int) -> np.ndarray:
    """calculate_cdf_kernel (:7-28).  erfc is evaluated in double and rounded to float32: CUDA's float erfc differs from it in
    the last bits, so byte-exact comparisons feed attr_encode / attr_decode the table the GPU produced."""
    from math import sqrt
    from scipy.special import erfc
    mean = np.asarray(mean, np.float32)
    Q = np.asarray(Q, np.float32)
    sc = np.maximum(np.asarray(scale, np.float32).astype(np.float64), 1e-9).astype(np.float32)
    i = np.arange(max_value - min_value + 2)
    sample = ((min_value + i - 0.5)[None, :] * Q.astype(np.float64)[:, None]).astype(np.float32)
    arg = (-(sample - mean[:, None]) / (sc * np.float32(sqrt(2.0)))[:, None]).astype(np.float32)
    return (0.5 * erfc(arg.astype(np.float64))).astype(np.float32)


def _attr_int_rows(cdf: np.ndarray) -> np.ndarray:
    n, Lp = cdf.shape
    v = np.rint(cdf.astype(np.float32) * np.float32(65536 - (Lp - 1))).astype(np.int64) + np.arange(Lp)
    return (v & 0xFFFF).astype(np.uint16)


def attr_encode(sym: np.ndarray, cdf: np.ndarray, chunk_size: int):
    """arithmetic_encode_cu (:186-232) -> (stream bytes, bytes per chunk int32[chunks])"""
    rows = _attr_int_rows(np.asarray(cdf, np.float32))
    sym = np.asarray(sym, np.int16)
    parts = [ac_encode(rows[o:o + chunk_size], sym[o:o + chunk_size]) for o in range(0, len(sym), chunk_size)]
    return b"".join(parts), np.array([len(p) for p in parts], dtype=np.int32)


def attr_decode(cdf: np.ndarray, stream: bytes, cnt: np.ndarray, chunk_size: int) -> np.ndarray:
    """arithmetic_decode_cu (:365-407)"""
    rows = _attr_int_rows(np.asarray(cdf, np.float32))
    out, pos = [], 0
    for c, o in enumerate(range(0, rows.shape[0], chunk_size)):
        out.append(ac_decode(rows[o:o + chunk_size], stream[pos:pos + int(cnt[c])]))
        pos += int(cnt[c])
    return np.concatenate(out) if out else np.zeros(0, np.int16)


# container -- kit/op.py:32-48
def pack_byte_stream_ls(streams: List[bytes]) -> bytes:
    out = np.array(len(streams), dtype=np.uint16).tobytes()
    for s in streams:
        out += np.array(len(s), dtype=np.uint32).tobytes()
        out += s
    return out


def unpack_byte_stream(stream: bytes) -> List[bytes]:
    n = int(np.frombuffer(stream[:2], dtype=np.uint16)[0])
    out, cur = [], 2
    for _ in range(n):
        ln = int(np.frombuffer(stream[cur:cur + 4], dtype=np.uint32)[0])
        out.append(stream[cur + 4:cur + 4 + ln])
        cur += 4 + ln
    return out


# ----------------------------------------------------------------------------- network pieces
def _res_stack(w: Dict[str, np.ndarray], prefix: str, km: np.ndarray, x: np.ndarray) -> np.ndarray:
    """nn.Sequential(Conv3d, ReLU, ResNet, ResNet) -- network_ue_4stage_conv.py:17-22, kit/nn.py:18-22."""
    x = np.maximum(conv(x, w[f"{prefix}.0.kernel"], km), 0)
    for blk in (2, 3):
        out = np.maximum(conv(x, w[f"{prefix}.{blk}.conv0.kernel"], km), 0)
        out = conv(out, w[f"{prefix}.{blk}.conv1.kernel"], km)
        x = np.maximum(out + x, 0)
    return x


def _stage(w: Dict[str, np.ndarray], i: int, km: np.ndarray, f: np.ndarray) -> np.ndarray:
    """spatial_conv_s{i} (Conv, ReLU, Conv) then pred_head_s{i} -- network_ue_4stage_conv.py:40-94."""
    f = np.maximum(conv(f, w[f"spatial_conv_s{i}.0.kernel"], km), 0)
    f = conv(f, w[f"spatial_conv_s{i}.2.kernel"], km)
    return head(f, w[f"pred_head_s{i}.0.weight"], w[f"pred_head_s{i}.0.bias"],
                w[f"pred_head_s{i}.2.weight"], w[f"pred_head_s{i}.2.bias"])


def split_symbols(occ: np.ndarray):
    """pcc_utils.py:112-115."""
    o = occ.astype(np.int64)
    return [(o >> 7) & 1, (o >> 6) & 1, (o >> 4) & 3, o & 15]


def build_pyramid(xyz: np.ndarray):
    """pcc_utils.py:83-89: FOG until a level has < 64 rows; returns list coarsest -> finest."""
    cur = np.unique(np.ascontiguousarray(xyz, dtype=np.int32), axis=0)     # duplicates merge
    levels = []
    while True:
        pc, po = fog(cur)
        levels.append((pc, po))
        cur = pc
        if pc.shape[0] < 64:
            break
    return levels[::-1]


def _level_features(w, K, x_C, x_O):
    """Shared encode/decode body: pcc_utils.py:99-109 == :300-311."""
    x_F = w["prior_embedding.weight"][x_O.astype(np.int64)]
    km_p = kmap(x_C, K)
    x_F = _res_stack(w, "prior_resnet", km_p, x_F)
    up_C, par = fcg(x_C, x_O)
    up_F = x_F[par]
    idx = (up_C[:, 0] & 1) + 2 * (up_C[:, 1] & 1) + 4 * (up_C[:, 2] & 1)           # kit/nn.py:114-116
    up_F = up_F + w["target_embedding.target_res_embedding.weight"][idx]
    km_t = kmap(up_C, K)
    up = _res_stack(w, "target_resnet", km_t, up_F)
    return up_C, up, km_t


def encode(xyz: np.ndarray, w: Dict[str, np.ndarray], K: int = 5, posQ: float = 1.0,
           collect: bool = False):
    """compress_point_cloud body (pcc_utils.py:73-203) -> file bytes (and per-level aux if collect)."""
    levels = build_pyramid(xyz)
    streams: List[bytes] = []
    aux = []
    for d in range(len(levels) - 1):
        x_C, x_O = levels[d]
        gt_C, gt_O = levels[d + 1]
        up_C, up, km_t = _level_features(w, K, x_C, x_O)
        assert np.array_equal(up_C, gt_C), "FCG children != FOG level (oracle self-check)"
        s = split_symbols(gt_O)
        ctx = [None, s[0], 2 * s[0] + s[1], 4 * (2 * s[0] + s[1]) + s[2]]            # :122,129,137
        lv = {"coords": up_C, "occ": gt_O, "probs": [], "cdfs": []}
        for i in range(4):
            f = up if i == 0 else up + w[f"pred_head_s{i}_emb.weight"][ctx[i]]
            p = _stage(w, i, km_t, f)
            q = cdf_u16(p)
            streams.append(ac_encode(q, s[i].astype(np.int16)))
            if collect:
                lv["probs"].append(p); lv["cdfs"].append(q)
        aux.append(lv)
    base_C, base_O = levels[0]
    blob = np.array(posQ, dtype=np.float16).tobytes()
    blob += np.array(base_C.shape[0], dtype=np.int32).tobytes()
    blob += np.ascontiguousarray(base_C, dtype=np.int32).tobytes()
    blob += np.ascontiguousarray(base_O, dtype=np.uint8).tobytes()
    blob += pack_byte_stream_ls(streams)
    return (blob, {"levels": levels, "aux": aux}) if collect else blob


def decode(blob: bytes, w: Dict[str, np.ndarray], K: int = 5) -> np.ndarray:
    """decompress_point_cloud body (pcc_utils.py:271-381) -> float32 [N,3] (coords * posQ)."""
    posQ = np.frombuffer(blob[:2], dtype=np.float16)[0]
    n0 = int(np.frombuffer(blob[2:6], dtype=np.int32)[0])
    x_C = np.frombuffer(blob[6:6 + 12 * n0], dtype=np.int32).reshape(-1, 3).copy()
    x_O = np.frombuffer(blob[6 + 12 * n0:6 + 13 * n0], dtype=np.uint8).copy()
    streams = unpack_byte_stream(blob[6 + 13 * n0:])
    for g in range(0, len(streams), 4):
        up_C, up, km_t = _level_features(w, K, x_C, x_O)
        s = []
        for i in range(4):
            if i == 0:
                f = up
            else:
                ctx = s[0] if i == 1 else (2 * s[0] + s[1] if i == 2 else 4 * (2 * s[0] + s[1]) + s[2])
                f = up + w[f"pred_head_s{i}_emb.weight"][ctx]
            q = cdf_u16(_stage(w, i, km_t, f))
            s.append(ac_decode(q, streams[g + i]).astype(np.int64))
        x_O = (s[0] * 128 + s[1] * 64 + s[2] * 16 + s[3]).astype(np.uint8)             # :369
        x_C = up_C
    scan = fcg_parent_major(x_C, x_O)                                                  # :375 (no sort_CF here)
    return scan.astype(np.float32) * np.float32(posQ)                                  # :379 (int32*f16 -> f32)
