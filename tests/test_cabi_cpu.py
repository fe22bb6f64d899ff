"""CPU-side checks of the product package: the C-ABI library loads and exports every symbol declared in
include/gpcgc.h, the host range coder and the container are bit-exact against the oracle / golden
vectors, and the product fails loudly without a GPU (no fallback).  No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from gauspcc_b200 import _lib, bitstream
from gauspcc_b200 import weights as W
from gauspcc_b200.synth import hac_like_cloud, uniform_unique_cloud
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "gpcgc.h")).read()
    declared = set(re.findall(r"\b(gpc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gpcgc.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.gpc_version() >= 1


def _enc(lib, cdf, sym):
    n, Lp = cdf.shape
    out = np.empty(4 * n + 64, dtype=np.uint8)
    ln = C.c_int64(0)
    rc = lib.gpc_ac_encode_h(cdf.ctypes.data_as(C.c_void_p), sym.ctypes.data_as(C.c_void_p), n, Lp,
                             out.ctypes.data_as(C.c_void_p), out.size, C.byref(ln))
    assert rc == 0
    return out[:ln.value].tobytes()


def _dec(lib, cdf, stream):
    n, Lp = cdf.shape
    buf = np.frombuffer(stream, dtype=np.uint8).copy() if stream else np.zeros(1, np.uint8)
    sym = np.empty(n, dtype=np.uint8)
    assert lib.gpc_ac_decode_h(cdf.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), len(stream), n, Lp,
                               sym.ctypes.data_as(C.c_void_p)) == 0
    return sym


def test_host_range_coder_known_answers():
    lib = _lib.load()
    cdf = np.tile(np.array([[0, 32768, 0]], dtype=np.uint16), (8, 1))
    assert _enc(lib, cdf, np.array([0, 1, 1, 0, 1, 0, 0, 1], np.uint8)).hex() == "6940"
    q = np.tile(np.array([[0, 45873, 52428, 58982, 0]], dtype=np.uint16), (6, 1))
    assert _enc(lib, q, np.array([0, 0, 3, 0, 1, 2], np.uint8)).hex() == "77c0"
    q = np.array([[*(4096 * np.arange(16)), 0]], dtype=np.uint16)
    assert _enc(lib, q, np.array([15], np.uint8)).hex() == "f4"


@pytest.mark.parametrize("A", [2, 4, 16])
@pytest.mark.parametrize("n", [0, 1, 7, 5000])
def test_host_range_coder_matches_oracle(A, n):
    lib = _lib.load()
    rng = np.random.default_rng(A * 1000 + n)
    conc = rng.choice([0.05, 1.0, 20.0])
    p = rng.dirichlet(np.full(A, conc), size=max(n, 1)).astype(np.float32)[:n]
    if n > 10:
        p[:3] = 0; p[:3, A - 1] = 1.0            # degenerate rows: p = 1 on the last symbol
        p[3:6] = 0; p[3:6, 0] = 1.0
    cdf = O.cdf_u16(p) if n else np.zeros((0, A + 1), np.uint16)
    # symbols drawn from p, plus some improbable ones (long carry chains / pending bits)
    sym = np.array([rng.choice(A, p=row / row.sum()) for row in p.astype(np.float64)], dtype=np.uint8) if n else np.zeros(0, np.uint8)
    if n > 100:
        sym[50:60] = rng.integers(0, A, 10)
    ref = O.ac_encode(cdf, sym.astype(np.int16)) if n else O.ac_encode(np.zeros((0, A + 1), np.uint16), np.zeros(0, np.int16))
    got = _enc(lib, np.ascontiguousarray(cdf), sym)
    assert got == ref
    assert np.array_equal(_dec(lib, cdf, got), sym)
    assert np.array_equal(O.ac_decode(cdf, got).astype(np.uint8), sym)


@pytest.mark.parametrize("A", [2, 4, 16])
def test_host_range_decoder_underflow_heavy(A):
    """Long streams whose intervals keep straddling the midpoint (boundaries one step off 1/2) and keep carrying: the decoder's
    one-step underflow renormalisation (4- and 16-ary) and its loop form (binary) against the oracle's bit-at-a-time decoder."""
    lib = _lib.load()
    rng = np.random.default_rng(77 + A)
    n = 60000
    cdf = np.zeros((n, A + 1), dtype=np.uint16)
    half = A // 2
    for i in range(n):
        mid = 32768 + int(rng.integers(-2, 3))                     # the boundary between the two halves of the alphabet
        lo = np.sort(rng.choice(np.arange(1, mid), size=half - 1, replace=False)) if half > 1 else np.zeros(0, int)
        hi = np.sort(rng.choice(np.arange(mid + 1, 65535), size=A - half - 1, replace=False)) if A - half > 1 else np.zeros(0, int)
        cdf[i, 1:A] = np.concatenate([lo, [mid], hi])
    sym = rng.integers(half - 1, half + 1, size=n).astype(np.uint8)      # always next to the midpoint
    sym[::97] = rng.integers(0, A, size=len(sym[::97]))
    stream = _enc(lib, cdf, sym)
    assert stream == O.ac_encode(cdf, sym.astype(np.int16))
    assert np.array_equal(_dec(lib, cdf, stream), sym)
    assert np.array_equal(O.ac_decode(cdf, stream).astype(np.uint8), sym)


@pytest.mark.parametrize("A", [2, 4, 16])
def test_host_range_decoder_in_pieces(A):
    """gpc_ac_decode_begin_h / _more_h: any split of the rows decodes to the symbols of the one-shot call (empty pieces, single
    symbols, a piece boundary right after the stream's last byte)."""
    lib = _lib.load()
    rng = np.random.default_rng(5 + A)
    n = 20000
    p = rng.dirichlet(np.ones(A) * 0.5, size=n)
    c = np.clip(np.cumsum(p, axis=1), 0, 1)
    cdf = np.zeros((n, A + 1), dtype=np.uint16)
    cdf[:, 1:] = (np.rint(c * (65536 - A)).astype(np.int64) + np.arange(1, A + 1)).astype(np.uint16)
    sym = (rng.random(n)[:, None] > c).sum(1).clip(0, A - 1).astype(np.uint8)
    stream = _enc(lib, cdf, sym)
    assert np.array_equal(_dec(lib, cdf, stream), sym)
    buf = (C.c_char * max(len(stream), 1)).from_buffer_copy(stream)
    for cuts in ([0, n], [0, 0, 1, 2, 7, 7, 4096, 4097, n - 1, n], sorted(rng.integers(0, n + 1, size=40).tolist() + [0, n])):
        state = (C.c_char * int(lib.gpc_ac_decode_state_bytes()))()
        assert lib.gpc_ac_decode_begin_h(C.cast(state, C.c_void_p), C.cast(buf, C.c_void_p), len(stream)) == 0
        out = np.full(n, 255, np.uint8)
        for a, b in zip(cuts[:-1], cuts[1:]):
            rows, dst = cdf[a:b], out[a:b]
            assert lib.gpc_ac_decode_more_h(C.cast(state, C.c_void_p), rows.ctypes.data_as(C.c_void_p), b - a, A + 1,
                                            dst.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(out, sym)


def test_host_range_coder_rejects_bad_symbol():
    lib = _lib.load()
    cdf = np.array([[0, 100, 0]], dtype=np.uint16)
    out = np.empty(64, np.uint8)
    ln = C.c_int64(0)
    rc = lib.gpc_ac_encode_h(cdf.ctypes.data_as(C.c_void_p), np.array([2], np.uint8).ctypes.data_as(C.c_void_p), 1, 3,
                             out.ctypes.data_as(C.c_void_p), 64, C.byref(ln))
    assert rc == -5 and b"out of range" in lib.gpc_last_error()


def test_container_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "op_golden.npz"))
    packed = g["packed"].tobytes()
    parts = bitstream.unpack_byte_stream(packed)
    assert [len(p) for p in parts] == list(g["stream_lens"])
    assert bitstream.pack_byte_stream_ls(parts) == packed
    assert bitstream.pack_byte_stream_ls([b"\x01\x02", b"", b"\xff"]).hex() == "03000200000001020000000001000000ff"


def test_file_layout_matches_reference_driver(golden_dir):
    cg = np.load(os.path.join(golden_dir, "codec_golden.npz"))
    blob = cg["hac600_bin"].tobytes()
    posQ, bx, bo, streams = bitstream.read_file(blob)
    assert float(posQ) == 1.0 and bx.shape[0] == bo.shape[0] < 64 and len(streams) % 4 == 0
    assert bitstream.write_file(1, bx, bo, streams) == blob
    with pytest.raises(ValueError):
        bitstream.read_file(blob[:20])


def test_make_xform_host():
    lib = _lib.load()
    mm = np.array([10, 20, 30, 10 + 255, 20 + 256, 30], dtype=np.uint32)
    xf = _lib.KeyXform()
    assert lib.gpc_make_xform_h(mm.ctypes.data_as(C.c_void_p), C.byref(xf)) == 0
    assert (xf.minx, xf.miny, xf.minz, xf.sy, xf.sz, xf.total_bits) == (10, 20, 30, 8, 17, 17)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gauspcc_b200 import pcc_utils
    with pytest.raises(_lib.GpcError):
        pcc_utils.calculate_morton_order(torch.zeros(4, 3))
    with pytest.raises(AssertionError):
        pcc_utils.calculate_morton_order(torch.zeros(4, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gauspcc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "gpcgc_oracle" not in src, f


def test_weights_layout_and_validation():
    sd = W.make_synthetic_state_dict()
    lay = W.reference_layout()
    assert list(sd) == list(lay) and sum(v.numel() for v in sd.values()) == 2318176   # 18*125*32*32 + embeddings + heads + fog
    W.validate_state_dict(sd)
    bad = dict(sd); bad.pop("pred_head_s2_emb.weight")
    with pytest.raises(RuntimeError):
        W.validate_state_dict(bad)
    bad = dict(sd); bad["prior_embedding.weight"] = torch.zeros(255, 32)
    with pytest.raises(RuntimeError):
        W.validate_state_dict(bad)
    assert torch.equal(W.make_synthetic_state_dict()["spatial_conv_s3.2.kernel"], sd["spatial_conv_s3.2.kernel"])


def test_synthetic_clouds():
    a = hac_like_cloud(20000, 0)
    assert a.shape == (20000, 3) and a.dtype == np.int32 and np.unique(a, axis=0).shape[0] == 20000
    assert a.min() < 0 < a.max() and np.abs(a).max() < (1 << 20) - 16
    assert np.array_equal(a, hac_like_cloud(20000, 0)) and not np.array_equal(a, hac_like_cloud(20000, 1))
    b = uniform_unique_cloud(5000, 1, extent_log2=10)
    assert np.unique(b, axis=0).shape[0] == 5000


@pytest.mark.parametrize("A", [2, 4, 16])
def test_lohi_encoder_same_bytes(A):
    """gpc_ac_encode_lohi_h (fed with c_low | c_high << 16, 0 == 0x10000) produces the torchac-compatible bytes."""
    lib = _lib.load()
    rng = np.random.default_rng(A)
    n = 20000
    p = rng.dirichlet(np.full(A, 0.3), size=n).astype(np.float32)
    p[:50] = 0; p[:50, A - 1] = 1.0
    cdf = O.cdf_u16(p)
    sym = rng.integers(0, A, n).astype(np.uint8)
    sym[:5000] = np.array([rng.choice(A, p=r / r.sum()) for r in p[:5000].astype(np.float64)], dtype=np.uint8)
    lo = cdf[np.arange(n), sym].astype(np.uint32)
    hi = np.where(sym == A - 1, 0, cdf[np.arange(n), np.minimum(sym + 1, A)]).astype(np.uint32)
    lohi = np.ascontiguousarray(lo | (hi << 16))
    out = np.empty(4 * n + 64, np.uint8)
    ln = C.c_int64(0)
    assert lib.gpc_ac_encode_lohi_h(lohi.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), out.size, C.byref(ln)) == 0
    ref = O.ac_encode(cdf, sym.astype(np.int16))
    assert out[:ln.value].tobytes() == ref == _enc(lib, cdf, sym)
    assert np.array_equal(_dec(lib, cdf, ref), sym)


def test_container_v2_trailer_and_chunk_rule():
    """Container version 2 (opt-in): the trailing stream that names the format, and the chunk length of a stream --
    host logic only (the coder itself is GPU code: tests/test_gpu_parity.py::test_container_v2_gpu_chunk_coder)."""
    from gauspcc_b200 import bitstream
    from gauspcc_b200.codec import GausPcgcCodec
    v1 = [b"\x01\x02", b"", b"\xff", b"abc"] * 3
    assert bitstream.split_v2(v1) == (v1, 0, None)
    v2 = v1 + [bitstream.v2_trailer(2048, 123456)]
    st, chunk, n_vox = bitstream.split_v2(v2)
    assert st == v1 and chunk == 2048 and n_vox == 123456
    # the outer layout is the reference's: a version-2 file parses with the same reader, one stream more
    blob = bitstream.write_file(1, np.zeros((2, 3), np.int32), np.array([1, 255], np.uint8), v2)
    _, bx, bo, back = bitstream.read_file(blob)
    assert bx.shape == (2, 3) and list(bo) == [1, 255] and back == v2
    assert bitstream.split_v2(v1 + [b"GPCGC-V2"]) == (v1 + [b"GPCGC-V2"], 0, None)          # a malformed trailer is not a trailer
    # chunk length of a stream of n symbols: the file's chunk size, at least 64 chunks per stream where 64-symbol chunks allow
    cl = GausPcgcCodec.chunk_len
    assert cl(1_000_000, 2048) == 2048 and cl(200_000, 2048) == 2048 and cl(100_000, 2048) == 1563
    assert cl(9_000, 2048) == 141 and cl(500, 2048) == 64 and cl(10, 2048) == 64 and cl(0, 2048) == 64
    assert cl(1_000_000, 32) == 32 and cl(5, 1) == 1
    for n in (1, 63, 64, 65, 4095, 4096, 4097, 131071):
        c = cl(n, 2048)
        assert 1 <= c <= 2048 and (n + c - 1) // c <= max(64, (n + 63) // 64 + 1)
