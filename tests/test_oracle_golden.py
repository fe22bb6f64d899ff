"""The CPU oracle against the committed golden fixtures (tests/golden/make_golden.py):
reference-owned kit/op.py + calculate_morton_order outputs, and a full run of the reference's own
compress/decompress driver + Network over stand-ins for torchsparse/torchac."""
import os

import numpy as np
import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def opg(golden_dir):
    return np.load(os.path.join(golden_dir, "op_golden.npz"))


@pytest.fixture(scope="module")
def cg(golden_dir):
    return np.load(os.path.join(golden_dir, "codec_golden.npz"))


def test_sort_cf_order(opg):
    C, F = opg["sort_C"], opg["sort_F"]
    perm = O.sort_zyx_perm(C[:, 1:])
    assert np.array_equal(C[perm], opg["sorted_C"])
    # the reference's first pass is a non-stable torch.sort (kit/op.py:18), so the payload order of
    # DUPLICATE coordinates is unspecified; rows with unique coordinates must match exactly
    _, inv, cnt = np.unique(opg["sorted_C"], axis=0, return_inverse=True, return_counts=True)
    uniq = cnt[inv.reshape(-1)] == 1
    assert uniq.sum() > 400
    assert np.array_equal(F[perm][uniq], opg["sorted_F"][uniq])


def test_cdf_int16(opg):
    # kit/op.py:50-79 takes the float CDF (already [0, cumsum]); feed the oracle the increments
    cdf_f = opg["cdf_float"]
    ref = opg["cdf_int16"].astype(np.int16).view(np.uint16)
    scale = np.float32(65536 - (cdf_f.shape[1] - 1))
    got = (np.rint(cdf_f * scale).astype(np.int64) + np.arange(cdf_f.shape[1])).astype(np.uint16)
    assert np.array_equal(got, ref)                          # the wrap rule the C code uses
    # The C path on probabilities: cumsum is SEQUENTIAL fp32 (what our CUDA head kernel does).  torch's own
    # cumsum accumulates in double on CPU and as a tree scan on CUDA, so an entry can differ by one
    # quantisation step in rare rows (documented in DESIGN.md); everything after the cumsum is exact.
    import torch
    p = np.random.default_rng(0).dirichlet(np.ones(16), size=200).astype(np.float32)
    p[:5] = 0; p[:5, 3] = 1.0
    seq = np.concatenate([np.zeros((200, 1), np.float32), np.add.accumulate(p, axis=1, dtype=np.float32)], 1)
    t = torch.clamp(torch.tensor(seq), 0, 1)
    f = t.mul(np.float32(65536 - 16)).round().to(torch.int16).add_(torch.arange(17, dtype=torch.int16))
    got = O.cdf_u16(p)
    assert np.array_equal(got, f.numpy().view(np.uint16))
    t = torch.tensor(p)
    c = torch.clamp(torch.cat((t[:, :1] * 0, t.cumsum(-1)), -1), 0, 1)
    f = c.mul(np.float32(65536 - 16)).round().to(torch.int16).add_(torch.arange(17, dtype=torch.int16))
    diff = np.abs(got.astype(np.int64) - f.numpy().view(np.uint16).astype(np.int64))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.01


def test_container(opg):
    lens = opg["stream_lens"]
    packed = opg["packed"].tobytes()
    parts = O.unpack_byte_stream(packed)
    assert [len(p) for p in parts] == list(lens)
    assert O.pack_byte_stream_ls(parts) == packed


def test_calculate_morton_order(opg):
    for i in range(3):
        x = opg[f"mo_in{i}"]
        assert np.array_equal(O.lexorder(x), opg[f"mo_out{i}"])
        xs = x[opg[f"mo_out{i}"]]
        assert np.array_equal(np.lexsort((xs[:, 0], xs[:, 1], xs[:, 2])), np.arange(len(xs)))


def test_ac_known_answers():
    cdf = np.tile(np.array([[0, 32768, 0]], dtype=np.uint16), (8, 1))
    assert O.ac_encode(cdf, np.array([0, 1, 1, 0, 1, 0, 0, 1])).hex() == "6940"
    q = O.cdf_u16(np.array([[.7, .1, .1, .1]] * 6, dtype=np.float32))
    assert list(q[0]) == [0, 45873, 52428, 58982, 0]
    assert O.ac_encode(q, np.array([0, 0, 3, 0, 1, 2])).hex() == "77c0"
    q = np.array([[*(4096 * np.arange(16)), 0]], dtype=np.uint16)
    assert O.ac_encode(q, np.array([15])).hex() == "f4"


def _split_file(blob):
    n0 = int(np.frombuffer(blob[2:6], dtype=np.int32)[0])
    base_c = np.frombuffer(blob[6:6 + 12 * n0], dtype=np.int32).reshape(-1, 3)
    base_o = np.frombuffer(blob[6 + 12 * n0:6 + 13 * n0], dtype=np.uint8)
    return blob[:2], base_c, base_o, O.unpack_byte_stream(blob[6 + 13 * n0:])


@pytest.mark.parametrize("name", ["hac600", "blob", "hac2500"])
def test_codec_vs_reference_driver(cg, weights_np, name):
    xyz = cg[f"{name}_xyz"]
    ref_blob = cg[f"{name}_bin"].tobytes()
    blob, info = O.encode(xyz, weights_np, collect=True)
    h_r, bc_r, bo_r, st_r = _split_file(ref_blob)
    h_o, bc_o, bo_o, st_o = _split_file(blob)
    assert h_r == h_o and np.array_equal(bc_r, bc_o) and np.array_equal(bo_r, bo_o)
    assert len(st_r) == len(st_o) == 4 * (len(info["levels"]) - 1)
    assert abs(len(blob) - len(ref_blob)) <= max(2, 0.002 * len(ref_blob))
    assert all(abs(len(a) - len(b)) <= 2 for a, b in zip(st_r, st_o))
    # decoded geometry of the reference run == the oracle's own round trip, row for row
    dec = O.decode(blob, weights_np)
    assert dec.dtype == np.float32
    assert np.array_equal(np.unique(dec.astype(np.int32), axis=0), np.unique(xyz, axis=0))
    assert np.array_equal(dec, cg[f"{name}_decoded"])
    if f"{name}_probs" not in cg:
        return
    # network math: every probability tensor of every (level, stage), in stream order
    ref_p = cg[f"{name}_probs"]
    got = np.concatenate([p.reshape(-1) for lv in info["aux"] for p in lv["probs"]])
    assert got.shape == ref_p.shape
    assert np.abs(got - ref_p).max() < 2e-4      # fp32 summation-order noise only
    # symbol split, row order, stream order and the range coder, BIT-EXACT: feed the reference run's own
    # probabilities through torch's CPU cumsum (what that run used) and the reference's quantisation rule;
    # the oracle's coder must then reproduce the reference stream bytes and decode them back.
    import torch
    cur = 0
    k = 0
    for lv in info["aux"]:
        sym = O.split_symbols(lv["occ"])
        for i, A in enumerate(O.STAGE_ALPHABETS):
            n = lv["coords"].shape[0]
            t = torch.tensor(ref_p[cur:cur + n * A].reshape(n, A)); cur += n * A
            c = torch.clamp(torch.cat((t[:, :1] * 0, t.cumsum(-1)), -1), 0, 1)
            q = c.mul(np.float32(65536 - A)).round().to(torch.int16).add_(torch.arange(A + 1, dtype=torch.int16))
            q = q.numpy().view(np.uint16)
            assert O.ac_encode(q, sym[i].astype(np.int16)) == st_r[k]
            assert np.array_equal(O.ac_decode(q, st_r[k]), sym[i].astype(np.int16))
            k += 1
    assert k == len(st_r)


def _golden_60k(golden_dir):
    import hashlib
    from gauspcc_b200.synth import hac_like_cloud
    g = np.load(os.path.join(golden_dir, "codec_golden_60k.npz"))
    xyz = hac_like_cloud(int(g["n"][0]), int(g["seed"][0]))
    assert hashlib.sha256(np.ascontiguousarray(xyz.astype(np.int32)).tobytes()).digest() == g["xyz_sha256"].tobytes(), \
        "synth.hac_like_cloud changed: regenerate tests/golden (make_golden.py)"
    return g, xyz


def test_codec_vs_reference_driver_60k(golden_dir, weights_np):
    """A cloud of realistic size (60 K anchors, 11 coded levels, 528 KB of stream) through the reference's own driver
    (make_golden.py section 3): stream sizes, base level, sampled probabilities and the decoded rows."""
    import hashlib
    g, xyz = _golden_60k(golden_dir)
    blob, info = O.encode(xyz, weights_np, collect=True)
    _, bc, bo, st = _split_file(blob)
    assert np.array_equal(bc, g["base_xyz"]) and np.array_equal(bo, g["base_occ"])
    ref_lens = g["stream_lens"]
    assert len(st) == len(ref_lens)
    assert abs(len(blob) - int(g["file_bytes"][0])) <= 0.002 * int(g["file_bytes"][0])
    assert all(abs(len(a) - int(b)) <= max(2, 0.002 * int(b)) for a, b in zip(st, ref_lens))
    every = int(g["probs_every"][0])
    probs = [p for lv in info["aux"] for p in lv["probs"]]
    assert [p.shape[0] for p in probs] == list(g["nprob"]) and [p.shape[1] for p in probs] == list(g["prob_cols"])
    got = np.concatenate([p[::every].reshape(-1) for p in probs])
    assert np.abs(got - g["probs"]).max() < 2e-4
    dec = O.decode(blob, weights_np)
    assert hashlib.sha256(np.ascontiguousarray(dec.astype(np.int32)).tobytes()).digest() == g["decoded_sha256"].tobytes()


def test_roundtrip_signed_and_tiny(weights_np):
    rng = np.random.default_rng(3)
    for xyz in (rng.integers(-50, 50, size=(40, 3)).astype(np.int32),          # < 64 points: base only
                rng.integers(-700, 700, size=(1500, 3)).astype(np.int32),
                np.array([[5, -7, 9]], dtype=np.int32)):
        blob = O.encode(xyz, weights_np)
        dec = O.decode(blob, weights_np).astype(np.int32)
        assert np.array_equal(np.unique(dec, axis=0), np.unique(xyz, axis=0))
