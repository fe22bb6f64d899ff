"""Generate the golden fixtures that pin the CPU oracle (run once, in the build container).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

What is REFERENCE-OWNED in these fixtures (imported unchanged from /root/reference):
  * kit/op.py: sort_CF, _convert_to_int_and_normalize, pack_byte_stream_ls/unpack_byte_stream
  * pcc_utils.py: calculate_morton_order, compress_point_cloud, decompress_point_cloud (the whole
    driver: pyramid loop, level loop, symbol split, contexts, CDF construction, stream order, header)
  * network_ue_4stage_conv.py: Network (layer list, state_dict keys/shapes via strict load_state_dict)
  * kit/nn.py: ResNet, FOG (codes), FCG (child table / mask), TargetEmbedding

What is STUBBED, because the packages are not vendored in the reference and cannot be installed
offline (requirements.txt:6,8): `torchsparse` (SparseTensor, spnn.Conv3d, spnn.ReLU, conv config)
and `torchac` (int16-CDF range coder).  The stand-ins below are pure torch / pure Python and
implement the published semantics:
  * spnn.Conv3d(C,C,K odd, stride 1): submanifold conv, kernel [K^3,Cin,Cout], no bias,
    offset index x-fastest; spnn.Conv3d(1,1,2,stride=2): parent = floor(c/2), sum of child feats.
  * torchac: the low/high/pending-bits coder, written here directly from the in-tree CUDA twin
    HAC/submodules/arithmetic.zip!arithmetic/arithmetic_kernel.cu:58-162,237-355 -- deliberately
    NOT sharing code with oracle/gpcgc_oracle.c so the fixture bytes check the C restatement.
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)


# ----------------------------------------------------------------------------- torchsparse stand-in
def _install_torchsparse_stub():
    ts = types.ModuleType("torchsparse")
    tsnn = types.ModuleType("torchsparse.nn")
    tsF = types.ModuleType("torchsparse.nn.functional")

    class SparseTensor:
        def __init__(self, coords=None, feats=None, stride=1, **kw):
            self.coords, self.feats, self.stride = coords, feats, stride

        C = property(lambda s: s.coords)
        F = property(lambda s: s.feats)

        def to(self, device):
            return SparseTensor(self.coords.to(device), self.feats.to(device), self.stride)

    def _key(c):  # c int64 [n,4] (b,x,y,z) -> python tuples
        return [tuple(r) for r in c.tolist()]

    class Conv3d(torch.nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, bias=False, **kw):
            super().__init__()
            self.k, self.s = kernel_size, stride
            self.kernel = torch.nn.Parameter(torch.zeros(kernel_size ** 3, in_channels, out_channels))
            assert not bias

        def forward(self, x):
            c = x.coords.long()
            f = x.feats.float()
            if self.s == 1:
                K, r = self.k, self.k // 2
                table = {t: i for i, t in enumerate(_key(c))}
                out = torch.zeros(c.shape[0], self.kernel.shape[2])
                cl = c.tolist()
                for dz in range(-r, r + 1):
                    for dy in range(-r, r + 1):
                        for dx in range(-r, r + 1):
                            k = ((dz + r) * K + (dy + r)) * K + (dx + r)
                            oi, ii = [], []
                            for o, (b, xx, yy, zz) in enumerate(cl):
                                j = table.get((b, xx + dx, yy + dy, zz + dz))
                                if j is not None:
                                    oi.append(o); ii.append(j)
                            if oi:
                                out.index_add_(0, torch.tensor(oi), f[torch.tensor(ii)] @ self.kernel[k])
                return SparseTensor(x.coords, out, x.stride)
            assert self.s == 2 and self.k == 2
            pc = torch.cat((c[:, :1], torch.div(c[:, 1:], 2, rounding_mode="floor")), dim=1)
            # canonical emission order: (b, z, y, x) ascending (torchsparse's own order is unspecified)
            uniq = sorted(set(_key(pc)), key=lambda t: (t[0], t[3], t[2], t[1]))
            table = {t: i for i, t in enumerate(uniq)}
            idx = torch.tensor([table[t] for t in _key(pc)])
            out = torch.zeros(len(uniq), 1).index_add_(0, idx, f * self.kernel.mean())
            return SparseTensor(torch.tensor(uniq, dtype=torch.int32), out, x.stride * 2)

    class ReLU(torch.nn.Module):
        def __init__(self, inplace=True):
            super().__init__()

        def forward(self, x):
            return SparseTensor(x.coords, torch.relu(x.feats), x.stride)

    class _Cfg:
        kmap_mode = "hashgrid"

    tsF.conv_config = types.SimpleNamespace(get_default_conv_config=lambda: _Cfg(),
                                            set_global_conv_config=lambda cfg: None)
    tsnn.Conv3d, tsnn.ReLU, tsnn.functional = Conv3d, ReLU, tsF
    ts.SparseTensor, ts.nn = SparseTensor, tsnn
    sys.modules.update({"torchsparse": ts, "torchsparse.nn": tsnn, "torchsparse.nn.functional": tsF})
    # `out + x` on SparseTensors inside ResNet.forward (kit/nn.py:21)
    SparseTensor.__add__ = lambda a, b: SparseTensor(a.coords, a.feats + b.feats, a.stride)


# ----------------------------------------------------------------------------- torchac stand-in
def _install_torchac_stub():
    tac = types.ModuleType("torchac")

    def encode(cdf, sym):
        q = cdf.numpy().astype(np.uint16).astype(np.int64).tolist()   # int16 reinterpreted as uint16
        s = sym.numpy().astype(np.int64).tolist()
        Lp = len(q[0]) if q else 0
        bits = []
        low, high, pending = 0, 0xFFFFFFFF, 0

        def emit(bit):
            nonlocal pending
            bits.append(bit)
            bits.extend([1 - bit] * pending)
            pending = 0

        for row, si in zip(q, s):
            span = high - low + 1
            c_low = row[si]
            c_high = 0x10000 if si == Lp - 2 else row[si + 1]
            high = ((low - 1) + ((span * c_high) >> 16)) & 0xFFFFFFFF
            low = (low + ((span * c_low) >> 16)) & 0xFFFFFFFF
            while True:
                if high < 0x80000000:
                    emit(0); low = (low << 1) & 0xFFFFFFFF; high = ((high << 1) | 1) & 0xFFFFFFFF
                elif low >= 0x80000000:
                    emit(1); low = (low << 1) & 0xFFFFFFFF; high = ((high << 1) | 1) & 0xFFFFFFFF
                elif low >= 0x40000000 and high < 0xC0000000:
                    pending += 1
                    low = (low << 1) & 0x7FFFFFFF
                    high = ((high << 1) | 0x80000001) & 0xFFFFFFFF
                else:
                    break
        pending += 1
        emit(0 if low < 0x40000000 else 1)
        while len(bits) % 8:
            bits.append(0)
        return np.packbits(np.array(bits, dtype=np.uint8)).tobytes()

    def decode(cdf, stream):
        q = cdf.numpy().astype(np.uint16).astype(np.int64).tolist()
        Lp = len(q[0]) if q else 0
        bits = np.unpackbits(np.frombuffer(stream, dtype=np.uint8)).tolist()
        pos = 0

        def get():
            nonlocal pos
            b = bits[pos] if pos < len(bits) else 0
            pos += 1
            return b

        low, high, value = 0, 0xFFFFFFFF, 0
        for _ in range(32):
            value = ((value << 1) | get()) & 0xFFFFFFFF
        out = []
        for row in q:
            span = high - low + 1
            count = ((((value - low + 1) << 16) - 1) // span) & 0xFFFF
            left, right, si = 0, Lp - 1, None
            while left + 1 < right:
                m = (left + right) // 2
                v = row[m]
                if v < count:
                    left = m
                elif v > count:
                    right = m
                else:
                    si = m
                    break
            if si is None:
                si = left
            out.append(si)
            c_low = row[si]
            c_high = 0x10000 if si == Lp - 2 else row[si + 1]
            high = ((low - 1) + ((span * c_high) >> 16)) & 0xFFFFFFFF
            low = (low + ((span * c_low) >> 16)) & 0xFFFFFFFF
            while True:
                if low >= 0x80000000 or high < 0x80000000:
                    low = (low << 1) & 0xFFFFFFFF; high = ((high << 1) | 1) & 0xFFFFFFFF
                    value = ((value << 1) | get()) & 0xFFFFFFFF
                elif low >= 0x40000000 and high < 0xC0000000:
                    low = (low << 1) & 0x7FFFFFFF
                    high = ((high << 1) | 0x80000001) & 0xFFFFFFFF
                    value = (value - 0x40000000) & 0xFFFFFFFF
                    value = ((value << 1) | get()) & 0xFFFFFFFF
                else:
                    break
        return torch.tensor(out, dtype=torch.int16)

    tac.encode_int16_normalized_cdf = encode
    tac.decode_int16_normalized_cdf = decode
    sys.modules["torchac"] = tac


# ----------------------------------------------------------------------------- helpers
def _force_cpu_factories():
    """kit/nn.py:36,73,75 hard-code device='cuda' in torch.tensor/arange; this container has no GPU."""
    for name in ("tensor", "arange"):
        orig = getattr(torch, name)

        def wrapped(*a, __orig=orig, **kw):
            if kw.get("device") == "cuda":
                kw["device"] = "cpu"
            return __orig(*a, **kw)

        setattr(torch, name, wrapped)


BIG_N, BIG_SEED, BIG_EVERY = 60_000, 21, 128


def main():
    from gauspcc_b200.synth import hac_like_cloud
    from gauspcc_b200.weights import make_synthetic_state_dict

    _install_torchsparse_stub()
    _install_torchac_stub()
    _force_cpu_factories()
    sys.path.insert(0, os.path.join(REF, "src/ai_pcc/GausPcgc"))
    sys.path.insert(0, os.path.join(REF, "src/gs_compress/HAC"))
    import kit.op as op                       # reference, unchanged
    from utils import pcc_utils               # reference, unchanged

    rng = np.random.default_rng(7)

    # ---- 1. kit/op.py and calculate_morton_order
    C = rng.integers(-40, 40, size=(500, 4)).astype(np.int32)
    C[:, 0] = 0
    C[100:120] = C[0:20]                                                   # duplicates: stability
    F = rng.normal(size=(500, 3)).astype(np.float32)
    sC, sF = op.sort_CF(torch.tensor(C), torch.tensor(F))
    cdf_f = np.sort(rng.random(size=(64, 5)).astype(np.float32), axis=1)
    cdf_f[:, 0] = 0
    cdf_f[:8, -1] = 1.0
    cdf_f[8:12] = np.array([0, 0.53723, 0.9, 1.0, 1.0], dtype=np.float32)  # > int16 range: wraps
    cdf_i = op._convert_to_int_and_normalize(torch.tensor(cdf_f), True).numpy()
    streams = [b"\x01\x02", b"", b"\xff", bytes(range(200))]
    packed = op.pack_byte_stream_ls(streams)
    assert op.unpack_byte_stream(packed) == streams
    mo_in = [rng.integers(-3000, 3000, size=(2000, 3)).astype(np.float32),
             rng.integers(0, 17, size=(300, 3)).astype(np.float32),
             hac_like_cloud(5000, 3).astype(np.float32)]
    mo_in = [np.unique(m, axis=0)[rng.permutation(np.unique(m, axis=0).shape[0])] for m in mo_in]
    mo_out = [pcc_utils.calculate_morton_order(torch.tensor(m)).numpy() for m in mo_in]
    np.savez_compressed(os.path.join(HERE, "op_golden.npz"),
                        sort_C=C, sort_F=F, sorted_C=sC.numpy(), sorted_F=sF.numpy(),
                        cdf_float=cdf_f, cdf_int16=cdf_i,
                        packed=np.frombuffer(packed, dtype=np.uint8),
                        stream_lens=np.array([len(s) for s in streams]),
                        **{f"mo_in{i}": m for i, m in enumerate(mo_in)},
                        **{f"mo_out{i}": m for i, m in enumerate(mo_out)})

    # ---- 2. full codec through the reference's own driver
    sd = make_synthetic_state_dict()
    tmp = tempfile.mkdtemp()
    ckpt = os.path.join(tmp, "ckpt.pt")
    torch.save(sd, ckpt)
    probs = []
    orig_softmax = torch.nn.Softmax.forward

    def rec(self, x):
        y = orig_softmax(self, x)
        probs.append(y.detach().numpy().copy())
        return y

    torch.nn.Softmax.forward = rec
    clouds = {
        "hac600": hac_like_cloud(600, 11, extent_log2=9),                  # signed, 5 coded levels
        "blob": np.unique(rng.integers(-9, 9, size=(900, 3)).astype(np.int32), axis=0),
        "hac2500": hac_like_cloud(2500, 5, extent_log2=12),
    }
    out = {}
    for name, xyz in clouds.items():
        probs.clear()
        order = pcc_utils.calculate_morton_order(torch.tensor(xyz.astype(np.float32)))
        xyz_sorted = torch.tensor(xyz.astype(np.float32))[order]          # as HAC does (gaussian_model.py:1108-1109)
        binp = os.path.join(tmp, name, "xyz_pcc.bin")
        r = pcc_utils.compress_point_cloud(xyz_sorted, ckpt, binp)
        enc_probs = [p.copy() for p in probs]
        probs.clear()
        d = pcc_utils.decompress_point_cloud(binp, ckpt)
        dec = d["point_cloud"]
        assert dec.dtype == torch.float32 and r["num_points"] == xyz.shape[0] == d["num_points"]
        blob = open(binp, "rb").read()
        assert r["file_size_bits"] == 8 * len(blob)
        for a, b in zip(enc_probs, probs):                                 # encoder CDFs == decoder CDFs
            assert np.array_equal(a, b)
        got = np.unique(dec.numpy().astype(np.int32), axis=0)
        assert np.array_equal(got, np.unique(xyz, axis=0)), "reference round trip not lossless?!"
        out[f"{name}_xyz"] = xyz
        out[f"{name}_bin"] = np.frombuffer(blob, dtype=np.uint8)
        out[f"{name}_decoded"] = dec.numpy()
        if name != "hac2500":                                              # keep the fixture small
            out[f"{name}_nprob"] = np.array([p.shape[0] for p in enc_probs])
            out[f"{name}_probs"] = np.concatenate([p.reshape(-1) for p in enc_probs])
        print(name, xyz.shape, "bytes", len(blob), "bpp %.3f" % r["bpp"], "streams", len(enc_probs))
    np.savez_compressed(os.path.join(HERE, "codec_golden.npz"), **out)

    # ---- 3. one cloud of realistic size (60 K anchors, 11 coded levels) through the same reference driver.  The fixture keeps what
    # is size-independent or small: stream lengths, base level, hashes of the cloud / the decoded rows, and every 128th
    # row of every probability tensor (the cloud itself is regenerated from its seed and checked against the hash).
    import hashlib
    from gauspcc_b200 import bitstream
    xyz = hac_like_cloud(BIG_N, BIG_SEED)
    probs.clear()
    order = pcc_utils.calculate_morton_order(torch.tensor(xyz.astype(np.float32)))
    xyz_sorted = torch.tensor(xyz.astype(np.float32))[order]
    binp = os.path.join(tmp, "big", "xyz_pcc.bin")
    r = pcc_utils.compress_point_cloud(xyz_sorted, ckpt, binp)
    enc_probs = [p.copy() for p in probs]
    probs.clear()
    d = pcc_utils.decompress_point_cloud(binp, ckpt)
    for a, b in zip(enc_probs, probs):
        assert np.array_equal(a, b)
    dec = d["point_cloud"].numpy()
    assert np.array_equal(np.unique(dec.astype(np.int32), axis=0), np.unique(xyz, axis=0))
    blob = open(binp, "rb").read()
    _, bx, bo, streams = bitstream.read_file(blob)
    np.savez_compressed(os.path.join(HERE, "codec_golden_60k.npz"),
                        n=np.array([BIG_N]), seed=np.array([BIG_SEED]),
                        xyz_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(xyz.astype(np.int32)).tobytes()).digest(), dtype=np.uint8),
                        decoded_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(dec.astype(np.int32)).tobytes()).digest(), dtype=np.uint8),
                        file_bytes=np.array([len(blob)]), stream_lens=np.array([len(s) for s in streams]),
                        base_xyz=bx, base_occ=bo,
                        nprob=np.array([p.shape[0] for p in enc_probs]), prob_cols=np.array([p.shape[1] for p in enc_probs]),
                        probs_every=np.array([BIG_EVERY]),
                        probs=np.concatenate([p[::BIG_EVERY].reshape(-1) for p in enc_probs]))
    print("big", xyz.shape, "bytes", len(blob), "bpp %.3f" % r["bpp"], "streams", len(enc_probs))
    torch.nn.Softmax.forward = orig_softmax
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
