"""Lossless round trip through the public API on clouds that stress different conv paths (run on the GPU box):
dense 3-D blob (up to 125 pairs per row), ultra-sparse cloud (every big level on the centre + stragglers conv, some with no
straggler at all), 2M-anchor HAC-like scene."""
import os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200 import pcc_utils
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import save_synthetic_checkpoint

def main():
    tmp = tempfile.mkdtemp(prefix="gpcgc_rt_")
    ckpt = save_synthetic_checkpoint(os.path.join(tmp, "GausPcgc", "best_model_ue_4stage_conv.pt"))
    rng = np.random.default_rng(0)
    clouds = {
        "dense blob 400K in 2^7 cube": np.unique(rng.integers(-64, 64, size=(500_000, 3)).astype(np.int32), axis=0)[:400_000],
        "ultra sparse 400K in 2^20 cube": np.unique(rng.integers(-(1 << 19), 1 << 19, size=(400_000, 3)).astype(np.int32), axis=0),
        "HAC-like 2M": hac_like_cloud(2_000_000, 7),
    }
    for name, xyz in clouds.items():
        x = torch.tensor(xyz, dtype=torch.float32, device="cuda")
        x = x[pcc_utils.calculate_morton_order(x)]
        binp = os.path.join(tmp, "xyz_pcc.bin")
        r = pcc_utils.compress_point_cloud(x, ckpt, binp)
        d = pcc_utils.decompress_point_cloud(binp, ckpt, sorted_output=True)
        dec = d["point_cloud"]
        ok = dec.shape == x.shape and bool(torch.equal(dec, x))          # x is already in (z,y,x) order: bit-exact rows
        print(f"{name}: n={xyz.shape[0]} bpp={r['bpp']:.2f} enc={r['enc_time']:.3f}s dec={d['dec_time']:.3f}s lossless={ok}", flush=True)
        assert ok

if __name__ == "__main__":
    main()
