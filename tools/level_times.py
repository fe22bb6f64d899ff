"""Per-level CUDA-event times of the encode's stages (kernel maps, 5-conv stacks, 4 x (2 convs + head)), for the default kernel choice and
with the big dense levels forced onto the mma.sync conv (A/B of the tcgen05 conv in situ).  Run on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200 import weights as W
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec, Level
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict
from gauspcc_b200.pcc_utils import calculate_morton_order

dev = torch.device("cuda:0")
w = DeviceWeights(make_synthetic_state_dict(), dev)
x = torch.tensor(hac_like_cloud(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 0), dtype=torch.float32, device=dev)
x = x[calculate_morton_order(x)]


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


for label, um_min in (("default (tcgen05 on big dense levels)", 150_000), ("mma.sync on big dense levels", 1 << 40)):
    codec = GausPcgcCodec(w, dev)
    codec.um_min_rows = um_min
    keys, meta = codec.pack_keys(x); mm = meta.cpu().numpy()[2:8]
    leaf = codec.sort_unique(keys, mm.astype(np.uint32))
    levels = codec.build_pyramid(leaf, mm.astype(np.int64))
    print("====", label)
    tot = {"kmap": 0.0, "prior": 0.0, "target": 0.0, "stages": 0.0}
    for d in range(len(levels) - 1):
        parent, gt = levels[d], levels[d + 1]
        t_km, km = timed(lambda: codec.build_kmap(gt.keys))
        gt.kmap = km
        if parent.kmap is None:
            parent.kmap = codec.build_kmap(parent.keys)
        pum = bool(parent.kmap.um_rows)
        f0 = torch.randn((parent.n, 32), device=dev)
        f0 = codec.split_rows(f0) if pum else f0
        t_prior, _ = timed(lambda: codec.res_stack(f0, W.PRIOR_CONVS, parent.kmap))
        cum = bool(km.um_rows)
        u0 = torch.randn((gt.n, 32), device=dev)
        u0 = codec.split_rows(u0) if cum else u0
        t_target, u = timed(lambda: codec.res_stack(u0, W.TARGET_CONVS, km, final="both" if cum else "f32"))
        lohi = torch.empty((gt.n,), dtype=torch.int32, device=dev)
        def stages():
            for i in range(4):
                codec.stage_cdf(u, gt.occ, i, km, None, lohi_out=lohi)
        t_st, _ = timed(stages)
        fam = "um" if km.um_rows else ("sparse" if km.sparse else f"v6/{km.v6_variant}/{km.tile_rows}")
        print(f"child n={gt.n:8d} {fam:12s} p/r={km.n_real / max(gt.n, 1):5.1f}  kmap {t_km:6.3f}  prior({parent.n}) {t_prior:6.3f}  target {t_target:6.3f}  4 stages {t_st:6.3f} ms")
        tot["kmap"] += t_km; tot["prior"] += t_prior; tot["target"] += t_target; tot["stages"] += t_st
    print("totals (one direction):", {k: round(v, 2) for k, v in tot.items()}, "sum", round(sum(tot.values()), 2))
