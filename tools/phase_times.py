"""Where does one bench step spend its time?  Wall + CUDA-event time per phase (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict
from gauspcc_b200.pcc_utils import calculate_morton_order

dev = torch.device("cuda:0")
codec = GausPcgcCodec(DeviceWeights(make_synthetic_state_dict(), dev), dev)
x = torch.tensor(hac_like_cloud(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 0), dtype=torch.float32, device=dev)
x = x[calculate_morton_order(x)]

class T:
    def __init__(s, name): s.name = name
    def __enter__(s):
        torch.cuda.synchronize(); s.t = time.perf_counter(); s.e0 = torch.cuda.Event(enable_timing=True); s.e1 = torch.cuda.Event(enable_timing=True); s.e0.record(); return s
    def __exit__(s, *a):
        s.e1.record(); torch.cuda.synchronize(); print(f"{s.name:28s} wall {1e3*(time.perf_counter()-s.t):8.2f} ms  gpu {s.e0.elapsed_time(s.e1):8.2f} ms")

for rep in range(3):
    print("--- rep", rep)
    codec.conv_profile = [] if rep == 2 else None
    codec._ev_next = 0
    with T("encode total"):
        bx, bo, _, aux = codec.encode(x, download=False)
    occs = [lv.occ for lv in aux["levels"][1:]]
    with T("decode total (forced)"):
        codec.decode(bx, bo, [b""] * (4 * len(occs)), forced_occ=occs)
    with T("pack+sort+unique"):
        keys, meta = codec.pack_keys(x); mm = meta.cpu().numpy()[2:8]; leaf = codec.sort_unique(keys, mm.astype(np.uint32))
    with T("pyramid"):
        levels = codec.build_pyramid(leaf, mm.astype(np.int64))
    with T("kmaps all levels"):
        kms = [codec.build_kmap(l.keys) for l in levels]
    big = levels[-3]
    with T(f"kmap n={big.n}"):
        codec.build_kmap(big.keys)
    f = torch.randn((big.n, 32), device=dev)
    with T("res_stack (5 convs)"):
        codec.res_stack(f, (0, 1, 2, 3, 4), kms[-3])
    with T("expand"):
        codec.expand(levels[-3], levels[-2].n)
    cdf = torch.empty((big.n, 17), dtype=torch.int16, device=dev)
    with T("stage_cdf i=3"):
        codec.stage_cdf(f, big.occ, 3, kms[-3], cdf)
