import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict
from gauspcc_b200.pcc_utils import calculate_morton_order
dev = torch.device("cuda:0")
codec = GausPcgcCodec(DeviceWeights(make_synthetic_state_dict(), dev), dev)
x = torch.tensor(hac_like_cloud(1_000_000, 0), dtype=torch.float32, device=dev)
x = x[calculate_morton_order(x)]
def step():
    bx, bo, _, aux = codec.encode(x, download=False)
    occs = [lv.occ for lv in aux["levels"][1:]]
    codec.decode(bx, bo, [b""] * (4 * len(occs)), forced_occ=occs)
for _ in range(3): step()
torch.cuda.synchronize()
import gc
for mode in ("plain", "profile", "plain2", "sampler+profile"):
    if mode == "gc_off": gc.disable()
    th = None
    codec.conv_profile = None
    if "profile" in mode:
        codec.prewarm_profile_events(4000)
    if "sampler" in mode:
        sys.path.insert(0, "/root/repo")
        import bench
        th = bench.ClockSampler(0); th.start()
    ts = []
    for i in range(8):
        if "profile" in mode:
            codec.conv_profile = []
        torch.cuda.synchronize(); t0 = time.perf_counter()
        step()
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    if th: th.stop_flag.set(); th.join()
    print(mode, " ".join(f"{t:.1f}" for t in ts), "mem reserved GB %.1f" % (torch.cuda.memory_reserved() / 2**30))

# ---- back-to-back steps as bench.py runs them (no sync between steps; optionally keeping the previous step's aux alive)
def step2():
    bx, bo, _, aux = codec.encode(x, download=False)
    occs = [lv.occ for lv in aux["levels"][1:]]
    out = codec.decode(bx, bo, [b""] * (4 * len(occs)), forced_occ=occs)
    return out, aux
for mode in ("b2b", "b2b+aux", "b2b+aux+profile"):
    codec.conv_profile = None
    if "profile" in mode:
        codec.prewarm_profile_events(6000); codec.conv_profile = []
    res = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        keep = None
        for i in range(5):
            if "aux" in mode:
                out, keep = step2()
            else:
                step()
        torch.cuda.synchronize(); res.append((time.perf_counter() - t0) * 1e3 / 5)
    print(mode, " ".join(f"{t:.1f}" for t in res))
