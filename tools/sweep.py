"""BASELINE config[2]: Morton-order + kernel-map build sweep, 1M-32M voxels on one B200 (HBM GB/s vs roofline).

    python tools/sweep.py [sizes in M, default 1,2,4,8,16,32]

Algorithmic bytes: order  N*(12+8) + passes*N*2*(8+4), passes = ceil(key bits / 8);
hash+kmap  n*8 + cap*16 (build) + n*124*16 (probe reads) + n*125*4 (dense map) + entries*8 (pair stream).
"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.pcc_utils import calculate_morton_order
from gauspcc_b200.weights import make_synthetic_state_dict

def timed(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

def main():
    sizes = [int(float(s) * 1e6) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2,4,8,16,32".split(","))]
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    codec = GausPcgcCodec(DeviceWeights(make_synthetic_state_dict(), dev), dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for n_target in sizes:
        ext = 1 << 17
        raw = torch.randint(-ext // 2, ext // 2, (int(n_target * 1.02) + 1024, 3), generator=g, device=dev, dtype=torch.int32)
        keys, meta = codec.pack_keys(raw)
        mm = meta.cpu().numpy()[2:8]
        leaf = codec.sort_unique(keys, mm.astype(np.uint32))
        perm = torch.randperm(leaf.shape[0], device=dev, generator=g)[:n_target]
        n = int(perm.shape[0])
        xyz = torch.empty((n, 3), dtype=torch.int32, device=dev)
        from gauspcc_b200.codec import _ptr
        shuffled = leaf[perm].contiguous()
        codec._call("gpc_unpack_keys_i32", _ptr(shuffled), n, _ptr(xyz), codec._stream())
        xf = xyz.float()
        # ---- a-2: calculate_morton_order (includes its min/max pass, one host sync, idx widening)
        ms_order, idx = timed(lambda: calculate_morton_order(xf))
        bits = 3 * 17
        passes = (bits + 7) // 8
        bytes_order = n * (12 + 8) + passes * n * 2 * 12
        srt = xf[idx].to(torch.int64)
        key = (srt[:, 2] + (1 << 20)) * (1 << 42) + (srt[:, 1] + (1 << 20)) * (1 << 21) + (srt[:, 0] + (1 << 20))
        assert bool((key[1:] > key[:-1]).all())
        # ---- kernel map on the sorted set
        skeys = torch.sort(shuffled)[0]
        def build():
            return codec.build_kmap(skeys)
        ms_kmap, km = timed(build, reps=2)
        cap = codec.lib.gpc_hash_capacity(n)
        bytes_kmap = n * 8 + cap * 16 + n * 124 * 16 + n * 125 * 4 * 2 + km.n_pairs * 8
        print(json.dumps({"n": n, "order_ms": round(ms_order, 3), "order_Mpts_s": round(n / ms_order / 1e3, 1),
                          "order_GBs_radix_model": round(bytes_order / ms_order / 1e6, 1), "order_frac_of_peak": round(bytes_order / ms_order / 1e6 / peak, 3),
                          "order_GBs_compulsory": round(n * 20 / ms_order / 1e6, 1), "radix_passes": passes,
                          "kmap_ms": round(ms_kmap, 3), "kmap_Mrows_s": round(n / ms_kmap / 1e3, 1), "kmap_Gprobes_s": round(n * 124 / ms_kmap / 1e6, 1),
                          "kmap_GBs_model": round(bytes_kmap / ms_kmap / 1e6, 1), "kmap_frac_of_peak": round(bytes_kmap / ms_kmap / 1e6 / peak, 3),
                          "pair_entries": km.n_pairs}), flush=True)
        del raw, keys, leaf, perm, xyz, xf, idx, srt, key, skeys, km
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
