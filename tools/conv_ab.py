"""A/B timing of the sparse-conv variants on the levels of a synthetic scene (run on the GPU box)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict

def main():
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    configs = [tuple(map(int, c.split(":"))) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0:256", "1:256", "1:512", "2:512", "2:1024", "3:256", "3:512"])]
    dev = torch.device("cuda:0")
    w = DeviceWeights(make_synthetic_state_dict(), dev)
    xyz = torch.tensor(hac_like_cloud(npts, 0), dtype=torch.float32, device=dev)
    base = GausPcgcCodec(w, dev, tile_rows=256)
    keys, meta = base.pack_keys(xyz)
    mm = meta.cpu().numpy()[2:8]
    leaf = base.sort_unique(keys, mm.astype(np.uint32))
    levels = base.build_pyramid(leaf, mm.astype(np.int64))
    sel = [l for l in levels if l.n >= int(os.environ.get("AB_MIN_ROWS", 2000)) and l.n <= int(os.environ.get("AB_MAX_ROWS", 1 << 30))]
    ref = {}
    print("levels:", [l.n for l in levels])
    for variant, tr in configs:
        codec = GausPcgcCodec(w, dev, tile_rows=tr)
        codec.conv_variant = variant
        if variant == 200:                  # centre + stragglers conv on every level
            codec.conv_variant, codec.sparse_min_rows, codec.sparse_max_density = 42, 1, 1e9
        elif variant >= 100:
            codec.tc_cta_rows, codec.tc_min_density = tr, 0.0
        tot_ms, tot_pairs = 0.0, 0
        line = []
        for li, lv in enumerate(sel):
            codec.conv_profile = []
            codec._seen_sparse = variant == 200
            km = codec.build_kmap(lv.keys)
            codec.conv_profile = None
            g = torch.Generator(device=dev).manual_seed(li)
            x = torch.randn((lv.n, 32), device=dev, generator=g)
            xin = codec.split_rows(x) if km.cta_rows else x
            y = codec.conv(xin, 3, km, relu=True)
            torch.cuda.synchronize()
            if li not in ref:
                ref[li] = y.clone()
            err = float((y - ref[li]).abs().max())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            if variant == 101:
                codec.lib.gpc_debug_conv_tc_profile(None, 1)
            e0.record()
            for _ in range(reps):
                codec.conv(xin, 3, km, relu=True, out=y)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            tot_ms += ms; tot_pairs += km.n_real
            clk = ms * 1e-3 * 1.9e9 * 148 / max(km.n_real, 1)
            if variant == 101 and km.cta_rows:
                import ctypes
                buf = (ctypes.c_uint64 * 32)()
                codec.lib.gpc_debug_conv_tc_profile(ctypes.cast(buf, ctypes.c_void_p), 1)
                ch = max(buf[9], 1)
                names = ["g.empty", "g.issue", "g.wait", "m.full", "m.dempty", "m.issue", "e.dfull", "e.ld", "e.rmw"]
                ctas = reps * ((lv.n + km.cta_rows - 1) // km.cta_rows)
                print(f"   prof n={lv.n}: chunks/cta={ch / ctas:.0f} " + " ".join(f"{nm}={buf[i] / ch:.0f}" for i, nm in enumerate(names))
                      + f" | per cta: total={buf[10] / ctas:.0f} setup={buf[11] / ctas:.0f} writeout={buf[12] / ctas:.0f}")
            line.append(f"n={lv.n} tiles={km.n_tiles} p/r={km.n_real / lv.n:.1f} {ms:.3f}ms {clk:.1f}clk/pair err={err:.1e}")
        print(f"variant {variant} tile {tr}: total {tot_ms:.2f} ms, {tot_ms * 1e-3 * 1.9e9 * 148 / tot_pairs:.1f} clk/pair/SM")
        for l in line:
            print("    ", l)

if __name__ == "__main__":
    main()
