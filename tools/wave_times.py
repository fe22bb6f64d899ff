"""Per-level wall time of a real decode (host range decoder in the loop), wavefront on / off (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict

dev = torch.device("cuda:0")
codec = GausPcgcCodec(DeviceWeights(make_synthetic_state_dict(), dev), dev)
x = torch.tensor(hac_like_cloud(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 0), dtype=torch.float32, device=dev)
bx, bo, streams, _ = codec.encode(x)
orig_lf, orig_wave = codec.level_features, codec._decode_level_wavefront
rec = []

def lf(parent, n_child, child_kmap=None):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = orig_lf(parent, n_child, child_kmap)
    torch.cuda.synchronize(); rec.append(["features", n_child, time.perf_counter() - t])
    return out

def wave(u, child, n, st, occ):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = orig_wave(u, child, n, st, occ)
    torch.cuda.synchronize(); rec.append(["wave", n, time.perf_counter() - t])
    return out

for mode in (True, False, True):
    codec.wave_decode = mode
    codec.level_features, codec._decode_level_wavefront = lf, wave
    codec.wave_log = []
    rec.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = codec.decode(bx, bo, streams)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(f"--- wavefront {mode}: decode {tot * 1e3:.1f} ms, stats {codec.last_stats}")
    feats = sum(r[2] for r in rec if r[0] == "features"); waves = sum(r[2] for r in rec if r[0] == "wave")
    print(f"    level_features {feats * 1e3:.1f} ms, wavefront levels {waves * 1e3:.1f} ms, rest {(tot - feats - waves) * 1e3:.1f} ms")
    for r, w in zip([r for r in rec if r[0] == "wave"], codec.wave_log):
        print(f"    wave n={r[1]:8d} chunks={w['chunks']:3d} sparse={w['sparse']} wall {r[2] * 1e3:6.1f} ms  decoder seconds per stage {w['ac_s']}")
