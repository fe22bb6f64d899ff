"""kernel_size = 3 (the default of the reference's stand-alone scripts, compress_ue_4stage_conv.py:44) on the bench scene: device-timed
encode + decode and the public-API round trip, next to kernel_size = 5.    python tools/k3_times.py [n_points]"""
import json, os, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200 import pcc_utils
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict, save_synthetic_checkpoint

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
x = torch.tensor(hac_like_cloud(n, 0), dtype=torch.float32, device=dev)
x = x[pcc_utils.calculate_morton_order(x)]
out = {"n_points": n}
tmp = tempfile.mkdtemp(prefix="gpcgc_k3_")
for K in (5, 3):
    codec = GausPcgcCodec(DeviceWeights(make_synthetic_state_dict(kernel_size=K), dev, kernel_size=K), dev)
    for _ in range(3):
        r = codec.encode(x, download=False)
        occs = [lv.occ for lv in r[3]["levels"][1:]]
        codec.decode(r[0], r[1], [b""] * (4 * len(occs)), forced_occ=occs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        r = codec.encode(x, download=False)
        occs = [lv.occ for lv in r[3]["levels"][1:]]
        codec.decode(r[0], r[1], [b""] * (4 * len(occs)), forced_occ=occs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    ckpt = save_synthetic_checkpoint(os.path.join(tmp, f"k{K}", "best_model_ue_4stage_conv.pt"), kernel_size=K)
    xh = x.cpu()
    walls = []
    for it in range(3):
        t0 = time.perf_counter()
        c = pcc_utils.compress_point_cloud(xh, ckpt, os.path.join(tmp, f"k{K}", "xyz_pcc.bin"), kernel_size=K)
        d = pcc_utils.decompress_point_cloud(c["output_path"], ckpt, kernel_size=K, sorted_output=True)
        pts = d["point_cloud"].cpu()
        walls.append(time.perf_counter() - t0)
    assert torch.equal(pts, xh)
    out[f"kernel_size_{K}"] = {"ms_per_step_device": round(ms, 2), "Mpoints_s_device": round(n / ms / 1e3, 3),
                               "e2e_Mpoints_s": round(n / min(walls[1:]) / 1e6, 3), "enc_s": round(c["enc_time"], 4),
                               "dec_s": round(d["dec_time"], 4), "bpp": round(c["bpp"], 3)}
print(json.dumps(out))
