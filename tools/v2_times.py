"""Where container version 2 (GPU chunk coder) spends its time: encode / decode wall and CUDA-event times per phase, chunk kernels
timed alone on the largest level.    python tools/v2_times.py [n_points] [chunk]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200.codec import DeviceWeights, GausPcgcCodec
from gauspcc_b200.pcc_utils import calculate_morton_order
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import make_synthetic_state_dict

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda:0")
codec = GausPcgcCodec(DeviceWeights(make_synthetic_state_dict(), dev), dev)
x = torch.tensor(hac_like_cloud(n, 0), dtype=torch.float32, device=dev)
x = x[calculate_morton_order(x)]


def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return min(t) * 1e3, r


t1, r1 = wall(lambda: codec.encode(x))
t2, r2 = wall(lambda: codec.encode(x, gpu_chunk=chunk))
print(f"encode  v1 {t1:.1f} ms   v2 {t2:.1f} ms   bytes v1 {sum(map(len, r1[2]))} v2 {sum(map(len, r2[2]))}")
d1, _ = wall(lambda: codec.decode(r1[0], r1[1], r1[2]))
d2, _ = wall(lambda: codec.decode(r2[0], r2[1], r2[2], gpu_chunk=chunk))
print(f"decode  v1 {d1:.1f} ms   v2 {d2:.1f} ms")
# chunk decode kernel alone on the largest level
_, _, _, aux = codec.encode(x, collect=True, gpu_chunk=chunk)
k = len(aux["cdfs"]) - 1
lv = aux["levels"][-1]
v2_dev, v2_off = codec._chunk_upload(r2[2])
for i, A in enumerate((2, 2, 4, 16)):
    cdf = aux["cdfs"][k - 3 + i]
    stream = r2[2][k - 3 + i]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    codec._chunk_decode(cdf, stream, v2_dev, v2_off[k - 3 + i], lv.n, A + 1, chunk); torch.cuda.synchronize()
    e0.record(); codec._chunk_decode(cdf, stream, v2_dev, v2_off[k - 3 + i], lv.n, A + 1, chunk); e1.record(); torch.cuda.synchronize()
    print(f"  level n={lv.n} stage {i} (A={A}): chunk decode {e0.elapsed_time(e1):.3f} ms, stream {len(stream)} B")
