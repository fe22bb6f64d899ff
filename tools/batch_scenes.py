"""BASELINE config[4] (sample): a batch of GausPcc-1K-shaped synthetic scenes sharded BY SCENE across the GPUs of one box.

    python tools/batch_scenes.py [n_scenes] [gpu_coder_chunk]     # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/batch_scenes.py [n_scenes] [gpu_coder_chunk]

Each rank runs the public API (compress_point_cloud / decompress_point_cloud, host buffers, files) on its scenes
(shard.assign_scenes, LPT by size), checks the lossless round trip, and one all_gather_into_tensor collects
{n_points, file_bytes, enc_us, dec_us} per scene.  Scene sizes are log-uniform in [100 K, 600 K].
"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauspcc_b200.shard import pin_rank
N_CPUS = pin_rank(int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))      # before the codec sizes its coder threads
import numpy as np, torch, torch.distributed as dist
from gauspcc_b200 import pcc_utils, shard
from gauspcc_b200.synth import hac_like_cloud
from gauspcc_b200.weights import save_synthetic_checkpoint

def main():
    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    gpu_chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # > 0: container version 2 (GPU chunk coder, opt-in format)
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local); torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)
    sizes = np.exp(rng.uniform(np.log(100e3), np.log(600e3), n_scenes)).astype(np.int64)
    assign = shard.assign_scenes(sizes.tolist(), world)
    tmp = tempfile.mkdtemp(prefix=f"gpcgc_batch_r{rank}_")
    ckpt = save_synthetic_checkpoint(os.path.join(tmp, "GausPcgc", "best_model_ue_4stage_conv.pt"))
    # scenes are generated up front (host) so the timed loop is codec work only
    clouds = {i: torch.tensor(hac_like_cloud(int(sizes[i]), seed=i), dtype=torch.float32) for i in assign[rank]}
    pcc_utils.compress_point_cloud(clouds[assign[rank][0]][:2048], ckpt, os.path.join(tmp, "warm.bin"))      # warm-up like the reference CLI
    if world > 1: dist.barrier()
    torch.cuda.synchronize(dev); t0 = time.perf_counter()
    rows = []
    for i in assign[rank]:
        x = clouds[i]
        order = pcc_utils.calculate_morton_order(x)
        xs = x[order]
        r = pcc_utils.compress_point_cloud(xs, ckpt, os.path.join(tmp, f"s{i}", "xyz_pcc.bin"), gpu_coder_chunk=gpu_chunk)
        d = pcc_utils.decompress_point_cloud(r["output_path"], ckpt)
        pc = d["point_cloud"]
        ok = torch.equal(pc[pcc_utils.calculate_morton_order(pc)].cpu(), xs)
        assert ok, f"scene {i}: round trip not lossless"
        rows.append([int(sizes[i]), r["file_size_bits"] // 8, int(r["enc_time"] * 1e6), int(d["dec_time"] * 1e6)])
    torch.cuda.synchronize(dev); t_local = time.perf_counter() - t0
    local_t = torch.tensor(rows, dtype=torch.int64, device=dev).reshape(-1, 4)
    res = shard.gather_results(local_t, n_scenes, assign)
    tmax = torch.tensor([t_local], dtype=torch.float64, device=dev)
    tall = [torch.zeros_like(tmax) for _ in range(world)]
    if world > 1:
        dist.all_gather(tall, tmax)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    else:
        tall = [tmax]
    if rank == 0:
        r = res.cpu().numpy()
        assert (r[:, 0] == sizes).all()
        print(json.dumps({"config": f"{n_scenes} synthetic scenes, sizes log-uniform 100K-600K, sharded by scene", "n_gpus": world,
                          "container": f"version 2, {gpu_chunk} symbols per chunk" if gpu_chunk else "reference bitstream",
                          "total_points": int(r[:, 0].sum()), "wall_s": round(float(tmax.item()), 3),
                          "Mpoints_s_e2e": round(float(r[:, 0].sum()) / float(tmax.item()) / 1e6, 3),
                          "mean_bpp": round(float((8 * r[:, 1] / r[:, 0]).mean()), 3), "lossless": True,
                          "sum_enc_s": round(float(r[:, 2].sum()) / 1e6, 3), "sum_dec_s": round(float(r[:, 3].sum()) / 1e6, 3),
                          "rank_wall_s": [round(float(t.item()), 3) for t in tall], "cpus_per_rank": N_CPUS}))
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
