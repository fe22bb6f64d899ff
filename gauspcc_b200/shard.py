"""Multi-GPU: the path shards by scene.  One process per GPU, weights replicated
(9.3 MB), scenes assigned by longest-processing-time-first, no collective inside the data path; one
small all_gather of per-scene results {n_points, file_bytes, t_enc_us, t_dec_us} at the end.

The reference has no distributed code at all (every script pins one GPU, e.g.
scripts/gs_compress/run_ours_hac.sh:7); this module is the B200-box equivalent of running its
per-scene loop on 8 GPUs.
"""
from __future__ import annotations

import os
from typing import List, Sequence


def pin_rank(local_rank: int, world_size: int) -> int:
    """Give this rank its own slice of the host CPUs (cores / world, contiguous in the affinity mask) and return its size.
    Call it before the codec is created: the range-coder thread pool is sized from the affinity mask.  Eight unpinned ranks
    x (16 coder threads + the launch thread) on one box's cores is what bent the end-to-end scaling curve in round 1."""
    cpus = sorted(os.sched_getaffinity(0))
    if world_size <= 1 or len(cpus) < 2 * world_size:
        return len(cpus)
    per = len(cpus) // world_size
    mine = cpus[local_rank * per:(local_rank + 1) * per]
    os.sched_setaffinity(0, mine)
    return len(mine)


def assign_scenes(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy LPT: scenes sorted by size descending, each to the currently lightest rank.
    Deterministic (ties by scene index), identical on every rank without communication."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i])
    return out


def gather_results(local, n_scenes: int, assignment: List[List[int]]):
    """local: int64 [len(assignment[rank]), F] on this rank's device -> int64 [n_scenes, F] on every rank.
    Ragged shards are padded to the longest one so a single all_gather_into_tensor suffices."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    F = local.shape[1]
    longest = max(len(a) for a in assignment) if assignment else 0
    pad = torch.zeros((longest, F), dtype=torch.int64, device=local.device)
    pad[: local.shape[0]] = local
    if world > 1:
        allr = torch.empty((world * longest, F), dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(allr, pad)
    else:
        allr = pad
    out = torch.zeros((n_scenes, F), dtype=torch.int64, device=local.device)
    for r in range(world):
        idx = assignment[r]
        if idx:
            out[torch.tensor(idx, device=local.device)] = allr[r * longest: r * longest + len(idx)]
    assert len(assignment[rank]) == local.shape[0]
    return out


# ----------------------------------------------------------------------------- inside one scene: Morton-ordered spatial blocks
# The drop-in file (xyz_pcc.bin) codes a scene as ONE octree: every 5^3 context crosses the whole scene, so one scene is one
# GPU's work.  For a scene that should be spread over several GPUs the rows -- already in
# calculate_morton_order = ascending (z, y, x) order -- are cut into n_blocks consecutive ranges, i.e. z slabs, and every block
# is coded as an independent point cloud with the unchanged codec: one file per block, each in the reference's own layout and
# decodable by the reference's decompress_point_cloud.  Not the reference's single file: contexts stop at the block faces (a
# few per cent more bits, measured in tools/block_split.py) and there are n_blocks headers; it is an opt-in beside the drop-in.
def block_ranges(n_rows: int, n_blocks: int) -> List[tuple]:
    """n_blocks consecutive row ranges of (almost) equal size covering [0, n_rows)"""
    n_blocks = max(1, min(int(n_blocks), max(int(n_rows), 1)))
    cuts = [(n_rows * b) // n_blocks for b in range(n_blocks + 1)]
    return [(cuts[b], cuts[b + 1]) for b in range(n_blocks)]


def block_path(output_path: str, b: int) -> str:
    stem, ext = os.path.splitext(output_path)
    return f"{stem}_blk{b}{ext}"


def compress_point_cloud_blocks(xyz_sorted, ckpt_path: str, output_path: str, n_blocks: int, rank: int = 0, world_size: int = 1, **kw):
    """Code the blocks b with b % world_size == rank of a scene given in calculate_morton_order order; block b goes to
    block_path(output_path, b).  Returns {"blocks": [(b, rows, file_size_bits, enc_time)], "file_size_bits": this rank's sum}."""
    from . import pcc_utils
    out = []
    for b, (r0, r1) in enumerate(block_ranges(xyz_sorted.shape[0], n_blocks)):
        if b % world_size != rank:
            continue
        r = pcc_utils.compress_point_cloud(xyz_sorted[r0:r1], ckpt_path, block_path(output_path, b), **kw)
        out.append((b, r1 - r0, int(r["file_size_bits"]), float(r["enc_time"])))
    return {"blocks": out, "file_size_bits": sum(o[2] for o in out)}


def decompress_point_cloud_blocks(output_path: str, ckpt_path: str, n_blocks: int, rank: int = 0, world_size: int = 1, **kw):
    """Decode this rank's blocks; each block comes back in (z, y, x) order, so the blocks of all ranks concatenated by block index
    are the scene in calculate_morton_order order.  Returns {"blocks": {b: point_cloud}, "dec_time": sum}."""
    from . import pcc_utils
    clouds, t = {}, 0.0
    for b in range(n_blocks):
        if b % world_size != rank:
            continue
        d = pcc_utils.decompress_point_cloud(block_path(output_path, b), ckpt_path, sorted_output=True, **kw)
        clouds[b] = d["point_cloud"]
        t += float(d["dec_time"])
    return {"blocks": clouds, "dec_time": t}
