"""GausPcgc `Network` state_dict layout, loader and seeded synthetic weights.

The reference rebuilds `Network(channels, kernel_size)` and `torch.load`s the checkpoint on
every call (reference: src/gs_compress/HAC/utils/pcc_utils.py:65-67, 266-268).  The key /
shape layout below is the one `Network.state_dict()` produces
(src/ai_pcc/GausPcgc/network_ue_4stage_conv.py:12-98, kit/nn.py:14-15,31,106); it is pinned by
tests/golden/make_golden.py, which instantiates the reference module itself.

No real GausPcgc checkpoint ships with the reference (README.md:73-77 is a Baidu-pan link), so
tests and the bench use `make_synthetic_state_dict` (SURVEY.md §8d recipe).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

STAGE_ALPHABETS = (2, 2, 4, 16)          # pred_head_s{0..3} outputs (1b, 1b, 2b, 4b)
STAGE_CONTEXTS = (0, 2, 4, 16)           # pred_head_s{1,2,3}_emb rows (stage 0 has none)

# order in which the 18 sparse-conv kernels are packed for the device
CONV_KEYS = (
    "prior_resnet.0.kernel",
    "prior_resnet.2.conv0.kernel", "prior_resnet.2.conv1.kernel",
    "prior_resnet.3.conv0.kernel", "prior_resnet.3.conv1.kernel",
    "target_resnet.0.kernel",
    "target_resnet.2.conv0.kernel", "target_resnet.2.conv1.kernel",
    "target_resnet.3.conv0.kernel", "target_resnet.3.conv1.kernel",
    "spatial_conv_s0.0.kernel", "spatial_conv_s0.2.kernel",
    "spatial_conv_s1.0.kernel", "spatial_conv_s1.2.kernel",
    "spatial_conv_s2.0.kernel", "spatial_conv_s2.2.kernel",
    "spatial_conv_s3.0.kernel", "spatial_conv_s3.2.kernel",
)
PRIOR_CONVS = tuple(range(0, 5))
TARGET_CONVS = tuple(range(5, 10))


def stage_convs(i: int) -> Tuple[int, int]:
    return 10 + 2 * i, 11 + 2 * i


def reference_layout(channels: int = 32, kernel_size: int = 5) -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of every tensor in the reference `Network.state_dict()`."""
    C, K3 = channels, kernel_size ** 3
    lay: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    lay["prior_embedding.weight"] = (256, C)
    for k in CONV_KEYS[0:5]:
        lay[k] = (K3, C, C)
    lay["target_embedding.target_res_embedding.weight"] = (8, C)
    for k in CONV_KEYS[5:10]:
        lay[k] = (K3, C, C)
    for k in CONV_KEYS[10:18]:
        lay[k] = (K3, C, C)
    for i, A in enumerate(STAGE_ALPHABETS):
        if i > 0:
            lay[f"pred_head_s{i}_emb.weight"] = (STAGE_CONTEXTS[i], C)
        lay[f"pred_head_s{i}.0.weight"] = (C, C)
        lay[f"pred_head_s{i}.0.bias"] = (C,)
        lay[f"pred_head_s{i}.2.weight"] = (A, C)
        lay[f"pred_head_s{i}.2.bias"] = (A,)
    lay["fog.conv.kernel"] = (8, 1, 1)
    return lay


def make_synthetic_state_dict(seed: int = 5, channels: int = 32, kernel_size: int = 5) -> Dict[str, torch.Tensor]:
    """Seeded stand-in weights in the reference's key/shape layout.

    Conv kernels are uniform with a per-offset scale: the centre offset carries a second-moment gain
    of 2 (He gain, so isolated voxels keep their signal through the ReLU stacks) and the other K^3-1
    offsets together add at most 2 more when every neighbour is occupied.  With the flat
    U(-sqrt(3/(8C)), +) of SURVEY.md §8d the 13-conv-deep path saturates every probability to 0/1 on
    the coarse levels (20-40 occupied neighbours) and decays to the uniform distribution on the fine
    ones (1-2 neighbours) -- measured with the oracle -- which makes a 1e-3 parity bound on
    probabilities meaningless; this profile keeps entropies mid-range on all levels, so the parity
    tests are sensitive to conv errors everywhere.  Embeddings N(0,1); nn.Linear default init;
    fog.conv.kernel = 1.
    """
    g = torch.Generator().manual_seed(seed)
    C = channels
    sd: Dict[str, torch.Tensor] = OrderedDict()
    K3 = kernel_size ** 3
    scale = torch.full((K3, 1, 1), math.sqrt(3.0 * 2.0 / (C * max(K3 - 1, 1))))
    scale[(K3 - 1) // 2] = math.sqrt(3.0 * 2.0 / C)
    for name, shape in reference_layout(channels, kernel_size).items():
        if name == "fog.conv.kernel":
            t = torch.ones(shape, dtype=torch.float32)
        elif name.endswith(".kernel"):
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * scale
        elif "embedding" in name or name.endswith("_emb.weight"):
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        else:  # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
            bound = 1.0 / math.sqrt(C)
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        sd[name] = t
    return sd


def save_synthetic_checkpoint(path: str, seed: int = 5, channels: int = 32, kernel_size: int = 5) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(make_synthetic_state_dict(seed, channels, kernel_size), path)
    return path


def validate_state_dict(sd: Dict[str, torch.Tensor], channels: int = 32, kernel_size: int = 5) -> None:
    """Same failure mode as `net.load_state_dict` in the reference (RuntimeError on mismatch)."""
    lay = reference_layout(channels, kernel_size)
    missing = [k for k in lay if k not in sd]
    unexpected = [k for k in sd if k not in lay]
    bad = [f"{k}: {tuple(sd[k].shape)} vs {lay[k]}" for k in lay if k in sd and tuple(sd[k].shape) != lay[k]]
    if missing or unexpected or bad:
        raise RuntimeError(
            "Error(s) in loading state_dict for Network: "
            f"missing={missing} unexpected={unexpected} size_mismatch={bad}")


def state_dict_to_numpy(sd: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
    return {k: v.detach().to("cpu", torch.float32).contiguous().numpy() for k, v in sd.items()}
