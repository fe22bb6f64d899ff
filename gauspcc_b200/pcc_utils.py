"""Drop-in replacement for the reference's `utils/pcc_utils.py` (the L3 boundary).

Same three entry points, argument meaning, return dicts and `xyz_pcc.bin` layout as
/root/reference/src/gs_compress/HAC/utils/pcc_utils.py:
    calculate_morton_order(x)                                             :12-22
    compress_point_cloud(xyz_quantized, ckpt_path, output_path, ...)      :24-217
    decompress_point_cloud(bin_file_path, ckpt_path, output_path=None...) :230-400
so HAC / HAC++ / CAT-3DGS / TC-GS `scene/gaussian_model.py` can switch with
`from gauspcc_b200.pcc_utils import ...`.  Underneath: libgpcgc.so (hand-written
sm_100a kernels behind the C ABI of include/gpcgc.h).  No torchsparse, no torchac, no Triton, no CPU
fallback -- without CUDA or without the built library these functions raise.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Dict

import numpy as np
import torch

from . import _lib, bitstream
from .codec import GausPcgcCodec, load_weights

_CODECS: Dict[tuple, GausPcgcCodec] = {}


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.GpcError("gauspcc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _codec(ckpt_path: str, channels: int, kernel_size: int) -> GausPcgcCodec:
    dev = _require_cuda()
    w = load_weights(ckpt_path, dev, channels, kernel_size)
    key = (id(w), str(dev))
    if key not in _CODECS:
        _CODECS[key] = GausPcgcCodec(w, dev)
    return _CODECS[key]


def calculate_morton_order(x: torch.Tensor) -> torch.Tensor:
    """
    Calculate Morton order of the input points.

    Same contract as the reference (pcc_utils.py:12-22): returns the int64 permutation, on `x.device`,
    that sorts the rows by x + y*M + z*M^2 after subtracting the per-axis minimum, i.e. ascending
    (z, y, x).  Ties (duplicate voxels) keep their input order (the reference leaves them unspecified).
    The sort runs on the GPU (key pack + LSD radix sort in libgpcgc); a CPU tensor is moved to the
    current CUDA device and the result moved back.
    """
    assert len(x.shape) == 2 and x.shape[1] == 3, f'Input data must be a 3D point cloud, but got {x.shape}.'
    lib = _lib.load()
    dev = _require_cuda()
    src_device = x.device
    xd = x.detach()
    if xd.is_floating_point():
        xd = xd.to(device=dev, dtype=torch.float32)
        is_f32 = 1
    else:
        xd = xd.to(device=dev, dtype=torch.int32)
        is_f32 = 0
    xd = xd.contiguous()
    n = xd.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=dev)
    if n:
        ws_b = lib.gpc_lexorder_workspace_bytes(n)
        ws = torch.empty(ws_b, dtype=torch.uint8, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.gpc_lexorder_zyx(C.c_void_p(xd.data_ptr()), is_f32, n, C.c_void_p(out.data_ptr()),
                                        C.c_void_p(ws.data_ptr()), ws_b, stream), "gpc_lexorder_zyx")
    return out.to(src_device)


def compress_point_cloud(
    xyz_quantized,            # Quantized point cloud coordinates, numpy array or torch tensor
    ckpt_path,                # Path to pre-trained weights file
    output_path,              # Output bin file path
    channels=32,              # Network channel count
    kernel_size=5,            # Convolution kernel size
    posQ=1,                   # Quantization scale
    gpu_coder_chunk=0         # extension (not in the reference): > 0 writes container version 2, see below
):
    """
    Compress point cloud into a bin file (reference: pcc_utils.py:24-217).

    Returns {'bpp', 'enc_time', 'file_size_bits', 'num_points', 'output_path'}; `enc_time` is wall-clock
    between device synchronisations, including the host range coder and excluding the checkpoint load
    and the file write, as in the reference (:78-79,188-189).  Extra key 'gpu_time': CUDA-event time
    of the device segments only.

    gpu_coder_chunk > 0 (e.g. 2048): the occupancy streams are coded on the GPU in chunks of that many symbols and the file is
    container version 2 (bitstream.py) -- smaller PCIe traffic and no host range coder on either side, but NOT readable by the
    reference's decompress_point_cloud.  The default (0) writes the reference's own bitstream.
    """
    if gpu_coder_chunk and not (32 <= int(gpu_coder_chunk) <= 16384):
        raise ValueError("gpu_coder_chunk must be in 32..16384 (chunk byte counts are stored as u16)")
    os.makedirs(os.path.dirname(output_path), exist_ok=True)
    codec = _codec(ckpt_path, channels, kernel_size)
    dev = codec.dev
    if isinstance(xyz_quantized, np.ndarray):
        xyz = torch.tensor(xyz_quantized)
    else:
        xyz = xyz_quantized.detach()          # never mutated (the reference clones, :61)
    N = xyz.shape[0]
    if xyz.is_floating_point():
        xyz = xyz.to(device=dev, dtype=torch.float32)
    else:
        xyz = xyz.to(device=dev, dtype=torch.int32)    # the reference casts with .int() (:73)

    torch.cuda.synchronize(dev)
    enc_time_start = time.time()
    base_xyz, base_occ, streams, _ = codec.encode(xyz, gpu_chunk=int(gpu_coder_chunk))
    torch.cuda.synchronize(dev)
    enc_time_end = time.time()
    if gpu_coder_chunk:
        streams = streams + [bitstream.v2_trailer(int(gpu_coder_chunk), int(codec.last_stats.get("n_unique", N)))]

    blob = bitstream.write_file(posQ, base_xyz, base_occ, streams)
    with open(output_path, 'wb') as f:
        f.write(blob)

    enc_time = enc_time_end - enc_time_start
    file_size_bits = os.stat(output_path).st_size * 8
    bpp = file_size_bits / N
    return {
        'bpp': bpp,
        'enc_time': enc_time,
        'file_size_bits': file_size_bits,
        'num_points': N,
        'output_path': output_path,
        'gpu_time': codec.last_stats.get("gpu_ms", 0.0) / 1e3,
    }


def decompress_point_cloud(
    bin_file_path,           # Path to compressed bin file
    ckpt_path,               # Path to pre-trained weights file
    output_path=None,        # Path for output ply file (optional)
    channels=32,             # Network channel count
    kernel_size=5,           # Convolution kernel size
    is_data_pre_quantized=True,  # Whether original point cloud is pre-quantized
    sorted_output=False      # extension (not in the reference): rows in calculate_morton_order order
):
    """
    Decompress point cloud from bin file (reference: pcc_utils.py:230-400).

    Returns {'dec_time', 'num_points', 'point_cloud': float32 [N,3] on the GPU, 'output_path'}.  Row
    order is the reference's: children of the (z,y,x)-sorted last level, octant ascending (:375).
    With sorted_output=True the rows come back in ascending (z,y,x) instead, which is the order
    calculate_morton_order produces: the re-sort HAC does right after this call
    (HAC/scene/gaussian_model.py:1253-1255) is then the identity permutation and can be dropped.
    """
    if output_path:
        os.makedirs(os.path.dirname(output_path), exist_ok=True)
    codec = _codec(ckpt_path, channels, kernel_size)
    dev = codec.dev
    with open(bin_file_path, 'rb') as f:
        blob = f.read()
    posQ, base_xyz, base_occ, streams = bitstream.read_file(blob)
    streams, gpu_chunk, n_coded = bitstream.split_v2(streams)    # container version 2 names itself in a trailing stream

    torch.cuda.synchronize(dev)
    dec_time_start = time.time()
    scan = codec.decode(base_xyz, base_occ, streams, scale=float(posQ), sorted_rows=bool(sorted_output), gpu_chunk=gpu_chunk)
    if n_coded is not None and scan.shape[0] != n_coded:
        raise ValueError(f"version-2 file decodes to {scan.shape[0]} voxels, its trailer says {n_coded}: corrupt file, wrong checkpoint "
                         "or a library whose kernels sum in another order than the writer's")
    if not is_data_pre_quantized:
        scan = (scan - 131072) * 0.001                 # pcc_utils.py:381
    torch.cuda.synchronize(dev)
    dec_time_end = time.time()
    dec_time = dec_time_end - dec_time_start

    point_cloud = scan
    if output_path:
        # the reference calls io.save_ply_ascii_geo here but never imports `io` (pcc_utils.py:392 would raise
        # NameError); write the same ASCII geometry PLY (kit/io.py:36-49) instead of failing.
        _save_ply_ascii_geo(point_cloud.cpu().numpy(), output_path)
    return {
        'dec_time': dec_time,
        'num_points': point_cloud.shape[0],
        'point_cloud': point_cloud,
        'output_path': output_path,
        'gpu_time': codec.last_stats.get("gpu_ms", 0.0) / 1e3,
    }


def _save_ply_ascii_geo(coords: np.ndarray, path: str) -> None:
    coords = coords.astype(np.float32)
    with open(path, 'w') as f:
        f.write('ply\nformat ascii 1.0\n')
        f.write(f'element vertex {coords.shape[0]}\n')
        f.write('property float x\nproperty float y\nproperty float z\nend_header\n')
        for p in coords:
            f.write(f'{p[0]} {p[1]} {p[2]}\n')
