"""Stand-alone file-level tools over the same codec: the B200 counterparts of the reference's
`src/ai_pcc/GausPcgc/compress_ue_4stage_conv.py` and `decompress_ue_4stage_conv.py`.

    python -m gauspcc_b200.cli compress   --input_glob DIR --output_folder OUT --ckpt CKPT [--posQ 16] [--is_data_pre_quantized]
    python -m gauspcc_b200.cli decompress --input_glob OUT --output_folder DEC --ckpt CKPT [--is_data_pre_quantized]

Same argument names, file discovery (recursive, sorted, h5 / ply / bin / npy; compress_ue_4stage_conv.py:57-63), coordinate
mapping (`x / 0.001 + 131072` unless pre-quantised, then `round(x / posQ)`, :90-95; inverse `(c * posQ - 131072) * 0.001`,
decompress_ue_4stage_conv.py:176-179), `.bin` container (one `<name>.bin` per input) and CSV report (`<prefix>_data<N>.csv` with a
final `avg` row, :253-276) as the reference scripts.  `--kernel_size` defaults to 3 as in the reference scripts
(compress_ue_4stage_conv.py:44; the HAC call sites use 5, pcc_utils.py:29): both are built, channels = 32 only.  Differences: point files are read in this process (the reference forks 64 workers); `h5` is not
read (the reference lists the suffix but `kit/io.py:read_points` cannot parse it either).  No torchsparse / torchac.
"""
from __future__ import annotations

import argparse
import csv
import os
import sys
import time
from glob import glob
from typing import Dict, List, Optional

import numpy as np
import torch

from . import pcc_utils

SUFFIXES = (".h5", ".ply", ".bin", ".npy")


def read_points(path: str) -> np.ndarray:
    """[N,3] float64 coordinates of one file (reference: kit/io.py:14-31): KITTI `.bin` = float32 x,y,z,i records; `.npy` = array
    whose first three columns are x,y,z; anything else = text with one point per line (ASCII ply: header lines do not parse as
    floats and are skipped, extra columns are ignored)."""
    ext = os.path.splitext(path)[-1].lower()
    if ext == ".bin":
        return np.fromfile(path, dtype=np.float32).reshape(-1, 4)[:, :3].astype(np.float64)
    if ext == ".npy":
        a = np.asarray(np.load(path), dtype=np.float64)
        return a.reshape(-1, a.shape[-1])[:, :3]
    rows: List[List[float]] = []
    with open(path, "r", errors="ignore") as f:
        for line in f:
            parts = line.split()
            if len(parts) < 3:
                continue
            try:
                rows.append([float(parts[0]), float(parts[1]), float(parts[2])])
            except ValueError:
                continue
    return np.asarray(rows, dtype=np.float64).reshape(-1, 3)


def list_inputs(input_glob: str, num_samples: int = -1, suffixes=SUFFIXES) -> List[str]:
    paths = sorted(glob(os.path.join(input_glob, "**", "*.*"), recursive=True))
    paths = [p for p in paths if p.lower().endswith(suffixes)]
    return paths[:num_samples] if num_samples and num_samples > 0 else paths


def quantise(xyz: np.ndarray, posQ: int, is_data_pre_quantized: bool) -> torch.Tensor:
    """compress_ue_4stage_conv.py:90-95: (pre-quantised ? x : x / 0.001 + 131072), then round-half-even(x / posQ) as int32."""
    t = torch.tensor(xyz if is_data_pre_quantized else xyz / 0.001 + 131072)
    return torch.round(t / posQ).int()


def _write_csv(path: str, rows: List[Dict[str, object]], numeric: List[str]) -> None:
    avg: Dict[str, object] = {k: float(np.mean([float(r[k]) for r in rows])) for k in numeric}
    avg["filedir"] = "avg"
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        for r in rows + [avg]:
            w.writerow(r)


def compress_files(input_glob: str, output_folder: str, ckpt: str, posQ: int = 16, is_data_pre_quantized: bool = False,
                   channels: int = 32, kernel_size: int = 5, num_samples: int = -1, resultdir: Optional[str] = None,
                   prefix: str = "ue_4stage_conv") -> List[Dict[str, object]]:
    os.makedirs(output_folder, exist_ok=True)
    paths = list_inputs(input_glob, num_samples)
    if not paths:
        raise FileNotFoundError(f"no point-cloud files (h5 / ply / bin / npy) under {input_glob}")
    rows: List[Dict[str, object]] = []
    for p in paths:
        name = os.path.split(p)[-1]
        xyz = quantise(read_points(p), posQ, is_data_pre_quantized)
        # the codec stores the SET of voxels; N of the report is the number of input points, as in the reference (:96, :246)
        r = pcc_utils.compress_point_cloud(xyz, ckpt, os.path.join(output_folder, name + ".bin"), channels=channels,
                                           kernel_size=kernel_size, posQ=posQ)
        rows.append({"filedir": name, "bpp": r["file_size_bits"] / max(xyz.shape[0], 1), "enc_time": r["enc_time"],
                     "file_size_bits": r["file_size_bits"], "num_points": int(xyz.shape[0])})
    if resultdir:
        os.makedirs(resultdir, exist_ok=True)
        _write_csv(os.path.join(resultdir, f"{prefix}_data{len(paths)}.csv"), rows, ["bpp", "enc_time", "file_size_bits", "num_points"])
    return rows


def decompress_files(input_glob: str, output_folder: str, ckpt: str, is_data_pre_quantized: bool = False, channels: int = 32,
                     kernel_size: int = 5, resultdir: Optional[str] = None, prefix: str = "ue_4stage_conv") -> List[Dict[str, object]]:
    os.makedirs(output_folder, exist_ok=True)
    paths = list_inputs(input_glob, suffixes=(".bin",))
    if not paths:
        raise FileNotFoundError(f"no .bin files under {input_glob}")
    rows: List[Dict[str, object]] = []
    for p in paths:
        name = os.path.split(p)[-1]
        out_path = os.path.join(output_folder, name + ".ply")
        r = pcc_utils.decompress_point_cloud(p, ckpt, out_path, channels=channels, kernel_size=kernel_size,
                                             is_data_pre_quantized=is_data_pre_quantized)
        rows.append({"filedir": name, "dec_time": r["dec_time"], "num_points": int(r["num_points"])})
    if resultdir:
        os.makedirs(resultdir, exist_ok=True)
        _write_csv(os.path.join(resultdir, f"{prefix}_dec_data{len(paths)}.csv"), rows, ["dec_time", "num_points"])
    return rows


def main(argv: Optional[List[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="python -m gauspcc_b200.cli", description=__doc__.split("\n\n")[0],
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    for cmd in ("compress", "decompress"):
        sp = sub.add_parser(cmd, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        sp.add_argument("--input_glob", required=True, help="folder searched recursively for input files")
        sp.add_argument("--output_folder", required=True)
        sp.add_argument("--ckpt", required=True, help="Network(32, 5).state_dict() checkpoint")
        sp.add_argument("--is_data_pre_quantized", action="store_true", help="inputs are integer voxel coordinates already")
        sp.add_argument("--channels", type=int, default=32)
        sp.add_argument("--kernel_size", type=int, default=3)      # the reference scripts' default; the checkpoint must match
        sp.add_argument("--resultdir", default=None, help="folder for the CSV report")
        sp.add_argument("--prefix", default="ue_4stage_conv")
        if cmd == "compress":
            sp.add_argument("--posQ", type=int, default=16, help="quantisation scale")
            sp.add_argument("--num_samples", type=int, default=-1, help="first N files only (-1 = all)")
    a = ap.parse_args(argv)
    t0 = time.time()
    if a.cmd == "compress":
        rows = compress_files(a.input_glob, a.output_folder, a.ckpt, a.posQ, a.is_data_pre_quantized, a.channels, a.kernel_size,
                              a.num_samples, a.resultdir, a.prefix)
        print("Total: {:d} | Average bitrate:{:.3f} | Encoding time:{:.3f}s | Max GPU memory:{:.2f}MB".format(
            len(rows), float(np.mean([r["bpp"] for r in rows])), float(np.mean([r["enc_time"] for r in rows])),
            torch.cuda.max_memory_allocated() / 1024 / 1024))
    else:
        rows = decompress_files(a.input_glob, a.output_folder, a.ckpt, a.is_data_pre_quantized, a.channels, a.kernel_size,
                                a.resultdir, a.prefix)
        print("Total: {:d} | Decoding time:{:.3f}s | Max GPU memory:{:.2f}MB".format(
            len(rows), float(np.mean([r["dec_time"] for r in rows])), torch.cuda.max_memory_allocated() / 1024 / 1024))
    print(f"wall {time.time() - t0:.2f} s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
