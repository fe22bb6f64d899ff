"""Host logic of the stand-alone file tools (gauspcc_b200/cli.py): readers, coordinate mapping, file discovery, report."""
import csv
import os

import numpy as np
import torch

from gauspcc_b200 import cli


def test_read_points_formats(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(50, 3)).astype(np.float32)
    kitti = np.concatenate([pts, rng.random((50, 1)).astype(np.float32)], axis=1)
    kitti.tofile(tmp_path / "a.bin")                              # KITTI records x, y, z, intensity (kit/io.py:15-16)
    np.save(tmp_path / "b.npy", pts.astype(np.float64))
    with open(tmp_path / "c.ply", "w") as f:                      # ASCII geometry ply as kit/io.py:36-49 writes it
        f.write("ply\nformat ascii 1.0\nelement vertex 50\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        for p in pts:
            f.write(f"{p[0]} {p[1]} {p[2]}\n")
    assert np.array_equal(cli.read_points(str(tmp_path / "a.bin")), pts.astype(np.float64))
    assert np.array_equal(cli.read_points(str(tmp_path / "b.npy")), pts.astype(np.float64))
    got = cli.read_points(str(tmp_path / "c.ply"))
    assert got.shape == (50, 3) and np.allclose(got, pts, rtol=1e-6)


def test_quantise_matches_reference_formula():
    """compress_ue_4stage_conv.py:90-95 verbatim"""
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-40, 40, size=(1000, 3))
    for posQ in (1, 16, 64):
        ref = torch.round(torch.tensor(xyz / 0.001 + 131072) / posQ).int()
        assert torch.equal(cli.quantise(xyz, posQ, False), ref)
        ints = np.round(xyz * 100)
        assert torch.equal(cli.quantise(ints, posQ, True), torch.round(torch.tensor(ints) / posQ).int())


def test_list_inputs_and_report(tmp_path):
    (tmp_path / "seq" / "x").mkdir(parents=True)
    for name in ("seq/x/2.ply", "seq/1.bin", "seq/x/3.txt", "0.npy"):
        (tmp_path / name).write_bytes(b"")
    got = [os.path.relpath(p, tmp_path) for p in cli.list_inputs(str(tmp_path))]
    assert got == ["0.npy", "seq/1.bin", "seq/x/2.ply"]           # recursive, sorted, suffix filter (:57-59)
    assert len(cli.list_inputs(str(tmp_path), num_samples=2)) == 2
    rows = [{"filedir": "a", "bpp": 2.0, "enc_time": 1.0}, {"filedir": "b", "bpp": 4.0, "enc_time": 3.0}]
    out = tmp_path / "r.csv"
    cli._write_csv(str(out), rows, ["bpp", "enc_time"])
    rd = list(csv.DictReader(open(out)))
    assert [r["filedir"] for r in rd] == ["a", "b", "avg"] and float(rd[-1]["bpp"]) == 3.0 and float(rd[-1]["enc_time"]) == 2.0
