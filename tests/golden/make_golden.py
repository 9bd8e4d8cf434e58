#!/usr/bin/env python
"""Writes tests/golden/reference_scenes.npz from the data files bundled with the reference (build/data/*).

These are DATA files of the AdelaideRMF / T-LESS examples the reference's notebooks run on (no source code): point
correspondences with a ground-truth instance label per row (0 = outlier), the only acceptance material the reference
ships (dataset_comparison/adelaideH.ipynb, adelaideF.ipynb, examples/example_multi_pose_6d.ipynb). /root/reference does
not exist on the GPU box, so the arrays travel as this small fixture.

    python tests/golden/make_golden.py [--ref /root/reference]
"""
import argparse
from pathlib import Path

import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--ref", default="/root/reference")
args = ap.parse_args()
data = Path(args.ref) / "build" / "data"
out = {}
for name in ("unionhouse", "oldclassicswing", "unihouse", "book", "breadcube", "cubetoy"):
    a = np.loadtxt(data / name / f"{name}.txt")            # x1 y1 1 x2 y2 1 label
    out[f"{name}_corrs"] = np.ascontiguousarray(a[:, [0, 1, 3, 4]], dtype=np.float64)
    out[f"{name}_labels"] = a[:, 6].astype(np.int32)
rows = np.loadtxt(data / "tless" / "tless.txt", skiprows=1)   # u v X Y Z
out["tless_points"] = np.ascontiguousarray(rows, dtype=np.float64)
out["tless_K"] = np.loadtxt(data / "tless" / "tless_intrinsics.txt")
out["tless_poses"] = np.loadtxt(data / "tless" / "tless_poses.txt", skiprows=1).reshape(-1, 3, 4)
np.savez_compressed(Path(__file__).resolve().parent / "reference_scenes.npz", **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
