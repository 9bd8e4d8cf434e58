#!/usr/bin/env python
"""Small ncu target: runs the bench kernels a few times on the headline grid (no timing, no CPU baseline).

    ncu --set full --clock-control none --import-source on -k regex:k_residual_matrix -s 2 -c 1 \
        -o gpurun_out/prof python tools/profile_target.py --kernel matrix
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="matrix", choices=["matrix", "mask", "score", "all"])
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--n", type=int, default=50_000)
ap.add_argument("--k", type=int, default=10_000)
ap.add_argument("--type", type=int, default=0)
args = ap.parse_args()

from pyprogressivex import _native
from pyprogressivex import synthetic as syn

t = args.type
if t == 0:
    pts, gt, _ = syn.multi_homography_scene(args.n, seed=0)
    thr = 2.0
elif t == 1:
    pts, gt, _ = syn.multi_motion_scene(args.n, seed=0)
    thr = 0.75
else:
    img, w, K, gt, _ = syn.multi_pose_scene(args.n, seed=0)
    pts = syn.normalize_pnp_points(img, w, K)
    thr = 4.0 / 1074.0
m = _native.SAMPLE_SIZE[t]
S = syn.minimal_samples(gt, args.k, m, seed=0)
T2 = (1.5 * thr) ** 2
ctx = _native.Context(0)
ctx.upload_points(t, pts)
models, n, _, _ = ctx.solve_minimal(S)
flat = np.ascontiguousarray(models.reshape(-1, _native.MODEL_SIZE[t])[: args.k])
flat[~np.isfinite(flat).all(1)] = 0.0
K, N = flat.shape[0], pts.shape[0]
words = (N + 31) // 32
d_models = ctx.alloc(flat.nbytes)
d_models.upload(flat)
d_r2 = ctx.alloc(K * N * 8)
d_mask = ctx.alloc(K * words * 4)
d_out = ctx.alloc(K * 8 * 3)
for _ in range(args.iters):
    if args.kernel in ("matrix", "all"):
        _native._check(ctx.lib.pxb_residual_matrix_dev(ctx.handle, d_models.ptr, K, T2, d_r2.ptr, d_mask.ptr))
    if args.kernel in ("mask", "all"):
        _native._check(ctx.lib.pxb_residual_matrix_dev(ctx.handle, d_models.ptr, K, T2, None, d_mask.ptr))
    if args.kernel in ("score", "all"):
        _native._check(ctx.lib.pxb_score_compound_dev(ctx.handle, d_models.ptr, K, T2, None, d_out.ptr,
                                                      d_out.ptr + K * 8, d_out.ptr + 2 * K * 8))
ctx.sync()
print("done", K, N, ctx.launch_count())
