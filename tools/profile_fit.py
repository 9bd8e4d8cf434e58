#!/usr/bin/env python
"""Wall time and kernel-launch count of complete findHomographies calls (config C2)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "progressive-x_b200"))
import numpy as np
import pyprogressivex
from pyprogressivex import synthetic as syn
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
pts, gt, _ = syn.multi_homography_scene(N, n_planes=5, outlier_ratio=0.4, noise=0.5, seed=42)
kw = dict(threshold=2.0, conf=0.5, spatial_coherence_weight=lam, neighborhood_ball_radius=60.0, maximum_tanimoto_similarity=0.4,
          max_iters=1000, minimum_point_number=max(50, N // 100), sampler_id=0 if lam == 0 else 3)
pyprogressivex.findHomographies(pts, 1024, 768, 1024, 768, seed=1, **kw)
ctx = pyprogressivex._ctx(0)
for s in range(2, 6):
    l0 = ctx.launch_count(); t0 = time.perf_counter()
    m, lab = pyprogressivex.findHomographies(pts, 1024, 768, 1024, 768, seed=s, **kw)
    dt = time.perf_counter() - t0
    print(f"N={N} lambda={lam} seed={s}: {dt*1e3:.1f} ms, {ctx.launch_count()-l0} launches, {m.shape[0]//3} models")
