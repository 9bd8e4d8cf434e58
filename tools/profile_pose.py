#!/usr/bin/env python
"""Host-side phase profile of one find6DPoses call on config C5 (PXB_PROFILE=1 makes the driver print it)."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
os.environ["PXB_PROFILE"] = "1"
import pyprogressivex  # noqa: E402
from pyprogressivex import synthetic as syn  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
img, w, K, gt, _ = syn.multi_pose_scene(N, n_objects=10, inlier_ratio_each=0.06, noise_px=1.0, seed=0)
kw = dict(threshold=4.0, conf=0.9, spatial_coherence_weight=lam, neighborhood_ball_radius=20.0, maximum_tanimoto_similarity=0.9,
          max_iters=5000, minimum_point_number=N // 100, maximum_model_number=-1)
for warm in (1, 2):
    pyprogressivex.find6DPoses(img, w, K, seed=warm, **kw)
print("---- timed call ----", file=sys.stderr)
t0 = time.perf_counter()
m, lab = pyprogressivex.find6DPoses(img, w, K, seed=3, **kw)
print(f"N={N} lambda={lam}: {1e3 * (time.perf_counter() - t0):.2f} ms, {m.shape[0] // 3} poses")
