#!/usr/bin/env python
"""Host-side phase profile of findHomographies on config C2 with the Python defaults (what bench.py's fits_c2 times)."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
os.environ.setdefault("PXB_PROFILE", "1")
import pyprogressivex  # noqa: E402
from pyprogressivex import synthetic as syn  # noqa: E402

lam = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
pts, gt, _ = syn.multi_homography_scene(10_000, n_planes=5, outlier_ratio=0.4, noise=0.5, seed=42)
kw = dict(threshold=4.0, conf=0.5, spatial_coherence_weight=lam, neighborhood_ball_radius=200.0,
          maximum_tanimoto_similarity=0.4, max_iters=1000, minimum_point_number=10, maximum_model_number=-1,
          sampler_id=3, scoring_exponent=2)
for warm in (1, 101, 102):
    pyprogressivex.findHomographies(pts, 1024, 768, 1024, 768, seed=warm, **kw)
print("---- timed call ----", file=sys.stderr)
t0 = time.perf_counter()
m, lab = pyprogressivex.findHomographies(pts, 1024, 768, 1024, 768, seed=2, **kw)
print(f"defaults lambda={lam}: {1e3 * (time.perf_counter() - t0):.2f} ms, {m.shape[0] // 3} models")
