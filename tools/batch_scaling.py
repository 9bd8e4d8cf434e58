#!/usr/bin/env python
"""fits/s of findHomographiesBatch versus the number of concurrent host threads / contexts (config C4 in miniature)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
import pyprogressivex  # noqa: E402
from pyprogressivex import synthetic as syn  # noqa: E402

n_pairs, n_pts = 48, int(sys.argv[1]) if len(sys.argv) > 1 else 5000
pairs = [syn.multi_homography_scene(n_pts, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=700 + p)[0] for p in range(n_pairs)]
kw = dict(threshold=2.0, conf=0.5, spatial_coherence_weight=0.0, max_iters=1000, minimum_point_number=60, sampler_id=0,
          scoring_exponent=2, seed=11)
pyprogressivex.findHomographiesBatch(pairs[:8], 1024, 768, 1024, 768, workers=8, **kw)
for w in (1, 2, 4, 8, 16, 32):
    t0 = time.perf_counter()
    pyprogressivex.findHomographiesBatch(pairs, 1024, 768, 1024, 768, workers=w, **kw)
    dt = time.perf_counter() - t0
    print(f"workers={w:2d}: {n_pairs / dt:7.1f} fits/s")
