#!/usr/bin/env python
"""fits/s of findHomographiesBatch versus host threads x problems in flight per thread (config C4 in miniature)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
import pyprogressivex  # noqa: E402
from pyprogressivex import synthetic as syn  # noqa: E402

n_pairs, n_pts = 192, int(sys.argv[1]) if len(sys.argv) > 1 else 5000
pairs = [syn.multi_homography_scene(n_pts, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=700 + p)[0] for p in range(n_pairs)]
kw = dict(threshold=2.0, conf=0.5, spatial_coherence_weight=0.0, max_iters=1000, minimum_point_number=60, sampler_id=0,
          scoring_exponent=2, seed=11)
pyprogressivex.findHomographiesBatch(pairs[:8], 1024, 768, 1024, 768, workers=2, in_flight=4, **kw)
for w, f in ((1, 1), (1, 2), (1, 4), (1, 8), (1, 16), (2, 4), (2, 8), (2, 16), (4, 4), (4, 8), (8, 4)):
    pyprogressivex.findHomographiesBatch(pairs[:3 * w * f], 1024, 768, 1024, 768, workers=w, in_flight=f, **kw)  # warm: every context sees, captures and replays its chains
    t0 = time.perf_counter()
    pyprogressivex.findHomographiesBatch(pairs, 1024, 768, 1024, 768, workers=w, in_flight=f, **kw)
    dt = time.perf_counter() - t0
    print(f"host threads={w:2d} x in flight={f:2d}: {n_pairs / dt:7.1f} fits/s")
