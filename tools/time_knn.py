#!/usr/bin/env python
"""Wall time of the neighbourhood-graph build through the C ABI (pxb_knn_graph: kernel + the copy of the lists back)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
import numpy as np  # noqa: E402
from pyprogressivex import _native  # noqa: E402
from pyprogressivex import synthetic as syn  # noqa: E402

ctx = _native.Context(0)
for t, N, radius in [(0, 10_000, 200.0), (0, 50_000, 200.0), (2, 100_000, 20.0 / 1074.0)]:
    if t == 0:
        pts, _, _ = syn.multi_homography_scene(N, seed=0)
    else:
        img, w, Kc, _, _ = syn.multi_pose_scene(N, seed=0)
        pts = syn.normalize_pnp_points(img, w, Kc)
    ctx.upload_points(t, np.ascontiguousarray(pts))
    ctx.knn_graph(radius, 5)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        off, idx = ctx.knn_graph(radius, 5)
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(f"type {t} N={N}: {ms:.3f} ms per build, mean degree {off[-1] / N:.2f}, {N * N / ms / 1e6:.1f} G pairs/s")
