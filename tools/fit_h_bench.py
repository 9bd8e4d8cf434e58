#!/usr/bin/env python
"""k_fit_h on the two shapes the driver produces -- the 50 inner-RANSAC refits of an LO step (28 points each) and the
per-instance refits of a PEARL iteration (5 x 1200 points) -- with the scalar FP64 block reduction (default) or the FP64
tensor-core accumulation (PXB_FIT_H_MMA=1). Prints the fitted matrices' checksum so that two runs can be compared, and
host-side timings; run under `ncu -k regex:k_fit_h --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active`
for kernel time and tensor-pipe utilisation (profiles/r02_fit_h_dmma.txt)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
from pyprogressivex import _native  # noqa: E402
from pyprogressivex import synthetic as syn  # noqa: E402

pts, gt, Hs = syn.multi_homography_scene(10_000, n_planes=5, outlier_ratio=0.4, noise=0.5, seed=42)
rng = np.random.default_rng(0)
inl = [np.flatnonzero(gt == k) for k in range(5)]
lo_sets = [rng.choice(inl[0], 28, replace=False) for _ in range(50)]
out = {"mma": os.environ.get("PXB_FIT_H_MMA", "0")}
with _native.Context(0) as ctx:
    ctx.upload_points(_native.MODEL_H, pts)
    for name, sets in (("lo_50x28", lo_sets), ("pearl_5x1200", inl)):
        H, ok = ctx.fit_nonminimal(sets)
        t0 = time.perf_counter()
        for _ in range(20):
            H, ok = ctx.fit_nonminimal(sets)
        out[name] = {"host_call_us": (time.perf_counter() - t0) / 20 * 1e6, "ok": int(ok.sum()),
                     "H": [float(x) for x in H.reshape(-1)]}
print(json.dumps(out))
