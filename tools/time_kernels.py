#!/usr/bin/env python
"""Times the headline kernels (CUDA events on the context's stream): the residual-and-inlier matrix and the fused score.
Environment: PXB_TYPE (0 H, 1 F, 2 PnP), PXB_N, PXB_K, PXB_RM_HYPS (32 / 64 hypotheses per block of the matrix kernel; read
once per process, hence the subprocess per run)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "progressive-x_b200"))
    import numpy as np
    import torch
    from pyprogressivex import _native
    from pyprogressivex import synthetic as syn
    t = int(os.environ.get("PXB_TYPE", "0"))
    N, K = int(os.environ.get("PXB_N", "50000")), int(os.environ.get("PXB_K", "10000"))
    if t == 0:
        pts, gt, _ = syn.multi_homography_scene(N, seed=0); thr = 2.0
    elif t == 1:
        pts, gt, _ = syn.multi_motion_scene(N, seed=0); thr = 0.75
    else:
        img, w, Kc, gt, _ = syn.multi_pose_scene(N, seed=0); pts = syn.normalize_pnp_points(img, w, Kc); thr = 4.0 / 1074.0
    S = syn.minimal_samples(gt, K, _native.SAMPLE_SIZE[t], seed=0)
    T2 = (1.5 * thr) ** 2
    ctx = _native.Context(0)
    ctx.upload_points(t, pts)
    models, n, _, _ = ctx.solve_minimal(S)
    filled = np.arange(models.shape[1])[None, :] < n[:, None]
    flat = models[filled]
    flat = flat[np.isfinite(flat).all(1)]
    flat = np.ascontiguousarray(np.resize(flat, (K, _native.MODEL_SIZE[t])))  # real solver outputs, tiled up to K
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    with torch.cuda.stream(stream):
        m = torch.from_numpy(flat).to(dev)
        r2 = torch.empty((K, N), dtype=torch.float64, device=dev)
        mask = torch.empty((K, (N + 31) // 32), dtype=torch.int32, device=dev)
        out = torch.empty((3, K), dtype=torch.float64, device=dev)
    stream.synchronize()

    def run(fn, iters=20):
        for _ in range(5):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for _ in range(iters):
            fn()
        ev[1].record(stream)
        stream.synchronize()
        return ev[0].elapsed_time(ev[1]) / iters

    f_m = lambda: ctx.lib.pxb_residual_matrix_dev(ctx.handle, m.data_ptr(), K, T2, r2.data_ptr(), mask.data_ptr())
    f_s = lambda: ctx.lib.pxb_score_compound_dev(ctx.handle, m.data_ptr(), K, T2, None, out[0].data_ptr(),
                                                 out[1].data_ptr(), out[2].data_ptr())
    tm, ts = run(f_m), run(f_s)
    print(f"hyps/block={os.environ.get('PXB_RM_HYPS','auto')} type={t} N={N} K={K} matrix {tm:.4f} ms "
          f"({N*K/tm/1e6:.1f} Gevals/s, {N*K*8.125/tm/1e6/6534.1*100:.1f}% of 6534 GB/s)  score {ts:.4f} ms ({N*K/ts/1e6:.1f} Gevals/s)")
else:
    for v in sys.argv[1:] or ["0"]:
        env = dict(os.environ) if v == "0" else dict(os.environ, PXB_RM_HYPS=v)
        subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
