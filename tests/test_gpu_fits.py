"""Non-minimal F and PnP fits ("next") and the two remaining task entry points.

These solvers replace PoseLib's LM / OpenCV's EPnP by simpler well-defined algorithms (see pxb_fit_fp.cu), so there is
no bit-level oracle; they are judged by what they must achieve: the fit explains its inliers, and the end-to-end
calls recover the planted structures (the reference's own acceptance measure, misclassification error)."""
import numpy as np
import pytest

import pyprogressivex
from pyprogressivex import synthetic as syn
from test_gpu_e2e import misclassification

pytestmark = pytest.mark.gpu
F, PNP = 1, 2


def test_fundamental_fit_explains_inliers(ctx, oracle):
    pts, gt, Fs = syn.multi_motion_scene(6000, noise=0.3, seed=31)
    ctx.upload_points(F, pts)
    sets = [np.flatnonzero(gt == k) for k in range(3)] + [np.flatnonzero(gt == 0)[:8], np.flatnonzero(gt == 0)[:7]]
    Fg, ok = ctx.fit_nonminimal(sets)
    assert ok.tolist() == [1, 1, 1, 1, 0]
    for k in range(3):
        Fm = Fg[k].reshape(3, 3)
        assert abs(np.linalg.norm(Fm) - 1.0) < 1e-12 and Fm[2, 2] >= 0       # fundamental_estimator.h:609-613
        assert np.linalg.svd(Fm, compute_uv=False)[2] < 1e-10                   # rank 2
        r2, _ = oracle.residual_matrix(F, pts[sets[k]], Fg[k], 1.0)
        assert np.median(np.sqrt(r2)) < 0.6                                     # ~ the 0.3 px noise level
        # the planted matrix is recovered up to scale
        Ft = Fs[k] / np.linalg.norm(Fs[k])
        Ft = Ft if Ft[2, 2] >= 0 else -Ft
        assert np.abs(Fm - Ft).max() < 5e-3 * np.abs(Ft).max() + 1e-6


def test_pnp_fit_recovers_pose(ctx, oracle):
    img, w, K, gt, poses = syn.multi_pose_scene(8000, n_objects=4, inlier_ratio_each=0.15, noise_px=1.0, seed=33)
    pts = syn.normalize_pnp_points(img, w, K)
    ctx.upload_points(PNP, pts)
    sets = [np.flatnonzero(gt == k) for k in range(4)] + [np.flatnonzero(gt == 0)[:6], np.flatnonzero(gt == 0)[:5]]
    Pg, ok = ctx.fit_nonminimal(sets)
    assert ok.tolist() == [1, 1, 1, 1, 1, 0]
    for k in range(4):
        P = Pg[k].reshape(3, 4)
        R = P[:, :3]
        assert np.abs(R.T @ R - np.eye(3)).max() < 1e-9 and abs(np.linalg.det(R) - 1) < 1e-9
        assert np.abs(R - poses[k][:, :3]).max() < 2e-2
        assert np.abs(P[:, 3] - poses[k][:, 3]).max() < 2e-2
        r2, _ = oracle.residual_matrix(PNP, pts[sets[k]], Pg[k], 1.0)
        assert np.median(np.sqrt(r2)) * 1074 < 2.0                              # pixels


def test_find_two_view_motions_synthetic():
    corrs, gt, Fs = syn.multi_motion_scene(3000, n_motions=2, ratios=(0.35, 0.3), noise=0.3, seed=41)
    models, labels = pyprogressivex.findTwoViewMotions(corrs, 1024, 768, 1024, 768, threshold=0.75, conf=0.95,
                                                       spatial_coherence_weight=0.0, neighborhood_ball_radius=50.0,
                                                       maximum_tanimoto_similarity=0.4, max_iters=3000,
                                                       minimum_point_number=100, maximum_model_number=4, sampler_id=0,
                                                       seed=11)
    M = models.shape[0] // 3
    assert models.shape == (3 * M, 3) and labels.dtype == np.int32
    assert 2 <= M <= 3
    assert misclassification(gt, labels, M) < 0.15
    assert pyprogressivex.findFundamentalMatrices is pyprogressivex.findTwoViewMotions


def test_find_6d_poses_synthetic():
    img, w, K, gt, poses = syn.multi_pose_scene(4000, n_objects=3, inlier_ratio_each=0.2, noise_px=1.0, seed=43)
    models, labels = pyprogressivex.find6DPoses(img, w, K, threshold=4.0, conf=0.95, spatial_coherence_weight=0.0,
                                                neighborhood_ball_radius=20.0, maximum_tanimoto_similarity=0.6,
                                                max_iters=1000, minimum_point_number=100, maximum_model_number=5, seed=13)
    M = models.shape[0] // 3
    assert models.shape == (3 * M, 4) and labels.shape == (4000,)
    assert 3 <= M <= 4
    assert misclassification(gt, labels, M) < 0.1
    # every planted pose is among the returned ones
    for k in range(3):
        errs = [np.abs(models[3 * j:3 * j + 3] - poses[k]).max() for j in range(M)]
        assert min(errs) < 3e-2


def test_find_6d_poses_with_spatial_coherence():
    img, w, K, gt, poses = syn.multi_pose_scene(2500, n_objects=2, inlier_ratio_each=0.25, noise_px=1.0, seed=47)
    models, labels = pyprogressivex.find6DPoses(img, w, K, threshold=4.0, conf=0.9, spatial_coherence_weight=0.1,
                                                neighborhood_ball_radius=20.0, maximum_tanimoto_similarity=0.9,
                                                max_iters=400, minimum_point_number=60, seed=17)
    M = models.shape[0] // 3
    assert M >= 2
    assert misclassification(gt, labels, M) < 0.15


def test_fit_h_tensor_core_variant_agrees():
    """k_fit_h<MMA>: the normal equations accumulated with FP64 DMMA m8n8k4 (PXB_FIT_H_MMA=1, read once per process) give
    the same homographies as the scalar block reduction to 1e-9 (the sums differ only in the order of additions)."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    tool = Path(__file__).resolve().parent.parent / "tools" / "fit_h_bench.py"
    runs = {}
    for mma in ("0", "1"):
        r = subprocess.run([sys.executable, str(tool)], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, PXB_FIT_H_MMA=mma))
        assert r.returncode == 0, r.stderr[-2000:]
        runs[mma] = json.loads(r.stdout.strip().splitlines()[-1])
    for shape in ("lo_50x28", "pearl_5x1200"):
        a, b = np.array(runs["0"][shape]["H"]), np.array(runs["1"][shape]["H"])
        assert runs["0"][shape]["ok"] == runs["1"][shape]["ok"] and a.shape == b.shape
        assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()
