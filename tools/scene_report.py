"""Misclassification / pose errors of the drop-in on the reference's bundled scenes over several seeds (tuning aid; the
asserting version is tests/test_gpu_reference_scenes.py)."""
import sys, itertools
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "progressive-x_b200"))
import pyprogressivex
G = np.load(ROOT / "tests" / "golden" / "reference_scenes.npz")
def mis(seg, ref):
    n = int(ref.max()) + 1
    return min(int(np.sum(np.asarray(p)[ref] != seg)) for p in itertools.permutations(range(n))) / len(ref)
for scene in ("book", "breadcube", "cubetoy"):
    corrs, ref = G[f"{scene}_corrs"], G[f"{scene}_labels"]
    for sampler in (2, 3):
        out = []
        for seed in range(1, 9):
            F, lab = pyprogressivex.findTwoViewMotions(corrs, 640, 480, 640, 480, threshold=0.75, conf=0.5, spatial_coherence_weight=0.5,
                neighborhood_ball_radius=50.0, maximum_tanimoto_similarity=0.4, max_iters=10000, minimum_point_number=7,
                maximum_model_number=4, sampler_id=sampler, scoring_exponent=1.0, seed=seed)
            out.append((F.shape[0] // 3, round(mis(lab, ref), 3)))
        print(scene, "sampler", sampler, out)
pts, K, gt = G["tless_points"], G["tless_K"], G["tless_poses"]
def perr(g, e):
    R = g[:, :3].T @ e[:, :3]
    return np.degrees(np.arccos(max(-1, min(1, 0.5 * (np.trace(R) - 1))))), np.linalg.norm(g[:, 3] - e[:, 3])
for seed in range(1, 6):
    poses, lab = pyprogressivex.find6DPoses(pts[:, :2], pts[:, 2:], K, 4.0, seed=seed)
    est = poses.reshape(-1, 3, 4)
    print("tless seed", seed, "M", len(est), [tuple(round(x, 1) for x in min(perr(g, e) for e in est)) for g in gt], np.bincount(lab))
