#!/usr/bin/env python
"""GPU driver vs the sequential CPU oracle (oracle/px_sequential.py), seed by seed, on the reference's bundled scenes:
instance count, number of differing labels, largest relative model difference, and the misclassification error of
both. Usage: python tools/parity_report.py [seeds...]   (default 1..5)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))
sys.path.insert(0, str(ROOT / "tests"))
import pyprogressivex  # noqa: E402
from pyprogressivex import _native  # noqa: E402
from oracle import px_sequential as seq  # noqa: E402
from test_gpu_reference_scenes import misclassification, _pose_error  # noqa: E402

G = np.load(ROOT / "tests" / "golden" / "reference_scenes.npz")
seeds = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]


def report(name, seed, models, labels, m_o, l_o, ms, ref=None):
    M = models.shape[0] * models.shape[1] // ms
    diff = int(np.sum(labels != l_o.astype(np.int32)))
    rel = float("nan")
    if M == m_o.shape[0] and M:
        a, b = models.reshape(M, ms), m_o
        rel = float((np.abs(a - b).max(1) / np.abs(b).max(1)).max())
    extra = ""
    if ref is not None:
        extra = f" err gpu {misclassification(labels, ref):.3f} oracle {misclassification(l_o.astype(np.int32), ref):.3f}"
    print(f"{name:16s} seed {seed}: M gpu {M} oracle {m_o.shape[0]}  labels differing {diff:4d}  max rel model diff {rel:.2e}{extra}",
          flush=True)


for scene in ("book", "breadcube", "cubetoy"):
    corrs, ref = G[f"{scene}_corrs"], G[f"{scene}_labels"]
    with _native.Context(0) as ctx:
        ctx.upload_points(_native.MODEL_F, corrs)
        graph = ctx.knn_graph(50.0, 5)
    kw = dict(threshold=0.75, conf=0.5, spatial_coherence_weight=0.5, neighborhood_ball_radius=50.0,
              maximum_tanimoto_similarity=0.4, max_iters=10000, minimum_point_number=7, maximum_model_number=4,
              sampler_id=2, scoring_exponent=1.0)
    for seed in seeds:
        models, labels = pyprogressivex.findTwoViewMotions(corrs, 640, 480, 640, 480, seed=seed, **kw)
        m_o, l_o = seq.find_two_view_motions(corrs, 0.75, 0.5, 0.5, 0.4, 10000, 7, 4, 2, 1.0, seed, graph,
                                             image_sizes=(640.0, 480.0, 640.0, 480.0))
        report(scene, seed, models, labels, m_o, l_o, 9, ref)

for scene in ("unionhouse", "oldclassicswing", "unihouse"):
    corrs, ref = G[f"{scene}_corrs"], G[f"{scene}_labels"]
    with _native.Context(0) as ctx:
        ctx.upload_points(_native.MODEL_H, corrs)
        graph = ctx.knn_graph(200.0, 5)
    for seed in seeds:
        models, labels = pyprogressivex.findHomographies(corrs, 640, 480, 640, 480, threshold=4.0, conf=0.5,
                                                         spatial_coherence_weight=0.05, neighborhood_ball_radius=200.0,
                                                         maximum_tanimoto_similarity=0.4, max_iters=1000, minimum_point_number=10,
                                                         maximum_model_number=6, scoring_exponent=2, sampler_id=3, seed=seed)
        m_o, l_o = seq.find_homographies(corrs, 4.0, 0.5, 0.05, 0.4, 1000, 10, 6, 3, 2, seed, graph)
        report(scene, seed, models, labels, m_o, l_o, 9, ref)

pts, K, gt = G["tless_points"], G["tless_K"], G["tless_poses"]
raw = np.ascontiguousarray(np.column_stack([pts[:, :2], pts[:, 2:]]))
with _native.Context(0) as ctx:
    ctx.upload_points(_native.MODEL_PNP, raw)
    graph = ctx.knn_graph(20.0, 5)
for seed in seeds:
    poses, labels = pyprogressivex.find6DPoses(pts[:, :2], pts[:, 2:], K, 4.0, seed=seed)
    m_o, l_o = seq.find_6d_poses(pts[:, :2], pts[:, 2:], K, 4.0, 0.9, 0.1, 0.9, 400, 6, -1, seed, graph)
    report("tless", seed, poses, labels, m_o, l_o, 12)
    est = poses.reshape(-1, 3, 4)
    best = [min(_pose_error(g, e) for e in est) for g in gt]
    print("                 gt pose errors (deg, m):", [(round(a, 1), round(t, 3)) for a, t in best], flush=True)
