"""The reference's own acceptance material, run through the drop-in Python surface on the GPU.

The reference ships no numeric tests; its only evidence are the notebooks under dataset_comparison/ and examples/, which
print a misclassification error per AdelaideRMF scene (dataset_comparison/utils.py:50-66) and pose errors for a T-LESS
image. The scenes bundled with the reference (build/data/*) travel as tests/golden/reference_scenes.npz
(tests/golden/make_golden.py); the calls below use the notebooks' parameters. The reference is stochastic
(std::random_device) -- its printed numbers are one draw -- so the bars are the printed values with a margin, checked over
several seeds, plus determinism per seed.

  adelaideH.ipynb   unionhouse 0.006, oldclassicswing 0.000, unihouse 0.186     (mean of the 19 scenes 0.064)
  adelaideF.ipynb   book 0.032, breadcube 0.017, cubetoy 0.012                   (mean 0.109)
  example_multi_pose_6d.ipynb   two ground-truth poses: 8.2 deg / 2.4 cm and 0.9 deg / 1.2 cm
"""
import itertools
from pathlib import Path

import numpy as np
import pytest

import pyprogressivex

pytestmark = pytest.mark.gpu

G = np.load(Path(__file__).resolve().parent / "golden" / "reference_scenes.npz")
IMAGE_SIZE = {"unionhouse": (640, 480), "oldclassicswing": (640, 480), "unihouse": (640, 480), "book": (640, 480),
              "breadcube": (640, 480), "cubetoy": (640, 480)}  # the entry points take, and ignore, the image sizes


def misclassification(segmentation, ref):
    """dataset_comparison/utils.py:50-66: best relabelling of the n reference labels onto 0..n-1, mismatches / N.
    The estimated labelling puts outliers last (label = number of models); the reference data puts them first (0)."""
    n = int(ref.max()) + 1
    best = len(ref)
    for p in itertools.permutations(range(n)):
        mapped = np.asarray(p)[ref]
        best = min(best, int(np.sum(mapped != segmentation)))
    return best / len(ref)


@pytest.mark.parametrize("scene,bar", [("unionhouse", 0.05), ("oldclassicswing", 0.05), ("unihouse", 0.25)])
def test_adelaide_h_scenes(scene, bar):
    corrs, ref = G[f"{scene}_corrs"], G[f"{scene}_labels"]
    w, h = IMAGE_SIZE[scene]
    errs = []
    for seed in (1, 2, 3):
        H, lab = pyprogressivex.findHomographies(corrs, w, h, w, h, threshold=4.0, conf=0.5, spatial_coherence_weight=0.05,
                                                 neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4,
                                                 max_iters=1000, minimum_point_number=10, maximum_model_number=6,
                                                 scoring_exponent=2, sampler_id=3, seed=seed)
        assert H.shape[1] == 3 and H.shape[0] % 3 == 0 and lab.shape == (len(corrs),)
        errs.append(misclassification(lab, ref))
    assert np.median(errs) <= bar, errs
    again = pyprogressivex.findHomographies(corrs, w, h, w, h, threshold=4.0, conf=0.5, spatial_coherence_weight=0.05,
                                            neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000,
                                            minimum_point_number=10, maximum_model_number=6, scoring_exponent=2,
                                            sampler_id=3, seed=3)
    assert np.array_equal(again[1], lab) and np.array_equal(again[0], H)


@pytest.mark.parametrize("scene,bar", [("book", 0.08), ("breadcube", 0.08)])
def test_adelaide_f_scenes(scene, bar):
    """Median over five seeds at the reference's level."""
    corrs, ref = G[f"{scene}_corrs"], G[f"{scene}_labels"]
    w, h = IMAGE_SIZE[scene]
    errs = []
    for seed in (1, 2, 3, 4, 5):
        F, lab = pyprogressivex.findTwoViewMotions(corrs, w, h, w, h, threshold=0.75, conf=0.5,
                                                   spatial_coherence_weight=0.5, neighborhood_ball_radius=50.0,
                                                   maximum_tanimoto_similarity=0.4, max_iters=10000,
                                                   minimum_point_number=7, maximum_model_number=4, sampler_id=2,
                                                   scoring_exponent=1.0, seed=seed)
        errs.append(misclassification(lab, ref))
    assert np.median(errs) <= bar, errs


def test_adelaide_f_cubetoy_is_bimodal():
    """cubetoy's two motions are plane dominated. Without DEGENSAC (fundamental_estimator.h:341-572,
    Driver::apply_degensac) the second motion is hardly ever proposed with enough support; with it a draw either recovers
    both motions (error <= 0.05; the reference's single printed draw: 0.012) or keeps one (0.29-0.36: the second motion's
    72 points count as outliers; one draw in nine finds no model that survives validation). tools/parity_report.py shows
    the sequential CPU oracle taking the same decisions draw by draw, so this is a property of the algorithm on this
    scene, not of the GPU path. Over nine seeds at least two draws must recover both motions (error <= 0.1)."""
    corrs, ref = G["cubetoy_corrs"], G["cubetoy_labels"]
    w, h = IMAGE_SIZE["cubetoy"]
    errs = []
    for seed in range(1, 10):
        F, lab = pyprogressivex.findTwoViewMotions(corrs, w, h, w, h, threshold=0.75, conf=0.5,
                                                   spatial_coherence_weight=0.5, neighborhood_ball_radius=50.0,
                                                   maximum_tanimoto_similarity=0.4, max_iters=10000,
                                                   minimum_point_number=7, maximum_model_number=4, sampler_id=2,
                                                   scoring_exponent=1.0, seed=seed)
        errs.append(misclassification(lab, ref))
    assert sum(e <= 0.10 for e in errs) >= 2, errs


def _pose_error(gt, est):
    R = gt[:, :3].T @ est[:, :3]
    ang = np.degrees(np.arccos(max(-1.0, min(1.0, 0.5 * (np.trace(R) - 1.0)))))
    return ang, float(np.linalg.norm(gt[:, 3] - est[:, 3]))


def test_tless_poses():
    """example_multi_pose_6d.ipynb: the ground-truth objects are among the returned instances (the reference's single
    printed draw: 8.2 deg / 24 mm and 0.9 deg / 12 mm). The proposal loop runs 400 iterations at conf = 0.9: the second
    object is recovered by every draw, the first (few, noisy matches) by about every second one -- the sequential CPU
    oracle takes the same decisions draw by draw (tools/parity_report.py). Over nine seeds: the easy object always, both
    in at least a third of the draws."""
    pts, K, gt = G["tless_points"], G["tless_K"], G["tless_poses"]
    both = 0
    for seed in range(1, 10):
        poses, lab = pyprogressivex.find6DPoses(pts[:, :2], pts[:, 2:], K, 4.0, seed=seed)
        M = poses.shape[0] // 3
        assert M >= 2 and lab.shape == (len(pts),) and poses.shape[1] == 4
        est = poses.reshape(M, 3, 4)
        best = [min(_pose_error(g, e) for e in est) for g in gt]
        assert best[1][0] < 15.0 and best[1][1] < 40.0, (seed, best)
        both += all(ang < 15.0 and tr < 40.0 for ang, tr in best)
    assert both >= 3
