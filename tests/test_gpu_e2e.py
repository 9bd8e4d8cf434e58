"""End-to-end through the pyprogressivex surface (same call a user of the reference makes) on synthetic scenes.

The reference is non-deterministic (std::random_device) and has no golden outputs, so the checks are the
reference's own acceptance measure: misclassification error against ground-truth labels (dataset_comparison/utils.py)
-- plus determinism for a fixed seed, the shape/dtype contract of bindings.cpp:152-165, and consistency of the
returned labeling with the returned models (a point labelled k lies within the truncated threshold of model k)."""
import itertools

import numpy as np
import pytest

import pyprogressivex
from pyprogressivex import synthetic as syn

pytestmark = pytest.mark.gpu


def misclassification(gt, labels, n_models):
    """best-permutation misclassification error (dataset_comparison/utils.py:50-66 does the same with sympy)"""
    gt = np.where(gt < 0, -1, gt)
    est = np.where(labels >= n_models, -1, labels)
    ks = sorted(set(gt[gt >= 0]))
    best = 1.0
    ids = list(range(n_models)) + [-2] * max(0, len(ks) - n_models)
    for perm in itertools.permutations(ids, len(ks)):
        mapped = np.full_like(gt, -1)
        for k, p in zip(ks, perm):
            if p >= 0:
                mapped[est == p] = k
        best = min(best, float(np.mean(mapped != gt)))
    return best


@pytest.mark.parametrize("sampler_id", [0, 3])
def test_find_homographies_synthetic(sampler_id):
    corrs, gt, Hs = syn.multi_homography_scene(3000, n_planes=3, outlier_ratio=0.3, noise=0.5, seed=11)
    models, labels = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, threshold=2.0, conf=0.95,
                                                     spatial_coherence_weight=0.0, neighborhood_ball_radius=60.0,
                                                     maximum_tanimoto_similarity=0.4, max_iters=2000,
                                                     minimum_point_number=50, maximum_model_number=-1,
                                                     sampler_id=sampler_id, scoring_exponent=2, seed=7)
    M = models.shape[0] // 3
    assert models.dtype == np.float64 and models.shape == (3 * M, 3)
    assert labels.dtype == np.int32 and labels.shape == (3000,)
    assert 3 <= M <= 4
    assert misclassification(gt, labels, M) < 0.08
    # labels are consistent with the models: an assigned point is within the truncated threshold of its model
    T = 9.0 / 4.0 * 2.0 * 2.0
    for k in range(M):
        H = models[3 * k:3 * k + 3]
        idx = np.flatnonzero(labels == k)
        p = np.c_[corrs[idx, :2], np.ones(idx.size)] @ H.T
        r2 = ((corrs[idx, 2:] - p[:, :2] / p[:, 2:3]) ** 2).sum(1)
        assert (r2 <= T * (1 + 1e-9)).mean() > 0.999


def test_find_homographies_with_spatial_coherence():
    """lambda > 0: GC-RANSAC's graph-cut local optimisation and PEARL's alpha-expansion run on the GPU min-cut engine
    (the AdelaideH configuration of the reference: lambda = 0.05, NAPSAC sampling)."""
    corrs, gt, Hs = syn.multi_homography_scene(2500, n_planes=3, outlier_ratio=0.3, noise=0.5, seed=21)
    models, labels = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, threshold=2.0, conf=0.95,
                                                     spatial_coherence_weight=0.05, neighborhood_ball_radius=60.0,
                                                     maximum_tanimoto_similarity=0.4, max_iters=1500,
                                                     minimum_point_number=50, maximum_model_number=6, sampler_id=3,
                                                     scoring_exponent=2, seed=5)
    M = models.shape[0] // 3
    assert 3 <= M <= 4
    assert misclassification(gt, labels, M) < 0.08


def test_batched_pairs_config_c4():
    """BASELINE config C4 in miniature: independent pairs through findHomographiesBatch (single process here; the
    distributed variant shards pairs over ranks and all-gathers the instances, covered at world size 2 in
    tests/test_sharding_cpu.py and by `bench.py --gpus N`)."""
    pairs, gts = [], []
    for p in range(6):
        c, gt, _ = syn.multi_homography_scene(1500, n_planes=2 + p % 2, outlier_ratio=0.4, seed=300 + p)
        pairs.append(c)
        gts.append(gt)
    kw = dict(threshold=2.0, conf=0.95, max_iters=1000, minimum_point_number=60, sampler_id=0, seed=9)
    out = pyprogressivex.findHomographiesBatch(pairs, 1024, 768, 1024, 768, workers=2, in_flight=3, **kw)
    assert len(out) == 6
    # the native batch driver (fibers that yield at every stream wait, own context each) returns exactly what
    # one-problem-at-a-time calls return, for any thread / in-flight split
    seq = [pyprogressivex.findHomographies(c, 1024, 768, 1024, 768, **kw) for c in pairs]
    one = pyprogressivex.findHomographiesBatch(pairs, 1024, 768, 1024, 768, workers=1, in_flight=6, **kw)
    for (ma, la), (mb, lb), (mc, lc) in zip(out, seq, one):
        assert np.array_equal(ma.view(np.uint64), mb.view(np.uint64)) and np.array_equal(la, lb)
        assert np.array_equal(mc.view(np.uint64), mb.view(np.uint64)) and np.array_equal(lc, lb)
    # with the spatial term (kNN graph, LO cuts and alpha-expansion inside the fibers): instance counts must agree
    # (the asynchronous max-flow may resolve exact energy ties differently between runs)
    kl = dict(kw, spatial_coherence_weight=0.05)
    bl = pyprogressivex.findHomographiesBatch(pairs[:3], 1024, 768, 1024, 768, workers=1, in_flight=3, **kl)
    sl = [pyprogressivex.findHomographies(c, 1024, 768, 1024, 768, **kl) for c in pairs[:3]]
    for (ma, la), (mb, lb) in zip(bl, sl):
        assert ma.shape == mb.shape and np.mean(la != lb) < 0.01
    for (models, labels), gt, p in zip(out, gts, range(6)):
        M = models.shape[0] // 3
        assert M >= 2 + p % 2
        assert misclassification(gt, labels, M) < 0.1


def test_find_homographies_is_deterministic_for_a_seed():
    corrs, gt, Hs = syn.multi_homography_scene(2000, n_planes=2, outlier_ratio=0.3, seed=5)
    a = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, threshold=2.0, conf=0.9, max_iters=500,
                                        minimum_point_number=40, sampler_id=0, seed=3)
    b = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, threshold=2.0, conf=0.9, max_iters=500,
                                        minimum_point_number=40, sampler_id=0, seed=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_block_replay_equals_the_sequential_loop(monkeypatch):
    """The driver evaluates blocks of 512 pre-drawn samples per launch and replays the reference's sequential
    bookkeeping over them. With PXB_BLOCK_SIZE=1 the driver degenerates to the reference's one-sample-at-a-time loop
    (sample -> solve -> score -> compare). Both must return bit-identical models and labels for the same seed, with and
    without the spatial term (LO graph cuts, alpha-expansion) -- that is the equivalence the block replay claims."""
    corrs, gt, Hs = syn.multi_homography_scene(1200, n_planes=2, outlier_ratio=0.35, seed=77)
    for lam, sampler in ((0.0, 0), (0.05, 3)):
        kw = dict(threshold=2.0, conf=0.9, spatial_coherence_weight=lam, neighborhood_ball_radius=60.0, max_iters=300,
                  minimum_point_number=40, sampler_id=sampler, seed=21)
        monkeypatch.delenv("PXB_BLOCK_SIZE", raising=False)
        blocked = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, **kw)
        monkeypatch.setenv("PXB_BLOCK_SIZE", "1")
        sequential = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, **kw)
        monkeypatch.setenv("PXB_BLOCK_SIZE", "7")
        odd = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, **kw)
        monkeypatch.delenv("PXB_BLOCK_SIZE", raising=False)
        assert blocked[0].shape[0] >= 6
        for other in (sequential, odd):
            assert np.array_equal(blocked[0].view(np.uint64), other[0].view(np.uint64))
            assert np.array_equal(blocked[1], other[1])


def test_shape_errors_match_the_reference_binding():
    with pytest.raises(ValueError):
        pyprogressivex.findHomographies(np.zeros((10, 3)), 10, 10, 10, 10)
    with pytest.raises(ValueError):
        pyprogressivex.findHomographies(np.zeros((3, 4)), 10, 10, 10, 10)


def test_unknown_sampler_returns_no_models(capfd):
    corrs, gt, Hs = syn.multi_homography_scene(500, n_planes=2, seed=5)
    models, labels = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, sampler_id=9)
    assert models.shape[0] == 0  # progressivex_python.cpp:240-245: message on stderr, return 0


def test_find_vanishing_points_synthetic():
    """findVanishingPoints through the Python surface: three planted vanishing points, 30 % random segments."""
    seg, gt, vps = syn.multi_vanishing_point_scene(1500, n_vps=3, outlier_ratio=0.3, noise=0.3, seed=13)
    models, labels = pyprogressivex.findVanishingPoints(seg, np.ones(len(seg)), 1024, 768, threshold=2.0, conf=0.95,
                                                        maximum_tanimoto_similarity=0.4, max_iters=1000,
                                                        minimum_point_number=50, sampler_id=0, seed=4)
    assert models.dtype == np.float64 and models.shape[1] == 3 and labels.dtype == np.int32 and labels.shape == (1500,)
    M = models.shape[0]
    assert 3 <= M <= 4
    assert misclassification(gt, labels, M) < 0.1
    # every planted vanishing point is found (directions agree up to sign)
    for v in vps:
        cosines = np.abs(models @ v) / np.linalg.norm(models, axis=1)
        assert cosines.max() > 0.9999
    # the reference's default sampler id (3) does not exist for this entry: no models, message on stderr
    m0, l0 = pyprogressivex.findVanishingPoints(seg, np.ones(len(seg)), 1024, 768)
    assert m0.shape == (0, 3)


def test_find_lines_synthetic():
    """findLines: the minimal solver carries the reference's `nx = y1 - x2`, so hypotheses only become lines through
    the local optimisation's least-squares fits -- exactly what the reference does; the planted lines are still found."""
    pts, gt, lines = syn.multi_line_scene(1200, n_lines=3, outlier_ratio=0.3, noise=0.5, seed=17)
    models, labels = pyprogressivex.findLines(pts, np.ones(len(pts)), 1024, 768, threshold=2.0, conf=0.99,
                                              maximum_tanimoto_similarity=0.4, max_iters=3000, minimum_point_number=60,
                                              sampler_id=0, seed=6)
    assert models.shape[1] == 3 and labels.shape == (1200,)
    M = models.shape[0]
    assert M >= 1
    found = 0
    for l in lines:
        d = np.abs(models[:, :2] @ l[:2])
        k = int(np.argmax(d))
        if d[k] > 0.999 and abs(abs(models[k, 2]) - abs(l[2])) < 5.0:
            found += 1
    assert found >= 1
    with pytest.raises(ValueError):
        pyprogressivex.findLines(np.zeros((10, 3)), None, 10, 10)


def test_plane_dominated_two_view_motion_uses_degensac(monkeypatch):
    """A rigid scene whose correspondences lie mostly on one plane (fundamental_estimator.h:341-572): seven-point samples
    are then H-degenerate most of the time. With DEGENSAC the returned epipolar geometry must also explain the off-plane
    points; the switch PXB_DEGENSAC=0 (not the reference's behaviour) is only exercised for determinism of the A/B."""
    rows, lab, F_true = syn.plane_dominated_pair(420, 80, 0.3, 21)
    rng = np.random.default_rng(1)
    outl = np.column_stack([rng.uniform(0, 1024, 150), rng.uniform(0, 768, 150), rng.uniform(0, 1024, 150), rng.uniform(0, 768, 150)])
    corrs = np.ascontiguousarray(np.concatenate([rows, outl]))
    off = np.flatnonzero(lab == 1)
    x1 = np.column_stack([corrs[:, :2], np.ones(len(corrs))])
    x2 = np.column_stack([corrs[:, 2:], np.ones(len(corrs))])

    def sampson(Fm):
        Fx1, Ftx2 = x1 @ Fm.T, x2 @ Fm
        num = np.einsum("ni,ni->n", x2, Fx1) ** 2
        return num / (Fx1[:, 0] ** 2 + Fx1[:, 1] ** 2 + Ftx2[:, 0] ** 2 + Ftx2[:, 1] ** 2)

    good = 0
    for seed in (1, 2, 3, 4):
        Fs, labels = pyprogressivex.findTwoViewMotions(corrs, 1024, 768, 1024, 768, threshold=1.0, conf=0.99,
                                                       spatial_coherence_weight=0.0, neighborhood_ball_radius=50.0,
                                                       maximum_tanimoto_similarity=0.4, max_iters=2000, minimum_point_number=50,
                                                       maximum_model_number=1, sampler_id=0, scoring_exponent=1.0, seed=seed)
        assert Fs.shape[0] == 3
        frac = float(np.mean(sampson(Fs[:3]) [off] < 1.5 ** 2))
        good += frac > 0.9
        again = pyprogressivex.findTwoViewMotions(corrs, 1024, 768, 1024, 768, threshold=1.0, conf=0.99,
                                                  spatial_coherence_weight=0.0, neighborhood_ball_radius=50.0,
                                                  maximum_tanimoto_similarity=0.4, max_iters=2000, minimum_point_number=50,
                                                  maximum_model_number=1, sampler_id=0, scoring_exponent=1.0, seed=seed)
        assert np.array_equal(again[0], Fs) and np.array_equal(again[1], labels)
    assert good >= 3, good


def test_statistics_and_mutable_settings():
    """ProgressiveX::getStatistics / getMutableSettings (progressive_x.h:210-217) through the ABI: one entry per accepted
    round with the four phase times (CUDA events), totals that add up, and engine settings that take effect."""
    corrs, gt, Hs = syn.multi_homography_scene(3000, n_planes=3, outlier_ratio=0.3, seed=21)
    kw = dict(threshold=2.0, conf=0.9, max_iters=600, minimum_point_number=100, sampler_id=0, seed=5)
    models, labels = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, **kw)
    st = pyprogressivex.getStatistics()
    its = st["iteration_statistics"]
    assert st["model_number"] == models.shape[0] // 3 >= 3 and len(its) >= 3
    assert its[-1]["number_of_instances"] == st["model_number"] and its[0]["number_of_instances"] == 1
    for it in its:
        assert it["time_of_proposal_engine"] > 0 and it["time_of_model_validation"] > 0 and it["time_of_compound_model_update"] > 0
        assert 20 <= it["ransac_iteration_number"] <= 700 and it["local_optimization_number"] >= 1  # failed generations count too (GCRANSAC.h:341)
        assert it["proposal_inlier_number"] >= 100
    assert abs(st["total_time_of_proposal_engine"] - sum(i["time_of_proposal_engine"] for i in its)) < 1e-12
    phases = (st["total_time_of_proposal_engine"] + st["total_time_of_model_validation"] + st["total_time_of_optimization"]
              + st["total_time_of_compound_model_calculation"])
    assert 0 < phases <= st["processing_time"] * 1.001 and st["kernel_launches"] > 50
    # one graph cut per local optimisation instead of up to ten: fewer cuts, still a valid result
    s = pyprogressivex.getMutableSettings()
    assert (s.max_graph_cut_number, s.max_local_optimization_number, s.min_iteration_number) == (10, 50, 20)
    s.max_graph_cut_number = 2
    pyprogressivex.setSettings(s)
    try:
        m2, l2 = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, **kw)
        st2 = pyprogressivex.getStatistics()
        assert all(i["graph_cut_number"] <= 2 * i["local_optimization_number"] for i in st2["iteration_statistics"])
        assert max(i["graph_cut_number"] for i in its) > 2
    finally:
        pyprogressivex.setSettings(None)
    again = pyprogressivex.findHomographies(corrs, 1024, 768, 1024, 768, **kw)
    assert np.array_equal(again[0], models) and np.array_equal(again[1], labels)
