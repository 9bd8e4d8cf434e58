"""GPU parity: every C-ABI operator of libpxb200.so against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): residuals, inlier masks, counts, labels and the four-point solver are BIT-EXACT;
sums agree to 1e-12 relative (fixed but different summation topology); F7 / P3P model parameters to
1e-6 relative (contract: 1e-5) because CUDA's cbrt/acos/cos differ from glibc's in the last ulp.
"""
import numpy as np
import pytest

from pyprogressivex import synthetic as syn

pytestmark = pytest.mark.gpu

H, F, PNP, VP, LINE = 0, 1, 2, 3, 4
ALL = [H, F, PNP, VP, LINE]
SUM_RTOL = 1e-12


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.uint64)
    return np.array_equal(a, b)


def scene(t, N, seed):
    """(points, models [K,ms], gt labels, threshold) with hypotheses from real minimal solves + planted models."""
    if t == H:
        pts, gt, Ms = syn.multi_homography_scene(N, seed=seed)
        thr = 2.0
    elif t == F:
        pts, gt, Ms = syn.multi_motion_scene(N, seed=seed)
        thr = 0.75
    elif t == PNP:
        img, w, K, gt, Ms = syn.multi_pose_scene(N, seed=seed)
        pts = syn.normalize_pnp_points(img, w, K)
        thr = 4.0 / (0.5 * (K[0, 0] + K[1, 1]))
    elif t == VP:
        pts, gt, Ms = syn.multi_vanishing_point_scene(N, seed=seed)
        thr = 2.0
    else:
        pts, gt, Ms = syn.multi_line_scene(N, seed=seed)
        thr = 2.0
    return pts, gt, Ms.reshape(Ms.shape[0], -1), thr


def hypotheses(oracle, t, pts, gt, planted, K, seed):
    m = {H: 4, F: 7, PNP: 3, VP: 2, LINE: 2}[t]
    S = syn.minimal_samples(gt, K, m, seed=seed)
    models, n, _, _ = oracle.solve_minimal(t, pts, S)
    flat = [models[k, j] for k in range(K) for j in range(n[k])]
    return np.concatenate([planted, np.asarray(flat)]) if flat else planted


def test_division_selftest(ctx):
    # the shared-reciprocal dual division of the hot loop must be bit-identical to div.rn.f64
    assert ctx.selftest_division(1, 200_000_000, 0) == 0
    assert ctx.selftest_division(2, 200_000_000, 1) == 0


@pytest.mark.parametrize("t", ALL)
@pytest.mark.parametrize("N", [1, 31, 64, 1000, 4097, 10000])
def test_residual_matrix_bit_exact(ctx, oracle, t, N):
    pts, gt, planted, thr = scene(t, max(N, 200), seed=N)
    pts, gt = pts[:N], gt[:N]
    models = hypotheses(oracle, t, *scene(t, 400, seed=N + 1)[:2], planted, 70, seed=N)
    T2 = (1.5 * thr) ** 2
    ctx.upload_points(t, pts)
    r2, mask = ctx.residual_matrix(models, T2)
    r2_o, mask_o = oracle.residual_matrix(t, pts, models, T2)
    assert bits_equal(r2, r2_o)
    assert np.array_equal(mask, mask_o)


def test_residual_matrix_nonfinite_models(ctx, oracle):
    """Degenerate hypotheses (zero row, NaN, inf) must produce the same NaN/inf pattern and an all-zero mask bit."""
    pts, gt, planted, thr = scene(H, 777, seed=5)
    bad = planted.copy()[:4]
    bad[0, 6:] = 0.0          # t3 == 0 -> division by zero
    bad[1, 0] = np.nan
    bad[2, 8] = np.inf
    bad[3, :] = 0.0
    ctx.upload_points(H, pts)
    r2, mask = ctx.residual_matrix(bad, 9.0)
    r2_o, mask_o = oracle.residual_matrix(H, pts, bad, 9.0)
    assert np.array_equal(np.isnan(r2), np.isnan(r2_o))
    fin = ~np.isnan(r2_o)
    assert bits_equal(r2[fin], r2_o[fin])
    assert np.array_equal(mask, mask_o)


@pytest.mark.parametrize("t", ALL)
@pytest.mark.parametrize("with_compound", [False, True])
def test_score_compound(ctx, oracle, t, with_compound):
    N = 9000
    pts, gt, planted, thr = scene(t, N, seed=21 + t)
    models = hypotheses(oracle, t, pts, gt, planted, 150, seed=3)
    T2 = (1.5 * thr) ** 2
    cp = None
    if with_compound:
        cp = oracle.compound_max(np.stack([oracle.preference_vector(t, pts, planted[k], 9 / 4 * thr * thr)
                                           for k in range(2)]))
    ctx.upload_points(t, pts)
    cnt, val, sh = ctx.score_compound(models, T2, cp)
    cnt_o, val_o, sh_o = oracle.score_batch(t, pts, models, T2, cp)
    assert np.array_equal(cnt, cnt_o)                       # inlier counts: exact
    np.testing.assert_allclose(val, val_o, rtol=SUM_RTOL, atol=1e-13)
    np.testing.assert_allclose(sh, sh_o, rtol=SUM_RTOL, atol=1e-13)
    # same sums whatever the batching (topology depends on N only)
    cnt2, val2, sh2 = ctx.score_compound(models[5:23], T2, cp)
    assert np.array_equal(cnt2, cnt[5:23]) and bits_equal(val2, val[5:23]) and bits_equal(sh2, sh[5:23])
    # the sequential reference scorer on the best hypothesis: same count, same inlier list
    k = int(np.argmax(cnt))
    s = oracle.get_score(t, pts, models[k], T2, cp, 2)
    assert s["count"] == cnt[k]
    assert np.array_equal(ctx.inliers(models[k], T2), s["inliers"])


@pytest.mark.parametrize("t", ALL)
def test_preference_tanimoto_compound(ctx, oracle, t):
    N = 5003
    pts, gt, planted, thr = scene(t, N, seed=31)
    T = 9.0 / 4.0 * thr * thr
    ctx.upload_points(t, pts)
    prefs = []
    for k in range(3):
        p = ctx.preference_vector(planted[k], T)
        assert bits_equal(p, oracle.preference_vector(t, pts, planted[k], T))
        prefs.append(p)
    prefs = np.stack(prefs)
    cm = ctx.compound_max(prefs)
    assert bits_equal(cm, oracle.compound_max(prefs))
    for a, b in ((prefs[0], cm), (prefs[0], prefs[1]), (prefs[2], prefs[2])):
        assert abs(ctx.tanimoto(a, b) - oracle.tanimoto(a, b)) <= 1e-12 * max(1.0, abs(oracle.tanimoto(a, b)))


def test_h4_solver_bit_exact(ctx, oracle):
    pts, gt, planted, thr = scene(H, 3000, seed=41)
    S = syn.minimal_samples(gt, 2000, 4, seed=41)
    S[0] = [5, 5, 9, 11]  # repeated point -> singular system, NaN/inf handling
    ctx.upload_points(H, pts)
    models, n, sv, mv = ctx.solve_minimal(S)
    models_o, n_o, sv_o, mv_o = oracle.solve_minimal(H, pts, S)
    assert np.array_equal(n, n_o) and np.array_equal(sv, sv_o) and np.array_equal(mv, mv_o)
    ok = n_o > 0
    assert ok.sum() > 1500
    assert bits_equal(models[ok, 0], models_o[ok, 0])


def _match_solutions(a, na, b, nb, rtol):
    """every solution of b appears in a (and vice versa) within rtol relative to the model's largest entry"""
    assert na == nb
    for j in range(nb):
        scale = np.abs(b[j]).max()
        errs = [np.abs(a[i] - b[j]).max() / scale for i in range(na)]
        assert min(errs) < rtol, (min(errs), a[:na], b[:nb])


def test_f7_solver(ctx, oracle):
    pts, gt, planted, thr = scene(F, 3000, seed=43)
    S = syn.minimal_samples(gt, 1500, 7, seed=43)
    ctx.upload_points(F, pts)
    models, n, _, _ = ctx.solve_minimal(S)
    models_o, n_o, _, _ = oracle.solve_minimal(F, pts, S)
    assert (n == n_o).mean() > 0.995  # root classification near a double root may differ by libm ulps
    for k in np.flatnonzero(n == n_o):
        _match_solutions(models[k], n[k], models_o[k], n_o[k], 1e-6)


def test_p3p_solver(ctx, oracle):
    img, w, K, gt, poses = syn.multi_pose_scene(3000, seed=47)
    pts = syn.normalize_pnp_points(img, w, K)
    S = syn.minimal_samples(gt, 1500, 3, seed=47)
    ctx.upload_points(PNP, pts)
    models, n, _, _ = ctx.solve_minimal(S)
    models_o, n_o, _, _ = oracle.solve_minimal(PNP, pts, S)
    assert (n == n_o).mean() > 0.995
    bad = 0
    for k in np.flatnonzero(n == n_o):
        try:
            _match_solutions(models[k], n[k], models_o[k], n_o[k], 1e-6)
        except AssertionError:
            bad += 1  # ill-conditioned triplets amplify the last-ulp libm differences
    assert bad <= 0.01 * len(n)


@pytest.mark.parametrize("t", ALL)
@pytest.mark.parametrize("lam", [0.0, 0.3])
def test_pearl_datacost_bit_exact(ctx, oracle, t, lam):
    pts, gt, planted, thr = scene(t, 6001, seed=51)
    ctx.upload_points(t, pts)
    D = ctx.pearl_datacost(planted, thr, lam)
    assert bits_equal(D, oracle.pearl_datacost(t, pts, planted, thr, lam))
    D0 = ctx.pearl_datacost(planted[:0], thr, lam)
    assert D0.shape == (6001, 1) and np.all(D0 == 1.0 - lam)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("label_cost", [10.0, 300.0])
def test_greedy_label_sweep_matches_reference_gco(ctx, oracle, seed, label_cost):
    """lambda = 0 (the Python default): labels identical to the reference's own gco-v3 solveGreedy."""
    rng = np.random.default_rng(seed)
    pts, gt, planted, thr = scene(H, 8000, seed=60 + seed)
    models = np.concatenate([planted, planted[:3] + rng.normal(0, 1e-5, (3, 9))])  # near-duplicate instances
    D = oracle.pearl_datacost(H, pts, models, thr, 0.0)
    ref = oracle.gco_pearl_label if oracle.have_gco_ref() else None
    for init in (None, rng.integers(0, D.shape[1], D.shape[0]).astype(np.int32)):
        lab, e = ctx.pearl_label(D, 0.0, label_cost, init_labels=init)
        if ref:
            lab_o, e_o, _ = ref(D, 0.0, label_cost, init_labels=init)
        else:
            lab_o, e_o = oracle.greedy_ufl(D, label_cost, init)
        assert np.array_equal(lab, lab_o)
        assert abs(e - e_o) <= 1e-11 * abs(e_o)


@pytest.mark.parametrize("t", ALL)
def test_segment_residual_sums(ctx, oracle, t):
    pts, gt, planted, thr = scene(t, 7000, seed=71)
    labels = np.where(gt < 0, planted.shape[0], gt).astype(np.int32)
    ctx.upload_points(t, pts)
    sums, counts = ctx.segment_residual_sums(planted, labels)
    sums_o, counts_o = oracle.segment_residual_sums(t, pts, planted, labels)
    assert np.array_equal(counts, counts_o)
    np.testing.assert_allclose(sums, sums_o, rtol=SUM_RTOL)


@pytest.mark.parametrize("t", ALL)
def test_lo_terms_and_tukey_bit_exact(ctx, oracle, t):
    pts, gt, planted, thr = scene(t, 4000, seed=81)
    ctx.upload_points(t, pts)
    d, e0, e1 = ctx.lo_unary_terms(planted[0], thr, 0.14)
    d_o, e0_o, e1_o = oracle.lo_unary_terms(t, pts, planted[0], thr, 0.14)
    assert bits_equal(d, d_o) and bits_equal(e0, e0_o) and bits_equal(e1, e1_o)
    T2 = (1.5 * thr) ** 2
    assert bits_equal(ctx.tukey_weights(planted[0], T2), oracle.tukey_weights(t, pts, planted[0], T2))


def test_errors_are_reported_not_fatal(ctx):
    from pyprogressivex import _native
    c2 = _native.Context(0)
    with pytest.raises(_native.PxbError):
        c2.lib.pxb_sync(None) and None
        _native._check(c2.lib.pxb_residual_matrix(c2.handle, None, 1, 1.0, None, None))
    c2.close()


@pytest.mark.parametrize("t,N,K", [(F, 50_000, 1_500), (PNP, 100_000, 1_000)])
def test_baseline_config_sizes_f_and_pnp(ctx, oracle, t, N, K):
    """BASELINE configs C3 (multi-F, 50k correspondences) and C5 (P3P, 100k 2D-3D matches) at full N: hypotheses from
    the GPU minimal solvers; mask == (r2 < T2) everywhere, popcount == fused count, a row sample is bit-exact against
    the oracle, and sharding the hypotheses into blocks (the multi-GPU mode) reproduces the unsharded sums bit for bit."""
    pts, gt, planted, thr = scene(t, N, seed=17 + t)
    m = {F: 7, PNP: 3}[t]
    S = syn.minimal_samples(gt, K, m, seed=17 + t)
    ctx.upload_points(t, pts)
    models, n, _, _ = ctx.solve_minimal(S)
    filled = np.arange(models.shape[1])[None, :] < n[:, None]
    hyps = np.ascontiguousarray(models[filled])
    assert hyps.shape[0] > K // 2
    T2 = (1.5 * thr) ** 2
    r2, mask = ctx.residual_matrix(hyps, T2)
    bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :N].astype(bool)
    assert np.array_equal(bits, r2 < T2)
    cnt, val, shr = ctx.score_compound(hyps, T2)
    assert np.array_equal(cnt, bits.sum(1))
    rows = np.random.default_rng(0).choice(hyps.shape[0], 24, replace=False)
    r2_o, mask_o = oracle.residual_matrix(t, pts, hyps[rows], T2)
    assert bits_equal(r2[rows], r2_o) and np.array_equal(mask[rows], mask_o)
    cnt_o, val_o, _ = oracle.score_batch(t, pts, hyps[rows], T2, None, threads=4)
    assert np.array_equal(cnt[rows], cnt_o)
    np.testing.assert_allclose(val[rows], val_o, rtol=SUM_RTOL, atol=1e-13)
    from pyprogressivex import sharding
    b = sharding.block_bounds(hyps.shape[0], 8)
    parts = [ctx.score_compound(hyps[b[r]:b[r + 1]], T2) for r in range(8)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), cnt)
    assert bits_equal(np.concatenate([p[1] for p in parts]), val)
    best = int(np.argmax(cnt))
    assert cnt[best] > 0.5 * (gt == gt[S[0][0]]).sum() or cnt[best] > 0.03 * N  # a planted structure was found


def test_full_size_properties(ctx, oracle):
    """BASELINE sizes (50k x 10k grid is too big for the CPU oracle): size-independent properties instead.
    (1) mask bit == (r2 < T2) for every entry, (2) popcount(mask row) == fused-score count, (3) a 64-row sample of
    the matrix is bit-exact against the oracle, (4) sum over rows of value_sum is invariant to hypothesis order."""
    N, K = 50_000, 2_000
    pts, gt, planted, thr = scene(H, N, seed=91)
    S = syn.minimal_samples(gt, K, 4, seed=91)
    ctx.upload_points(H, pts)
    models, n, sv, mv = ctx.solve_minimal(S)
    models = models[:, 0][n > 0]
    T2 = (1.5 * thr) ** 2
    r2, mask = ctx.residual_matrix(models, T2)
    bits = np.unpackbits(mask.view(np.uint8), axis=1, bitorder="little")[:, :N].astype(bool)
    assert np.array_equal(bits, r2 < T2)
    cnt, val, sh = ctx.score_compound(models, T2)
    assert np.array_equal(cnt, bits.sum(1))
    rows = np.random.default_rng(0).choice(models.shape[0], 64, replace=False)
    r2_o, mask_o = oracle.residual_matrix(H, pts, models[rows], T2)
    assert bits_equal(r2[rows], r2_o) and np.array_equal(mask[rows], mask_o)
    perm = np.random.default_rng(1).permutation(models.shape[0])
    cnt_p, val_p, _ = ctx.score_compound(models[perm], T2)
    assert np.array_equal(cnt_p, cnt[perm]) and bits_equal(val_p, val[perm])


# ---- float32 screening of the fused score kernel (pxb_screen.cuh): it may only ever dismiss certain outliers ----------
def _assert_score_equals_oracle(ctx, oracle, t, pts, models, T2, cp=None):
    ctx.upload_points(t, pts)
    cnt, val, sh = ctx.score_compound(models, T2, cp)
    cnt_o, val_o, sh_o = oracle.score_batch(t, pts, models, T2, cp)
    assert np.array_equal(cnt, cnt_o)
    np.testing.assert_allclose(val, val_o, rtol=SUM_RTOL, atol=1e-13)
    np.testing.assert_allclose(sh, sh_o, rtol=SUM_RTOL, atol=1e-13)
    return cnt


@pytest.mark.parametrize("t", ALL)
def test_score_screening_model_scale_and_garbage(ctx, oracle, t):
    """The inlier tests are homogeneous in the model, the screening rescales it: hypotheses multiplied by 1e+-150,
    zero / NaN / inf hypotheses and random garbage must give the oracle's counts."""
    pts, gt, planted, thr = scene(t, 5000, seed=40 + t)
    models = hypotheses(oracle, t, pts, gt, planted, 60, seed=4)
    rng = np.random.default_rng(7)
    ms = models.shape[1]
    extra = [models[:20] * 1e150, models[:20] * 1e-150, models[:20] * -3.0, rng.normal(size=(20, ms)),
             rng.normal(size=(20, ms)) * np.exp(rng.normal(size=(20, ms)) * 8), np.zeros((1, ms)),
             np.full((1, ms), np.nan), np.full((1, ms), np.inf)]
    bad = models[:3].copy()
    bad[0, -3:] = 0.0  # H: t3 == 0 everywhere; F/PnP: a zero last row
    bad[1, 0] = np.nan
    bad[2, -1] = np.inf
    allm = np.concatenate([models] + extra + [bad])
    T2 = (1.5 * thr) ** 2
    cnt = _assert_score_equals_oracle(ctx, oracle, t, pts, allm, T2)
    assert cnt[:len(models)].max() > 100  # the planted structures are found


@pytest.mark.parametrize("t", ALL)
def test_score_screening_wild_points_and_offsets(ctx, oracle, t):
    """Points far from the origin (normalisation must absorb a 1e7 offset), a few non-finite / astronomically large
    points (they must take the exact path), and a set whose bounding box is degenerate."""
    pts, gt, planted, thr = scene(t, 3000, seed=50 + t)
    models = hypotheses(oracle, t, pts, gt, planted, 40, seed=5)
    T2 = (1.5 * thr) ** 2
    wild = pts.copy()
    wild[5, 0] = np.nan
    wild[6, 1] = np.inf
    wild[7, :] = 1e200
    wild[8, -1] = -1e30
    wild[9, :] = 0.0
    _assert_score_equals_oracle(ctx, oracle, t, wild, models, T2)
    if t in (H, F):
        # translate both images by 1e7 px: H' = T2 H T1^-1, F' = T2^-T F T1^-1 keep the geometry; the counts of the
        # translated problem are whatever the oracle says they are (float64 loses digits too) -- parity is the bar
        off = pts + 1.0e7
        _assert_score_equals_oracle(ctx, oracle, t, off, models, T2)
    same = np.repeat(pts[:1], 700, axis=0)  # zero-extent bounding box
    _assert_score_equals_oracle(ctx, oracle, t, same, models, T2)


@pytest.mark.parametrize("T2", [0.0, -1.0, 1e-30, 1e-9, 1e9, 1e300, float("inf"), float("nan")])
def test_score_screening_threshold_extremes(ctx, oracle, T2):
    pts, gt, planted, thr = scene(H, 2000, seed=61)
    models = hypotheses(oracle, H, pts, gt, planted, 30, seed=6)
    _assert_score_equals_oracle(ctx, oracle, H, pts, models, T2)


@pytest.mark.parametrize("t", ALL)
def test_score_screening_near_threshold_band(ctx, oracle, t):
    """Points placed a hair inside / outside the threshold of a model (relative 1e-9 .. 1e-3): float32 cannot tell them
    apart, so they must reach the exact path and be classified exactly like the oracle does."""
    pts, gt, planted, thr = scene(t, 4000, seed=70 + t)
    T2 = (1.5 * thr) ** 2
    r2, _ = oracle.residual_matrix(t, pts, planted[:1], T2)
    r2 = r2[0]
    # thresholds chosen as residuals of actual points, nudged by tiny relative amounts
    order = np.argsort(r2)
    picks = r2[order[[50, 400, 900, 1500]]]
    for base in picks:
        for eps in (0.0, 1e-15, -1e-15, 1e-9, -1e-9, 1e-4, -1e-4):
            _assert_score_equals_oracle(ctx, oracle, t, pts, planted[:3], float(base * (1.0 + eps)))


@pytest.mark.parametrize("t", [VP, LINE])
def test_vp_and_line_minimal_solvers_bit_exact(ctx, oracle, t):
    pts, gt, planted, thr = scene(t, 3000, seed=101)
    S = syn.minimal_samples(gt, 2000, 2, seed=101)
    S[0] = [7, 7]  # degenerate: NaN model, produced all the same (the reference pushes it unconditionally)
    ctx.upload_points(t, pts)
    models, n, sv, mv = ctx.solve_minimal(S)
    models_o, n_o, sv_o, mv_o = oracle.solve_minimal(t, pts, S)
    assert np.array_equal(n, n_o) and np.array_equal(sv, sv_o) and np.array_equal(mv, mv_o)
    assert np.array_equal(np.isnan(models), np.isnan(models_o))
    fin = ~np.isnan(models_o)
    assert bits_equal(models[fin], models_o[fin])


@pytest.mark.parametrize("t", [VP, LINE])
def test_vp_and_line_nonminimal_fits(ctx, oracle, t):
    """Eigenvector / QR based fits: 1e-9 relative to the oracle up to the sign of the model (the residual is sign
    invariant; Eigen's own sign convention is not pinned), with and without per-point weights for VP."""
    pts, gt, planted, thr = scene(t, 4000, seed=103)
    rng = np.random.default_rng(3)
    sets = [np.flatnonzero(gt == k) for k in range(planted.shape[0])]
    sets += [rng.choice(s, 14, replace=False) for s in sets]
    sets += [np.flatnonzero(gt == 0)[:2]]
    ctx.upload_points(t, pts)
    weights = rng.uniform(0.2, 1.0, pts.shape[0]) if t == VP else None
    for w in ([None, weights] if t == VP else [None]):
        got, ok = ctx.fit_nonminimal(sets, w)
        for k, st in enumerate(sets):
            ref, ok_o = oracle.fit_nonminimal(t, pts, st, w)
            assert bool(ok[k]) == ok_o
            if not ok_o:
                continue
            sign = 1.0 if np.dot(got[k], ref) >= 0 else -1.0
            np.testing.assert_allclose(sign * got[k], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    # the fit recovers the planted structure
    got, ok = ctx.fit_nonminimal(sets[:planted.shape[0]])
    T2 = (1.5 * thr) ** 2
    cnt, _, _ = ctx.score_compound(got, T2)
    for k in range(planted.shape[0]):
        assert cnt[k] > 0.9 * len(sets[k])


@pytest.mark.parametrize("t", ALL)
@pytest.mark.parametrize("N", [1, 5, 33, 127, 129, 1023, 1025, 2049])
def test_score_compound_ragged_sizes(ctx, oracle, t, N):
    """Point counts around the warp (32), chunk (128) and block (1024) boundaries of the score kernel, K not a multiple
    of the hypothesis tile, with and without a compound preference vector."""
    pts, gt, planted, thr = scene(t, max(N, 300), seed=200 + N)
    models = hypotheses(oracle, t, pts, gt, planted, 37, seed=7)
    pts = np.ascontiguousarray(pts[:N])
    T2 = (1.5 * thr) ** 2
    cp = oracle.preference_vector(t, pts, planted[0], 9 / 4 * thr * thr)
    for comp in (None, cp):
        _assert_score_equals_oracle(ctx, oracle, t, pts, models, T2, comp)
    _assert_score_equals_oracle(ctx, oracle, t, pts, models[:1], T2, cp)


def test_plane_parallax_solver_bit_exact(ctx, oracle):
    """DEGENSAC's two-point solver over a fixed homography: bit-exact against the oracle restatement, and the algebra the
    reference relies on: the plane's points and both sample points satisfy the epipolar constraint of F = [e]x H."""
    rows, lab, _ = syn.plane_dominated_pair(400, 400, 0.0, 11)
    plane, off = np.flatnonzero(lab == 0), np.flatnonzero(lab == 1)
    Hm, ok = oracle.fit_h_nonminimal(rows, plane)
    assert ok
    rng = np.random.default_rng(2)
    S = np.stack([rng.choice(off, 2, replace=False) for _ in range(3000)]).astype(np.int64)
    S[0] = [off[3], off[3]]  # the same correspondence twice: both lines coincide, epipole = 0 -> no model
    ctx.upload_points(F, rows)
    models, n = ctx.solve_plane_parallax(S, Hm)
    models_o, n_o = oracle.solve_plane_parallax(rows, S, Hm)
    assert np.array_equal(n, n_o) and n[0] == 0 and n[1:].all()
    assert bits_equal(models, models_o)
    x1 = np.column_stack([rows[:, :2], np.ones(len(rows))])
    x2 = np.column_stack([rows[:, 2:], np.ones(len(rows))])
    for k in (1, 17, 2999):
        Fm = models[k].reshape(3, 3)
        Fm = Fm / np.linalg.norm(Fm)
        alg = np.abs(np.einsum("ni,ij,nj->n", x2, Fm, x1))
        assert alg[plane].max() < 1e-6 and alg[S[k]].max() < 1e-6
        assert np.median(alg[off]) < 1e-6  # noise-free rigid scene: every point is on the true epipolar geometry
    with pytest.raises(Exception):
        ctx.solve_plane_parallax(np.array([[0, len(rows)]]), Hm)


@pytest.mark.parametrize("t", [0, 1, 2, 3, 4], ids=["H", "F", "PnP", "VP", "Line"])
def test_screened_bit_matrix_equals_exact_mask(ctx, t):
    from pyprogressivex import _native
    """pxb_residual_matrix(r2 = NULL): the float32-screened bit-matrix kernel (k_mask_screened) must give exactly the mask
    the float64 matrix kernel writes beside r2 -- on planted models, solver output incl. empty solution slots, degenerate
    models (zero, NaN, inf, huge), a threshold with a non-zero low word, and a ragged N."""
    rng = np.random.default_rng(5 + t)
    N = 7013
    if t == 0:
        pts, gt, planted = syn.multi_homography_scene(N, seed=40)
        thr = 2.0
    elif t == 1:
        pts, gt, planted = syn.multi_motion_scene(N, seed=41)
        thr = 0.75
    elif t == 2:
        img, w, K, gt, planted = syn.multi_pose_scene(N, n_objects=4, inlier_ratio_each=0.15, seed=42)
        pts = syn.normalize_pnp_points(img, w, K)
        thr = 4.0 / 1074.0
    elif t == 3:
        pts, gt, planted = syn.multi_vanishing_point_scene(N, seed=43)
        thr = 2.0
    else:
        pts, gt, planted = syn.multi_line_scene(N, seed=44)
        thr = 2.0
    ms = _native.MODEL_SIZE[t]
    ctx.upload_points(t, pts)
    S = syn.minimal_samples(gt, 150, _native.SAMPLE_SIZE[t], seed=t)
    models, n, _, _ = ctx.solve_minimal(S)
    hyps = [planted.reshape(-1, ms), models.reshape(-1, ms)]  # all slots, filled or not
    bad = np.tile(planted.reshape(-1, ms)[:1], (6, 1))
    bad[0] = 0.0
    bad[1, 0] = np.nan
    bad[2, -1] = np.inf
    bad[3] *= 1e200
    bad[4] *= 1e-200
    bad[5] = rng.normal(size=ms)
    hyps = np.ascontiguousarray(np.concatenate(hyps + [bad]))
    for T2 in ((1.5 * thr) ** 2, (1.5 * thr) ** 2 * (1 + 2.0 ** -40), 1e-30, 1e30):
        _, exact = ctx.residual_matrix(hyps, T2, want_r2=True, want_mask=True)
        _, screened = ctx.residual_matrix(hyps, T2, want_r2=False, want_mask=True)
        assert np.array_equal(exact, screened)
