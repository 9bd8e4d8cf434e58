"""GPU min-cut engine against the REFERENCE's own gco-v3 / Boykov-Kolmogorov build (oracle/_ref/libgco_ref.so):
alpha-expansion labels of PEARL (lambda > 0), the st-cut of the GC-RANSAC local optimisation, plus the two "next"
operators the driver uses (kNN graph, non-minimal homography fit)."""
import numpy as np
import pytest

from pyprogressivex import synthetic as syn

pytestmark = pytest.mark.gpu
H = 0


def _need_ref(oracle):
    if not oracle.have_gco_ref():
        pytest.fail("oracle/_ref/libgco_ref.so missing: it is built in the build container and travels with the repo")


def _assert_equal_up_to_exact_ties(oracle, D, lam, label_cost, off, idx, lab, lab_ref):
    """Labels must be identical, except at sites where the reference's own energy is EXACTLY tied between the two
    labels (e.g. a far outlier that costs 2(1-lambda) under both instances and has as many disagreeing edges either
    way). The reference resolves such ties by the rounding noise of BK's flow value (GCoptimization.cpp:1286 compares
    Econst + flow with a separately accumulated sum), which no other max-flow order can reproduce."""
    diff = np.flatnonzero(lab != lab_ref)
    assert diff.size <= max(1, lab.size // 1000), f"{diff.size} of {lab.size} labels differ"
    e_ref = oracle.gco_energy(D, lam, label_cost, off, idx, lab_ref)
    for i in diff:
        swapped = lab_ref.copy()
        swapped[i] = lab[i]
        assert oracle.gco_energy(D, lam, label_cost, off, idx, swapped) == e_ref, f"site {i} differs without a tie"


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("lam", [0.05, 0.3])
def test_alpha_expansion_labels_match_reference_gco(ctx, oracle, seed, lam, monkeypatch):
    _need_ref(oracle)
    monkeypatch.setenv("PXB_CHECK_ENERGY", "1")  # the incremental energy bookkeeping must equal the full edge walk
    rng = np.random.default_rng(seed)
    N = [600, 1500, 4000, 9000][seed]
    pts, gt, Hs = syn.multi_homography_scene(N, n_planes=3 + seed % 2, outlier_ratio=0.35, seed=100 + seed)
    models = Hs.reshape(-1, 9)
    models = np.concatenate([models, models[:1] + rng.normal(0, 1e-5, (1, 9))])  # a redundant instance to eliminate
    D = oracle.pearl_datacost(H, pts, models, 2.0, lam)
    off, idx = syn.knn_graph(pts, 80.0, 5)
    for label_cost, init in ((10.0, None), (40.0, rng.integers(0, D.shape[1], N).astype(np.int32))):
        lab_o, e_o, cyc = oracle.gco_pearl_label(D, lam, label_cost, off, idx, init)
        lab, e = ctx.pearl_label(D, lam, label_cost, off, idx, init)
        assert e == e_o  # energies are evaluated in the reference's summation order: bit-exact
        assert abs(oracle.gco_energy(D, lam, label_cost, off, idx, lab) - e) == 0.0
        _assert_equal_up_to_exact_ties(oracle, D, lam, label_cost, off, idx, lab, lab_o)


def test_alpha_expansion_with_ties_and_duplicates(ctx, oracle):
    """Structural ties: far outliers cost 2(1-lambda) under every instance; mutual neighbours give doubled edges."""
    _need_ref(oracle)
    pts, gt, Hs = syn.multi_homography_scene(2500, n_planes=2, outlier_ratio=0.6, seed=9)
    D = oracle.pearl_datacost(H, pts, Hs.reshape(-1, 9), 1.0, 0.2)
    off, idx = syn.knn_graph(pts, 300.0, 6)
    lab_o, e_o, _ = oracle.gco_pearl_label(D, 0.2, 25.0, off, idx)
    lab, e = ctx.pearl_label(D, 0.2, 25.0, off, idx)
    assert e == e_o
    _assert_equal_up_to_exact_ties(oracle, D, 0.2, 25.0, off, idx, lab, lab_o)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("lam", [0.14, 0.6, 0.975])
def test_lo_graph_cut_matches_reference_bk(ctx, oracle, seed, lam):
    _need_ref(oracle)
    N = 5000
    pts, gt, Hs = syn.multi_homography_scene(N, n_planes=3, seed=200 + seed)
    ctx.upload_points(H, pts)
    model = Hs[seed % 3].reshape(-1) + np.random.default_rng(seed).normal(0, 1e-6, 9)
    d, e0, e1 = ctx.lo_unary_terms(model, 2.0, lam)
    off, idx = ctx.knn_graph(60.0, 8)
    seg = ctx.lo_graph_cut(e0, e1, d, lam, off, idx)
    seg_o, _ = oracle.gco_lo_labeling(e0, e1, d, lam, off, idx)
    assert np.array_equal(seg, seg_o)
    assert 0 < seg.sum() < N
    # the device-resident form (unary terms, graph capacities and cut without leaving the GPU; cached arc skeleton)
    assert np.array_equal(ctx.lo_labeling(model, 2.0, lam, off, idx), seg_o)
    other = Hs[(seed + 1) % 3].reshape(-1)  # second model on the same (cached) skeleton
    d2, e02, e12 = oracle.lo_unary_terms(H, pts, other, 2.0, lam)
    assert np.array_equal(ctx.lo_labeling(other, 2.0, lam, off, idx), oracle.gco_lo_labeling(e02, e12, d2, lam, off, idx)[0])


def test_knn_graph_matches_kdtree(ctx):
    pts, gt, Hs = syn.multi_homography_scene(4000, seed=3)
    ctx.upload_points(H, pts)
    off, idx = ctx.knn_graph(50.0, 8)
    off_o, idx_o = syn.knn_graph(pts, 50.0, 8)
    assert np.array_equal(off, off_o)
    # identical neighbour sets per point (order may differ only between exactly equidistant neighbours)
    for i in range(0, 4000, 37):
        assert set(idx[off[i]:off[i + 1]]) == set(idx_o[off_o[i]:off_o[i + 1]])


def _knn_bruteforce(rows, radius, k):
    """the rule k_knn_graph implements: the k smallest (squared distance, index) pairs within the radius, self excluded;
    squared distances summed coordinate by coordinate in float64 (no fused multiply-add), ties to the lower index"""
    N = rows.shape[0]
    out = []
    for i in range(N):
        d2 = np.zeros(N)
        for c in range(rows.shape[1]):
            d = rows[i, c] - rows[:, c]
            d2 = d2 + d * d
        d2[i] = np.inf
        order = np.argsort(d2, kind="stable")[:k]
        out.append([int(j) for j in order if d2[j] <= radius * radius])
    return out


@pytest.mark.parametrize("t,N,radius,k", [(0, 3001, 60.0, 5), (0, 1237, 400.0, 12), (4, 2050, 40.0, 8), (2, 1500, 0.35, 5), (0, 7, 1e9, 8)])
def test_knn_graph_exact_order(ctx, t, N, radius, k):
    """every neighbour list is exactly the sequential scan's (order included): 64 x SLICES threads per block split the
    candidates of a query into index ranges and merge their lists; N not a multiple of anything, k <= 8 and k > 8
    instantiations, 2 / 4 / 5 coordinates, duplicated points (exact distance ties)"""
    rng = np.random.default_rng(N)
    if t == 0:
        rows, _, _ = syn.multi_homography_scene(N, seed=5)
    elif t == 4:
        rows = np.round(rng.uniform(0, 500, size=(N, 2)))  # integer coordinates: many exactly equal distances
    else:
        rows = np.concatenate([rng.normal(size=(N, 2)), rng.uniform(-1, 1, size=(N, 3))], axis=1)
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    rows[N // 2] = rows[N // 3]  # a duplicated point
    ctx.upload_points(t, rows)
    off, idx = ctx.knn_graph(radius, k)
    want = _knn_bruteforce(rows, radius, k)
    assert off[-1] == sum(len(w) for w in want)
    for i in range(N):
        assert idx[off[i]:off[i + 1]].tolist() == want[i], i


def test_nonminimal_homography_fit(ctx, oracle):
    """normal equations on the device vs column-pivoted Householder QR in the oracle: same least-squares solution"""
    pts, gt, Hs = syn.multi_homography_scene(6000, n_planes=3, noise=0.5, seed=13)
    ctx.upload_points(H, pts)
    rng = np.random.default_rng(0)
    sets = [np.flatnonzero(gt == 0), np.flatnonzero(gt == 1)[:28], rng.choice(np.flatnonzero(gt == 2), 4, replace=False),
            np.flatnonzero(gt == 2)[:3]]
    Hg, ok = ctx.fit_homographies(sets)
    assert ok.tolist() == [1, 1, 1, 0]
    for k in range(3):
        Ho, oko = oracle.fit_h_nonminimal(pts, sets[k])
        assert oko
        np.testing.assert_allclose(Hg[k] / Hg[k][8], Ho / Ho[8], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(Hg[k], Ho, rtol=1e-6, atol=1e-9)
    # weighted fit (IRLS): weights follow the reference's row indexing
    w = rng.uniform(0.1, 1.0, sets[0].size)
    Hw, okw = ctx.fit_homographies([sets[0]], w)
    How, _ = oracle.fit_h_nonminimal(pts, sets[0], w)
    np.testing.assert_allclose(Hw[0], How, rtol=1e-6, atol=1e-9)
    # the fit explains its own inliers
    r2, _ = oracle.residual_matrix(H, pts[sets[0]], Hg[0], 9.0)
    assert np.median(r2) < 1.0


@pytest.mark.parametrize("layout", ["offsets_in_smem", "heights_and_queue_only", "grid_wide", "blocks_leave_on_their_own"])
def test_global_relabel_layouts_give_the_same_cut(ctx, oracle, layout, monkeypatch):
    """k_maxflow's global relabel has three forms: block 0 alone with heights + queue + CSR offsets in
    shared memory, the same without the offsets (larger graphs), and the grid-wide level-synchronous BFS (graphs that do
    not fit). All three must reproduce the reference's cut and labels."""
    _need_ref(oracle)
    N = 4000  # 12 B/node -> 47 KB, 8 B/node -> 31 KB
    if layout == "heights_and_queue_only":
        monkeypatch.setenv("PXB_MF_SMEM_KB", "40")
    elif layout == "grid_wide":
        monkeypatch.setenv("PXB_MF_GRID_BFS", "1")
    elif layout == "blocks_leave_on_their_own":  # the other rule for ending the asynchronous phase
        monkeypatch.setenv("PXB_MF_LOCAL_EXIT", "1")
    pts, gt, Hs = syn.multi_homography_scene(N, n_planes=3, outlier_ratio=0.4, seed=77)
    lam = 0.1
    D = oracle.pearl_datacost(H, pts, Hs.reshape(-1, 9), 2.0, lam)
    off, idx = syn.knn_graph(pts, 80.0, 5)
    lab_o, e_o, _ = oracle.gco_pearl_label(D, lam, 20.0, off, idx)
    lab, e = ctx.pearl_label(D, lam, 20.0, off, idx)
    assert e == e_o
    _assert_equal_up_to_exact_ties(oracle, D, lam, 20.0, off, idx, lab, lab_o)
    ctx.upload_points(H, pts)
    model = Hs[0].reshape(-1)
    d, e0, e1 = oracle.lo_unary_terms(H, pts, model, 2.0, 0.6)
    assert np.array_equal(ctx.lo_labeling(model, 2.0, 0.6, off, idx), oracle.gco_lo_labeling(e0, e1, d, 0.6, off, idx)[0])
