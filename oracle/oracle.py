"""ctypes loader of the parity oracle. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module
(see oracle/pxo_oracle.h). The product package never does.

  libpxoracle.so        CPU restatement of rows a1..a9, a12, a13 (+ a restated greedy UFL)
  _ref/libgco_ref.so    the reference's own gco-v3 / BK max-flow sources (rows a10/a11, and the LO st-cut)
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "libpxoracle.so"
GCO_REF_LIB = HERE / "_ref" / "libgco_ref.so"

MODEL_H, MODEL_F, MODEL_PNP, MODEL_VP, MODEL_LINE = 0, 1, 2, 3, 4
DIM = {0: 4, 1: 4, 2: 5, 3: 4, 4: 2}
MSIZE = {0: 9, 1: 9, 2: 12, 3: 3, 4: 3}
SSIZE = {0: 4, 1: 7, 2: 3, 3: 2, 4: 2}
MAXSOL = {0: 1, 1: 3, 2: 4, 3: 1, 4: 1}


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference is present, the reference gco build)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    if ref and Path("/root/reference").is_dir():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


_lib = None
_gco = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not ORACLE_LIB.exists():
            build(ref=False)
        L = C.CDLL(str(ORACLE_LIB))
        vp, i64, f64 = C.c_void_p, C.c_int64, C.c_double
        L.pxo_squared_residual.restype = f64
        L.pxo_squared_residual.argtypes = [C.c_int, vp, vp]
        L.pxo_residual_matrix.argtypes = [C.c_int, vp, i64, vp, i64, f64, vp, vp]
        L.pxo_residual_matrix.restype = None
        L.pxo_get_score.restype = f64
        L.pxo_get_score.argtypes = [C.c_int, vp, i64, vp, f64, vp, C.c_int, i64, C.POINTER(i64), C.POINTER(f64),
                                    C.POINTER(f64), vp]
        L.pxo_score_batch.argtypes = [C.c_int, vp, i64, vp, i64, f64, vp, vp, vp, vp, C.c_int]
        L.pxo_score_batch.restype = None
        L.pxo_preference_vector.argtypes = [C.c_int, vp, i64, vp, f64, vp]
        L.pxo_preference_vector.restype = None
        L.pxo_tanimoto.restype = f64
        L.pxo_tanimoto.argtypes = [vp, vp, i64]
        L.pxo_compound_max.argtypes = [vp, i64, i64, vp]
        L.pxo_compound_max.restype = None
        L.pxo_h4_solve.argtypes = [vp, vp, vp]
        L.pxo_h4_is_valid_sample.argtypes = [vp, vp]
        L.pxo_h_is_valid_model.argtypes = [vp]
        L.pxo_f7_solve.argtypes = [vp, vp, vp, C.c_int]
        L.pxo_p3p_solve.argtypes = [vp, vp, vp]
        L.pxo_pearl_datacost.argtypes = [C.c_int, vp, i64, vp, i64, f64, f64, vp]
        L.pxo_pearl_datacost.restype = None
        L.pxo_segment_residual_sums.argtypes = [C.c_int, vp, i64, vp, i64, vp, vp, vp]
        L.pxo_segment_residual_sums.restype = None
        L.pxo_lo_unary_terms.argtypes = [C.c_int, vp, i64, vp, f64, f64, vp, vp, vp]
        L.pxo_lo_unary_terms.restype = None
        L.pxo_tukey_weights.argtypes = [C.c_int, vp, vp, i64, vp, f64, vp]
        L.pxo_tukey_weights.restype = None
        L.pxo_vp2_solve.argtypes = [vp, vp, vp]
        L.pxo_line2_solve.argtypes = [vp, vp, vp]
        L.pxo_fit_vp_nonminimal.argtypes = [vp, vp, i64, vp, vp]
        L.pxo_fit_line_nonminimal.argtypes = [vp, vp, i64, vp]
        L.pxo_greedy_ufl.restype = f64
        L.pxo_greedy_ufl.argtypes = [vp, i64, C.c_int32, f64, vp, vp]
        _lib = L
    return _lib


def have_gco_ref() -> bool:
    return GCO_REF_LIB.exists()


def gco() -> C.CDLL:
    global _gco
    if _gco is None:
        if not GCO_REF_LIB.exists():
            build(ref=True)
        G = C.CDLL(str(GCO_REF_LIB))
        vp, f64 = C.c_void_p, C.c_double
        G.gco_ref_pearl_label.restype = f64
        G.gco_ref_pearl_label.argtypes = [C.c_int, C.c_int, vp, f64, f64, vp, vp, vp, vp, C.POINTER(C.c_int)]
        G.gco_ref_energy.restype = f64
        G.gco_ref_energy.argtypes = [C.c_int, C.c_int, vp, f64, f64, vp, vp, vp]
        G.gco_ref_lo_labeling.restype = f64
        G.gco_ref_lo_labeling.argtypes = [C.c_int, vp, vp, vp, f64, vp, vp, vp]
        _gco = G
    return _gco


REF_BODIES_LIB = HERE / "_ref" / "libpx_refbodies.so"
_refb = None


def have_ref_bodies() -> bool:
    return REF_BODIES_LIB.exists()


def refb() -> C.CDLL:
    """oracle/_ref/libpx_refbodies.so: the reference's own function bodies, compiled verbatim by extract_ref.py."""
    global _refb
    if _refb is None:
        if not REF_BODIES_LIB.exists():
            subprocess.run(["python", str(HERE / "extract_ref.py")], check=True)
        B = C.CDLL(str(REF_BODIES_LIB))
        vp, i64, f64 = C.c_void_p, C.c_int64, C.c_double
        B.pxr_squared_residual.restype = f64
        B.pxr_squared_residual.argtypes = [C.c_int, vp, vp]
        B.pxr_residuals.argtypes = [C.c_int, vp, i64, vp, vp]
        B.pxr_residuals.restype = None
        B.pxr_get_score.restype = f64
        B.pxr_get_score.argtypes = [C.c_int, vp, i64, vp, f64, vp, C.c_int, i64, C.POINTER(i64), vp]
        B.pxr_preference_vector.argtypes = [C.c_int, vp, i64, vp, f64, vp]
        B.pxr_preference_vector.restype = None
        B.pxr_pearl_datacost.argtypes = [C.c_int, vp, i64, vp, i64, f64, f64, vp]
        B.pxr_pearl_datacost.restype = None
        B.pxr_h4_solve.argtypes = [vp, i64, vp, vp]
        B.pxr_h4_is_valid_sample.argtypes = [vp, i64, vp]
        B.pxr_h_is_valid_model.argtypes = [vp]
        B.pxr_f_orientation_valid.argtypes = [vp, vp, i64, vp, C.c_int]
        B.pxr_vp2_solve.argtypes = [vp, i64, vp, vp]
        B.pxr_line2_solve.argtypes = [vp, i64, vp, vp]
        B.pxr_fpp_solve.argtypes = [vp, i64, vp, vp, vp]
        _refb = B
    return _refb


def ref_residuals(t, pts, model):
    pts, model = _f(pts), _f(model)
    out = np.empty(pts.shape[0])
    refb().pxr_residuals(t, _p(pts), pts.shape[0], _p(model), _p(out))
    return out


def ref_get_score(t, pts, model, T2, compound_pref=None, exponent=2, best_inlier_number=0):
    pts, model = _f(pts), _f(model)
    N = pts.shape[0]
    cp = None if compound_pref is None else _f(compound_pref)
    cnt = C.c_int64()
    inl = np.full(N, -1, dtype=np.int64)
    val = refb().pxr_get_score(t, _p(pts), N, _p(model), float(T2), _p(cp), int(exponent), int(best_inlier_number),
                               C.byref(cnt), _p(inl))
    return dict(value=val, count=cnt.value, inliers=inl[: cnt.value].copy())


def ref_preference_vector(t, pts, model, T):
    pts, model = _f(pts), _f(model)
    out = np.empty(pts.shape[0])
    refb().pxr_preference_vector(t, _p(pts), pts.shape[0], _p(model), float(T), _p(out))
    return out


def ref_pearl_datacost(t, pts, models, thr, lam):
    pts, models = _f(pts), _f(models).reshape(-1, MSIZE[t])
    N, L = pts.shape[0], models.shape[0]
    D = np.empty((N, L + 1))
    refb().pxr_pearl_datacost(t, _p(pts), N, _p(models), L, float(thr), float(lam), _p(D))
    return D


def ref_h4(pts, sample):
    pts = _f(pts)
    s = np.ascontiguousarray(sample, dtype=np.int64)
    H = np.zeros(9)
    B = refb()
    ok = B.pxr_h4_solve(_p(pts), pts.shape[0], _p(s), _p(H))
    sv = B.pxr_h4_is_valid_sample(_p(pts), pts.shape[0], _p(s))
    mv = B.pxr_h_is_valid_model(_p(H)) if ok else 0
    return H, int(ok), int(sv), int(mv)


def ref_minimal_vp_or_line(t, pts, sample):
    """The reference's own two-segment vanishing-point solver / two-point line solver (verbatim bodies)."""
    pts = _f(pts)
    s = np.ascontiguousarray(sample, dtype=np.int64)
    out = np.zeros(3)
    fn = refb().pxr_vp2_solve if t == MODEL_VP else refb().pxr_line2_solve
    ok = fn(_p(pts), pts.shape[0], _p(s), _p(out))
    return out, int(ok)


def ref_plane_parallax(pts, sample, H):
    """The reference's own FundamentalMatrixPlaneParallaxSolver::estimateModel body (matrix products through the shim)."""
    pts, H = _f(pts), _f(H).reshape(9)
    s = np.ascontiguousarray(sample, dtype=np.int64)
    out = np.zeros(9)
    ok = refb().pxr_fpp_solve(_p(pts), pts.shape[0], _p(s), _p(H), _p(out))
    return out, int(ok)


def ref_f_orientation_valid(F, pts, sample):
    F, pts = _f(F), _f(pts)
    s = np.ascontiguousarray(sample, dtype=np.int64)
    return int(refb().pxr_f_orientation_valid(_p(F), _p(pts), pts.shape[0], _p(s), s.shape[0]))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def residual_matrix(t, pts, models, T2, want_r2=True, want_mask=True):
    pts, models = _f(pts), _f(models).reshape(-1, MSIZE[t])
    N, K = pts.shape[0], models.shape[0]
    words = (N + 31) // 32
    r2 = np.empty((K, N)) if want_r2 else None
    mask = np.zeros((K, words), dtype=np.uint32) if want_mask else None
    lib().pxo_residual_matrix(t, _p(pts), N, _p(models), K, float(T2), _p(r2), _p(mask))
    return r2, mask


def get_score(t, pts, model, T2, compound_pref=None, exponent=2, best_inlier_number=0):
    pts, model = _f(pts), _f(model)
    N = pts.shape[0]
    cp = None if compound_pref is None else _f(compound_pref)
    cnt, val, sh = C.c_int64(), C.c_double(), C.c_double()
    inl = np.empty(N, dtype=np.int64)
    final = lib().pxo_get_score(t, _p(pts), N, _p(model), float(T2), _p(cp), int(exponent), int(best_inlier_number),
                                C.byref(cnt), C.byref(val), C.byref(sh), _p(inl))
    return dict(value=final, count=cnt.value, value_sum=val.value, shared=sh.value, inliers=inl[: cnt.value].copy())


def score_batch(t, pts, models, T2, compound_pref=None, threads=1):
    pts, models = _f(pts), _f(models).reshape(-1, MSIZE[t])
    N, K = pts.shape[0], models.shape[0]
    cp = None if compound_pref is None else _f(compound_pref)
    cnt = np.empty(K, dtype=np.int64)
    val = np.empty(K)
    sh = np.empty(K)
    lib().pxo_score_batch(t, _p(pts), N, _p(models), K, float(T2), _p(cp), _p(cnt), _p(val), _p(sh), int(threads))
    return cnt, val, sh


def preference_vector(t, pts, model, T):
    pts, model = _f(pts), _f(model)
    out = np.empty(pts.shape[0])
    lib().pxo_preference_vector(t, _p(pts), pts.shape[0], _p(model), float(T), _p(out))
    return out


def tanimoto(a, b):
    a, b = _f(a), _f(b)
    return lib().pxo_tanimoto(_p(a), _p(b), a.shape[0])


def compound_max(prefs):
    prefs = _f(prefs)
    out = np.empty(prefs.shape[1])
    lib().pxo_compound_max(_p(prefs), prefs.shape[0], prefs.shape[1], _p(out))
    return out


def solve_minimal(t, pts, samples):
    """Same output convention as Context.solve_minimal: (models [K,maxsol,ms], n [K], sample_valid, model_valid)."""
    pts = _f(pts)
    s = np.ascontiguousarray(samples, dtype=np.int64).reshape(-1, SSIZE[t])
    K = s.shape[0]
    models = np.zeros((K, MAXSOL[t], MSIZE[t]))
    n = np.zeros(K, dtype=np.int32)
    sv = np.ones(K, dtype=np.uint8)
    mv = np.ones(K, dtype=np.uint8)
    L = lib()
    buf = np.zeros(MAXSOL[t] * MSIZE[t])
    for k in range(K):
        row = np.ascontiguousarray(s[k])
        if t == MODEL_H:
            sv[k] = L.pxo_h4_is_valid_sample(_p(pts), _p(row))
            n[k] = L.pxo_h4_solve(_p(pts), _p(row), _p(buf))
            if n[k]:
                models[k, 0] = buf[:9]
                mv[k] = L.pxo_h_is_valid_model(_p(buf))
            else:
                mv[k] = 0
        elif t == MODEL_F:
            n[k] = L.pxo_f7_solve(_p(pts), _p(row), _p(buf), 1)
            models[k].reshape(-1)[: n[k] * 9] = buf[: n[k] * 9]
        elif t == MODEL_PNP:
            n[k] = L.pxo_p3p_solve(_p(pts), _p(row), _p(buf))
            models[k].reshape(-1)[: n[k] * 12] = buf[: n[k] * 12]
        else:
            n[k] = (L.pxo_vp2_solve if t == MODEL_VP else L.pxo_line2_solve)(_p(pts), _p(row), _p(buf))
            models[k, 0] = buf[:3]
            if not np.isfinite(buf[:3]).all():  # the reference pushes the model whatever it contains
                pass
    return models, n, sv, mv


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _matvec3(M, v):
    """row-by-row (a*b + c*d) + e*f, the order Eigen's 3x3 product evaluates in (no FMA on the reference's x86-64 build)"""
    return np.array([(M[r, 0] * v[0] + M[r, 1] * v[1]) + M[r, 2] * v[2] for r in range(3)])


def solve_plane_parallax(pts, samples, H):
    """FundamentalMatrixPlaneParallaxSolver::estimateModel
    (gcr/estimators/solver_fundamental_matrix_plane_and_parallax.h:105-163), restated; pinned bit for bit against the
    reference's own body in tests/test_oracle_pinned.py (its two Eigen 3x3 products go through the shim's operator*)."""
    pts = _f(pts)
    H = _f(H).reshape(3, 3)
    s = np.ascontiguousarray(samples, dtype=np.int64).reshape(-1, 2)
    models = np.zeros((s.shape[0], 9))
    n = np.zeros(s.shape[0], dtype=np.int32)
    for k, (a, b) in enumerate(s):
        lines = []
        for i in (a, b):
            src = np.array([pts[i, 0], pts[i, 1], 1.0])
            dst = np.array([pts[i, 2], pts[i, 3], 1.0])
            lines.append(_cross(_matvec3(H, src), dst))                      # :136-142
        e = _cross(lines[0], lines[1])                                       # :145
        if not abs(e[2]) >= np.finfo(np.float64).eps:                         # :148-149
            continue
        ex = np.array([[0.0, -e[2], e[1]], [e[2], 0.0, -e[0]], [-e[1], e[0], 0.0]])   # :152-155
        F = np.array([[(ex[r, 0] * H[0, c] + ex[r, 1] * H[1, c]) + ex[r, 2] * H[2, c] for c in range(3)] for r in range(3)])
        models[k] = F.reshape(9)                                             # :158-160
        n[k] = 1
    return models, n


def h_degenerate_sample(rows, sample7, F, homography_threshold=2.0):
    """The seven-point H-degeneracy test of FundamentalMatrixEstimator::applyDegensac
    (gcr/estimators/fundamental_estimator.h:352-476), restated with numpy's SVD for the epipole.
    Returns (degenerate, H, margins): margins = for every triplet the sorted transfer errors of the other four points
    (so that tests can skip samples that sit on the 2 px decision boundary)."""
    rows = _f(rows).reshape(-1, 4)
    F = _f(F).reshape(3, 3)
    U, _, _ = np.linalg.svd(F)
    e = U[:, 2] / U[2, 2]                                                    # :370-372
    ex = np.array([[0.0, -e[2], e[1]], [e[2], 0.0, -e[0]], [-e[1], e[0], 0.0]])
    A = ex @ F                                                               # :380-381
    triplets = [(0, 1, 2), (3, 4, 5), (0, 1, 6), (3, 4, 6), (2, 5, 6)]       # :352-357
    margins = []
    for trip in triplets:
        ids = [int(sample7[j]) for j in trip]
        x1 = np.array([[rows[i, 0], rows[i, 1], 1.0] for i in ids])
        x2 = np.array([[rows[i, 2], rows[i, 3], 1.0] for i in ids])
        b = np.empty(3)
        for k in range(3):
            c = np.cross(x2[k], e)
            b[k] = np.dot(np.cross(x2[k], A @ x1[k]), c) / np.dot(c, c)      # :419-422
        try:
            Hm = A - np.outer(e, np.linalg.solve(x1, b))                     # :424-430
        except np.linalg.LinAlgError:  # collinear triplet: M.inverse() is inf/NaN in the reference, no point passes the test
            margins.append([float("inf")] * 4)
            continue
        errs = []
        for j in range(7):
            i = int(sample7[j])
            if i in ids:
                continue
            t = Hm @ np.array([rows[i, 0], rows[i, 1], 1.0])
            with np.errstate(divide="ignore", invalid="ignore"):
                errs.append((rows[i, 2] - t[0] / t[2]) ** 2 + (rows[i, 3] - t[1] / t[2]) ** 2)   # :452-462
        margins.append(sorted(errs))
        if 3 + sum(x < homography_threshold ** 2 for x in errs) >= 5:        # :466-476
            return True, Hm, margins
    return False, None, margins


def pearl_datacost(t, pts, models, thr, lam):
    pts, models = _f(pts), _f(models).reshape(-1, MSIZE[t])
    N, L = pts.shape[0], models.shape[0]
    D = np.empty((N, L + 1))
    lib().pxo_pearl_datacost(t, _p(pts), N, _p(models), L, float(thr), float(lam), _p(D))
    return D


def segment_residual_sums(t, pts, models, labels):
    pts, models = _f(pts), _f(models).reshape(-1, MSIZE[t])
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    L = models.shape[0]
    sums = np.zeros(L)
    counts = np.zeros(L, dtype=np.int64)
    lib().pxo_segment_residual_sums(t, _p(pts), pts.shape[0], _p(models), L, _p(lab), _p(sums), _p(counts))
    return sums, counts


def lo_unary_terms(t, pts, model, thr, lam):
    pts, model = _f(pts), _f(model)
    N = pts.shape[0]
    d, e0, e1 = np.empty(N), np.empty(N), np.empty(N)
    lib().pxo_lo_unary_terms(t, _p(pts), N, _p(model), float(thr), float(lam), _p(d), _p(e0), _p(e1))
    return d, e0, e1


def tukey_weights(t, pts, model, T2):
    pts, model = _f(pts), _f(model)
    N = pts.shape[0]
    idx = np.arange(N, dtype=np.int64)
    w = np.zeros(N)
    lib().pxo_tukey_weights(t, _p(pts), _p(idx), N, _p(model), float(T2), _p(w))
    return w


def fit_h_nonminimal(pts, idx, weights_by_row=None):
    pts = _f(pts)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    w = None if weights_by_row is None else _f(weights_by_row)
    H = np.zeros(9)
    L = lib()
    L.pxo_fit_h_nonminimal.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    ok = L.pxo_fit_h_nonminimal(_p(pts), _p(idx), idx.shape[0], _p(w), _p(H))
    return H, bool(ok)


def fit_nonminimal(t, pts, idx, weights=None):
    """Non-minimal fit of one problem for the VP (weights indexed by point) and 2D-line families."""
    pts = _f(pts)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.zeros(3)
    if t == MODEL_VP:
        w = None if weights is None else _f(weights)
        ok = lib().pxo_fit_vp_nonminimal(_p(pts), _p(idx), idx.size, _p(w) if w is not None else None, _p(out))
    elif t == MODEL_LINE:
        ok = lib().pxo_fit_line_nonminimal(_p(pts), _p(idx), idx.size, _p(out))
    else:
        raise ValueError("fit_nonminimal: VP and LINE only (H has fit_h_nonminimal)")
    return out, bool(ok)


def greedy_ufl(D, label_cost, init_labels=None):
    D = _f(D)
    N, L1 = D.shape
    init = None if init_labels is None else np.ascontiguousarray(init_labels, dtype=np.int32)
    out = np.empty(N, dtype=np.int32)
    e = lib().pxo_greedy_ufl(_p(D), N, L1, float(label_cost), _p(init), _p(out))
    return out, e


def gco_pearl_label(D, lam, label_cost, csr_off=None, csr_idx=None, init_labels=None):
    """The reference's own gco-v3 driven like PEARL::labeling (oracle/_ref/libgco_ref.so)."""
    D = _f(D)
    N, L1 = D.shape
    if csr_off is None:
        csr_off = np.zeros(N + 1, dtype=np.int32)
        csr_idx = np.zeros(1, dtype=np.int32)
    off = np.ascontiguousarray(csr_off, dtype=np.int32)
    idx = np.ascontiguousarray(csr_idx, dtype=np.int32)
    init = None if init_labels is None else np.ascontiguousarray(init_labels, dtype=np.int32)
    out = np.empty(N, dtype=np.int32)
    cyc = C.c_int()
    e = gco().gco_ref_pearl_label(N, L1, _p(D), float(lam), float(label_cost), _p(off), _p(idx), _p(init), _p(out),
                                  C.byref(cyc))
    return out, e, cyc.value


def gco_energy(D, lam, label_cost, csr_off, csr_idx, labels):
    D = _f(D)
    N, L1 = D.shape
    off = np.ascontiguousarray(csr_off, dtype=np.int32)
    idx = np.ascontiguousarray(csr_idx, dtype=np.int32)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    return gco().gco_ref_energy(N, L1, _p(D), float(lam), float(label_cost), _p(off), _p(idx), _p(lab))


def gco_lo_labeling(e0, e1, d, lam, csr_off, csr_idx):
    e0, e1, d = _f(e0), _f(e1), _f(d)
    N = e0.shape[0]
    off = np.ascontiguousarray(csr_off, dtype=np.int32)
    idx = np.ascontiguousarray(csr_idx, dtype=np.int32)
    out = np.empty(N, dtype=np.uint8)
    en = gco().gco_ref_lo_labeling(N, _p(e0), _p(e1), _p(d), float(lam), _p(off), _p(idx), _p(out))
    return out, en
