"""Sequential CPU restatement of the Progressive-X control flow (TEST INFRASTRUCTURE ONLY, see oracle/pxo_oracle.h).

One hypothesis at a time, exactly as the reference runs it -- no blocks, no batching -- on top of the oracle's
operators (pinned against the reference's own function bodies) and the reference's own gco / max-flow build:

    ProgressiveX::run              src/pyprogressivex/include/progressive_x.h:251-489     -> ProgressiveXOracle.run
    isPutativeModelValid           progressive_x.h:565-591                               -> putative_model_valid
    updateCompoundModel            progressive_x.h:597-624                               -> (inside run)
    getPredictedUnseenInliers      progressive_x.h:495-513                               -> predicted_unseen_inliers
    GCRANSAC::run                  graph-cut-ransac/.../GCRANSAC.h:203-628               -> propose
    graphCutLocalOptimization      GCRANSAC.h:781-911                                    -> local_optimization
    labeling                       GCRANSAC.h:914-1022                                   -> lo_labeling (reference BK build)
    iteratedLeastSquaresFitting    GCRANSAC.h:631-759                                    -> irls
    PEARL::run / labeling / parameterEstimation / rejectInstances   PEARL.h:275-555      -> pearl (reference gco build)

tests/test_gpu_sequential_oracle.py runs the GPU driver (blocks of 512 hypotheses per launch, device-resident
chains) and this loop on the same inputs and seeds and compares instance counts, labels and models.

What is shared with the GPU driver BY CONSTRUCTION, because the reference leaves it unspecified or non-reproducible
(DESIGN.md section 8): the random streams (splitmix64, one stream per proposal for the main sampler and one for the
local-optimisation sampler), the neighbourhood graph (handed in by the caller), and "the inlier list of the current
best model" in place of the reference's accidental two-buffer ping-pong. Families: homographies (non-minimal fit:
column-pivoted Householder QR here, 8x8 normal equations on the GPU -- ~1e-9 apart, which is why models are compared with
a tolerance and labels exactly), vanishing points and 2D lines (findVanishingPoints_ / findLines_,
progressivex_python.cpp:306-535; weights of the vanishing-point solver are indexed by point), fundamental matrices
(findTwoViewMotions_, :535-666: every model of a seven-point sample is scored, isValidModel = symmetric-epipolar recount +
DEGENSAC with its nested plane-and-parallax GC-RANSAC, fundamental_estimator.h:268-572) and 6D poses (find6DPoses_, :41-171:
up to four P3P poses per sample). The reference's non-minimal F / pose solvers (PoseLib, OpenCV EPnP) are not on disk: for
those two the fits restate the GPU engine's own algorithms (eight-point + LM on the Sampson error, DLT + LM on the
reprojection error) in numpy, so the comparison checks the control flow around them, not the solvers against the reference."""
from __future__ import annotations

import math
import os

import numpy as np

from . import oracle as O

H, F, PNP, VP, LINE = 0, 1, 2, 3, 4
M64 = (1 << 64) - 1


class Rng:
    """splitmix64 as in progressive-x_b200/csrc/pxb_driver.cu (the reference seeds std::mt19937 from std::random_device)."""

    def __init__(self, seed):
        self.s = (seed & M64) or 0x9E3779B97F4A7C15

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    def uniform(self, mx):  # inclusive, rejection sampled like std::uniform_int_distribution(0, mx)
        rng = mx + 1
        limit = M64 - (M64 % rng)
        while True:
            r = self.next()
            if r < limit:
                return r % rng

    def unique_set(self, n, mx, skip=None):  # gcr/uniform_random_generator.h:76-122
        out = []
        while len(out) < n:
            v = self.uniform(mx)
            if skip is not None and v == skip:
                continue
            if v in out:
                continue
            out.append(v)
        return out


def lo_substream(seed, event, trial):
    """Generator seed of trial `trial` of the `event`-th LO labelling of a proposal (pxb_driver.cu lo_substream: the GPU
    driver draws all inner-RANSAC samples of one LO step at once, one generator per trial)."""
    z = (seed + 0x9E3779B97F4A7C15 * (event * 64 + trial + 1)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


class UniformSampler:  # gcr/samplers/uniform_sampler.h:118-134
    def __init__(self, seed):
        self.rng = Rng(seed)

    def sample(self, pool, m):
        if m > len(pool):
            return None
        return [pool[i] for i in self.rng.unique_set(m, len(pool) - 1)]


class NapsacSampler:  # gcr/samplers/napsac_sampler.h:102-151 (incl. the point-index-as-list-index skip, :139-142)
    def __init__(self, seed, off, idx):
        self.rng, self.off, self.idx = Rng(seed), off, idx

    def sample(self, pool, m):
        if m > len(pool):
            return None
        attempts = 0
        subset = None
        while True:
            attempts += 1
            if not attempts - 1 < 100:
                break
            c = self.rng.unique_set(1, len(pool) - 1)[0]
            nb = self.idx[self.off[c]:self.off[c + 1]]
            if len(nb) < m:
                continue
            if len(nb) == m:
                subset = [int(v) for v in nb[:m]]
                break
            rest = self.rng.unique_set(m - 1, len(nb) - 1, skip=c)
            subset = [c] + [int(nb[j]) for j in rest]
            break
        return subset if attempts < 100 else None


class ProsacSampler:  # gcr/samplers/prosac_sampler.h (reset() state per proposal, progressive_x.h:290-291)
    def __init__(self, seed, m, N, convergence=100000):
        self.rng, self.m, self.N = Rng(seed), m, N
        self.convergence, self.kth, self.subset_size, self.gen_max = convergence, 1, m, m - 1
        self.growth = [0] * N
        T_n = float(self.convergence)
        for i in range(m):
            T_n *= (m - i) / (N - i)
        T_n_prime = 1
        for i in range(N):
            if i + 1 <= m:
                self.growth[i] = T_n_prime
                continue
            Tn_plus1 = (i + 1) * T_n / (i + 1 - m)
            self.growth[i] = T_n_prime + (int(math.ceil(Tn_plus1 - T_n)) & 0xFFFFFFFF)
            T_n = Tn_plus1
            T_n_prime = self.growth[i]

    def _increment(self):
        self.kth += 1
        if self.kth > self.convergence:
            self.gen_max = self.N - 1
        elif self.kth > self.growth[self.subset_size - 1]:
            self.subset_size = min(self.subset_size + 1, self.N)
            self.gen_max = self.subset_size - 2

    def set_sample_number(self, k):  # setSampleNumber
        self.kth = k
        if self.kth > self.convergence:
            self.gen_max = self.N - 1
        else:
            while self.kth > self.growth[self.subset_size - 1] and self.subset_size != self.N:
                self.subset_size = min(self.subset_size + 1, self.N)
                self.gen_max = self.subset_size - 2

    def sample(self, pool, m):
        if m != self.m:
            self._increment()
            return None
        if self.kth > self.convergence:
            return self.rng.unique_set(m, self.gen_max)
        out = self.rng.unique_set(m - 1, self.gen_max) + [self.subset_size - 1]
        self._increment()
        return out


class GridLayer:  # gcr/neighborhood/grid_neighborhood_graph.h (out-of-image coordinates clamped to the border cells)
    def __init__(self, rows, sizes, cells):
        idx = np.zeros(len(rows), dtype=np.int64)
        offset = 1
        for d in range(rows.shape[1]):
            f = np.floor(rows[:, d] / (sizes[d] / cells))
            f = np.where(f >= 0, f, 0)
            f = np.minimum(f, cells - 1)
            idx += offset * f.astype(np.int64)
            offset *= cells
        self.cell_of = idx
        self.cells = {}
        for i, c in enumerate(idx):
            self.cells.setdefault(int(c), []).append(i)

    def neighbors(self, i):
        return self.cells[int(self.cell_of[i])]


class ProgressiveNapsacSampler:  # gcr/samplers/progressive_napsac_sampler.h; layers {16, 8, 4, 2}, length 0.5
    def __init__(self, seed, m, N, layers, sampler_length=0.5):
        self.rng, self.layers, self.m, self.N, self.kth = Rng(seed), layers, m, N, 0
        self.one_point = ProsacSampler(seed ^ 0x5851F42D4C957F2D, 1, N, N)
        self.prosac = ProsacSampler(seed ^ 0x14057B7EF767814F, m, N, N)
        self.max_local = int(sampler_length * N)
        self.current_layer, self.hits, self.subset_size_of = [0] * N, [0] * N, [m] * N
        self.growth = [0] * N
        local = m - 1
        T_n = float(self.max_local)
        for i in range(local):
            T_n *= (local - i) / (N - i)
        T_n_prime = 1
        for i in range(N):
            if i + 1 <= local:
                self.growth[i] = T_n_prime
                continue
            Tn_plus1 = (i + 1) * T_n / (i + 1 - local)
            self.growth[i] = T_n_prime + int(math.ceil(Tn_plus1 - T_n))
            T_n = Tn_plus1
            T_n_prime = self.growth[i] & 0xFFFFFFFF

    def sample(self, pool, m):
        self.kth += 1
        if m != self.m or m > len(pool):
            return None
        if self.kth > self.max_local:
            self.prosac.set_sample_number(self.kth)
            return self.prosac.sample(pool, m)
        c = self.one_point.sample(pool, 1)
        if c is None:
            return None
        centre = c[0]
        self.hits[centre] += 1
        h, ss = self.hits[centre], self.subset_size_of[centre]
        while h > self.growth[ss - 1] and ss < self.N:
            ss = min(ss + 1, self.N)
        self.subset_size_of[centre] = ss
        last = False
        while True:
            if self.current_layer[centre] >= len(self.layers):
                last = True
                break
            if len(self.layers[self.current_layer[centre]].neighbors(centre)) < ss:
                self.current_layer[centre] += 1
                continue
            break
        if last:
            self.prosac.set_sample_number(self.kth)
            out = self.prosac.sample(pool, m)
            if out is None:
                return None
            out[m - 1] = centre
            return out
        nb = self.layers[self.current_layer[centre]].neighbors(centre)
        picks = self.rng.unique_set(m - 2, ss - 2, skip=centre)
        out = [nb[j] for j in picks] + [nb[ss - 1], centre]
        for v in out[:m - 2]:
            self.hits[v] += 1
        self.hits[out[m - 2]] += 1
        return out


# ---- fundamental matrices ------------------------------------------------------------------------------------------
def _skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def _factorized_F(U, V, sigma):  # FactorizedFundamentalMatrix::F (relative_pose/jacobian_impl.h:484-486)
    return np.outer(U[:, 0], V[:, 0]) + sigma * np.outer(U[:, 1], V[:, 1])


def _lm_f_cost(Fm, p1, p2, w, sq_thr=1.0):  # FundamentalJacobianAccumulator::residual (jacobian_impl.h:504-532)
    Fx = p1 @ Fm.T
    Ft = p2 @ Fm
    C = np.einsum("ni,ni->n", p2, Fx)
    nJ = Fx[:, 0] ** 2 + Fx[:, 1] ** 2 + Ft[:, 0] ** 2 + Ft[:, 1] ** 2
    with np.errstate(all="ignore"):
        r2 = C * C / nJ
    return float(np.sum(w * np.fmin(r2, sq_thr)))


def _lm_f_system(Fm, U, V, p1, p2, w, sq_thr=1.0):  # ...::accumulate (jacobian_impl.h:534-606): J^T J, J^T r
    n = len(p1)
    E = np.eye(3)
    dP = [(_skew(E[k]) @ Fm) for k in range(3)] + [(-Fm @ _skew(E[k])) for k in range(3)] + [np.outer(U[:, 1], V[:, 1])]
    Fx = p1 @ Fm.T
    Ft = p2 @ Fm
    C = np.einsum("ni,ni->n", p2, Fx)
    nJ = np.sqrt(Ft[:, 0] ** 2 + Ft[:, 1] ** 2 + Fx[:, 0] ** 2 + Fx[:, 1] ** 2)
    with np.errstate(all="ignore"):
        inv = 1.0 / nJ
    res = C * inv
    wgt = np.where(res * res < sq_thr, 1.0, 0.0) / n * w
    sC = C * inv * inv
    G = np.empty((n, 3, 3))
    for r in range(3):
        for c in range(3):
            G[:, r, c] = (p2[:, r] * p1[:, c] - sC * ((Ft[:, c] * p2[:, r] if c < 2 else 0.0) + (Fx[:, r] * p1[:, c] if r < 2 else 0.0))) * inv
    J = np.stack([np.einsum("nrc,rc->n", G, d) for d in dP], 1)
    keep = wgt != 0.0
    J, wgt, res = J[keep], wgt[keep], res[keep]
    return (J * wgt[:, None]).T @ J, J.T @ (wgt * res)


def fit_f_nonminimal(pts, idx, weights_by_row=None):
    """FundamentalMatrixEstimator::estimateModelNonminimal (fundamental_estimator.h:574-618) with the reference's
    FundamentalMatrixBundleAdjustmentSolver (solver_fundamental_matrix_bundle_adjustment.h:114-178) restated with numpy:
    Hartley normalisation, eight-point estimate on the normalised points, then PoseLib's refine_fundamental = lm_F_impl
    (relative_pose/bundle.cpp:253-340,454-500; 7-parameter factorised F, truncated loss with loss_scale 1, LLT steps, at most
    25 iterations, tolerances 1e-8), denormalisation, unit norm, f33 >= 0. The same restatement as k_fit_f
    (progressive-x_b200/csrc/pxb_fit_fp.cu); sums are numpy's, so results agree to ~1e-9, not bit for bit. As there, the
    eight-point null vector is the smallest eigenvector of A^T A (the reference: last column of a FullPivHouseholderQR) and
    n == 7 (seven-point initialisation) is not covered."""
    q = pts[np.asarray(idx, dtype=np.int64)]
    n = len(q)
    if n < 8:
        return None, False
    have_w = weights_by_row is not None
    w = np.ones(n) if not have_w else np.asarray(weights_by_row, dtype=np.float64)[:n]
    m1, m2 = q[:, :2].mean(0), q[:, 2:].mean(0)
    r1 = math.sqrt(2.0) / np.mean(np.sqrt(((m1 - q[:, :2]) ** 2).sum(1)))
    r2 = math.sqrt(2.0) / np.mean(np.sqrt(((m2 - q[:, 2:]) ** 2).sum(1)))
    a = (q[:, :2] - m1) * r1
    b = (q[:, 2:] - m2) * r2
    rows = np.column_stack([b[:, 0] * a[:, 0], b[:, 0] * a[:, 1], b[:, 0], b[:, 1] * a[:, 0], b[:, 1] * a[:, 1], b[:, 1],
                            a[:, 0], a[:, 1], np.ones(n)]) * w[:, None]
    evals, evecs = np.linalg.eigh(rows.T @ rows)
    U, sv, Vt = np.linalg.svd(evecs[:, 0].reshape(3, 3))
    V = Vt.T
    sigma = sv[1] / sv[0] if sv[0] > 0 else 0.0
    p1 = np.column_stack([a, np.ones(n)])
    p2 = np.column_stack([b, np.ones(n)])
    Fc = _factorized_F(U, V, sigma)
    cost = _lm_f_cost(Fc, p1, p2, w)
    lam, recompute = 1e-3, True
    JtJ = Jtr = None
    for _ in range(25):
        if recompute:
            JtJ, Jtr = _lm_f_system(Fc, U, V, p1, p2, w)
            if np.linalg.norm(Jtr) < 1e-8:
                break
        try:
            L = np.linalg.cholesky(JtJ + lam * np.eye(7))
        except np.linalg.LinAlgError:
            break
        sol = -np.linalg.solve(L.T, np.linalg.solve(L, Jtr))
        if np.linalg.norm(sol) < 1e-8:
            break
        with np.errstate(all="ignore"):
            new = []
            for M, wv in ((U, sol[:3]), (V, sol[3:6])):
                theta = np.linalg.norm(wv)
                sw = _skew(wv / theta)
                new.append(M + (math.sin(theta) * sw + (1 - math.cos(theta)) * sw @ sw) @ M)
        Un, Vn, sn = new[0], new[1], sigma + sol[6]
        Fn = _factorized_F(Un, Vn, sn)
        cost_new = _lm_f_cost(Fn, p1, p2, w)
        if cost_new < cost:
            U, V, sigma, Fc, cost = Un, Vn, sn, Fn, cost_new
            lam /= 10
            recompute = True
        else:
            lam *= 10
            recompute = False
    T1 = np.array([[r1, 0, -r1 * m1[0]], [0, r1, -r1 * m1[1]], [0, 0, 1.0]])
    T2m = np.array([[r2, 0, -r2 * m2[0]], [0, r2, -r2 * m2[1]], [0, 0, 1.0]])
    Fm = T2m.T @ Fc @ T1
    nrm = np.linalg.norm(Fm)
    if not (nrm > 0.0) or not np.isfinite(nrm):
        return None, False
    sgn = -1.0 if Fm[2, 2] < 0 else 1.0  # fundamental_estimator.h:611-613
    return (sgn * Fm / nrm).reshape(9), True


def sym_epipolar_sq(pts, Fm):
    """squaredSymmetricEpipolarDistance (fundamental_estimator.h:224-252) for all points"""
    e = np.asarray(Fm, dtype=np.float64).reshape(3, 3)
    x1, y1, x2, y2 = pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3]
    rxc = e[0, 0] * x2 + e[1, 0] * y2 + e[2, 0]
    ryc = e[0, 1] * x2 + e[1, 1] * y2 + e[2, 1]
    rwc = e[0, 2] * x2 + e[1, 2] * y2 + e[2, 2]
    r = x1 * rxc + y1 * ryc + rwc
    rx = e[0, 0] * x1 + e[0, 1] * y1 + e[0, 2]
    ry = e[1, 0] * x1 + e[1, 1] * y1 + e[1, 2]
    a, b = rxc * rxc + ryc * ryc, rx * rx + ry * ry
    with np.errstate(divide="ignore", invalid="ignore"):
        return r * r * (a + b) / (a * b)



# ---- 6D poses -------------------------------------------------------------------------------------------------------
def _pnp_cost(R, t, q, sq_thr=1.0):  # CameraJacobianAccumulator::residual (jacobian_impl.h:26-59), calibrated camera
    Z = q[:, 2:] @ R.T + t
    front = Z[:, 2] >= 0
    z = Z[front, :2] / Z[front, 2:3]
    r = z - q[front, :2]
    return float(np.sum(np.fmin((r * r).sum(1), sq_thr)))


def _pnp_system(R, t, q, sq_thr=1.0):  # ...::accumulate (jacobian_impl.h:63-155): J^T J and J^T r, truncated-loss weights
    X = q[:, 2:]
    Z = X @ R.T + t
    keep = Z[:, 2] >= 0
    X, Z, x = X[keep], Z[keep], q[keep, :2]
    z = Z[:, :2] / Z[:, 2:3]
    r = z - x
    on = (r * r).sum(1) < sq_thr
    X, Z, z, r = X[on], Z[on], z[on], r[on]
    iz = 1.0 / Z[:, 2]
    d0 = (R[0][None, :] - z[:, 0:1] * R[2][None, :]) * iz[:, None]  # rows of dZ = [I | -z] R / Z_z
    d1 = (R[1][None, :] - z[:, 1:2] * R[2][None, :]) * iz[:, None]

    def rot(d):  # columns of -dZ [X]x
        return np.column_stack([X[:, 1] * d[:, 2] - X[:, 2] * d[:, 1], X[:, 2] * d[:, 0] - X[:, 0] * d[:, 2],
                                X[:, 0] * d[:, 1] - X[:, 1] * d[:, 0]])
    J0 = np.column_stack([rot(d0), d0])
    J1 = np.column_stack([rot(d1), d1])
    return J0.T @ J0 + J1.T @ J1, J0.T @ r[:, 0] + J1.T @ r[:, 1]


def fit_pnp_nonminimal(pts, idx):
    """PerspectiveNPointEstimator::estimateModelNonminimal -> PnPBundleAdjustment (solver_pnp_bundle_adjustment.h:108-225)
    restated with numpy, as k_fit_pnp (pxb_fit_fp.cu) does: the initial pose is a DLT on normalised 3D points projected onto
    SO(3) (the reference: cv::solvePnP(EPNP), OpenCV -- not on disk), the refinement is PoseLib's refine_pnp = lm_pnp_impl
    (relative_pose/bundle.cpp:24-100: truncated loss with loss_scale 1, lambda0 = 1e-3, LLT steps, R <- R exp([w]x),
    t <- t + R dt, tolerances 1e-8, at most 25 iterations) over the sample with unit weights."""
    q = pts[np.asarray(idx, dtype=np.int64)]
    n = len(q)
    if n < 6:
        return None, False
    c = q[:, 2:].mean(0)
    md = float(np.mean(np.sqrt(((q[:, 2:] - c) ** 2).sum(1))))
    sc = math.sqrt(3.0) / md if md > 0 else 1.0
    X = (q[:, 2:] - c) * sc
    u, v = q[:, 0], q[:, 1]
    one, zero = np.ones(n), np.zeros((n, 4))
    Xh = np.column_stack([X, one])
    A = np.concatenate([np.column_stack([Xh, zero, -u[:, None] * Xh]), np.column_stack([zero, Xh, -v[:, None] * Xh])])
    evals, evecs = np.linalg.eigh(A.T @ A)
    Pn = evecs[:, 0].reshape(3, 4)
    M = Pn[:, :3]
    sg = -1.0 if np.linalg.det(M) < 0 else 1.0
    U, S, Vt = np.linalg.svd(sg * M)
    scale = float(S.mean())
    if not scale > 0:
        return None, False
    R = U @ np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt)) or 1.0]) @ Vt
    tn = sg * Pn[:, 3] / scale
    t = tn / sc - R @ c
    if not (np.all(np.isfinite(R)) and np.all(np.isfinite(t))):
        return None, False
    cost = _pnp_cost(R, t, q)
    lam, recompute = 1e-3, True
    JtJ = Jtr = None
    for _ in range(25):
        if recompute:
            JtJ, Jtr = _pnp_system(R, t, q)
            if np.linalg.norm(Jtr) < 1e-8:
                break
        try:
            L = np.linalg.cholesky(JtJ + lam * np.eye(6))
        except np.linalg.LinAlgError:
            break
        sol = -np.linalg.solve(L.T, np.linalg.solve(L, Jtr))
        if np.linalg.norm(sol) < 1e-8:
            break
        with np.errstate(all="ignore"):
            theta = np.linalg.norm(sol[:3])
            sw = _skew(sol[:3] / theta)
            Rn = R + R @ (math.sin(theta) * sw + (1 - math.cos(theta)) * sw @ sw)
        tn2 = t + R @ sol[3:]
        cost_new = _pnp_cost(Rn, tn2, q)
        if cost_new < cost:
            R, t, cost = Rn, tn2, cost_new
            lam /= 10
            recompute = True
        else:
            lam *= 10
            recompute = False
    out = np.column_stack([R, t]).reshape(12)
    return (out, True) if np.all(np.isfinite(out)) else (None, False)


class Score:
    __slots__ = ("inliers", "value")

    def __init__(self, inliers=0, value=0.0):
        self.inliers, self.value = inliers, value


class ProgressiveXOracle:
    def __init__(self, pts, *, threshold, confidence, lam, max_tanimoto, max_iters, min_inliers, max_models, napsac, exponent,
                 seed, graph, family=H, point_weights=None, prosac=False, pnapsac_sizes=None):
        self.pts = np.ascontiguousarray(pts, dtype=np.float64)
        self.N = self.pts.shape[0]
        self.t = family
        self.m = {H: 4, F: 7, PNP: 3}.get(family, 2)           # Estimator::sampleSize()
        self.nonminimal_size = {H: 4, F: 7, PNP: 4}.get(family, 2)  # Estimator::nonMinimalSampleSize()
        # FundamentalMatrixEstimator(minimum_inlier_ratio_in_validity_check = 0.5, use_degensac = true); the nested
        # estimator of DEGENSAC uses the plane-and-parallax solver over a fixed homography, ratio 0 and no DEGENSAC
        self.sym_ratio, self.use_degensac, self.pp_H = 0.5, family == F, None
        self.degensac_stats = [0, 0]                           # H-degenerate samples seen / models replaced
        self.point_weights = None if point_weights is None else np.ascontiguousarray(point_weights, dtype=np.float64)
        self.thr, self.conf, self.lam, self.max_tanimoto = threshold, confidence, lam, max_tanimoto
        self.max_iters, self.min_inliers = max_iters, min_inliers
        self.max_models = max_models if max_models > 0 else 1 << 62
        self.napsac, self.exponent, self.seed = napsac, int(exponent), seed
        self.prosac = prosac
        self.grid_layers = None
        if pnapsac_sizes is not None:
            self.grid_layers = [GridLayer(self.pts, pnapsac_sizes, c) for c in (16, 8, 4, 2)]
        self.off, self.idx = graph if graph is not None else (np.zeros(self.N + 1, np.int32), np.zeros(0, np.int32))
        # gcransac::utils::Settings as overridden by progressive_x.h:64-71
        self.min_iteration_number = 20
        self.min_iteration_number_before_lo = 20
        self.max_local_optimization_number = 50
        self.max_graph_cut_number = 10
        self.max_least_squares_iterations = 10
        self.max_unsuccessful_model_generations = 100
        self.models, self.prefs = [], []
        self.compound = np.zeros(self.N)
        self.labeling = np.zeros(self.N, dtype=np.int64)
        self.pearl_outliers = 0

    # ---- operators: every N-point loop is an oracle call -----------------------------------------------------------
    def _score(self, model, T2, best_inliers):
        cp = self.compound if self.models else None
        cnt, val, shr = O.score_batch(self.t, self.pts, model, T2, cp)
        cnt, val, shr = int(cnt[0]), float(val[0]), float(shr[0])
        if cnt + 1 < best_inliers:  # scoring_function_with_compound_model.h:105-106
            return Score()
        value = val - (math.pow(shr, self.exponent) if self.models else 0.0)  # :110-121
        return Score(cnt, value)

    def _inliers_of(self, model, T2):
        r2, _ = O.residual_matrix(self.t, self.pts, model, T2, want_mask=False)
        return [int(i) for i in np.flatnonzero(r2[0] < T2)]

    def _fit(self, idx, weights=None):
        """Estimator::estimateModelNonminimal. H reads weights_[row of the gathered sample]; the vanishing-point solver
        reads weights_[point index] (solver_vanishing_point_two_lines.h:203); the line solver never reads them."""
        if self.t == H:
            return O.fit_h_nonminimal(self.pts, idx, weights)
        if self.t == F:
            return fit_f_nonminimal(self.pts, idx, weights)
        if self.t == PNP:  # PerspectiveNPointEstimator::isWeightingApplicable() is false: weights never reach the solver
            return fit_pnp_nonminimal(self.pts, idx)
        return O.fit_nonminimal(self.t, self.pts, idx, weights if self.t == VP else None)

    # ---- FundamentalMatrixEstimator::isValidModel (fundamental_estimator.h:268-334) + applyDegensac (:341-572) --------
    def _model_is_valid(self, model, solver_valid, sample, nested_seed):
        """returns (valid, model): the model may have been replaced by DEGENSAC"""
        if self.t != F or not solver_valid:
            return bool(solver_valid), model
        tt = 3.0 / 2.0 * self.thr
        r2, _ = O.residual_matrix(self.t, self.pts, model, tt * tt, want_mask=False)
        sampson_inliers = np.flatnonzero(r2[0] < tt * tt)
        sym_count = int(np.sum(sym_epipolar_sq(self.pts[sampson_inliers], model) < tt * tt))
        minimum = max(7, int(len(sampson_inliers) * self.sym_ratio))  # :303-304
        if sym_count < minimum:
            return False, model
        if not (self.use_degensac and self.pp_H is None):
            return True, model
        degenerate, Hm, _ = O.h_degenerate_sample(self.pts, sample, model)
        if not degenerate:
            return True, model
        self.degensac_stats[0] += 1
        Hflat = Hm.reshape(9)
        f_inliers = [int(i) for i in sampson_inliers]
        q = self.pts[f_inliers]
        t1 = Hflat[0] * q[:, 0] + Hflat[1] * q[:, 1] + Hflat[2]
        t2 = Hflat[3] * q[:, 0] + Hflat[4] * q[:, 1] + Hflat[5]
        t3 = Hflat[6] * q[:, 0] + Hflat[7] * q[:, 1] + Hflat[8]
        with np.errstate(divide="ignore", invalid="ignore"):
            err = (q[:, 2] - t1 / t3) ** 2 + (q[:, 3] - t2 / t3) ** 2
        h_inliers = [f_inliers[k] for k in np.flatnonzero(err < 4.0)]  # homography_threshold_ = 2 px (:93)
        if len(h_inliers) < 4:
            return False, model
        Hfit, ok = O.fit_h_nonminimal(self.pts, h_inliers)
        if not ok:
            return False, model
        nested = ProgressiveXOracle(self.pts, threshold=tt, confidence=0.99, lam=0.0, max_tanimoto=self.max_tanimoto,
                                    max_iters=5000, min_inliers=self.min_inliers, max_models=1, napsac=False,
                                    exponent=self.exponent, seed=0, graph=None, family=F)
        nested.pp_H, nested.m, nested.sym_ratio, nested.use_degensac = Hfit, 2, 0.0, False
        nested.max_local_optimization_number = 10  # gcr/settings.h:72 (progressive_x.h's 50 is not applied here)
        model2 = nested.propose(nested_seed & M64)
        if model2 is not None and len(nested.proposal_inliers) > len(f_inliers):
            self.degensac_stats[1] += 1
            return True, model2
        return True, model

    def _lo_labeling(self, model):  # GCRANSAC.h:914-1022
        d, e0, e1 = O.lo_unary_terms(self.t, self.pts, model, self.thr, self.lam)
        if not (self.lam > 0) or self.idx.size == 0:
            return [int(i) for i in np.flatnonzero(e1 - e0 < 0)]
        seg, _ = O.gco_lo_labeling(e0, e1, d, self.lam, self.off, self.idx)
        return [int(i) for i in np.flatnonzero(seg)]

    def _iteration_number_for(self, inliers, log_probability):  # GCRANSAC.h:158-173
        q = math.pow(inliers / self.N, self.m)
        log2 = math.log(1 - q) if q < 1.0 else -math.inf
        if abs(log2) < np.finfo(np.float64).eps:
            return 1 << 62
        return int(log_probability / log2) + 1

    # ---- GCRANSAC.h:781-911 -----------------------------------------------------------------------------------------
    def local_optimization(self, lo_seed, best_model, best_score, T2):
        inlier_limit = 7 * self.m
        max_score, lo_model = Score(best_score.inliers, best_score.value), best_model
        self.lo_number += 1
        while True:
            self.graph_cut_number += 1
            if not self.graph_cut_number < self.max_graph_cut_number:
                break
            updated = False
            inliers = self._lo_labeling(lo_model)
            event = self.lo_events
            self.lo_events += 1
            sample_size = min(inlier_limit, len(inliers))
            if sample_size < len(inliers):
                sets = [UniformSampler(lo_substream(lo_seed, event, t)).sample(inliers, sample_size)
                        for t in range(self.max_local_optimization_number)]
            elif self.m < len(inliers):
                sets = [inliers]
            else:
                break
            for st in sets:
                model, ok = self._fit(st)
                if not ok:
                    continue
                sc = self._score(model, T2, max_score.inliers)
                if max_score.value < sc.value:
                    updated, max_score, lo_model = True, sc, model
            if not updated:
                break
        if best_score.value < max_score.value:
            return lo_model, max_score
        return best_model, best_score

    # ---- GCRANSAC.h:631-759 -----------------------------------------------------------------------------------------
    def irls(self, inliers, model, T2):
        if len(inliers) <= self.m:
            return inliers, model, False
        iterations = 0
        while True:
            iterations += 1
            if not iterations < self.max_least_squares_iterations:
                break
            weights = O.tukey_weights(self.t, self.pts, model, T2)
            w_point = np.zeros(self.N)
            w_point[inliers] = weights[inliers]
            w_row = w_point[:len(inliers)].copy()  # the H solver reads weights_[row] (reference quirk)
            fitted, ok = self._fit(inliers, w_point if self.t == VP else w_row)
            if not ok:
                break
            sc = self._score(fitted, T2, 0)
            if sc.inliers < self.m or sc.inliers <= len(inliers):
                break
            model = fitted
            inliers = self._inliers_of(model, T2)
        return inliers, model, iterations > 1

    # ---- GCRANSAC.h:203-628 -----------------------------------------------------------------------------------------
    def propose(self, round_seed):
        self.iteration_number = self.graph_cut_number = self.lo_number = 0
        self.proposal_inliers = []
        log_probability = math.log(1.0 - self.conf)
        max_iteration = self._iteration_number_for(1, log_probability)
        tt = 3.0 / 2.0 * self.thr
        T2 = tt * tt
        if self.napsac and self.idx.size:
            main = NapsacSampler((round_seed * 2 + 1) & M64, self.off, self.idx)
        elif self.grid_layers is not None and self.N > self.m:
            main = ProgressiveNapsacSampler((round_seed * 2 + 1) & M64, self.m, self.N, self.grid_layers)
        elif self.prosac and self.N > self.m:
            main = ProsacSampler((round_seed * 2 + 1) & M64, self.m, self.N)
        else:
            main = UniformSampler((round_seed * 2 + 1) & M64)
        lo_seed = (round_seed * 2 + 2) & M64
        self.lo_events = 0
        pool = list(range(self.N))
        best_score, best_model = Score(), None
        while self.min_iteration_number > self.iteration_number or self.iteration_number < min(max_iteration, self.max_iters):
            do_lo = False
            self.iteration_number += 1
            unsuccessful, found = -1, None
            while True:  # :296-339 select a sample that yields at least one model (<= 100 attempts)
                unsuccessful += 1
                if not unsuccessful < self.max_unsuccessful_model_generations:
                    break
                sample = main.sample(pool, self.m)
                if sample is None:
                    continue
                if self.pp_H is not None:  # DEGENSAC's nested estimator: FundamentalMatrixPlaneParallaxSolver
                    pm, pn = O.solve_plane_parallax(self.pts, np.asarray([sample], dtype=np.int64), self.pp_H)
                    if pn[0] > 0:
                        found = ([pm[0].copy()], 1, sample)
                        break
                    continue
                models, n, sv, mv = O.solve_minimal(self.t, self.pts, np.asarray([sample], dtype=np.int64))
                if not sv[0]:
                    continue
                if n[0] > 0:
                    found = ([models[0, j].copy() for j in range(int(n[0]))], int(mv[0]), sample)
                    break
            self.iteration_number += unsuccessful
            if found is not None:
                candidates, solver_valid, sample = found
                for j, model in enumerate(candidates):  # :373-470, every model the sample produced
                    sc = self._score(model, T2, best_score.inliers)
                    if not best_score.value < sc.value:
                        continue
                    nested_seed = round_seed * 7919 + self.iteration_number * 31 + j
                    valid, replaced = self._model_is_valid(model, solver_valid, sample, nested_seed)
                    if not valid:
                        continue
                    if replaced is not model:  # :450-457 re-score the model DEGENSAC put in its place
                        model = replaced
                        sc = self._score(model, T2, best_score.inliers)
                    best_model, best_score = model, sc
                    do_lo = self.iteration_number > self.min_iteration_number_before_lo and best_score.inliers > self.m
                    max_iteration = self._iteration_number_for(best_score.inliers, log_probability)
            if do_lo:  # :482-503
                self.lo_number += 1
                best_model, best_score = self.local_optimization(lo_seed, best_model, best_score, T2)
                max_iteration = self._iteration_number_for(best_score.inliers, log_probability)
        if best_score.inliers <= self.m:
            return None
        if self.lo_number == 0:  # :531-544
            self.lo_number += 1
            best_model, best_score = self.local_optimization(lo_seed, best_model, best_score, T2)
        best_inliers = self._inliers_of(best_model, T2)
        best_score.inliers = len(best_inliers)
        refit_applied = False
        inl, model, success = self.irls(list(best_inliers), best_model, T2)  # :561-590
        if success:
            sc = self._score(model, T2, 0)
            if best_score.value < sc.value:
                refit_applied = True
                best_model = model
                best_inliers = self._inliers_of(best_model, T2)
        if not refit_applied:  # :592-618
            fitted, ok = self._fit(best_inliers)
            if ok:
                sc = self._score(fitted, T2, 0)
                if best_score.value < sc.value:
                    best_model = fitted
                    best_inliers = self._inliers_of(best_model, T2)
        self.proposal_inliers = best_inliers
        return best_model

    # ---- progressive_x.h:565-591 ------------------------------------------------------------------------------------
    def putative_model_valid(self, model):
        if len(self.proposal_inliers) < max(self.m, self.min_inliers):
            return False, None
        T = 9.0 / 4.0 * self.thr * self.thr
        pref = O.preference_vector(self.t, self.pts, model, T)
        tanimoto = O.tanimoto(pref, self.compound)
        if self.max_tanimoto < tanimoto:  # NaN compares false -> accepted
            return False, pref
        return True, pref

    # ---- PEARL.h:405-472 / 476-555 / 319-401 / 275-315 --------------------------------------------------------------
    def pearl(self):
        iteration_number, energy, previous_energy = 0, np.finfo(np.float64).max, -1.0
        model_rejected, convergence, have_labels = False, False, False
        labels = np.zeros(self.N, dtype=np.int32)
        label_cost = float(self.min_inliers)
        smooth = self.lam > 0.0 and self.idx.size > 0
        while not convergence:
            iteration_number += 1
            if not iteration_number - 1 < 100:
                break
            init_with_previous = iteration_number > 1 and not model_rejected
            L = len(self.models)
            if L == 0:
                break
            flat = np.stack(self.models)
            D = O.pearl_datacost(self.t, self.pts, flat, self.thr, self.lam)
            init = labels.copy() if (init_with_previous and have_labels) else None
            labels, energy, _ = O.gco_pearl_label(D, self.lam, label_cost, self.off if smooth else None,
                                                  self.idx if smooth else None, init)
            labels = np.asarray(labels, dtype=np.int32)
            have_labels = True
            changed, model_rejected = False, False
            per_instance = [[int(i) for i in np.flatnonzero(labels == l)] for l in range(L)]
            outliers = int(np.sum(labels >= L))
            before, _ = O.segment_residual_sums(self.t, self.pts, flat, labels)
            cand = flat.copy()
            fitted_ok = [False] * L
            for l in range(L):
                if len(per_instance[l]) >= self.nonminimal_size:  # nonMinimalSampleSize() (:363-365)
                    model, ok = self._fit(per_instance[l], self.point_weights if self.t == VP else None)  # :373-380
                    if ok:
                        cand[l], fitted_ok[l] = model, True
            after, _ = O.segment_residual_sums(self.t, self.pts, cand, labels)
            for l in range(L):
                if fitted_ok[l] and after[l] < before[l]:  # :393-399
                    self.models[l] = cand[l].copy()
                    changed = True
            for l in range(L - 1, -1, -1):  # rejectInstances, back to front
                if len(per_instance[l]) < self.min_inliers:
                    outliers += len(per_instance[l])
                    del self.models[l]
                    del self.prefs[l]
                    del per_instance[l]
                    model_rejected = True
            self.pearl_outliers = outliers
            if not model_rejected and not changed and abs(energy - previous_energy) < 1e-5 and iteration_number > 1:
                convergence = True
            previous_energy = energy
        self.labeling = labels.astype(np.int64)

    def predicted_unseen_inliers(self, iterations, compound_inliers):  # progressive_x.h:495-513
        unseen = self.N - compound_inliers
        ratio = math.pow(1.0 - math.pow(1.0 - self.conf, 1.0 / iterations), 1.0 / self.m)
        return int(round(unseen * ratio))

    # ---- progressive_x.h:251-489 ------------------------------------------------------------------------------------
    def run(self):
        total_iterations, unaccepted = 0, 0
        first_model_stat = 0  # statistics.inliers_of_each_model.size(): grows whenever an instance is added as the only one
        for it in range(10):  # :272 hard cap
            model = self.propose((self.seed * 1000003 + it) & M64)
            if os.environ.get("PXO_LOG"):  # same fields as the driver's do_logging line
                print(f"[pxo] proposal {it + 1}: {'found' if model is not None else 'none'}, {len(self.proposal_inliers)} inliers, "
                      f"{self.iteration_number} iterations, {self.lo_number} LO runs, {self.graph_cut_number} graph cuts, "
                      f"DEGENSAC {self.degensac_stats[1]}/{self.degensac_stats[0]}")
            if model is None:
                continue
            total_iterations += self.iteration_number
            valid, pref = self.putative_model_valid(model)
            if not valid:  # :334-346 (the counter is never reset)
                unaccepted += 1
                if unaccepted == 10:
                    break
                continue
            self.models.append(model.copy())
            self.prefs.append(pref)
            if len(self.models) == 1:  # :375-385
                self.labeling[:] = 1
                self.labeling[self.proposal_inliers] = 0
                first_model_stat += 1
            else:
                self.pearl()
            if self.models:  # updateCompoundModel: max over the stored (stale) preference vectors
                self.compound = O.compound_max(np.stack(self.prefs))
            # :447-457 branches on models.size() == 1 AFTER the optimisation (PEARL may have pruned the compound set back to
            # one instance) and passes inliers_of_each_model.size() -- a model count -- as the inlier number (:451 quirk)
            if len(self.models) == 1:
                unseen = self.predicted_unseen_inliers(total_iterations, first_model_stat)
            else:
                unseen = self.predicted_unseen_inliers(total_iterations, self.N - self.pearl_outliers)
            if unseen < self.min_inliers:
                break
            if len(self.models) >= self.max_models:
                break
        models = np.stack(self.models) if self.models else np.zeros((0, O.MSIZE[self.t]))
        return models, self.labeling.copy()


def find_homographies(corrs, threshold, conf, spatial_coherence_weight, maximum_tanimoto_similarity, max_iters,
                      minimum_point_number, maximum_model_number, sampler_id, scoring_exponent, seed, graph=None,
                      image_sizes=None):
    """findHomographies_ (src/pyprogressivex/src/progressivex_python.cpp:173-304) on the sequential loop above."""
    px = ProgressiveXOracle(corrs, threshold=threshold, confidence=conf, lam=spatial_coherence_weight,
                            max_tanimoto=maximum_tanimoto_similarity, max_iters=max_iters, min_inliers=minimum_point_number,
                            max_models=maximum_model_number, napsac=(sampler_id == 3), exponent=scoring_exponent, seed=seed,
                            graph=graph, prosac=(sampler_id == 1),
                            pnapsac_sizes=image_sizes if sampler_id == 2 else None)
    return px.run()


def find_points_family(family, rows, weights, threshold, conf, spatial_coherence_weight, maximum_tanimoto_similarity, max_iters,
                       minimum_point_number, maximum_model_number, sampler_id, scoring_exponent, seed, graph=None):
    """findVanishingPoints_ / findLines_ (progressivex_python.cpp:306-535) on the sequential loop: family = VP or LINE."""
    napsac = family == LINE and sampler_id == 2
    px = ProgressiveXOracle(rows, threshold=threshold, confidence=conf, lam=spatial_coherence_weight,
                            max_tanimoto=maximum_tanimoto_similarity, max_iters=max_iters, min_inliers=minimum_point_number,
                            max_models=maximum_model_number, napsac=napsac, exponent=scoring_exponent, seed=seed, graph=graph,
                            family=family, point_weights=weights if family == VP else None, prosac=(sampler_id == 1))
    return px.run()


def find_two_view_motions(corrs, threshold, conf, spatial_coherence_weight, maximum_tanimoto_similarity, max_iters,
                          minimum_point_number, maximum_model_number, sampler_id, scoring_exponent, seed, graph=None,
                          image_sizes=None):
    """findTwoViewMotions_ (src/pyprogressivex/src/progressivex_python.cpp:535-666) on the sequential loop: seven-point
    solver with the oriented-epipolar filter, symmetric-epipolar validity, DEGENSAC, eight-point + LM non-minimal fit.
    The reference takes `scoring_exponent` (:554) and never applies it -- there is no setScoringExponent call in this entry,
    unlike :276 / :399 / :513 -- so the scorer keeps its default exponent 2."""
    px = ProgressiveXOracle(corrs, threshold=threshold, confidence=conf, lam=spatial_coherence_weight,
                            max_tanimoto=maximum_tanimoto_similarity, max_iters=max_iters, min_inliers=minimum_point_number,
                            max_models=maximum_model_number, napsac=(sampler_id == 3), exponent=2, seed=seed,
                            graph=graph, family=F, prosac=(sampler_id == 1),
                            pnapsac_sizes=image_sizes if sampler_id == 2 else None)
    return px.run()


def find_6d_poses(image_points, world_points, K, threshold, conf, spatial_coherence_weight, maximum_tanimoto_similarity,
                  max_iters, minimum_point_number, maximum_model_number, seed, graph=None):
    """find6DPoses_ (src/pyprogressivex/src/progressivex_python.cpp:41-171) on the sequential loop: image points are
    normalised by K^-1, the threshold by the mean focal length (:96-98); P3P minimal solver (up to four poses per
    sample), uniform sampler, default exponent. `graph` is the neighbourhood graph of the RAW [u v X Y Z] rows."""
    Km = np.asarray(K, dtype=np.float64).reshape(3, 3)
    m = Km.reshape(9)
    c00, c01, c02 = m[4] * m[8] - m[5] * m[7], m[5] * m[6] - m[3] * m[8], m[3] * m[7] - m[4] * m[6]
    idet = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02)  # Eigen Matrix3d::inverse(): cofactors / determinant
    Kinv = np.array([c00, m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
                     c01, m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
                     c02, m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]]) * idet
    img = np.asarray(image_points, dtype=np.float64)
    rows = np.empty((len(img), 5))
    rows[:, 0] = Kinv[0] * img[:, 0] + Kinv[1] * img[:, 1] + Kinv[2] * 1
    rows[:, 1] = Kinv[3] * img[:, 0] + Kinv[4] * img[:, 1] + Kinv[5] * 1
    rows[:, 2:] = np.asarray(world_points, dtype=np.float64)
    f = 0.5 * (Km[0, 0] + Km[1, 1])
    px = ProgressiveXOracle(rows, threshold=threshold / f, confidence=conf, lam=spatial_coherence_weight,
                            max_tanimoto=maximum_tanimoto_similarity, max_iters=max_iters, min_inliers=minimum_point_number,
                            max_models=maximum_model_number, napsac=False, exponent=2, seed=seed, graph=graph, family=PNP)
    return px.run()
