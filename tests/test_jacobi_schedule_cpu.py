"""CPU restatement of the round-robin Jacobi order used by k_fit_f / k_fit_pnp (csrc/pxb_fit_fp.cu, jacobi_eig_warp):
the tournament schedule visits every index pair exactly once per sweep, the rotations of a round touch disjoint
rows / columns (so applying them from one snapshot of the matrix equals applying them one after the other), and the
iteration converges to the eigen-decomposition numpy returns. No GPU needed: this pins the ALGORITHM the kernel runs."""
import numpy as np
import pytest


def schedule(nn):
    """pairs of round r, as the kernel's lanes compute them (NP = nn rounded up to even; index NP - 1 may be a bye)"""
    np_ = (nn + 1) // 2 * 2
    half = np_ // 2
    rounds = []
    for r in range(np_ - 1):
        pairs = []
        for lane in range(half):
            a = np_ - 1 if lane == 0 else (r + lane) % (np_ - 1)
            b = r if lane == 0 else (r - lane + (np_ - 1)) % (np_ - 1)
            p, q = min(a, b), max(a, b)
            if q < nn:
                pairs.append((p, q))
        rounds.append(pairs)
    return rounds


@pytest.mark.parametrize("nn", [3, 8, 9, 12])
def test_every_pair_once_and_rounds_disjoint(nn):
    seen = set()
    for pairs in schedule(nn):
        used = [i for pq in pairs for i in pq]
        assert len(used) == len(set(used)), "two rotations of a round share an index"
        for pq in pairs:
            assert pq not in seen
            seen.add(pq)
    assert seen == {(p, q) for p in range(nn) for q in range(p + 1, nn)}


def rotation(app, aqq, apq):
    """the kernel's angle: one sqrt, one division, one reciprocal square root"""
    if apq == 0.0:
        return 1.0, 0.0
    d = aqq - app
    r = np.sqrt(d * d + 4.0 * apq * apq)
    t = (2.0 * apq if d >= 0 else -2.0 * apq) / (abs(d) + r)
    c = 1.0 / np.sqrt(t * t + 1.0)
    return c, t * c


def jacobi_round_robin(A):
    nn = A.shape[0]
    A = A.copy()
    V = np.eye(nn)
    sweeps = 0
    for sweeps in range(60):
        off = np.sum(np.triu(A, 1) ** 2)
        diag = np.sum(np.diag(A) ** 2)
        if not off > 1e-30 * diag:
            break
        for pairs in schedule(nn):
            rots = [(p, q) + rotation(A[p, p], A[q, q], A[p, q]) for p, q in pairs]  # all angles from one snapshot
            for p, q, c, s in rots:  # columns, then rows: the two phases of the kernel
                ap, aq = A[:, p].copy(), A[:, q].copy()
                A[:, p], A[:, q] = c * ap - s * aq, s * ap + c * aq
                vp, vq = V[:, p].copy(), V[:, q].copy()
                V[:, p], V[:, q] = c * vp - s * vq, s * vp + c * vq
            for p, q, c, s in rots:
                ap, aq = A[p, :].copy(), A[q, :].copy()
                A[p, :], A[q, :] = c * ap - s * aq, s * ap + c * aq
    return np.diag(A).copy(), V, sweeps


@pytest.mark.parametrize("nn,seed", [(9, 0), (9, 1), (12, 2), (12, 3)])
def test_converges_to_the_eigen_decomposition(nn, seed):
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(3 * nn, nn)) * rng.uniform(0.01, 100.0, size=nn)  # a design matrix with badly scaled columns
    A = M.T @ M
    w, V, sweeps = jacobi_round_robin(A)
    assert sweeps <= 12
    w_ref, V_ref = np.linalg.eigh(A)
    order = np.argsort(w)
    np.testing.assert_allclose(w[order], w_ref, rtol=1e-10, atol=1e-12 * w_ref[-1])
    np.testing.assert_allclose(V.T @ V, np.eye(nn), atol=1e-12)
    v_min = V[:, order[0]]  # the eigenvector the fits use: the seed of the LM refinement
    assert min(np.linalg.norm(v_min - V_ref[:, 0]), np.linalg.norm(v_min + V_ref[:, 0])) < 1e-8
