"""Seeded synthetic scenes for the BASELINE.json configs. numpy only.

  multi_homography_scene   C2 / C4 / headline grid: planted planes + uniform outliers
  multi_motion_scene       C3: rigid motions -> fundamental matrices
  multi_pose_scene         C5: 2D-3D matches of several rigid objects (T-LESS intrinsics)
  multi_vanishing_point_scene  line segments converging to a few vanishing points + random segments
  multi_line_scene         2D points on a few lines + uniform outliers
  minimal_samples          hypothesis samples drawn half within-structure, half at random
  knn_graph                exact radius graph truncated to the k nearest, as directed CSR lists

Everything is float64 and C-contiguous, the layout the reference's drivers receive
(px/src/progressivex_python.cpp:203: cv::Mat(N, 4, CV_64F, ptr)).
"""
from __future__ import annotations

import numpy as np

TLESS_K = np.array([[1075.65, 0.0, 370.07], [0.0, 1073.90, 278.72], [0.0, 0.0, 1.0]])


def _random_homography(rng, w, h):
    """A mild projective warp of the image rectangle (keeps points finite and inside a sane range)."""
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], dtype=np.float64)
    dst = src + rng.uniform(-0.18, 0.18, size=(4, 2)) * np.array([w, h])
    A = []
    for (x, y), (u, v) in zip(src, dst):
        A.append([-x, -y, -1, 0, 0, 0, u * x, u * y, u])
        A.append([0, 0, 0, -x, -y, -1, v * x, v * y, v])
    _, _, vt = np.linalg.svd(np.asarray(A))
    H = vt[-1].reshape(3, 3)
    return H / H[2, 2]


def multi_homography_scene(N, n_planes=5, outlier_ratio=0.4, noise=0.5, w=1024, h=768, seed=0):
    """Returns (corrs [N,4], gt_labels [N] with -1 = outlier, Hs [n_planes,3,3])."""
    rng = np.random.default_rng(seed)
    n_out = int(round(N * outlier_ratio))
    per = (N - n_out) // n_planes
    corrs = np.empty((N, 4), dtype=np.float64)
    labels = np.full(N, -1, dtype=np.int64)
    Hs = np.stack([_random_homography(rng, w, h) for _ in range(n_planes)])
    pos = 0
    for k in range(n_planes):
        n_k = per if k < n_planes - 1 else (N - n_out) - per * (n_planes - 1)
        # each plane occupies its own image region so that spatial coherence is meaningful
        cx, cy = rng.uniform(0.2, 0.8) * w, rng.uniform(0.2, 0.8) * h
        x1 = np.stack([np.clip(rng.normal(cx, 0.18 * w, n_k), 0, w), np.clip(rng.normal(cy, 0.18 * h, n_k), 0, h)], 1)
        p = np.concatenate([x1, np.ones((n_k, 1))], 1) @ Hs[k].T
        x2 = p[:, :2] / p[:, 2:3] + rng.normal(0, noise, (n_k, 2))
        corrs[pos:pos + n_k, :2] = x1
        corrs[pos:pos + n_k, 2:] = x2
        labels[pos:pos + n_k] = k
        pos += n_k
    corrs[pos:, 0] = rng.uniform(0, w, N - pos)
    corrs[pos:, 1] = rng.uniform(0, h, N - pos)
    corrs[pos:, 2] = rng.uniform(0, w, N - pos)
    corrs[pos:, 3] = rng.uniform(0, h, N - pos)
    perm = rng.permutation(N)
    return np.ascontiguousarray(corrs[perm]), labels[perm], Hs


def _random_rotation(rng, max_angle):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    a = rng.uniform(-max_angle, max_angle)
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx


def multi_motion_scene(N, n_motions=3, ratios=(0.25, 0.25, 0.20), noise=0.3, f=800.0, w=1024, h=768, seed=0):
    """Two-view correspondences of n rigid motions + uniform outliers. Returns (corrs, gt_labels, Fs)."""
    rng = np.random.default_rng(seed)
    Kc = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1.0]])
    Kinv = np.linalg.inv(Kc)
    corrs = np.empty((N, 4))
    labels = np.full(N, -1, dtype=np.int64)
    Fs = []
    pos = 0
    for k in range(n_motions):
        n_k = int(round(N * ratios[k]))
        R = _random_rotation(rng, 0.25)
        t = rng.normal(size=3)
        t = t / np.linalg.norm(t) * rng.uniform(0.3, 0.8)
        X = np.stack([rng.uniform(-2.5, 2.5, n_k), rng.uniform(-2, 2, n_k), rng.uniform(4, 8, n_k)], 1)
        p1 = X @ Kc.T
        X2 = X @ R.T + t
        p2 = X2 @ Kc.T
        x1 = p1[:, :2] / p1[:, 2:3] + rng.normal(0, noise, (n_k, 2))
        x2 = p2[:, :2] / p2[:, 2:3] + rng.normal(0, noise, (n_k, 2))
        corrs[pos:pos + n_k, :2], corrs[pos:pos + n_k, 2:] = x1, x2
        labels[pos:pos + n_k] = k
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        F = Kinv.T @ tx @ R @ Kinv
        Fs.append(F / F[2, 2])
        pos += n_k
    corrs[pos:, 0] = rng.uniform(0, w, N - pos)
    corrs[pos:, 1] = rng.uniform(0, h, N - pos)
    corrs[pos:, 2] = rng.uniform(0, w, N - pos)
    corrs[pos:, 3] = rng.uniform(0, h, N - pos)
    perm = rng.permutation(N)
    return np.ascontiguousarray(corrs[perm]), labels[perm], np.stack(Fs)


def plane_dominated_pair(n_plane, n_off, noise, seed):
    """Two views of a plane plus off-plane points: rows [n, 4], labels (0 = plane, 1 = off-plane), F."""
    rng = np.random.default_rng(seed)
    f, w, h = 800.0, 1024, 768
    Kc = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1.0]])
    a = rng.normal(0, 0.15, 3)
    R, _ = np.linalg.qr(np.eye(3) + np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]]))
    R *= np.sign(np.linalg.det(R))
    t = np.array([0.6, 0.1, 0.05])
    xy = rng.uniform(-2, 2, (n_plane, 2))
    Xp = np.column_stack([xy, 6.0 + 0.2 * xy[:, 0] - 0.1 * xy[:, 1]])
    Xo = np.column_stack([rng.uniform(-2, 2, (n_off, 2)), rng.uniform(3.5, 9.0, n_off)])
    X = np.concatenate([Xp, Xo])
    p1, p2 = X @ Kc.T, (X @ R.T + t) @ Kc.T
    rows = np.column_stack([p1[:, :2] / p1[:, 2:], p2[:, :2] / p2[:, 2:]]) + rng.normal(0, noise, (len(X), 4))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Kinv = np.linalg.inv(Kc)
    return np.ascontiguousarray(rows), np.r_[np.zeros(n_plane, int), np.ones(n_off, int)], Kinv.T @ tx @ R @ Kinv


def multi_pose_scene(N, n_objects=10, inlier_ratio_each=0.06, noise_px=1.0, seed=0, Kc=TLESS_K):
    """2D-3D matches. Returns (image_points [N,2] px, world_points [N,3], K, gt_labels, poses [n,3,4])."""
    rng = np.random.default_rng(seed)
    img = np.empty((N, 2))
    wpts = np.empty((N, 3))
    labels = np.full(N, -1, dtype=np.int64)
    poses = []
    pos = 0
    for k in range(n_objects):
        n_k = int(round(N * inlier_ratio_each))
        R = _random_rotation(rng, np.pi)
        t = np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.2, 0.2), rng.uniform(0.6, 1.2)])
        X = rng.uniform(-0.08, 0.08, (n_k, 3))
        pc = X @ R.T + t
        p = pc @ Kc.T
        img[pos:pos + n_k] = p[:, :2] / p[:, 2:3] + rng.normal(0, noise_px, (n_k, 2))
        wpts[pos:pos + n_k] = X
        labels[pos:pos + n_k] = k
        poses.append(np.concatenate([R, t[:, None]], 1))
        pos += n_k
    img[pos:, 0] = rng.uniform(0, 2 * Kc[0, 2], N - pos)
    img[pos:, 1] = rng.uniform(0, 2 * Kc[1, 2], N - pos)
    wpts[pos:] = rng.uniform(-0.08, 0.08, (N - pos, 3))
    perm = rng.permutation(N)
    return (np.ascontiguousarray(img[perm]), np.ascontiguousarray(wpts[perm]), Kc.copy(), labels[perm],
            np.stack(poses))


def normalize_pnp_points(image_points, world_points, Kc):
    """[u v X Y Z] rows with (u,v) = K^-1 (x,y,1), as px/src/progressivex_python.cpp:64-98 builds them."""
    Kinv = np.linalg.inv(Kc)
    n = image_points.shape[0]
    hom = np.concatenate([image_points, np.ones((n, 1))], 1)
    out = np.empty((n, 5))
    out[:, 0] = hom @ Kinv[0]
    out[:, 1] = hom @ Kinv[1]
    out[:, 2:] = world_points
    return np.ascontiguousarray(out)


def multi_vanishing_point_scene(N, n_vps=3, outlier_ratio=0.3, noise=0.3, w=1024, h=768, seed=0):
    """Line segments [xs ys xe ye] whose supporting lines pass (up to `noise` px at the endpoints) through one of
    n_vps vanishing points placed outside the image. Returns (segments [N,4], gt [N] (-1 outlier), vps [n_vps,3])."""
    rng = np.random.default_rng(seed)
    n_out = int(round(N * outlier_ratio))
    per = (N - n_out) // n_vps
    seg = np.empty((N, 4))
    gt = np.full(N, -1, dtype=np.int64)
    vps = np.empty((n_vps, 3))
    pos = 0
    for k in range(n_vps):
        ang = rng.uniform(0, 2 * np.pi)
        rad = rng.uniform(1.5, 4.0) * w
        v = np.array([w / 2 + rad * np.cos(ang), h / 2 + rad * np.sin(ang)])
        vps[k] = np.array([v[0], v[1], 1.0]) / np.linalg.norm([v[0], v[1], 1.0])
        n_k = per if k < n_vps - 1 else (N - n_out) - per * (n_vps - 1)
        mid = np.stack([rng.uniform(0, w, n_k), rng.uniform(0, h, n_k)], 1)
        d = v[None, :] - mid
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        half = rng.uniform(15, 60, n_k)[:, None]
        seg[pos:pos + n_k, :2] = mid - half * d + rng.normal(0, noise, (n_k, 2))
        seg[pos:pos + n_k, 2:] = mid + half * d + rng.normal(0, noise, (n_k, 2))
        gt[pos:pos + n_k] = k
        pos += n_k
    n_k = N - pos
    mid = np.stack([rng.uniform(0, w, n_k), rng.uniform(0, h, n_k)], 1)
    a = rng.uniform(0, np.pi, n_k)
    d = np.stack([np.cos(a), np.sin(a)], 1)
    half = rng.uniform(15, 60, n_k)[:, None]
    seg[pos:, :2] = mid - half * d
    seg[pos:, 2:] = mid + half * d
    perm = rng.permutation(N)
    return np.ascontiguousarray(seg[perm]), gt[perm], vps


def multi_line_scene(N, n_lines=4, outlier_ratio=0.4, noise=0.5, w=1024, h=768, seed=0):
    """2D points [x y] on n_lines lines + uniform outliers. Returns (points [N,2], gt [N], lines [n_lines,3] with
    unit normals, n . p + c = 0)."""
    rng = np.random.default_rng(seed)
    n_out = int(round(N * outlier_ratio))
    per = (N - n_out) // n_lines
    pts = np.empty((N, 2))
    gt = np.full(N, -1, dtype=np.int64)
    lines = np.empty((n_lines, 3))
    pos = 0
    for k in range(n_lines):
        a = rng.uniform(0, np.pi)
        nrm = np.array([np.cos(a), np.sin(a)])
        p0 = np.array([rng.uniform(0.3, 0.7) * w, rng.uniform(0.3, 0.7) * h])
        lines[k] = np.array([nrm[0], nrm[1], -nrm @ p0])
        n_k = per if k < n_lines - 1 else (N - n_out) - per * (n_lines - 1)
        t = rng.uniform(-0.45 * w, 0.45 * w, n_k)[:, None]
        pts[pos:pos + n_k] = p0 + t * np.array([-nrm[1], nrm[0]]) + rng.normal(0, noise, (n_k, 2))
        gt[pos:pos + n_k] = k
        pos += n_k
    pts[pos:, 0] = rng.uniform(0, w, N - pos)
    pts[pos:, 1] = rng.uniform(0, h, N - pos)
    perm = rng.permutation(N)
    return np.ascontiguousarray(pts[perm]), gt[perm], lines


def minimal_samples(gt_labels, K, m, within_ratio=0.5, seed=0):
    """K samples of m distinct point indices: `within_ratio` of them inside one ground-truth structure."""
    rng = np.random.default_rng(seed + 7919)
    N = gt_labels.shape[0]
    structures = [np.flatnonzero(gt_labels == k) for k in range(int(gt_labels.max()) + 1)]
    structures = [s for s in structures if s.size >= m]
    out = np.empty((K, m), dtype=np.int64)
    for k in range(K):
        if structures and rng.random() < within_ratio:
            pool = structures[rng.integers(len(structures))]
            out[k] = rng.choice(pool, size=m, replace=False)
        else:
            out[k] = rng.choice(N, size=m, replace=False)
    return out


def knn_graph(points, radius, k=5):
    """Directed neighbour lists: the k nearest points within `radius` (self excluded), as CSR (off, idx) int32.

    Mimics the degree the reference's FLANN graph has in practice (~5 matches/point). The graph is
    an *input* to both the oracle and the GPU path."""
    from scipy.spatial import cKDTree

    tree = cKDTree(points)
    dist, idx = tree.query(points, k=k + 1, distance_upper_bound=radius)
    N = points.shape[0]
    off = np.zeros(N + 1, dtype=np.int32)
    rows = []
    for i in range(N):
        nb = [int(j) for j, d in zip(idx[i], dist[i]) if j != i and j < N and np.isfinite(d)]
        rows.append(nb)
        off[i + 1] = off[i] + len(nb)
    flat = np.fromiter((j for r in rows for j in r), dtype=np.int32, count=int(off[-1]))
    return off, flat
