"""pyprogressivex -- drop-in Python surface of danini/progressive-x backed by libpxb200.so (B200, sm_100a).

Same function names, positional order, keyword names and defaults as the reference's pybind11 module
(src/pyprogressivex/src/bindings.cpp:410-491); same return convention `(models, labeling)` with models stacked
as float64 [M*3, 3] (homographies / fundamental matrices) or [M*3, 4] (poses) and labeling as int32 [N]
(bindings.cpp:152-165). `findFundamentalMatrices` is an alias of `findTwoViewMotions` (the reference only has the
latter name). One extra keyword everywhere: `seed` (the reference seeds from std::random_device and cannot be
reproduced; 0 keeps that behaviour), and `device`.

All compute runs through the C ABI in include/pxb200.h; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as _C

import numpy as _np

from . import _native
from ._native import Context, PxbError  # noqa: F401

__all__ = ["findHomographies", "findTwoViewMotions", "findFundamentalMatrices", "find6DPoses", "findVanishingPoints",
           "findLines", "findHomographiesBatch", "distributed", "getStatistics", "getMutableSettings", "setSettings", "Context"]

import threading as _threading

_contexts = {}
_contexts_guard = _threading.Lock()

_worker_pool = {}

# every find* call may add at most this many instances (the outer loop of ProgressiveX::run is capped at 10 proposals,
# progressive_x.h:272); the native side refuses to truncate if it ever found more
_MODEL_CAP = 16


def _worker_contexts(device: int, n: int):
    pool = _worker_pool.setdefault(device, [])
    while len(pool) < n:
        pool.append(Context(device))
    return pool


def _ctx(device: int) -> Context:
    """The cached per-device context. It is shared by every caller of the find* functions on that device; the calls
    serialise on Context.lock (a pxb_ctx is not re-entrant). Use findHomographiesBatch for concurrent problems."""
    with _contexts_guard:
        if device not in _contexts:
            _contexts[device] = Context(device)
        return _contexts[device]


def getStatistics(device: int = 0) -> dict:
    """ProgressiveX::getStatistics (progressive_x.h:210-213) of the last find* call on `device`: processing_time, the four
    total_time_of_* sums and one entry per accepted round (times from CUDA events on the context's stream, seconds)."""
    return _ctx(device).statistics()


def getMutableSettings(device: int = 0):
    """A MultiModelSettings structure holding the reference's defaults (progressive_x.h:60-75); change fields and hand it
    to setSettings. Fields that the find* argument lists carry are overwritten by those arguments, as in the reference."""
    return _ctx(device).default_settings()


def setSettings(settings, device: int = 0) -> None:
    _ctx(device).set_settings(settings)


_shards = {}


def _nccl_shard(device: int):
    """This rank's NCCL communicator inside libpxb200.so (created on first use; collective over the torch.distributed
    job: every rank must reach this call)."""
    from . import sharding
    if device not in _shards:
        _shards[device] = sharding.NcclShard(_ctx(device))
    return _shards[device]


class distributed:
    """Context manager: inside `with pyprogressivex.distributed(device):` the find* calls of this process are
    COLLECTIVE over the torch.distributed job -- every rank makes the same call with the same arguments, the hypothesis
    blocks of the problem are split over the ranks' GPUs (pxb_ctx_set_shard) and every rank returns the same result, which
    is bit-identical to the single-GPU result for the same seed."""

    def __init__(self, device: int = 0):
        self.shard = _nccl_shard(device)

    def __enter__(self):
        self.shard.attach()
        return self.shard

    def __exit__(self, *exc):
        self.shard.detach()


def _two_view(fn_name, corrs, w1, h1, w2, h2, threshold, conf, spatial_coherence_weight, neighborhood_ball_radius,
              maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id,
              scoring_exponent, do_logging, seed, device):
    corrs = _np.ascontiguousarray(corrs, dtype=_np.float64)
    # bindings.cpp:119-124 / :26-50: shape checks raise std::invalid_argument -> ValueError in Python
    if corrs.ndim != 2 or corrs.shape[1] != 4:
        raise ValueError("corrs should be an array with dims [n,4], n>=4")
    if corrs.shape[0] < 4:
        raise ValueError("corrs should be an array with dims [n,4], n>=4")
    ctx = _ctx(device)
    N = corrs.shape[0]
    labeling = _np.zeros(N, dtype=_np.int64)
    cap = _MODEL_CAP
    models = _np.zeros((cap, 9), dtype=_np.float64)
    fn = getattr(ctx.lib, fn_name)
    with ctx.lock:
        rc = fn(ctx.handle, corrs.ctypes.data_as(_C.c_void_p), N, labeling.ctypes.data_as(_C.c_void_p),
                models.ctypes.data_as(_C.c_void_p), cap, int(w1), int(h1), int(w2), int(h2),
                float(spatial_coherence_weight), float(threshold), float(conf), float(neighborhood_ball_radius),
                float(maximum_tanimoto_similarity), int(max_iters), int(minimum_point_number), int(maximum_model_number),
                int(sampler_id), float(scoring_exponent), int(bool(do_logging)), int(seed))
        M = _native._check(rc)
    assert M <= cap
    return models[:M].reshape(M * 3, 3).copy(), labeling.astype(_np.int32)


def findHomographies(corrs, w1, h1, w2, h2, threshold=4.0, conf=0.5, spatial_coherence_weight=0.0,
                     neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000,
                     minimum_point_number=10, maximum_model_number=-1, sampler_id=3, scoring_exponent=2,
                     do_logging=False, seed=0, device=0):
    """bindings.cpp:99-168 / :410-426 -> findHomographies_ (progressivex_python.cpp:173-304)."""
    return _two_view("pxb_find_homographies", corrs, w1, h1, w2, h2, threshold, conf, spatial_coherence_weight,
                     neighborhood_ball_radius, maximum_tanimoto_similarity, max_iters, minimum_point_number,
                     maximum_model_number, sampler_id, scoring_exponent, do_logging, seed, device)


def findTwoViewMotions(corrs, w1, h1, w2, h2, threshold=4.0, conf=0.5, spatial_coherence_weight=0.0,
                       neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000,
                       minimum_point_number=10, maximum_model_number=-1, sampler_id=3, scoring_exponent=3,
                       do_logging=False, seed=0, device=0):
    """bindings.cpp:324-393 / :444-460 -> findTwoViewMotions_ (progressivex_python.cpp:535-666)."""
    return _two_view("pxb_find_two_view_motions", corrs, w1, h1, w2, h2, threshold, conf, spatial_coherence_weight,
                     neighborhood_ball_radius, maximum_tanimoto_similarity, max_iters, minimum_point_number,
                     maximum_model_number, sampler_id, scoring_exponent, do_logging, seed, device)


findFundamentalMatrices = findTwoViewMotions


def findHomographiesBatch(pairs, w1, h1, w2, h2, distributed=False, workers=2, in_flight=8, threshold=4.0, conf=0.5,
                          spatial_coherence_weight=0.0, neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4,
                          max_iters=1000, minimum_point_number=10, maximum_model_number=-1, sampler_id=3, scoring_exponent=2,
                          do_logging=False, seed=0, per_pair_seed=False, device=0):
    """BASELINE config C4: many independent image pairs. `pairs` is a list of [N_p, 4] correspondence arrays; the other
    arguments are findHomographies'. Returns [(models, labeling), ...], identical to calling findHomographies per pair.

    One fit is a chain of short kernel sequences separated by host decisions, so a single problem leaves the GPU mostly
    idle. The native batch driver (pxb_find_homographies_batch) runs `workers` host threads with `in_flight` problems each
    -- every problem is a fiber with its own stream that yields wherever the sequential driver would block -- entirely
    inside libpxb200.so (no Python threads, no GIL).
    With distributed=True (inside a torch.distributed job, one process per GPU) pair p is solved on rank p mod world and
    the surviving instances of every pair are merged by one ncclAllGather (pxb_allgather_instances); every rank returns
    the full list. Pairs must then share one N (padding is the caller's business)."""
    import ctypes as C

    def solve_many(indices):
        if not indices:
            return {}
        lib = _native.load_library()
        arrs = [_np.ascontiguousarray(pairs[p], dtype=_np.float64) for p in indices]
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != 4 or a.shape[0] < 4:
                raise ValueError("corrs should be an array with dims [n,4], n>=4")
        n = len(arrs)
        cap = _MODEL_CAP
        labs = [_np.zeros(a.shape[0], dtype=_np.int64) for a in arrs]
        mods = [_np.zeros((cap, 9), dtype=_np.float64) for _ in arrs]
        counts = _np.zeros(n, dtype=_np.int32)
        npts = _np.array([a.shape[0] for a in arrs], dtype=_np.int64)
        ptrs = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])  # noqa: E731
        rc = lib.pxb_find_homographies_batch(int(device), n, ptrs(arrs), npts.ctypes.data_as(C.c_void_p), ptrs(labs), ptrs(mods),
                                             cap, counts.ctypes.data_as(C.c_void_p), int(w1), int(h1), int(w2), int(h2),
                                             float(spatial_coherence_weight), float(threshold), float(conf),
                                             float(neighborhood_ball_radius), float(maximum_tanimoto_similarity), int(max_iters),
                                             int(minimum_point_number), int(maximum_model_number), int(sampler_id),
                                             float(scoring_exponent), int(seed), int(bool(per_pair_seed)), int(workers),
                                             int(in_flight))
        _native._check(rc)
        return {p: (mods[i][:counts[i]].reshape(int(counts[i]) * 3, 3).copy(), labs[i].astype(_np.int32))
                for i, p in enumerate(indices)}

    if not distributed:
        res = solve_many(list(range(len(pairs))))
        return [res[p] for p in range(len(pairs))]
    from . import sharding
    shard = _nccl_shard(device)  # this rank's communicator (created on first use from the torch.distributed job)
    rank, world = shard.rank, shard.world
    n_points = int(pairs[0].shape[0])
    mine = list(sharding.pairs_of_rank(len(pairs), rank, world))
    res = solve_many(mine)
    local = [(p, res[p][0].reshape(-1, 9), res[p][1]) for p in mine]
    # the exchange step: one ncclAllGather inside libpxb200.so (pxb_allgather_instances) on this rank's context
    gathered = shard.gather_instances(local, len(pairs), n_points, 9, 10)
    return [(m.reshape(-1, 3), lab) for m, lab in gathered]


def _find_with_context(ctx, corrs, w1, h1, w2, h2, threshold=4.0, conf=0.5, spatial_coherence_weight=0.0,
                       neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000,
                       minimum_point_number=10, maximum_model_number=-1, sampler_id=3, scoring_exponent=2,
                       do_logging=False, seed=0, device=0):
    """findHomographies on an explicit context (used by the concurrent batch)."""
    corrs = _np.ascontiguousarray(corrs, dtype=_np.float64)
    if corrs.ndim != 2 or corrs.shape[1] != 4 or corrs.shape[0] < 4:
        raise ValueError("corrs should be an array with dims [n,4], n>=4")
    N = corrs.shape[0]
    labeling = _np.zeros(N, dtype=_np.int64)
    cap = _MODEL_CAP
    models = _np.zeros((cap, 9), dtype=_np.float64)
    with ctx.lock:
        rc = ctx.lib.pxb_find_homographies(ctx.handle, corrs.ctypes.data_as(_C.c_void_p), N, labeling.ctypes.data_as(_C.c_void_p),
                                           models.ctypes.data_as(_C.c_void_p), cap, int(w1), int(h1), int(w2), int(h2),
                                           float(spatial_coherence_weight), float(threshold), float(conf),
                                           float(neighborhood_ball_radius), float(maximum_tanimoto_similarity), int(max_iters),
                                           int(minimum_point_number), int(maximum_model_number), int(sampler_id),
                                           float(scoring_exponent), int(bool(do_logging)), int(seed))
        M = _native._check(rc)
    assert M <= cap
    return models[:M].reshape(M * 3, 3).copy(), labeling.astype(_np.int32)


def find6DPoses(x1y1, x2y2z2, K, threshold=4.0, conf=0.90, spatial_coherence_weight=0.1,
                neighborhood_ball_radius=20.0, maximum_tanimoto_similarity=0.9, max_iters=400,
                minimum_point_number=2 * 3, maximum_model_number=-1, seed=0, device=0):
    """bindings.cpp:9-97 / :462-473 -> find6DPoses_ (progressivex_python.cpp:41-171)."""
    x1y1 = _np.ascontiguousarray(x1y1, dtype=_np.float64)
    xyz = _np.ascontiguousarray(x2y2z2, dtype=_np.float64)
    Km = _np.ascontiguousarray(K, dtype=_np.float64)
    if x1y1.ndim != 2 or x1y1.shape[1] != 2 or x1y1.shape[0] < 3:
        raise ValueError("x1y1 should be an array with dims [n,2], n>=3")
    if xyz.ndim != 2 or xyz.shape[1] != 3 or xyz.shape[0] != x1y1.shape[0]:
        raise ValueError("x2y2z2 should be an array with dims [n,3], n>=3")
    if Km.shape != (3, 3):
        raise ValueError("K should be an array with dims [3,3]")
    ctx = _ctx(device)
    N = x1y1.shape[0]
    labeling = _np.zeros(N, dtype=_np.int64)
    cap = _MODEL_CAP
    poses = _np.zeros((cap, 12), dtype=_np.float64)
    with ctx.lock:
        rc = ctx.lib.pxb_find_6d_poses(ctx.handle, x1y1.ctypes.data_as(_C.c_void_p), xyz.ctypes.data_as(_C.c_void_p),
                                       Km.ctypes.data_as(_C.c_void_p), N, labeling.ctypes.data_as(_C.c_void_p),
                                       poses.ctypes.data_as(_C.c_void_p), cap, float(spatial_coherence_weight),
                                       float(threshold), float(conf), float(neighborhood_ball_radius),
                                       float(maximum_tanimoto_similarity), int(max_iters), int(minimum_point_number),
                                       int(maximum_model_number), int(seed))
        M = _native._check(rc)
    assert M <= cap
    return poses[:M].reshape(M * 3, 4).copy(), labeling.astype(_np.int32)


def _points_family(fn_name, rows, dim, what, weights, w, h, threshold, conf, spatial_coherence_weight,
                   neighborhood_ball_radius, maximum_tanimoto_similarity, max_iters, minimum_point_number,
                   maximum_model_number, sampler_id, scoring_exponent, do_logging, seed, device):
    rows = _np.ascontiguousarray(rows, dtype=_np.float64)
    if rows.ndim != 2 or rows.shape[1] != dim or rows.shape[0] < 2:  # bindings.cpp:189-194 / :266-269
        raise ValueError(what)
    wts = None
    if weights is not None and _np.ndim(weights) > 0:  # bindings.cpp:201-209: a 0-d array means "no weights"
        wts = _np.ascontiguousarray(weights, dtype=_np.float64).reshape(-1)
        if wts.size == 0:
            wts = None
        elif wts.size != rows.shape[0]:
            raise ValueError("weights should hold one entry per row")
    ctx = _ctx(device)
    N = rows.shape[0]
    labeling = _np.zeros(N, dtype=_np.int64)
    cap = _MODEL_CAP
    models = _np.zeros((cap, 3), dtype=_np.float64)
    fn = getattr(ctx.lib, fn_name)
    with ctx.lock:
        rc = fn(ctx.handle, rows.ctypes.data_as(_C.c_void_p), None if wts is None else wts.ctypes.data_as(_C.c_void_p), N,
                labeling.ctypes.data_as(_C.c_void_p), models.ctypes.data_as(_C.c_void_p), cap, int(w), int(h),
                float(spatial_coherence_weight), float(threshold), float(conf), float(neighborhood_ball_radius),
                float(maximum_tanimoto_similarity), int(max_iters), int(minimum_point_number), int(maximum_model_number),
                int(sampler_id), float(scoring_exponent), int(bool(do_logging)), int(seed))
        M = _native._check(rc)
    assert M <= cap
    return models[:M].copy(), labeling.astype(_np.int32)


def findVanishingPoints(lines, weights, w, h, threshold=4.0, conf=0.5, spatial_coherence_weight=0.0,
                        neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000,
                        minimum_point_number=10, maximum_model_number=-1, sampler_id=3, scoring_exponent=2,
                        do_logging=False, seed=0, device=0):
    """bindings.cpp:168-245 / :428-442 -> findVanishingPoints_ (progressivex_python.cpp:306-423). Returns
    (vanishing_points float64 [M, 3], labeling int32 [N]). As in the reference only sampler_id 0 and 1 exist for this
    entry: the DEFAULT sampler_id = 3 prints "Unknown sampler identifier" and returns no models."""
    return _points_family("pxb_find_vanishing_points", lines, 4, "lines should be an array with dims [n,4], n>=2", weights,
                          w, h, threshold, conf, spatial_coherence_weight, neighborhood_ball_radius,
                          maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id,
                          scoring_exponent, do_logging, seed, device)


def findLines(points, weights, w, h, threshold=2.0, conf=0.5, spatial_coherence_weight=0.0,
              neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000, minimum_point_number=10,
              maximum_model_number=-1, sampler_id=3, scoring_exponent=2, do_logging=False, seed=0, device=0):
    """bindings.cpp:247-322 / :476-491 -> findLines_ (progressivex_python.cpp:425-535). Returns (lines float64 [M, 3] as
    (nx, ny, c), labeling int32 [N]). Samplers 0, 1, 2 (= NAPSAC); the default 3 is unknown to the reference too."""
    return _points_family("pxb_find_lines", points, 2, "Points should be an array with dims [n,3], n>=2", weights, w, h,
                          threshold, conf, spatial_coherence_weight, neighborhood_ball_radius,
                          maximum_tanimoto_similarity, max_iters, minimum_point_number, maximum_model_number, sampler_id,
                          scoring_exponent, do_logging, seed, device)
