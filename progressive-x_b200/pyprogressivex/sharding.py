"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for plumbing.

Two modes, both embarrassingly parallel on the data path:

  * hypothesis blocks of ONE large pair (C3, C5): the points are replicated on every rank, rank r scores the
    hypotheses [block_bounds(K, world)[r], ...[r+1]) and the per-hypothesis summaries (count, value, shared) -- 24
    bytes each -- are all-gathered so that every rank's host replay sees them in sample order.
  * independent pairs (C4): pair p goes to rank p mod world; at the end (M, models, labels) of every pair are
    gathered to all ranks.

The collectives are backend agnostic (NCCL over NVLink on the GPU box, gloo in the CPU tests). Messages are KB-MB,
i.e. latency bound: no fused compute+collective kernel is warranted; callers overlap the gather with the next
block's solve.
"""
from __future__ import annotations

import ctypes as _C
from typing import Callable, List, Sequence, Tuple

import numpy as np


def block_bounds(K: int, world: int) -> List[int]:
    """Contiguous hypothesis blocks: rank r owns [b[r], b[r+1]); hypothesis k lives on rank floor(k*world/K)."""
    return [(K * r + world - 1) // world if r else 0 for r in range(world)] + [K]


def owner_of(k: int, K: int, world: int) -> int:
    b = block_bounds(K, world)
    return int(np.searchsorted(b, k, side="right") - 1)


def pairs_of_rank(n_pairs: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_pairs, world))


def _dist():
    import torch.distributed as dist
    return dist


def allgather_hypothesis_summaries(count, value, shared, K: int, device=None):
    """count/value/shared: this rank's block (numpy or torch, length b[r+1]-b[r]). Returns three numpy arrays of
    length K in sample order, identical on every rank."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    b = block_bounds(K, world)
    width = max(b[r + 1] - b[r] for r in range(world))
    dev = device if device is not None else torch.device("cpu")
    local = torch.zeros((width, 3), dtype=torch.float64, device=dev)
    n = b[rank + 1] - b[rank]
    local[:n, 0] = torch.as_tensor(np.asarray(count), dtype=torch.float64, device=dev)  # counts < 2^53: exact
    local[:n, 1] = torch.as_tensor(np.asarray(value), dtype=torch.float64, device=dev)
    local[:n, 2] = torch.as_tensor(np.asarray(shared), dtype=torch.float64, device=dev)
    out = torch.empty((world * width, 3), dtype=torch.float64, device=dev)  # gloo wants the dim-0 concatenation
    dist.all_gather_into_tensor(out, local)
    out = out.cpu().numpy().reshape(world, width, 3)
    cnt = np.concatenate([out[r, : b[r + 1] - b[r], 0] for r in range(world)]).astype(np.int64)
    val = np.concatenate([out[r, : b[r + 1] - b[r], 1] for r in range(world)])
    shr = np.concatenate([out[r, : b[r + 1] - b[r], 2] for r in range(world)])
    return cnt, val, shr


def score_hypotheses_sharded(score_fn: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]],
                             models: np.ndarray, device=None):
    """Every rank holds all K models (they are tiny); rank r evaluates its block with score_fn (normally
    Context.score_compound on its own GPU) and the summaries are all-gathered."""
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    K = models.shape[0]
    b = block_bounds(K, world)
    cnt, val, shr = score_fn(models[b[rank]:b[rank + 1]])
    return allgather_hypothesis_summaries(cnt, val, shr, K, device)


def gather_instances(local_results: Sequence[Tuple[int, np.ndarray, np.ndarray]], n_pairs: int, n_points: int,
                     model_size: int = 9, max_models: int = 10, device=None):
    """local_results: [(pair_index, models [M, model_size], labels [n_points] int32), ...] of this rank's pairs.
    Returns a list of length n_pairs with (models, labels) for every pair, identical on every rank. The per-pair
    record is {M:int32} + max_models*model_size float64 + int32 labels[n_points]."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_pairs + world - 1) // world
    dev = device if device is not None else torch.device("cpu")
    models = torch.zeros((per_rank, max_models, model_size), dtype=torch.float64, device=dev)
    labels = torch.full((per_rank, n_points), -1, dtype=torch.int32, device=dev)
    counts = torch.full((per_rank,), -1, dtype=torch.int32, device=dev)
    for pair, m, lab in local_results:
        slot = pair // world
        assert pair % world == rank, "pair does not belong to this rank"
        M = min(int(m.shape[0]), max_models)
        counts[slot] = M
        if M:
            models[slot, :M] = torch.as_tensor(np.asarray(m[:M]).reshape(M, model_size), dtype=torch.float64, device=dev)
        labels[slot] = torch.as_tensor(np.asarray(lab), dtype=torch.int32, device=dev)
    g_models = torch.empty((world * per_rank, max_models, model_size), dtype=models.dtype, device=dev)
    g_labels = torch.empty((world * per_rank, n_points), dtype=labels.dtype, device=dev)
    g_counts = torch.empty((world * per_rank,), dtype=counts.dtype, device=dev)
    dist.all_gather_into_tensor(g_models, models)
    dist.all_gather_into_tensor(g_labels, labels)
    dist.all_gather_into_tensor(g_counts, counts)
    g_models = g_models.cpu().numpy().reshape(world, per_rank, max_models, model_size)
    g_labels = g_labels.cpu().numpy().reshape(world, per_rank, n_points)
    g_counts = g_counts.cpu().numpy().reshape(world, per_rank)
    out = []
    for pair in range(n_pairs):
        r, slot = pair % world, pair // world
        M = int(g_counts[r, slot])
        out.append((g_models[r, slot, :max(M, 0)].copy(), g_labels[r, slot].copy()))
    return out


# ---- the native exchange steps (include/pxb200.h "multi-GPU exchange steps"): NCCL inside libpxb200.so --------------------
class NcclShard:
    """An ncclComm_t owned by libpxb200.so, one rank per GPU, created from a unique id that rank 0 draws and
    torch.distributed hands to the other ranks (any backend: this is the only thing torch does on this path).

        shard = NcclShard(ctx)              # inside a torch.distributed job, ctx = this rank's Context
        shard.attach()                      # find* calls on ctx are now collective, hypothesis blocks split over ranks
        models, labels = pyprogressivex.find6DPoses(..., seed=1)   # same call, same arguments on every rank
        shard.detach(); shard.close()
    """

    def __init__(self, ctx, world: int | None = None, rank: int | None = None, unique_id: bytes | None = None):
        from . import _native
        _native.preload_nccl()
        self.ctx, self.lib = ctx, ctx.lib
        if world is None:
            dist = _dist()
            world, rank = dist.get_world_size(), dist.get_rank()
        self.world, self.rank = int(world), int(rank)
        if unique_id is None:
            buf = (_C.c_char * 128)()
            if self.rank == 0:
                _native._check(self.lib.pxb_nccl_unique_id(buf))
            box = [bytes(buf)]
            if self.world > 1:
                _dist().broadcast_object_list(box, src=0)
            unique_id = box[0]
        comm = _C.c_void_p()
        _native._check(self.lib.pxb_nccl_comm_init(ctx.handle, unique_id, self.world, self.rank, _C.byref(comm)))
        self.comm = comm

    def attach(self):
        from . import _native
        _native._check(self.lib.pxb_ctx_set_shard(self.ctx.handle, self.comm))

    def detach(self):
        self.lib.pxb_ctx_set_shard(self.ctx.handle, None)

    def close(self):
        if getattr(self, "comm", None):
            self.detach()
            self.lib.pxb_nccl_comm_destroy(self.comm)
            self.comm = None

    def __enter__(self):
        self.attach()
        return self

    def __exit__(self, *exc):
        self.detach()

    def gather_instances(self, local_results, n_pairs: int, n_points: int, model_size: int = 9, max_models: int = 10):
        """Same contract as gather_instances() above, through pxb_allgather_instances (one ncclAllGather of fixed-size
        records on the context's stream)."""
        from . import _native
        world, rank = self.world, self.rank
        per_rank = (n_pairs + world - 1) // world
        counts = np.full(per_rank, -1, dtype=np.int32)
        models = np.zeros((per_rank, max_models, model_size), dtype=np.float64)
        labels = np.full((per_rank, n_points), -1, dtype=np.int32)
        for pair, m, lab in local_results:
            assert pair % world == rank, "pair does not belong to this rank"
            slot = pair // world
            M = min(int(m.shape[0]), max_models)
            counts[slot] = M
            if M:
                models[slot, :M] = np.asarray(m[:M], dtype=np.float64).reshape(M, model_size)
            labels[slot] = np.asarray(lab, dtype=np.int32)
        g_counts = np.empty(world * per_rank, dtype=np.int32)
        g_models = np.empty((world * per_rank, max_models, model_size), dtype=np.float64)
        g_labels = np.empty((world * per_rank, n_points), dtype=np.int32)
        p = lambda a: a.ctypes.data_as(_C.c_void_p)  # noqa: E731
        _native._check(self.lib.pxb_allgather_instances(self.ctx.handle, self.comm, per_rank, n_points, model_size, max_models,
                                                        p(counts), p(models), p(labels), p(g_counts), p(g_models), p(g_labels)))
        out = []
        for pair in range(n_pairs):
            at = (pair % world) * per_rank + pair // world
            M = int(g_counts[at])
            out.append((g_models[at, :max(M, 0)].copy(), g_labels[at].copy()))
        return out
