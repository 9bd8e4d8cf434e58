"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: hypothesis-block sharding with the summary all-gather
and pair sharding with the instance gather. The per-rank compute is the oracle here (the GPU kernels
cannot run in this container); on the GPU box the same functions are driven with Context.score_compound."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from pyprogressivex import sharding
from pyprogressivex import synthetic as syn


def test_block_bounds_partition():
    for K in (0, 1, 7, 10, 10_000):
        for world in (1, 2, 3, 8):
            b = sharding.block_bounds(K, world)
            assert b[0] == 0 and b[-1] == K and all(x <= y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
            for k in range(0, K, max(1, K // 13)):
                r = sharding.owner_of(k, K, world)
                assert b[r] <= k < b[r + 1]
    assert sharding.pairs_of_rank(7, 1, 3) == [1, 4]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    sys.path.insert(0, str(root / "progressive-x_b200"))
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- hypothesis blocks of one pair ----
        pts, gt, Hs = syn.multi_homography_scene(3000, seed=4)
        S = syn.minimal_samples(gt, 101, 4, seed=4)
        models, n, _, _ = O.solve_minimal(0, pts, S)
        models = models[:, 0][n > 0]
        cp = O.preference_vector(0, pts, Hs[0].reshape(-1), 9.0)
        cnt, val, shr = sharding.score_hypotheses_sharded(lambda m: O.score_batch(0, pts, m, 9.0, cp), models)
        # ---- independent pairs ----
        n_pairs, N = 5, 400
        local = []
        for p in sharding.pairs_of_rank(n_pairs, rank, world):
            rng = np.random.default_rng(100 + p)
            M = p % 3
            local.append((p, rng.normal(size=(M, 9)), rng.integers(0, M + 1, N).astype(np.int32)))
        gathered = sharding.gather_instances(local, n_pairs, N)
        q.put((rank, cnt, val, shr, [(m.copy(), l.copy()) for m, l in gathered]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    from oracle import oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process ground truth
    pts, gt, Hs = syn.multi_homography_scene(3000, seed=4)
    S = syn.minimal_samples(gt, 101, 4, seed=4)
    models, n, _, _ = O.solve_minimal(0, pts, S)
    models = models[:, 0][n > 0]
    cp = O.preference_vector(0, pts, Hs[0].reshape(-1), 9.0)
    cnt, val, shr = O.score_batch(0, pts, models, 9.0, cp)
    for rank, c, v, s, gathered in results:
        assert np.array_equal(c, cnt) and np.array_equal(v, val) and np.array_equal(s, shr)  # sample order, bit-exact
        assert len(gathered) == 5
        for p, (m, lab) in enumerate(gathered):
            rng = np.random.default_rng(100 + p)
            M = p % 3
            assert m.shape == (M, 9) and np.array_equal(m, rng.normal(size=(M, 9)))
            assert np.array_equal(lab, rng.integers(0, M + 1, 400).astype(np.int32))
