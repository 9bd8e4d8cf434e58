"""Sharded path on the GPU. The NCCL exchange code runs with a ONE-rank communicator on any box (same
record packing, broadcast and all-gather calls as with N ranks); with two or more GPUs the real two-rank gates of
tools/shard_parity.py run under torchrun."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _same(a, b):
    return a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64)) and np.array_equal(a[1], b[1])


@pytest.fixture(scope="module")
def shard1():
    import pyprogressivex as px
    from pyprogressivex import sharding
    ctx = px._ctx(0)
    sh = sharding.NcclShard(ctx, world=1, rank=0)
    yield sh
    sh.close()


def test_one_rank_communicator_runs_the_sharded_driver(shard1):
    import ctypes as C
    import pyprogressivex as px
    from pyprogressivex import synthetic as syn
    v = C.c_int()
    assert shard1.lib.pxb_nccl_version(C.byref(v)) == 0 and v.value >= 21800
    corr, _, _ = syn.multi_homography_scene(6000, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=3)
    kw = dict(threshold=2.0, conf=0.5, max_iters=1000, minimum_point_number=100, sampler_id=0)
    for lam in (0.0, 0.05):
        local = px.findHomographies(corr, 1024, 768, 1024, 768, spatial_coherence_weight=lam, seed=4, **kw)
        with shard1:
            w, r = C.c_int(), C.c_int()
            shard1.lib.pxb_shard_info(shard1.ctx.handle, C.byref(w), C.byref(r))
            assert (w.value, r.value) == (1, 0)
            sharded = px.findHomographies(corr, 1024, 768, 1024, 768, spatial_coherence_weight=lam, seed=4, **kw)
        assert local[0].shape[0] >= 3 and _same(local, sharded)
    img, wpts, K, _, _ = syn.multi_pose_scene(8000, n_objects=3, inlier_ratio_each=0.15, noise_px=1.0, seed=2)
    pk = dict(threshold=4.0, conf=0.9, spatial_coherence_weight=0.0, max_iters=1500, minimum_point_number=200, seed=9)
    local = px.find6DPoses(img, wpts, K, **pk)
    with shard1:
        sharded = px.find6DPoses(img, wpts, K, **pk)
    assert local[0].shape[0] >= 3 and _same(local, sharded)
    cf, _, _ = syn.multi_motion_scene(3000, seed=1)
    fk = dict(threshold=0.75, conf=0.5, max_iters=400, minimum_point_number=150, sampler_id=0, seed=5)
    local = px.findTwoViewMotions(cf, 1024, 768, 1024, 768, **fk)
    with shard1:
        sharded = px.findTwoViewMotions(cf, 1024, 768, 1024, 768, **fk)
    assert local[0].shape[0] >= 3 and _same(local, sharded)


def test_allgather_instances_one_rank(shard1):
    rng = np.random.default_rng(0)
    n_pairs, N = 5, 333
    local = []
    for p in range(n_pairs):
        M = p % 4
        local.append((p, rng.normal(size=(M, 9)), rng.integers(0, M + 1, N).astype(np.int32)))
    out = shard1.gather_instances(local, n_pairs, N, 9, 10)
    for (p, m, lab), (gm, gl) in zip(local, out):
        assert np.array_equal(m.reshape(-1, 9), gm) and np.array_equal(lab, gl)


def test_two_rank_parity_gates():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run by `gpurun --gpus 2` and by bench.py --gpus N)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tools" / "shard_parity.py"), "--n-pose", "12000", "--pairs", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0 and '"shard_parity": "ok"' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
