#!/usr/bin/env python
"""Multi-GPU parity gates of SURVEY.md 8(e), run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_parity.py [--n-pose 20000] [--pairs 16] [--n-pair 2000]

  (1) hypothesis-block sharding of ONE problem: find6DPoses / findHomographies / findTwoViewMotions inside
      `pyprogressivex.distributed()` must return, on every rank, exactly the models and labels of the single-GPU run with
      the same seed (every rank also computes that local run on its own GPU);
  (2) pair sharding: findHomographiesBatch(distributed=True) must equal the undistributed batch.
Prints one JSON line on rank 0 and exits non-zero on any mismatch. Also imported by bench.py and tests/test_gpu_sharded.py."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))


def same(a, b):
    return a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64)) and np.array_equal(a[1], b[1])


def check_single_problem(px, syn, dev, n_pose, n_h, n_f, seeds=(1, 2)):
    """-> dict of timings; raises AssertionError on any difference between the sharded and the local run."""
    out = {}
    img, world_pts, K, gt, _ = syn.multi_pose_scene(n_pose, n_objects=4, inlier_ratio_each=0.12, noise_px=1.0, seed=5)
    corr_h, _, _ = syn.multi_homography_scene(n_h, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=6)
    corr_f, _, _ = syn.multi_motion_scene(n_f, seed=7)
    cases = {
        "find6DPoses": lambda seed: px.find6DPoses(img, world_pts, K, threshold=4.0, conf=0.9, spatial_coherence_weight=0.0,
                                                    maximum_tanimoto_similarity=0.9, max_iters=2000,
                                                    minimum_point_number=max(6, n_pose // 50), seed=seed, device=dev),
        "findHomographies": lambda seed: px.findHomographies(corr_h, 1024, 768, 1024, 768, threshold=2.0, conf=0.5,
                                                              max_iters=1000, minimum_point_number=max(10, n_h // 50),
                                                              sampler_id=0, seed=seed, device=dev),
        "findTwoViewMotions": lambda seed: px.findTwoViewMotions(corr_f, 1024, 768, 1024, 768, threshold=0.75, conf=0.5,
                                                                  max_iters=600, minimum_point_number=max(14, n_f // 20),
                                                                  sampler_id=0, seed=seed, device=dev),
    }
    for name, fn in cases.items():
        for seed in seeds:
            local = fn(seed)
            t0 = time.perf_counter()
            local = fn(seed)
            t_local = time.perf_counter() - t0
            with px.distributed(dev):
                sharded = fn(seed)
                t0 = time.perf_counter()
                sharded = fn(seed)
                t_shard = time.perf_counter() - t0
            assert same(local, sharded), f"{name} seed {seed}: sharded result differs from the single-GPU run"
            assert local[0].shape[0] > 0, f"{name} seed {seed}: no model found (the gate would be vacuous)"
            out[f"{name}/seed{seed}"] = {"models": int(local[0].shape[0] // 3), "local_ms": t_local * 1e3, "sharded_ms": t_shard * 1e3}
    return out


def check_pairs(px, syn, dev, pairs, n_pair):
    scenes = [syn.multi_homography_scene(n_pair, n_planes=3, outlier_ratio=0.4, noise=0.5, seed=900 + p)[0] for p in range(pairs)]
    kw = dict(threshold=2.0, conf=0.5, max_iters=1000, minimum_point_number=max(10, n_pair // 40), sampler_id=0, seed=11, device=dev)
    local = px.findHomographiesBatch(scenes, 1024, 768, 1024, 768, workers=4, **kw)
    gathered = px.findHomographiesBatch(scenes, 1024, 768, 1024, 768, workers=4, distributed=True, **kw)
    assert len(local) == len(gathered) == pairs
    for p in range(pairs):
        assert same(local[p], gathered[p]), f"pair {p}: gathered result differs from the single-rank one"
    return {"pairs": pairs, "models_mean": float(np.mean([m.shape[0] // 3 for m, _ in local]))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-pose", type=int, default=20000)
    ap.add_argument("--n-h", type=int, default=8000)
    ap.add_argument("--n-f", type=int, default=4000)
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--n-pair", type=int, default=2000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    import pyprogressivex as px
    from pyprogressivex import synthetic as syn
    res = {"world": world}
    res["single_problem"] = check_single_problem(px, syn, dev, args.n_pose, args.n_h, args.n_f)
    res["pairs"] = check_pairs(px, syn, dev, args.pairs, args.n_pair)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"shard_parity": "ok", **res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
