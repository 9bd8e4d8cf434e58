#!/usr/bin/env python
"""bench.py -- residual evaluations/s of the Progressive-X hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the residual-and-inlier-matrix kernel (exact float64 r2 + 1 mask bit per entry) over the
headline grid of SURVEY.md section 8(d): N = 50 000 correspondences x K = 10 000 four-point homography hypotheses,
inputs resident in HBM.  With --gpus N every rank owns a disjoint block of 10 000 hypotheses over the same
(replicated) points -- weak scaling, no collective on the data path (the matrix has no exchange step; the
hypothesis-summary all-gather of the sharded RANSAC loop is timed separately and reported under "exchange").

Printed JSON (one line, rank 0): the driver contract + "roofline", "cpu_baseline", "e2e", "clocks",
"gpu_launches" and explanatory extras: "score_kernel" (the fused score path the RANSAC loop uses), "screening_f32" (the
f32-output matrix), and the second half of the BASELINE metric -- "fits" (C2, one problem at a time), "fits_lambda" (C2
with the spatial term), "fits_batch" (C4 in miniature: independent pairs on concurrent contexts).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "progressive-x_b200"))

N_POINTS = 50_000
K_HYPS = 10_000
THRESHOLD = 2.0
T2 = (1.5 * THRESHOLD) ** 2
BYTES_PER_EVAL = 8.125  # 8 B float64 r2 + 1 mask bit (SURVEY.md 8d / BASELINE.md section 3)
METRIC = "residual evals/sec (N pts x K hyps)"
UNIT = "evals/s"


def workload(seed=0, n_points=N_POINTS, k_hyps=K_HYPS):
    from pyprogressivex import synthetic as syn
    pts, gt, _ = syn.multi_homography_scene(n_points, n_planes=5, outlier_ratio=0.4, noise=0.5, seed=seed)
    samples = syn.minimal_samples(gt, k_hyps, 4, within_ratio=0.5, seed=seed)
    return pts, gt, samples


def workload_name(n, k):
    return (f"synthetic multi-H grid: N={n} correspondences (5 planes x 12% + 40% outliers) x K={k} four-point "
            f"hypotheses per GPU, thr={THRESHOLD}px; exact f64 r2 + inlier bit")


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def committed_traffic():
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("k_residual_matrix_f64_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md 'clocks' line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.nvml_rows, self._stop, self._thr = [], False, None

    def _init_nvml(self):
        # NVML is initialised synchronously, BEFORE the timed region: importing pynvml inside the polling thread can take
        # longer than the whole timed region (tens of milliseconds), which left the N > 1 runs without samples.
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self._mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _sample_nvml(self):
        nv, h = self._nv, self._h
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.nvml_rows.append((sm, self._mx, reasons))

    def _poll_nvml(self):
        # in-process NVML polling every ~2 ms: the timed region is tens of milliseconds, far shorter than the 200 ms
        # period of `nvidia-smi -lms`, so the subprocess alone would miss it.
        try:
            while not self._stop:
                self._sample_nvml()
                time.sleep(0.002)
        except Exception:
            pass

    def start(self):
        self._init_nvml()
        if self._nv is not None:
            try:
                self._sample_nvml()  # at least one sample, taken right at the start of the timed region
            except Exception:
                pass
            self._thr = threading.Thread(target=self._poll_nvml, daemon=True)
            self._thr.start()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        self._stop = True
        if self._thr:
            self._thr.join(timeout=1)
        if self.nvml_rows:
            # NVML bit masks (nvml.h nvmlClocksEventReasons*): 0x4 sw_power_cap, 0x8 hw_slowdown,
            # 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
            bits = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
            for _, _, r in self.nvml_rows:
                for b, n in bits.items():
                    if r & b:
                        reasons.add(n)
            sm = [float(r[0]) for r in self.nvml_rows]
            mx = [float(r[1]) for r in self.nvml_rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "pynvml polled every 2 ms during the timed region" if self.nvml_rows else "nvidia-smi -lms 200"}


def cpu_baseline(pts, models, target_seconds=10.0):
    """The oracle's restated getScore loops (kind 'port') on all host cores, bounded sample of the same workload."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    k_cal = min(models.shape[0], max(cores * 4, 16))
    t0 = time.perf_counter()
    O.score_batch(0, pts, models[:k_cal], T2, None, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    k = int(min(models.shape[0], max(k_cal, k_cal * target_seconds / dt)))
    O.score_batch(0, pts, models[:k], T2, None, threads=cores)  # warm-up pass (thread pool, page faults)
    t0 = time.perf_counter()
    O.score_batch(0, pts, models[:k], T2, None, threads=cores)
    one = max(time.perf_counter() - t0, 1e-4)
    reps = int(max(1, min(200, round(target_seconds / one))))
    t0 = time.perf_counter()
    for _ in range(reps):
        O.score_batch(0, pts, models[:k], T2, None, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": pts.shape[0] * k * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle pxo_score_batch (restated getScore loop, OpenMP over hypotheses): {reps} passes over "
                      f"N={pts.shape[0]} x K={k} of the {models.shape[0]} bench hypotheses, {dt:.2f} s wall"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference itself cannot be compiled here, see
    DESIGN.md) on all host threads, bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    pts, gt, samples = workload()
    cores = os.cpu_count() or 1
    models, n, _, _ = O.solve_minimal(0, pts, samples[:4096])
    models = models[:, 0][n > 0]
    # size one step to ~1.5 s
    k_cal = min(models.shape[0], cores * 4)
    t0 = time.perf_counter()
    O.score_batch(0, pts, models[:k_cal], T2, None, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    k = int(min(models.shape[0], max(k_cal, k_cal * 1.5 / dt)))
    O.score_batch(0, pts, models[:k], T2, None, threads=cores)
    t0 = time.perf_counter()
    O.score_batch(0, pts, models[:k], T2, None, threads=cores)
    one = max(time.perf_counter() - t0, 1e-4)
    reps = int(max(1, min(100, round(1.0 / one))))  # ~1 s of host work per step

    def step():
        for _ in range(reps):
            O.score_batch(0, pts, models[:k], T2, None, threads=cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = pts.shape[0] * k * reps * args.steps / dt
    sample = (f"oracle port of getScore (scoring_function_with_compound_model.h:61-125) with OpenMP over hypotheses; "
              f"each step = {reps} passes over N={pts.shape[0]} x K={k} hypotheses of the bench workload")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N_POINTS, K_HYPS),
                   "sample": "bounded sample of that workload per step (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pxb200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from pyprogressivex import _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    ctx = _native.Context(local_rank)
    lib = ctx.lib
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- workload: same points everywhere, a disjoint hypothesis block per rank ------------------------------
    pts, gt, _ = workload(seed=0)
    from pyprogressivex import synthetic as syn
    samples = syn.minimal_samples(gt, K_HYPS, 4, within_ratio=0.5, seed=1000 + rank)
    ctx.upload_points(_native.MODEL_H, pts)
    models_np, n, sv, mv = ctx.solve_minimal(samples)          # hypotheses come from the GPU four-point solver
    models_np = np.ascontiguousarray(np.where((n > 0)[:, None], models_np[:, 0], np.eye(3).reshape(1, 9)))
    K, N = models_np.shape[0], pts.shape[0]
    words = (N + 31) // 32

    with torch.cuda.stream(stream):
        models = torch.from_numpy(models_np).to(dev)
        r2 = torch.empty((K, N), dtype=torch.float64, device=dev)
        mask = torch.empty((K, words), dtype=torch.int32, device=dev)
        cnt = torch.empty(K, dtype=torch.int64, device=dev)
        val = torch.empty(K, dtype=torch.float64, device=dev)
        shr = torch.empty(K, dtype=torch.float64, device=dev)
        r2f = None
    stream.synchronize()

    def step_matrix():
        rc = lib.pxb_residual_matrix_dev(ctx.handle, models.data_ptr(), K, T2, r2.data_ptr(), mask.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.pxb_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        launches0 = ctx.launch_count()
        with torch.cuda.stream(stream):
            evs[0].record(stream)
            for i in range(steps):
                fn()
                evs[i + 1].record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        total_ms = evs[0].elapsed_time(evs[steps])
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, per, ctx.launch_count() - launches0, clocks

    total_ms, per_launch_ms, launches, clocks = timed(step_matrix, args.steps, args.warmup, sample_clocks=True)
    if clocks and any(r in clocks["reasons"] for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")):
        total_ms, per_launch_ms, launches, clocks = timed(step_matrix, args.steps, args.warmup, sample_clocks=True)
        clocks["remeasured"] = True
    evals_per_step = N * K
    value = evals_per_step * world * args.steps / (total_ms * 1e-3)
    kernel_ms = float(np.mean(per_launch_ms))
    peak, peak_src = measured_peak()
    achieved = evals_per_step * BYTES_PER_EVAL / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": committed_traffic(), "kernel": "k_residual_matrix<H,f64>",
                "algorithmic_bytes_per_launch": evals_per_step * BYTES_PER_EVAL, "launch_ms": kernel_ms,
                "peak_source": peak_src + "; burst figure (kernel timed alone)"}

    # ---- explanatory extras: fused score kernel and the f32-screening matrix ---------------------------------
    def step_score():
        rc = lib.pxb_score_compound_dev(ctx.handle, models.data_ptr(), K, T2, None, cnt.data_ptr(), val.data_ptr(),
                                        shr.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.pxb_last_error().decode())

    sc_ms, _, _, _ = timed(step_score, max(3, args.steps // 2), 3)
    sc_steps = max(3, args.steps // 2)
    extras = {"score_kernel": {"evals_per_s": evals_per_step * world * sc_steps / (sc_ms * 1e-3),
                               "note": "k_screen_prepare + k_score_screened + k_score_finalize: count / sum score / shared "
                                       "support, no matrix written; float32 screening proves outliers, exact float64 only "
                                       "for queued candidates (issue bound)"}}
    del r2
    with torch.cuda.stream(stream):
        r2f = torch.empty((K, N), dtype=torch.float32, device=dev)

    def step_f32():
        rc = lib.pxb_residual_matrix_f32_dev(ctx.handle, models.data_ptr(), K, T2, r2f.data_ptr(), mask.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.pxb_last_error().decode())

    f32_ms, f32_per, _, _ = timed(step_f32, sc_steps, 3)
    f32_achieved = evals_per_step * 4.125 / (float(np.mean(f32_per)) * 1e-3) / 1e9
    extras["screening_f32"] = {"evals_per_s": evals_per_step * world * sc_steps / (f32_ms * 1e-3),
                               "roofline_frac": f32_achieved / peak, "bytes_per_eval": 4.125}
    del r2f

    # ---- hypothesis-summary exchange of the sharded RANSAC loop (N > 1 only) ---------------------------------
    if world > 1:
        summ = torch.stack([cnt.to(torch.float64), val, shr], 1).contiguous()
        gathered = torch.empty((world * summ.shape[0], summ.shape[1]), dtype=summ.dtype, device=dev)

        def step_exchange():
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, summ)

        ex_ms, _, _, _ = timed(step_exchange, 10, 3)
        extras["exchange"] = {"op": "ncclAllGather of (count,value,shared) per hypothesis", "bytes_per_rank": summ.numel() * 8,
                              "ms": ex_ms / 10}

    # ---- end to end through the host-pointer C ABI (the call a user of the plugin makes) ----------------------
    # per step: H2D of the points and the hypothesis block from pinned host memory, fused getScore batch
    # (pxb_score_compound = the ScoringFunction seam), D2H of the K scores, then the best hypothesis' inlier list.
    pin_pts = torch.from_numpy(pts).pin_memory()
    pin_models = torch.from_numpy(models_np).pin_memory()
    out_cnt = torch.empty(K, dtype=torch.int64).pin_memory()
    out_val = torch.empty(K, dtype=torch.float64).pin_memory()
    out_shr = torch.empty(K, dtype=torch.float64).pin_memory()
    inl = np.empty(N, dtype=np.int64)
    n_inl = C.c_int64()

    def step_e2e():
        _native._check(lib.pxb_upload_points(ctx.handle, _native.MODEL_H, pin_pts.data_ptr(), N))
        _native._check(lib.pxb_score_compound(ctx.handle, pin_models.data_ptr(), K, T2, None, out_cnt.data_ptr(),
                                              out_val.data_ptr(), out_shr.data_ptr()))
        best = int(torch.argmax(out_val))
        _native._check(lib.pxb_inliers(ctx.handle, pin_models[best].data_ptr(), T2, inl.ctypes.data_as(C.c_void_p),
                                       C.byref(n_inl)))

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": evals_per_step * world * e2e_steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(pts.nbytes + models_np.nbytes + 72),
           "d2h_bytes_per_step": int(K * 24 + words * 4),
           "call": "pxb_upload_points + pxb_score_compound + pxb_inliers (host pointers, copies inside the timed region)"}

    # ---- second half of the BASELINE metric: complete multi-model fits per second (config C2) ---------------------
    import pyprogressivex
    c2_pts, c2_gt, _ = syn.multi_homography_scene(10_000, n_planes=5, outlier_ratio=0.4, noise=0.5, seed=42 + rank)
    pyprogressivex._contexts[local_rank] = ctx  # reuse this process' context
    fit_kwargs = dict(threshold=2.0, conf=0.5, spatial_coherence_weight=0.0, neighborhood_ball_radius=200.0,
                      maximum_tanimoto_similarity=0.4, max_iters=1000, minimum_point_number=100,
                      maximum_model_number=-1, sampler_id=0, scoring_exponent=2, device=local_rank)
    pyprogressivex.findHomographies(c2_pts, 1024, 768, 1024, 768, seed=1, **fit_kwargs)
    n_fits = 5
    barrier()
    t0 = time.perf_counter()
    n_models = []
    for i in range(n_fits):
        m_, l_ = pyprogressivex.findHomographies(c2_pts, 1024, 768, 1024, 768, seed=2 + i, **fit_kwargs)
        n_models.append(m_.shape[0] // 3)
    barrier()
    fit_s = time.perf_counter() - t0
    extras["fits"] = {"fits_per_s": n_fits * world / fit_s, "config": "C2: synthetic multi-H, 10k correspondences, 5 planted "
                      "planes + 40% outliers, findHomographies(max_iters=1000, conf=0.5, lambda=0), one problem per GPU "
                      "at a time", "models_found": n_models}

    # same problem with the spatial coherence term (AdelaideH's lambda): GC-RANSAC LO cuts + alpha-expansion on the GPU max-flow
    lam_kwargs = dict(fit_kwargs, spatial_coherence_weight=0.05)
    pyprogressivex.findHomographies(c2_pts, 1024, 768, 1024, 768, seed=1, **lam_kwargs)
    barrier()
    t0 = time.perf_counter()
    for i in range(2):
        pyprogressivex.findHomographies(c2_pts, 1024, 768, 1024, 768, seed=2 + i, **lam_kwargs)
    barrier()
    extras["fits_lambda"] = {"fits_per_s": 2 * world / (time.perf_counter() - t0),
                             "config": "C2 with spatial_coherence_weight=0.05 (kNN graph, LO st-cuts, alpha-expansion)"}

    # ---- config C4 in miniature: independent pairs solved concurrently (8 host threads, one context each) -----------
    n_pairs, n_pts = 32, 5_000
    c4 = [syn.multi_homography_scene(n_pts, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=700 + 97 * rank + p)[0]
          for p in range(n_pairs)]
    c4_kwargs = dict(fit_kwargs, minimum_point_number=60, seed=11)
    pyprogressivex.findHomographiesBatch(c4[:8], 1024, 768, 1024, 768, workers=8, **c4_kwargs)  # warm-up (contexts, JIT-free)
    barrier()
    t0 = time.perf_counter()
    res = pyprogressivex.findHomographiesBatch(c4, 1024, 768, 1024, 768, workers=8, **c4_kwargs)
    barrier()
    c4_s = time.perf_counter() - t0
    extras["fits_batch"] = {"fits_per_s": n_pairs * world / c4_s, "config": f"C4 in miniature: {n_pairs} independent pairs x "
                            f"{n_pts} correspondences per GPU (4 planes + 40% outliers), findHomographiesBatch with 8 host "
                            "threads / contexts per GPU, lambda=0", "models_found_mean": float(np.mean([m.shape[0] // 3 for m, _ in res]))}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N, K),
                       "cache": "4.06 GB written per step with streaming stores: output far exceeds the 126 MB L2; "
                                "inputs (2.3 MB) are meant to stay in L2",
                       "sharding": "points replicated, disjoint hypothesis block per GPU, no data-path collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        out.update(extras)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(pts, models_np)
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
