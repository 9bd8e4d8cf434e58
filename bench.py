#!/usr/bin/env python
"""bench.py -- the Progressive-X hot path on B200 (BASELINE.json metric: residual evals/s and multi-model fits/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--quick]

Headline ("value", "roofline"): one step = one pass of the residual-and-inlier-matrix kernel (exact float64 r2 + 1 mask
bit per entry) over the grid of SURVEY.md 8(d): N = 50 000 correspondences x K = 10 000 four-point homography hypotheses,
inputs resident in HBM. With --gpus N every rank owns a disjoint block of 10 000 hypotheses over the same (replicated)
points -- weak scaling, no collective on the data path.

The same JSON line carries every BASELINE config as specified, each with a parity gate:
  "matrix_F" / "matrix_PnP"  C3 / C5 residual-matrix roofline lines (50k x <=30k Sampson, 100k x <=40k reprojection)
  "score_kernel"             the fused getScore batch the RANSAC loop runs, with its issue-slot roofline
  "fits_c2"                  C2: findHomographies, 10k correspondences, Python defaults + max_iters=1000, lambda in {0, 0.05},
                             beside the sequential CPU oracle's fits/s on the same problem (rank 0, N = 1)
  "labeling_cpu_vs_gpu"      PEARL label sweep and LO st-cut: reference gco/BK build (oracle/_ref) vs the GPU engine
  "batch_c4"                 C4: 256 independent pairs x 5k correspondences sharded over the ranks, instances merged by one
                             ncclAllGather (pxb_allgather_instances); gate: gathered result == single-rank result
  "pose_c5"                  C5: find6DPoses on 100k 2D-3D matches, 10 objects, hypothesis blocks sharded over the ranks
                             (pxb_ctx_set_shard); gate: labels and poses bit-identical to the 1-GPU run with the same seed
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

# NCCL prints its version banner to stdout when NCCL_DEBUG is VERSION/WARN/INFO: keep stdout for the one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

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


def committed(key):
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(key)
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


# ---- CPU baselines (the oracle as the thing timed: allowed here and nowhere else) ------------------------------------------
def cpu_baseline(pts, models, target_seconds=10.0):
    """The oracle's restated getScore loops (kind 'port') on all host cores, bounded sample of the same workload; plus the
    single-thread figure (the reference itself is single-threaded)."""
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
    k1 = max(8, min(models.shape[0], k // max(cores, 1)))
    O.score_batch(0, pts, models[:k1], T2, None, threads=1)
    t0 = time.perf_counter()
    O.score_batch(0, pts, models[:k1], T2, None, threads=1)
    dt1 = max(time.perf_counter() - t0, 1e-6)
    return {"value": pts.shape[0] * k * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": pts.shape[0] * k1 / dt1,
            "sample": f"oracle pxo_score_batch (restated getScore loop, OpenMP over hypotheses): {reps} passes over "
                      f"N={pts.shape[0]} x K={k} of the {models.shape[0]} bench hypotheses, {dt:.2f} s wall; single thread: "
                      f"one pass over K={k1}"}


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


C2_KW = dict(threshold=4.0, conf=0.5, neighborhood_ball_radius=200.0, maximum_tanimoto_similarity=0.4, max_iters=1000,
             minimum_point_number=10, maximum_model_number=-1, sampler_id=3, scoring_exponent=2)  # bindings.cpp:410-426


def same_result(a, b):
    return (a[0].shape == b[0].shape and np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
            and np.array_equal(a[1], b[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pxb200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + score kernel only (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import pyprogressivex
    from pyprogressivex import _native, sharding
    from pyprogressivex import synthetic as syn

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
    pyprogressivex._contexts[local_rank] = ctx  # the find* calls below reuse this process' context
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

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
        total_ms = max_over_ranks(evs[0].elapsed_time(evs[steps]))
        return total_ms, per, ctx.launch_count() - launches0, clocks

    def check(rc):
        if rc != 0:
            raise RuntimeError(lib.pxb_last_error().decode())

    # ---- headline workload: same points everywhere, a disjoint hypothesis block per rank ------------------------
    pts, gt, _ = workload(seed=0)
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
    stream.synchronize()

    def step_matrix():
        check(lib.pxb_residual_matrix_dev(ctx.handle, models.data_ptr(), K, T2, r2.data_ptr(), mask.data_ptr()))

    total_ms, per_launch_ms, launches, clocks = timed(step_matrix, args.steps, args.warmup, sample_clocks=True)
    if clocks and any(r in clocks["reasons"] for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")):
        total_ms, per_launch_ms, launches, clocks = timed(step_matrix, args.steps, args.warmup, sample_clocks=True)
        clocks["remeasured"] = True
    evals_per_step = N * K
    value = evals_per_step * world * args.steps / (total_ms * 1e-3)
    kernel_ms = float(np.mean(per_launch_ms))
    achieved = evals_per_step * BYTES_PER_EVAL / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": committed("k_residual_matrix_f64_bytes_per_launch"), "kernel": "k_residual_matrix<H,f64>",
                "algorithmic_bytes_per_launch": evals_per_step * BYTES_PER_EVAL, "launch_ms": kernel_ms,
                "peak_source": peak_src + "; burst figure (kernel timed alone)"}
    extras = {}
    sc_steps = max(3, args.steps // 2)

    # ---- fused score kernel (what the RANSAC loop and the e2e path run): issue-slot roofline ------------------------
    def step_score():
        check(lib.pxb_score_compound_dev(ctx.handle, models.data_ptr(), K, T2, None, cnt.data_ptr(), val.data_ptr(),
                                         shr.data_ptr()))

    sc_ms, sc_per, _, sc_clk = timed(step_score, sc_steps, 3, sample_clocks=True)
    sc_launch_ms = float(np.mean(sc_per))
    inst = committed("k_score_screened_warp_instructions_per_launch")
    sm_mhz = (sc_clk or {}).get("sm_mhz") or (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6  # one warp instruction per cycle per SM sub-partition
    extras["score_kernel"] = {
        "evals_per_s": evals_per_step * world * sc_steps / (sc_ms * 1e-3), "launch_ms": sc_launch_ms,
        "roofline": None if not inst else {
            "bound": "issue", "achieved": inst / (sc_launch_ms * 1e-3), "peak": issue_peak, "unit": "warp-instructions/s",
            "frac": inst / (sc_launch_ms * 1e-3) / issue_peak,
            "note": "warp instructions per launch from the committed ncu capture (profiles/roofline_traffic.json) / measured "
                    "launch time, against 148 SMs x 4 sub-partitions x 1 warp instruction per cycle at the sampled SM clock"},
        "note": "k_screen_prepare + k_score_screened + k_score_finalize: count / sum score / shared support, no matrix "
                "written; float32 screening proves outliers, exact float64 only for queued candidates (issue bound)"}

    if not args.quick:
        # ---- inlier bit matrix only (float32-screened kernel) -------------------------------------------------------
        del r2

        def step_mask():
            check(lib.pxb_residual_matrix_dev(ctx.handle, models.data_ptr(), K, T2, None, mask.data_ptr()))

        mk_ms, mk_per, _, _ = timed(step_mask, sc_steps, 3)
        extras["mask_only"] = {"evals_per_s": evals_per_step * world * sc_steps / (mk_ms * 1e-3),
                               "launch_ms": float(np.mean(mk_per)), "bytes_per_eval": 0.125,
                               "note": "inlier bit matrix only (pxb_residual_matrix_dev with r2 = NULL): float32-screened kernel, "
                                       "exact float64 only for the pairs screening cannot dismiss; bit-identical masks. Replaces "
                                       "round 1's float32-r2 variant (removed: it ran at the float64 kernel's speed)"}

        # ---- C3 / C5 residual-matrix roofline lines --------------------------------------------------------------------
        def matrix_line(model_type, rows, sample_rows, t2, label, traffic_key):
            c = _native.Context(local_rank)
            try:
                c.upload_points(model_type, rows)
                m, nn, _, _ = c.solve_minimal(sample_rows)
                flat = np.ascontiguousarray(m.reshape(-1, m.shape[-1])[(np.arange(m.shape[1])[None, :] < nn[:, None]).reshape(-1)])
                Kx, Nx = flat.shape[0], rows.shape[0]
                st = torch.cuda.ExternalStream(c.stream, device=dev)
                with torch.cuda.stream(st):
                    md = torch.from_numpy(flat).to(dev)
                    out = torch.empty((Kx, Nx), dtype=torch.float64, device=dev)
                    mk = torch.empty((Kx, (Nx + 31) // 32), dtype=torch.int32, device=dev)
                st.synchronize()

                def fn():
                    check(c.lib.pxb_residual_matrix_dev(c.handle, md.data_ptr(), Kx, t2, out.data_ptr(), mk.data_ptr()))

                for _ in range(3):
                    fn()
                barrier()
                reps = max(3, min(10, args.steps))
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
                with torch.cuda.stream(st):
                    evs[0].record(st)
                    for i in range(reps):
                        fn()
                        evs[i + 1].record(st)
                barrier()
                ms = float(np.mean([evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]))
                tot = max_over_ranks(evs[0].elapsed_time(evs[reps]))
                inl = int(np.unpackbits(mk[: min(Kx, 64)].cpu().numpy().view(np.uint8)).sum())
                ach = Nx * Kx * BYTES_PER_EVAL / (ms * 1e-3) / 1e9
                del out, mk, md
                return {"workload": label, "N": Nx, "K": Kx, "launch_ms": ms, "evals_per_s": Nx * Kx * world * reps / (tot * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                     "traffic": (committed(traffic_key) * Nx * Kx) if committed(traffic_key) else None,
                                     "algorithmic_bytes_per_launch": Nx * Kx * BYTES_PER_EVAL},
                        "inliers_in_first_64_hypotheses": inl}
            finally:
                c.close()
                torch.cuda.empty_cache()

        f_pts, f_gt, _ = syn.multi_motion_scene(50_000, seed=0)
        f_s = syn.minimal_samples(f_gt, 10_000, 7, within_ratio=0.5, seed=2000 + rank)
        extras["matrix_F"] = matrix_line(_native.MODEL_F, f_pts, f_s, (1.5 * 0.75) ** 2,
                                         "C3: synthetic multi-F, N=50 000 correspondences (3 motions 25/25/20% + 30% outliers), "
                                         "hypotheses = all solutions of 10 000 seven-point samples, thr=0.75 px; Sampson r2 f64 + bit",
                                         "k_residual_matrix_F_traffic_bytes_per_eval")
        p_img, p_w, p_K, p_gt, _ = syn.multi_pose_scene(100_000, n_objects=10, inlier_ratio_each=0.06, noise_px=1.0, seed=0)
        p_rows = syn.normalize_pnp_points(p_img, p_w, p_K)
        p_s = syn.minimal_samples(p_gt, 10_000, 3, within_ratio=0.5, seed=3000 + rank)
        p_thr = 4.0 / (0.5 * (p_K[0, 0] + p_K[1, 1]))
        extras["matrix_PnP"] = matrix_line(_native.MODEL_PNP, p_rows, p_s, (1.5 * p_thr) ** 2,
                                           "C5: synthetic multi-pose, N=100 000 2D-3D matches (10 objects x 6% + 40% outliers), "
                                           "hypotheses = all poses of 10 000 P3P samples, thr=4 px / f; reprojection r2 f64 + bit",
                                           "k_residual_matrix_PnP_traffic_bytes_per_eval")
        ctx.upload_points(_native.MODEL_H, pts)

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
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": evals_per_step * world * e2e_steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(pts.nbytes + models_np.nbytes + 72),
           "d2h_bytes_per_step": int(K * 24 + words * 4),
           "operator": "fused compound score (the gcransac::ScoringFunction::getScore seam): count / score / shared support per "
                       "hypothesis + the winner's inlier list; the N x K matrix is NOT materialised on this path (`value` and "
                       "`roofline` time the matrix-writing kernel; see e2e_matrix for that operator through host buffers)",
           "call": "pxb_upload_points + pxb_score_compound + pxb_inliers (host pointers, copies inside the timed region)"}

    if not args.quick:
        # the matrix operator through host buffers: inlier bit matrix of every hypothesis back on the host
        host_mask = torch.empty((K, words), dtype=torch.int32).pin_memory()

        def step_e2e_matrix():
            _native._check(lib.pxb_upload_points(ctx.handle, _native.MODEL_H, pin_pts.data_ptr(), N))
            _native._check(lib.pxb_residual_matrix(ctx.handle, pin_models.data_ptr(), K, T2, None, host_mask.data_ptr()))

        for _ in range(2):
            step_e2e_matrix()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            step_e2e_matrix()
        barrier()
        em_s = max_over_ranks(time.perf_counter() - t0)
        extras["e2e_matrix"] = {"value": evals_per_step * world * 3 / em_s, "unit": UNIT,
                                "h2d_bytes_per_step": int(pts.nbytes + models_np.nbytes),
                                "d2h_bytes_per_step": int(K * words * 4),
                                "call": "pxb_upload_points + pxb_residual_matrix(r2 = NULL, mask -> pinned host): the exact inlier "
                                        "bit matrix of all K hypotheses returned to the host (the f64 r2 matrix, 4 GB, is PCIe "
                                        "bound at ~80 ms and is kept on the device by design)"}

        # ---- C2: complete multi-model fits per second, SURVEY 8(d) parameters ----------------------------------------
        c2_pts, c2_gt, c2_H = syn.multi_homography_scene(10_000, n_planes=5, outlier_ratio=0.4, noise=0.5, seed=42 + rank)
        fits = {"config": "C2: synthetic multi-H, 10k correspondences, 5 planted planes + 40% outliers; findHomographies with the "
                          "Python defaults (threshold 4, conf 0.5, NAPSAC sampler, min 10 points) and max_iters=1000; one problem "
                          "per GPU at a time"}
        for lam, tag, n_fits in ((0.0, "lambda0", 8), (0.05, "lambda0.05", 3)):
            kw = dict(C2_KW, spatial_coherence_weight=lam, device=local_rank)
            for warm in (1, 101, 102):  # first sight of every device chain, its capture into a CUDA graph, one replayed run
                pyprogressivex.findHomographies(c2_pts, 1024, 768, 1024, 768, seed=warm, **kw)
            barrier()
            t0 = time.perf_counter()
            found = []
            for i in range(n_fits):
                m_, l_ = pyprogressivex.findHomographies(c2_pts, 1024, 768, 1024, 768, seed=2 + i, **kw)
                found.append(m_.shape[0] // 3)
            barrier()
            s_ = max_over_ranks(time.perf_counter() - t0)
            fits[tag] = {"fits_per_s": n_fits * world / s_, "ms_per_fit": s_ / n_fits * 1e3, "models_found": found}
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            from oracle import px_sequential as S
            g = syn.knn_graph(c2_pts, 200.0, 5)
            for lam, tag in ((0.0, "lambda0"), (0.05, "lambda0.05")):
                t0 = time.perf_counter()
                m_, l_ = S.find_homographies(c2_pts, 4.0, 0.5, lam, 0.4, 1000, 10, -1, 3, 2, seed=2, graph=g)
                dt = time.perf_counter() - t0
                fits[tag]["cpu_baseline"] = {
                    "value": 1.0 / dt, "unit": "fits/s", "cores": 1, "kind": "port", "models_found": int(m_.shape[0]),
                    "sample": "one fit of the same problem by oracle/px_sequential.py (the reference's sequential control flow on "
                              "the C oracle operators and the reference's own gco/BK build; neighbourhood graph excluded)"}
        extras["fits_c2"] = fits

        # ---- PEARL label sweep and LO st-cut: reference gco/BK build vs the GPU engine (rank 0, N = 1) -----------------
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            from oracle import oracle as O
            if O.have_gco_ref():
                ctx.upload_points(_native.MODEL_H, c2_pts)
                off, idx = syn.knn_graph(c2_pts, 200.0, 5)
                Hs = c2_H.reshape(-1, 9)
                lab_cmp = {}
                for lam in (0.0, 0.05):
                    D = ctx.pearl_datacost(Hs, 2.0, lam)
                    a = (off, idx) if lam > 0 else (None, None)
                    ctx.pearl_label(D, lam, 10.0, *a)
                    t0 = time.perf_counter()
                    for _ in range(5):
                        g_lab, g_e = ctx.pearl_label(D, lam, 10.0, *a)
                    g_ms = (time.perf_counter() - t0) / 5 * 1e3
                    O.gco_pearl_label(D, lam, 10.0, *a)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        c_lab, c_e, _ = O.gco_pearl_label(D, lam, 10.0, *a)
                    c_ms = (time.perf_counter() - t0) / 3 * 1e3
                    lab_cmp[f"pearl_label_lambda{lam}"] = {"gpu_ms": g_ms, "cpu_ms": c_ms, "labels_equal": bool(np.array_equal(g_lab, c_lab)),
                                                           "host_io": "D [N, 6] f64 up, labels down, inside both timings"}
                d_, e0_, e1_ = ctx.lo_unary_terms(Hs[0], 2.0, 0.05)
                ctx.lo_labeling(Hs[0], 2.0, 0.05, off, idx)
                t0 = time.perf_counter()
                for _ in range(10):
                    g_in = ctx.lo_labeling(Hs[0], 2.0, 0.05, off, idx)
                g_ms = (time.perf_counter() - t0) / 10 * 1e3
                t0 = time.perf_counter()
                for _ in range(3):
                    c_in, _ = O.gco_lo_labeling(e0_, e1_, d_, 0.05, off, idx)
                c_ms = (time.perf_counter() - t0) / 3 * 1e3
                lab_cmp["lo_graph_cut_lambda0.05"] = {"gpu_ms": g_ms, "cpu_ms": c_ms, "labels_equal": bool(np.array_equal(g_in, c_in))}
                lab_cmp["note"] = ("C2 scene (N = 10 000, 5 ground-truth homographies as instances, 5-nearest graph): cpu = the "
                                   "reference's GCoptimization.cpp / maxflow.cpp compiled unchanged (oracle/_ref/libgco_ref.so, one "
                                   "thread) driven like PEARL::labeling (GCoptimization.cpp:1003-1086) and GCRANSAC::labeling "
                                   "(GCRANSAC.h:914-1022); gpu = pxb_pearl_label / pxb_lo_labeling through host pointers")
                extras["labeling_cpu_vs_gpu"] = lab_cmp
                ctx.upload_points(_native.MODEL_H, pts)

        # ---- C4 as specified: 256 independent pairs x 5k, pairs sharded over the ranks, NCCL gather of the instances ------
        n_pairs, n_pts = 256, 5_000
        c4 = [syn.multi_homography_scene(n_pts, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=700 + p)[0] for p in range(n_pairs)]
        c4_kw = dict(C2_KW, spatial_coherence_weight=0.0, sampler_id=0, seed=11, device=local_rank)
        workers = int(os.environ.get("PXB_BENCH_WORKERS", "2"))
        in_flight = int(os.environ.get("PXB_BENCH_IN_FLIGHT", "8"))
        if world == 1:
            pyprogressivex._shards[local_rank] = sharding.NcclShard(ctx, world=1, rank=0)
        pyprogressivex.findHomographiesBatch(c4[:3 * world * workers * in_flight], 1024, 768, 1024, 768, workers=workers, in_flight=in_flight,
                                             distributed=True, **c4_kw)
        barrier()
        t0 = time.perf_counter()
        res = pyprogressivex.findHomographiesBatch(c4, 1024, 768, 1024, 768, workers=workers, in_flight=in_flight, distributed=True, **c4_kw)
        barrier()
        c4_s = max_over_ranks(time.perf_counter() - t0)
        # gate: the gathered result of every pair equals the single-rank result (every rank re-solves a stripe locally)
        stripe = list(range(rank, n_pairs, max(world, 4)))
        local = pyprogressivex.findHomographiesBatch([c4[p] for p in stripe], 1024, 768, 1024, 768, workers=workers, in_flight=in_flight, **c4_kw)
        bad = [p for p, r_ in zip(stripe, local) if not same_result(r_, res[p])]
        if bad:
            raise AssertionError(f"C4 parity gate: gathered result differs from the single-rank result for pairs {bad[:8]}")
        extras["batch_c4"] = {"fits_per_s": n_pairs / c4_s, "seconds": c4_s, "pairs": n_pairs, "points_per_pair": n_pts,
                              "host_threads_per_gpu": workers, "problems_in_flight_per_thread": in_flight, "models_found_mean": float(np.mean([m.shape[0] // 3 for m, _ in res])),
                              "parity_gate": f"gathered == single-rank result on {len(stripe)} pairs per rank: ok",
                              "config": "C4: 256 independent pairs x 5 000 correspondences (4 planes + 40% outliers), pair p on rank "
                                        "p mod N, findHomographies(Python defaults, uniform sampler, lambda=0), instances merged by one "
                                        "ncclAllGather (pxb_allgather_instances); strong scaling over N"}

        # weak-scaling companion: 64 pairs PER GPU (different scenes per rank), same call, no exchange step
        weak = [syn.multi_homography_scene(n_pts, n_planes=4, outlier_ratio=0.4, noise=0.5, seed=5000 + 64 * rank + p)[0]
                for p in range(64)]
        pyprogressivex.findHomographiesBatch(weak[:3 * workers * in_flight], 1024, 768, 1024, 768, workers=workers, in_flight=in_flight, **c4_kw)
        weak_runs = []
        for _ in range(3):  # 64 pairs take ~0.1 s: one host-scheduling hiccup is a third of that, so the median of three runs
            barrier()
            t0 = time.perf_counter()
            pyprogressivex.findHomographiesBatch(weak, 1024, 768, 1024, 768, workers=workers, in_flight=in_flight, **c4_kw)
            barrier()
            weak_runs.append(max_over_ranks(time.perf_counter() - t0))
        weak_s = sorted(weak_runs)[1]
        extras["batch_c4_weak"] = {"fits_per_s": 64 * world / weak_s, "pairs_per_gpu": 64, "points_per_pair": n_pts,
                                   "seconds_of_three_runs": [round(x, 4) for x in weak_runs],
                                   "host_threads_per_gpu": workers, "problems_in_flight_per_thread": in_flight,
                                   "note": "same call as batch_c4 with a fixed load per GPU (weak scaling, no exchange): "
                                           "efficiency at N GPUs = fits_per_s(N) / (N x fits_per_s(1))"}

        # ---- C5 as specified: one 100k-match 6D-pose problem, hypothesis blocks sharded over the ranks -------------------
        c5_kw = dict(threshold=4.0, conf=0.9, spatial_coherence_weight=0.0, neighborhood_ball_radius=20.0,
                     maximum_tanimoto_similarity=0.9, max_iters=5000, minimum_point_number=1000, maximum_model_number=-1,
                     device=local_rank)
        for _ in range(2):
            local_pose = pyprogressivex.find6DPoses(p_img, p_w, p_K, seed=3, **c5_kw)
        barrier()
        t0 = time.perf_counter()
        local_pose = pyprogressivex.find6DPoses(p_img, p_w, p_K, seed=3, **c5_kw)
        t_local = time.perf_counter() - t0
        barrier()
        with pyprogressivex.distributed(local_rank):
            sharded_pose = pyprogressivex.find6DPoses(p_img, p_w, p_K, seed=3, **c5_kw)
            barrier()
            t0 = time.perf_counter()
            sharded_pose = pyprogressivex.find6DPoses(p_img, p_w, p_K, seed=3, **c5_kw)
            barrier()
            t_shard = max_over_ranks(time.perf_counter() - t0)
        if not same_result(local_pose, sharded_pose):
            raise AssertionError("C5 parity gate: the sharded find6DPoses differs from the 1-GPU run with the same seed")
        lab = local_pose[1]
        purity = []
        for k in range(local_pose[0].shape[0] // 3):
            own = p_gt[lab == k]
            own = own[own >= 0]
            purity.append(float(np.bincount(own).max() / max(1, (lab == k).sum())) if own.size else 0.0)
        extras["pose_c5"] = {"fits_per_s_sharded": 1.0 / t_shard, "ms_sharded": t_shard * 1e3, "ms_one_gpu": t_local * 1e3,
                             "poses_found": int(local_pose[0].shape[0] // 3), "label_purity_mean": float(np.mean(purity)) if purity else None,
                             "parity_gate": "poses and per-point labels of the sharded run bit-identical to the 1-GPU run (same seed) on every rank: ok",
                             "config": "C5: 100 000 2D-3D matches, 10 objects x 6% + 40% outliers, find6DPoses(conf 0.9, max_iters=5000, "
                                       "min 1000 points, lambda=0); blocks of 512 x N minimal P3P samples, slice r solved and scored on "
                                       "rank r, one ncclAllGather per block (pxb_ctx_set_shard); strong scaling over N"}

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
    for sh in list(pyprogressivex._shards.values()):
        sh.close()
    pyprogressivex._shards.clear()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
