"""ctypes binding of libpxb200.so (the C ABI declared in include/pxb200.h).

This is the only place Python touches the native library. There is no CPU fallback: if the shared library is
missing, or no sm_100 device is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import numpy as np

_PKG_ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = _PKG_ROOT / "libpxb200.so"

MODEL_H, MODEL_F, MODEL_PNP, MODEL_VP, MODEL_LINE = 0, 1, 2, 3, 4
POINT_DIM = {MODEL_H: 4, MODEL_F: 4, MODEL_PNP: 5, MODEL_VP: 4, MODEL_LINE: 2}
MODEL_SIZE = {MODEL_H: 9, MODEL_F: 9, MODEL_PNP: 12, MODEL_VP: 3, MODEL_LINE: 3}
SAMPLE_SIZE = {MODEL_H: 4, MODEL_F: 7, MODEL_PNP: 3, MODEL_VP: 2, MODEL_LINE: 2}
MAX_SOLUTIONS = {MODEL_H: 1, MODEL_F: 3, MODEL_PNP: 4, MODEL_VP: 1, MODEL_LINE: 1}


class MultiModelSettings(C.Structure):
    """pxb_multi_model_settings (include/pxb200.h) = progx::MultiModelSettings + the proposal engine's settings."""
    _fields_ = [("minimum_number_of_inliers", C.c_size_t), ("max_proposal_number_without_change", C.c_size_t),
                ("cell_number_in_neighborhood_graph", C.c_size_t), ("maximum_model_number", C.c_size_t),
                ("maximum_tanimoto_similarity", C.c_double), ("confidence", C.c_double),
                ("inlier_outlier_threshold", C.c_double), ("spatial_coherence_weight", C.c_double),
                ("max_iteration_number", C.c_size_t), ("min_iteration_number", C.c_size_t),
                ("min_iteration_number_before_lo", C.c_size_t), ("max_local_optimization_number", C.c_size_t),
                ("max_graph_cut_number", C.c_size_t), ("max_least_squares_iterations", C.c_size_t),
                ("max_unsuccessful_model_generations", C.c_size_t), ("scoring_exponent", C.c_int)]


class IterationStatistics(C.Structure):
    _fields_ = [("time_of_proposal_engine", C.c_double), ("time_of_model_validation", C.c_double),
                ("time_of_optimization", C.c_double), ("time_of_compound_model_update", C.c_double),
                ("number_of_instances", C.c_size_t), ("ransac_iteration_number", C.c_size_t),
                ("local_optimization_number", C.c_size_t), ("graph_cut_number", C.c_size_t),
                ("proposal_inlier_number", C.c_size_t)]


class MultiModelStatistics(C.Structure):
    """pxb_multi_model_statistics = progx::MultiModelStatistics of the last find* call on a context."""
    _fields_ = [("processing_time", C.c_double), ("total_time_of_proposal_engine", C.c_double),
                ("total_time_of_model_validation", C.c_double), ("total_time_of_optimization", C.c_double),
                ("total_time_of_compound_model_calculation", C.c_double), ("iteration_statistics_size", C.c_size_t),
                ("iteration_statistics", IterationStatistics * 10), ("model_number", C.c_size_t),
                ("inliers_of_each_model_size", C.c_size_t), ("kernel_launches", C.c_size_t)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "iteration_statistics"}
        d["iteration_statistics"] = [{k: getattr(it, k) for k, _ in IterationStatistics._fields_}
                                     for it in self.iteration_statistics[: self.iteration_statistics_size]]
        return d


class PxbError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libpxb200 error {status}: {message}")
        self.status = status


_lib = None


def _p(dtype):
    return np.ctypeslib.ndpointer(dtype=dtype, flags="C_CONTIGUOUS")


def load_library() -> C.CDLL:
    """Load libpxb200.so and declare every prototype of include/pxb200.h. Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    vp, i64, i32, f64, sz, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_size_t, C.c_uint64
    lib.pxb_last_error.restype = C.c_char_p
    lib.pxb_version.restype = C.c_char_p
    lib.pxb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.pxb_ctx_destroy.argtypes = [vp]
    lib.pxb_ctx_destroy.restype = None
    lib.pxb_ctx_stream.argtypes = [vp]
    lib.pxb_ctx_stream.restype = vp
    lib.pxb_sync.argtypes = [vp]
    lib.pxb_launch_count.argtypes = [vp]
    lib.pxb_launch_count.restype = i64
    lib.pxb_dev_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.pxb_dev_free.argtypes = [vp, vp]
    lib.pxb_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    lib.pxb_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    lib.pxb_host_alloc_pinned.argtypes = [sz, C.POINTER(vp)]
    lib.pxb_host_free_pinned.argtypes = [vp]
    lib.pxb_upload_points.argtypes = [vp, C.c_int, vp, i64]
    lib.pxb_point_count.argtypes = [vp]
    lib.pxb_point_count.restype = i64
    lib.pxb_residual_matrix.argtypes = [vp, vp, i64, f64, vp, vp]
    lib.pxb_residual_matrix_dev.argtypes = [vp, vp, i64, f64, vp, vp]
    lib.pxb_score_compound.argtypes = [vp, vp, i64, f64, vp, vp, vp, vp]
    lib.pxb_score_compound_dev.argtypes = [vp, vp, i64, f64, vp, vp, vp, vp]
    lib.pxb_inliers.argtypes = [vp, vp, f64, vp, C.POINTER(i64)]
    lib.pxb_preference_vector.argtypes = [vp, vp, f64, vp]
    lib.pxb_tanimoto.argtypes = [vp, vp, vp, i64, C.POINTER(f64)]
    lib.pxb_compound_max.argtypes = [vp, vp, i64, i64, vp]
    lib.pxb_solve_minimal.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.pxb_solve_plane_parallax.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.pxb_h_degenerate_sample.argtypes = [vp, vp, vp, vp, vp]
    lib.pxb_pearl_datacost.argtypes = [vp, vp, i64, f64, f64, vp]
    lib.pxb_pearl_label.argtypes = [vp, vp, i64, i32, f64, f64, vp, vp, vp, vp, C.POINTER(f64)]
    lib.pxb_segment_residual_sums.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.pxb_lo_unary_terms.argtypes = [vp, vp, f64, f64, vp, vp, vp]
    lib.pxb_tukey_weights.argtypes = [vp, vp, f64, vp]
    lib.pxb_selftest_division.argtypes = [vp, u64, i64, C.c_int, C.POINTER(i64)]
    lib.pxb_lo_graph_cut.argtypes = [vp, vp, vp, vp, i64, f64, vp, vp, vp]
    lib.pxb_lo_labeling.argtypes = [vp, vp, f64, f64, vp, vp, vp]
    lib.pxb_knn_graph.argtypes = [vp, f64, C.c_int, vp, vp]
    lib.pxb_fit_homographies.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.pxb_fit_nonminimal.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.pxb_find_homographies.argtypes = [vp, vp, i64, vp, vp, i64, sz, sz, sz, sz, f64, f64, f64, f64, f64, sz, sz,
                                          C.c_int, sz, f64, C.c_int, u64]
    lib.pxb_find_two_view_motions.argtypes = lib.pxb_find_homographies.argtypes
    lib.pxb_find_6d_poses.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, f64, f64, f64, f64, f64, sz, sz, C.c_int, u64]
    lib.pxb_find_vanishing_points.argtypes = [vp, vp, vp, i64, vp, vp, i64, sz, sz, f64, f64, f64, f64, f64, sz, sz,
                                              C.c_int, sz, f64, C.c_int, u64]
    lib.pxb_find_lines.argtypes = lib.pxb_find_vanishing_points.argtypes
    lib.pxb_find_homographies_batch.argtypes = [C.c_int, i64, vp, vp, vp, vp, i64, vp, sz, sz, sz, sz, f64, f64, f64, f64, f64,
                                                sz, sz, C.c_int, sz, f64, u64, C.c_int, C.c_int, C.c_int]
    lib.pxb_batch_release.argtypes = []
    lib.pxb_batch_release.restype = None
    lib.pxb_settings_default.argtypes = [C.POINTER(MultiModelSettings)]
    lib.pxb_ctx_set_settings.argtypes = [vp, C.POINTER(MultiModelSettings)]
    lib.pxb_ctx_get_statistics.argtypes = [vp, C.POINTER(MultiModelStatistics)]
    lib.pxb_nccl_version.argtypes = [C.POINTER(C.c_int)]
    lib.pxb_nccl_unique_id.argtypes = [vp]
    lib.pxb_nccl_comm_init.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pxb_nccl_comm_destroy.argtypes = [vp]
    lib.pxb_ctx_set_shard.argtypes = [vp, vp]
    lib.pxb_shard_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pxb_allgather_instances.argtypes = [vp, vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


_nccl_preloaded = False


def preload_nccl() -> None:
    """libpxb200.so resolves NCCL at run time and takes the copy of libnccl.so.2 the process already holds. A Python
    process that has not imported torch yet holds none -- and mapping the SYSTEM libnccl first would later break
    `import torch` (torch's libtorch_cuda.so needs the newer copy it ships with: the loader reuses whatever already
    carries the soname). So map the wheel-bundled copy (nvidia/nccl/lib) first when there is one."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    import sys
    if "torch" in sys.modules:
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec is not None and spec.submodule_search_locations:
            for base in spec.submodule_search_locations:
                cand = Path(base) / "lib" / "libnccl.so.2"
                if cand.exists():
                    C.CDLL(str(cand), mode=C.RTLD_GLOBAL)
                    return
    except Exception:
        pass


def _check(rc: int):
    if rc < 0:
        raise PxbError(rc, load_library().pxb_last_error().decode("utf-8", "replace"))
    return rc


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceBuffer:
    """A raw device allocation owned by a Context (used by bench.py to keep the residual matrix resident)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        _check(ctx.lib.pxb_dev_alloc(ctx.handle, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        _check(self.ctx.lib.pxb_memcpy_h2d(self.ctx.handle, self.ptr, _ptr(arr), arr.nbytes))

    def download(self, dtype, count: int, offset_bytes: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        _check(self.ctx.lib.pxb_memcpy_d2h(self.ctx.handle, _ptr(out), self.ptr + offset_bytes, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.pxb_dev_free(self.ctx.handle, self.ptr)
            self.ptr = None


class Context:
    """One pxb_ctx: a device, a stream, the resident point set and scratch buffers."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.pxb_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.model_type = None
        self.N = 0
        # A pxb_ctx is NOT re-entrant (one stream, one resident point set, shared scratch and staging buffers) and ctypes
        # releases the GIL for the whole C call -- unlike the reference's pybind11 module, which holds it. The find* entry
        # points take this lock around the native call so that Python threads sharing a context are serialised.
        self.lock = threading.RLock()

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pxb_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ----------------------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        return self.lib.pxb_ctx_stream(self.handle)

    def sync(self):
        _check(self.lib.pxb_sync(self.handle))

    def launch_count(self) -> int:
        return int(self.lib.pxb_launch_count(self.handle))

    # -- ProgressiveX::getMutableSettings / getStatistics (progressive_x.h:210-217) -----------------------
    def default_settings(self) -> MultiModelSettings:
        s = MultiModelSettings()
        _check(self.lib.pxb_settings_default(C.byref(s)))
        return s

    def set_settings(self, settings) -> None:
        """Installs the engine settings no find* argument carries (None restores the defaults)."""
        _check(self.lib.pxb_ctx_set_settings(self.handle, None if settings is None else C.byref(settings)))

    def statistics(self) -> dict:
        """Statistics of the last find* call on this context (per-round times from CUDA events, in seconds)."""
        st = MultiModelStatistics()
        _check(self.lib.pxb_ctx_get_statistics(self.handle, C.byref(st)))
        return st.as_dict()

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    # -- data --------------------------------------------------------------------------------------------
    def upload_points(self, model_type: int, pts) -> None:
        pts = _f64(pts)
        if pts.ndim != 2 or pts.shape[1] != POINT_DIM[model_type]:
            raise ValueError(f"points must be [N, {POINT_DIM[model_type]}]")
        _check(self.lib.pxb_upload_points(self.handle, model_type, _ptr(pts), pts.shape[0]))
        self.model_type, self.N = model_type, pts.shape[0]

    def _models(self, models):
        ms = MODEL_SIZE[self.model_type]
        m = _f64(models).reshape(-1, ms)
        return m, m.shape[0]

    # -- a1/a2/a3 ------------------------------------------------------------------------------------------
    def residual_matrix(self, models, T2: float, want_r2=True, want_mask=True):
        m, K = self._models(models)
        words = (self.N + 31) // 32
        r2 = np.empty((K, self.N), dtype=np.float64) if want_r2 else None
        mask = np.empty((K, words), dtype=np.uint32) if want_mask else None
        _check(self.lib.pxb_residual_matrix(self.handle, _ptr(m), K, float(T2), _ptr(r2), _ptr(mask)))
        return r2, mask

    # -- a4 ------------------------------------------------------------------------------------------------
    def score_compound(self, models, T2: float, compound_pref=None):
        m, K = self._models(models)
        cp = None if compound_pref is None else _f64(compound_pref)
        count = np.empty(K, dtype=np.int64)
        value = np.empty(K, dtype=np.float64)
        shared = np.empty(K, dtype=np.float64)
        _check(self.lib.pxb_score_compound(self.handle, _ptr(m), K, float(T2), _ptr(cp), _ptr(count), _ptr(value),
                                           _ptr(shared)))
        return count, value, shared

    def inliers(self, model, T2: float) -> np.ndarray:
        m, _ = self._models(model)
        out = np.empty(self.N, dtype=np.int64)
        n = C.c_int64()
        _check(self.lib.pxb_inliers(self.handle, _ptr(m), float(T2), _ptr(out), C.byref(n)))
        return out[: n.value].copy()

    # -- a5 ------------------------------------------------------------------------------------------------
    def preference_vector(self, model, T: float) -> np.ndarray:
        m, _ = self._models(model)
        out = np.empty(self.N, dtype=np.float64)
        _check(self.lib.pxb_preference_vector(self.handle, _ptr(m), float(T), _ptr(out)))
        return out

    def tanimoto(self, a, b) -> float:
        a, b = _f64(a), _f64(b)
        s = C.c_double()
        _check(self.lib.pxb_tanimoto(self.handle, _ptr(a), _ptr(b), a.shape[0], C.byref(s)))
        return s.value

    def compound_max(self, prefs) -> np.ndarray:
        prefs = _f64(prefs)
        L, N = prefs.shape
        out = np.empty(N, dtype=np.float64)
        _check(self.lib.pxb_compound_max(self.handle, _ptr(prefs), L, N, _ptr(out)))
        return out

    # -- a6/a7/a8 ------------------------------------------------------------------------------------------
    def solve_minimal(self, samples):
        t = self.model_type
        s = np.ascontiguousarray(samples, dtype=np.int64).reshape(-1, SAMPLE_SIZE[t])
        K = s.shape[0]
        models = np.zeros((K, MAX_SOLUTIONS[t], MODEL_SIZE[t]), dtype=np.float64)
        n = np.zeros(K, dtype=np.int32)
        sv = np.zeros(K, dtype=np.uint8)
        mv = np.zeros(K, dtype=np.uint8)
        _check(self.lib.pxb_solve_minimal(self.handle, _ptr(s), K, _ptr(models), _ptr(n), _ptr(sv), _ptr(mv)))
        return models, n, sv, mv

    def solve_plane_parallax(self, samples, H):
        """DEGENSAC's two-point solver over a fixed homography: (models [K, 9], n [K])."""
        s = np.ascontiguousarray(samples, dtype=np.int64).reshape(-1, 2)
        Hm = np.ascontiguousarray(H, dtype=np.float64).reshape(9)
        K = s.shape[0]
        models = np.zeros((K, 9), dtype=np.float64)
        n = np.zeros(K, dtype=np.int32)
        _check(self.lib.pxb_solve_plane_parallax(self.handle, _ptr(s), K, _ptr(Hm), _ptr(models), _ptr(n)))
        return models, n

    # -- a9..a12 -------------------------------------------------------------------------------------------
    def pearl_datacost(self, models, thr: float, lam: float) -> np.ndarray:
        m, L = self._models(models)
        D = np.empty((self.N, L + 1), dtype=np.float64)
        _check(self.lib.pxb_pearl_datacost(self.handle, _ptr(m), L, float(thr), float(lam), _ptr(D)))
        return D

    def pearl_label(self, D, lam: float, label_cost: float, csr_off=None, csr_idx=None, init_labels=None):
        D = _f64(D)
        N, L1 = D.shape
        off = None if csr_off is None else np.ascontiguousarray(csr_off, dtype=np.int32)
        idx = None if csr_idx is None else np.ascontiguousarray(csr_idx, dtype=np.int32)
        init = None if init_labels is None else np.ascontiguousarray(init_labels, dtype=np.int32)
        labels = np.empty(N, dtype=np.int32)
        e = C.c_double()
        _check(self.lib.pxb_pearl_label(self.handle, _ptr(D), N, L1, float(lam), float(label_cost), _ptr(off),
                                        _ptr(idx), _ptr(init), _ptr(labels), C.byref(e)))
        return labels, e.value

    def segment_residual_sums(self, models, labels):
        m, L = self._models(models)
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        sums = np.zeros(L, dtype=np.float64)
        counts = np.zeros(L, dtype=np.int64)
        _check(self.lib.pxb_segment_residual_sums(self.handle, _ptr(m), L, _ptr(lab), _ptr(sums), _ptr(counts)))
        return sums, counts

    # -- a13 -----------------------------------------------------------------------------------------------
    def lo_unary_terms(self, model, thr: float, lam: float):
        m, _ = self._models(model)
        d = np.empty(self.N)
        e0 = np.empty(self.N)
        e1 = np.empty(self.N)
        _check(self.lib.pxb_lo_unary_terms(self.handle, _ptr(m), float(thr), float(lam), _ptr(d), _ptr(e0), _ptr(e1)))
        return d, e0, e1

    def tukey_weights(self, model, T2: float) -> np.ndarray:
        m, _ = self._models(model)
        w = np.empty(self.N)
        _check(self.lib.pxb_tukey_weights(self.handle, _ptr(m), float(T2), _ptr(w)))
        return w

    def lo_graph_cut(self, e0, e1, d, lam: float, csr_off, csr_idx) -> np.ndarray:
        e0, e1, d = _f64(e0), _f64(e1), _f64(d)
        off = np.ascontiguousarray(csr_off, dtype=np.int32)
        idx = np.ascontiguousarray(csr_idx, dtype=np.int32)
        out = np.zeros(e0.shape[0], dtype=np.uint8)
        _check(self.lib.pxb_lo_graph_cut(self.handle, _ptr(e0), _ptr(e1), _ptr(d), e0.shape[0], float(lam), _ptr(off),
                                         _ptr(idx), _ptr(out)))
        return out

    def lo_labeling(self, model, thr: float, lam: float, csr_off, csr_idx) -> np.ndarray:
        """GCRANSAC::labeling in one device-resident call (unary terms + pairwise graph + st-cut)."""
        m, _ = self._models(model)
        off = np.ascontiguousarray(csr_off, dtype=np.int32)
        idx = np.ascontiguousarray(csr_idx, dtype=np.int32)
        out = np.zeros(self.N, dtype=np.uint8)
        _check(self.lib.pxb_lo_labeling(self.handle, _ptr(m), float(thr), float(lam), _ptr(off), _ptr(idx), _ptr(out)))
        return out

    # -- next rows -----------------------------------------------------------------------------------------
    def knn_graph(self, radius: float, k: int = 8):
        """Directed neighbour lists as CSR (off, idx), the format pearl_label / lo_graph_cut take."""
        nbr = np.empty((self.N, k), dtype=np.int32)
        deg = np.empty(self.N, dtype=np.int32)
        _check(self.lib.pxb_knn_graph(self.handle, float(radius), int(k), _ptr(nbr), _ptr(deg)))
        off = np.zeros(self.N + 1, dtype=np.int32)
        np.cumsum(deg, out=off[1:])
        idx = nbr[np.arange(k)[None, :] < deg[:, None]]
        return off, np.ascontiguousarray(idx, dtype=np.int32)

    def fit_homographies(self, index_sets, weights_by_row=None):
        off = np.zeros(len(index_sets) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(s) for s in index_sets])
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in index_sets]))
        w = None if weights_by_row is None else _f64(weights_by_row)
        H = np.zeros((len(index_sets), MODEL_SIZE[self.model_type]))
        ok = np.zeros(len(index_sets), dtype=np.int32)
        _check(self.lib.pxb_fit_nonminimal(self.handle, len(index_sets), _ptr(off), _ptr(idx), _ptr(w), _ptr(H),
                                           _ptr(ok)))
        return H, ok

    fit_nonminimal = fit_homographies

    def selftest_division(self, seed: int, n: int, mode: int) -> int:
        bad = C.c_int64()
        _check(self.lib.pxb_selftest_division(self.handle, seed, n, mode, C.byref(bad)))
        return bad.value


def h_degenerate_sample(rows, sample7, F):
    """DEGENSAC's seven-point test (host arithmetic, no context): (degenerate, H [3, 3] or None)."""
    r = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 4)
    s = np.ascontiguousarray(sample7, dtype=np.int64).reshape(7)
    Fm = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
    H = np.zeros(9, dtype=np.float64)
    flag = np.zeros(1, dtype=np.int32)
    _check(load_library().pxb_h_degenerate_sample(_ptr(r), _ptr(s), _ptr(Fm), _ptr(H), _ptr(flag)))
    return bool(flag[0]), (H.reshape(3, 3) if flag[0] else None)
