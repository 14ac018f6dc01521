"""ctypes binding of libnpdr_b200.so (the C ABI declared in include/npdr_b200.h).

There is no CPU fallback: importing this module loads the CUDA shared library, and every
compute call fails loudly (NpdrB200Error) if no sm_100 device is usable.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("NPDRB_LIB_PATH") or os.path.join(_HERE, "libnpdr_b200.so")   # override: A/B runs of two builds

STATS_PER_ATTR = 8
MAX_COVARS = 60
DIFF = {"numeric-abs": 0, "manhattan": 0, "numeric-sqr": 1, "allele-sharing": 2, "match-mismatch": 3, "raw": 4}
METRIC = {"manhattan": 0, "euclidean": 1, "allele-sharing-manhattan": 2, "relief-scaled-manhattan": 3,
          "relief-scaled-euclidean": 4, "precomputed": 5, "hamming": 6}
NBD = {"multisurf": 0, "surf": 1, "relieff": 2}
REG = {"lm": 0, "binomial": 1}
ENODEVICE = 2


class NpdrB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnpdr_b200 error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [("regression_type", C.c_int32), ("attr_diff_type", C.c_int32), ("nbd_method", C.c_int32),
                ("nbd_metric", C.c_int32), ("knn", C.c_int64), ("msurf_sd_frac", C.c_double),
                ("n_covars", C.c_int32), ("covar_diff_type", C.c_int32 * MAX_COVARS),
                ("neighbor_sampling_unique", C.c_int32), ("fast_euclidean", C.c_int32),
                ("return_pairs", C.c_int32), ("separate_hitmiss_nbds", C.c_int32), ("reserved", C.c_int32 * 7)]


class Result(C.Structure):
    _fields_ = [("n_attr", C.c_int64), ("stats", C.POINTER(C.c_double)), ("n_pairs", C.c_int64),
                ("n_unique_pairs", C.c_int64), ("n_pairs_used", C.c_int64), ("k_ave_empirical", C.c_double),
                ("n_empty", C.c_int64), ("knn_used", C.c_int64), ("Ri", C.POINTER(C.c_int32)),
                ("NN", C.POINTER(C.c_int32)), ("ms_h2d", C.c_double), ("ms_distance", C.c_double),
                ("ms_neighbors", C.c_double), ("ms_prepare", C.c_double), ("ms_regress", C.c_double),
                ("ms_total", C.c_double), ("irls_passes", C.c_int32), ("reserved", C.c_int32 * 7)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32

# name -> (restype, argtypes); every symbol include/npdr_b200.h declares
SIGNATURES = {
    "npdrb_version": (C.c_char_p, []),
    "npdrb_last_error": (C.c_char_p, []),
    "npdrb_free": (None, [_vp]),
    "npdrb_opts_default": (None, [C.POINTER(Opts)]),
    "npdrb_result_release": (None, [C.POINTER(Result)]),
    "npdrb_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "npdrb_destroy": (None, [_vp]),
    "npdrb_set_stream": (C.c_int, [_vp, _vp]),
    "npdrb_synchronize": (C.c_int, [_vp]),
    "npdrb_measure_fp64_peak": (C.c_int, [_vp, _dp, _dp]),
    "npdrb_launch_count": (_i64, [_vp]),
    "npdrb_release_cache": (C.c_int, [_vp]),
    "npdrb_upload_columns": (C.c_int, [_vp, _vp, _i64, _dp, _i64, _i64]),
    "npdrb_diff": (C.c_int, [_vp, _dp, _dp, _i64, _i32, C.c_double, _dp]),
    "npdrb_distances": (C.c_int, [_vp, _dp, _i64, _i64, _i32, _i32, _dp]),
    "npdrb_distances2": (C.c_int, [_vp, _dp, _i64, _dp, _i64, _i64, _i32, _dp]),
    "npdrb_nearest_neighbors2": (C.c_int, [_vp, _dp, _i64, _dp, _i64, _i64, _i32, _i32, _dp, C.c_double, _i64,
                                           C.POINTER(C.POINTER(_i64)), C.POINTER(_ip), C.POINTER(_i64), _dp]),
    "npdrb_knn_surf": (_i64, [_i64, C.c_double]),
    "npdrb_nearest_neighbors": (C.c_int, [_vp, _dp, _i64, _i64, _i32, _i32, _dp, C.c_double, _i64, _i32,
                                          C.POINTER(_ip), C.POINTER(_ip), C.POINTER(_i64), C.POINTER(_i64),
                                          C.POINTER(_i64), _dp]),
    "npdrb_nearest_neighbors_hitmiss": (C.c_int, [_vp, _dp, _i64, _i64, _dp, _i32, _i32, C.c_double, _i64, _i32,
                                                  C.POINTER(_ip), C.POINTER(_ip), C.POINTER(_i64), C.POINTER(_i64), _dp, _dp]),
    "npdrb_regress": (C.c_int, [_vp, _dp, _i64, _i64, _ip, _ip, _i64, _dp, _dp, _i32, _ip, _i32, _i32, _dp]),
    "npdrb_npdr": (C.c_int, [_vp, _dp, _i64, _i64, _dp, _dp, C.POINTER(Opts), C.POINTER(Result)]),
    "npdrb_npdr_multi": (C.c_int, [_i32, _ip, _dp, _i64, _i64, _dp, _dp, C.POINTER(Opts), C.POINTER(Result)]),
    "npdrb_multi_shutdown": (None, []),
    "npdrb_unireg": (C.c_int, [_vp, _dp, _i64, _i64, _dp, _dp, _i32, _i32, _dp]),
    "npdrb_npdr_k_grid": (C.c_int, [_vp, _dp, _i64, _i64, _dp, _dp, C.POINTER(Opts), C.POINTER(_i64), _i32, _dp, C.POINTER(_i64)]),
    "npdrb_session_create": (C.c_int, [_vp, _i64, _i64, _i32, C.POINTER(_vp)]),
    "npdrb_session_destroy": (None, [_vp]),
    "npdrb_session_upload": (C.c_int, [_vp, _dp, _dp, _dp]),
    "npdrb_session_set_device_inputs": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "npdrb_session_distances": (C.c_int, [_vp, _i32, _i32, _i64, _i64]),
    "npdrb_session_set_device_distances": (C.c_int, [_vp, _vp, _i64, _i64, _i64]),
    "npdrb_session_distances_balanced": (C.c_int, [_vp, _i32, _i32, C.POINTER(_i64), _i32, _i32]),
    "npdrb_session_balanced_counts": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "npdrb_session_balanced_pack": (C.c_int, [_vp, _vp]),
    "npdrb_session_balanced_unpack": (C.c_int, [_vp, _vp]),
    "npdrb_session_set_sd_vec": (C.c_int, [_vp, _dp]),
    "npdrb_session_neighbors": (C.c_int, [_vp, _i32, C.c_double, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "npdrb_session_neighbors_hitmiss": (C.c_int, [_vp, _i32, C.c_double, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "npdrb_session_set_device_slab_events": (C.c_int, [_vp, _i32, _i64, C.POINTER(_vp), C.POINTER(_i32)]),
    "npdrb_session_set_device_slab_events_v": (C.c_int, [_vp, _i32, C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_i32)]),
    "npdrb_session_copy_radius2": (C.c_int, [_vp, _dp]),
    "npdrb_session_sort_rows": (C.c_int, [_vp]),
    "npdrb_session_neighbors_sorted": (C.c_int, [_vp, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "npdrb_session_surf_partial": (C.c_int, [_vp, C.c_double, _dp]),
    "npdrb_session_set_surf_radius": (C.c_int, [_vp, C.c_double]),
    "npdrb_surf_radius_from_moments": (C.c_int, [_dp, _dp, _i64, C.c_double, _dp, _dp]),
    "npdrb_session_surf_band_count": (C.c_int, [_vp, C.c_double, C.c_double, C.POINTER(_i64)]),
    "npdrb_surf_radius_exact_host": (C.c_int, [_dp, _i64, _i64, C.c_double, _dp]),
    "npdrb_session_local_pairs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "npdrb_session_copy_local_pairs_device": (C.c_int, [_vp, _vp, _vp]),
    "npdrb_session_copy_radius": (C.c_int, [_vp, _dp]),
    "npdrb_session_copy_distances": (C.c_int, [_vp, _dp]),
    "npdrb_session_import_pairs": (C.c_int, [_vp, _vp, _vp, _i64]),
    "npdrb_session_import_pairs_host": (C.c_int, [_vp, _ip, _ip, _i64]),
    "npdrb_session_unique_pairs": (C.c_int, [_vp, _i32, C.POINTER(_i64)]),
    "npdrb_session_regress": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _ip, _dp, C.POINTER(_i32)]),
    "npdrb_session_copy_pairs": (C.c_int, [_vp, _ip, _ip]),
    "npdrb_session_stage_ms": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "npdrb_session_regress_design": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "npdrb_session_kernel_ms": (C.c_int, [_vp, _dp]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libnpdr_b200.so (needs no GPU) and bind every declared symbol."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(SO_PATH):
                raise ImportError(f"{SO_PATH} is missing: build it with ./build.sh or __graft_entry__.build(); "
                                  "npdr_b200 has no CPU fallback")
            L = C.CDLL(SO_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)            # AttributeError if the library does not export it
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def last_error():
    return load().npdrb_last_error().decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise NpdrB200Error(rc, last_error())


def f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64,
                      requirements=["F" if order == "F" else "C", "A"])


def dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class Context:
    """One per device (npdrb_ctx)."""

    def __init__(self, device=0):
        L = load()
        h = _vp()
        check(L.npdrb_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            load().npdrb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        check(load().npdrb_set_stream(self.handle, _vp(int(cuda_stream_ptr))))

    def synchronize(self):
        check(load().npdrb_synchronize(self.handle))

    def release_cache(self):
        check(load().npdrb_release_cache(self.handle))

    def launch_count(self):
        return int(load().npdrb_launch_count(self.handle))

    def upload_columns(self, dst_device_ptr, ld_dst, src):
        """Host columns (Fortran-ordered float64 [rows, ncols], pageable or pinned) -> device memory at dst_device_ptr."""
        a = f64(src)
        check(load().npdrb_upload_columns(self.handle, _vp(int(dst_device_ptr)), int(ld_dst), dptr(a), a.shape[0], a.shape[1]))

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        check(load().npdrb_measure_fp64_peak(self.handle, C.byref(a), C.byref(b)))
        return {"dfma_tflops": a.value, "dadd_tflops": b.value}


_contexts = {}


def context(device=0):
    """Process-wide context cache."""
    ctx = _contexts.get(device)
    if ctx is None or ctx.handle is None:
        ctx = Context(device)
        _contexts[device] = ctx
    return ctx


def surf_radius_from_moments(m0, m1, N, sd_frac):
    """-> (radius, eps): SURF radius from combined partial moments and the bound on its distance from R's value."""
    a = f64(np.asarray(m0, dtype=np.float64), "C"); b = f64(np.asarray(m1, dtype=np.float64), "C")
    r, e = C.c_double(), C.c_double()
    check(load().npdrb_surf_radius_from_moments(dptr(a), dptr(b), int(N), float(sd_frac), C.byref(r), C.byref(e)))
    return r.value, e.value


def surf_radius_exact_host(D, sd_frac):
    """R's SURF radius replayed exactly on the host from the full symmetric distance matrix (numpy, any order)."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    r = C.c_double()
    check(load().npdrb_surf_radius_exact_host(dptr(D), D.shape[1], D.shape[0], float(sd_frac), C.byref(r)))
    return r.value


def take_int_array(ptr, n):
    """Copy a library-malloc'ed int32 buffer into numpy and free it."""
    if n > 0:
        out = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
    else:
        out = np.zeros(0, dtype=np.int32)
    load().npdrb_free(C.cast(ptr, _vp))
    return out


class Session:
    """One NPDR problem resident on one device (npdrb_session): the sharding building block."""

    def __init__(self, ctx, N, p, q=0):
        self.ctx, self.N, self.p, self.q = ctx, int(N), int(p), int(q)
        h = _vp()
        check(load().npdrb_session_create(ctx.handle, self.N, self.p, self.q, C.byref(h)))
        self.handle = h
        self._keep = []

    def close(self):
        if getattr(self, "handle", None):
            load().npdrb_session_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, X=None, y=None, C_=None):
        Xa = f64(X) if X is not None else None
        ya = f64(y) if y is not None else None
        Ca = f64(C_) if C_ is not None else None
        self._host = (Xa, ya, Ca)       # a large X is uploaded in slabs that are still in flight when this call returns
        check(load().npdrb_session_upload(self.handle, dptr(Xa), dptr(ya), dptr(Ca)))

    def set_device_inputs(self, dX=0, ldx=0, dy=0, dC=0):
        check(load().npdrb_session_set_device_inputs(self.handle, _vp(dX or None), int(ldx), _vp(dy or None), _vp(dC or None)))

    def set_device_slab_events(self, slab_cols, events, ready=None):
        """events: raw cudaEvent_t handles (ints), one per column slab of the adopted device matrix; ready: optional int32
        numpy array of host flags set by the thread that issues the transfers (see npdr_b200.h)."""
        arr = (_vp * len(events))(*[int(e) for e in events])
        self._slab_events = (arr, ready)
        rp = ready.ctypes.data_as(C.POINTER(_i32)) if ready is not None else None
        if np.ndim(slab_cols) == 0:
            check(load().npdrb_session_set_device_slab_events(self.handle, len(events), int(slab_cols), arr, rp))
        else:                                                  # slab bounds (n + 1 values): slabs of different widths
            st = (_i64 * len(slab_cols))(*[int(v) for v in slab_cols])
            check(load().npdrb_session_set_device_slab_events_v(self.handle, len(events), st, arr, rp))

    def distances(self, metric="manhattan", fast_euclidean=False, row0=0, nrows=None):
        nrows = self.N if nrows is None else nrows
        check(load().npdrb_session_distances(self.handle, METRIC[metric], int(bool(fast_euclidean)), int(row0), int(nrows)))
        self.nrows = int(nrows)

    def distances_balanced(self, metric, row_starts, rank):
        """This rank's share of the symmetric matrix (diagonal square + half of every off-diagonal square);
        -> (send_counts, recv_counts) in doubles per peer for the all-to-all that completes the block."""
        world = len(row_starts) - 1
        rs = (_i64 * (world + 1))(*[int(v) for v in row_starts])
        check(load().npdrb_session_distances_balanced(self.handle, METRIC[metric], 0, rs, int(rank), world))
        self.nrows = int(row_starts[rank + 1] - row_starts[rank])
        sc, rc = (_i64 * world)(), (_i64 * world)()
        check(load().npdrb_session_balanced_counts(self.handle, sc, rc))
        return list(sc), list(rc)

    def balanced_pack(self, d_send):
        check(load().npdrb_session_balanced_pack(self.handle, _vp(d_send)))

    def balanced_unpack(self, d_recv):
        check(load().npdrb_session_balanced_unpack(self.handle, _vp(d_recv)))

    def set_device_distances(self, dD, ldd, row0, nrows):
        check(load().npdrb_session_set_device_distances(self.handle, _vp(dD), int(ldd), int(row0), int(nrows)))
        self.nrows = int(nrows)

    def set_sd_vec(self, sd_vec):
        a = f64(sd_vec) if sd_vec is not None else None
        check(load().npdrb_session_set_sd_vec(self.handle, dptr(a)))

    def neighbors(self, method="multisurf", sd_frac=0.5, k=0):
        M, ne = _i64(), _i64()
        check(load().npdrb_session_neighbors(self.handle, NBD[method], float(sd_frac), int(k), C.byref(M), C.byref(ne)))
        return M.value, ne.value

    def neighbors_hitmiss(self, method="relieff", sd_frac=0.5, k=0):
        M, ne = _i64(), _i64()
        check(load().npdrb_session_neighbors_hitmiss(self.handle, NBD[method], float(sd_frac), int(k), C.byref(M), C.byref(ne)))
        return M.value, ne.value

    def sort_rows(self):
        check(load().npdrb_session_sort_rows(self.handle))

    def neighbors_sorted(self, k):
        M, ne = _i64(), _i64()
        check(load().npdrb_session_neighbors_sorted(self.handle, int(k), C.byref(M), C.byref(ne)))
        return M.value, ne.value

    def copy_radius2(self):
        out = np.empty(self.nrows)
        check(load().npdrb_session_copy_radius2(self.handle, dptr(out)))
        return out

    def surf_partial(self, centre=0.0):
        out = np.zeros(4)
        check(load().npdrb_session_surf_partial(self.handle, float(centre), dptr(out)))
        return out

    def set_surf_radius(self, r):
        check(load().npdrb_session_set_surf_radius(self.handle, float(r)))

    def surf_band_count(self, lo, hi):
        n = _i64()
        check(load().npdrb_session_surf_band_count(self.handle, float(lo), float(hi), C.byref(n)))
        return n.value

    def local_pairs(self):
        a, b, M = _vp(), _vp(), _i64()
        check(load().npdrb_session_local_pairs(self.handle, C.byref(a), C.byref(b), C.byref(M)))
        return a.value, b.value, M.value

    def copy_local_pairs_device(self, dRi_dst, dNN_dst):
        check(load().npdrb_session_copy_local_pairs_device(self.handle, _vp(dRi_dst), _vp(dNN_dst)))

    def copy_radius(self):
        out = np.empty(self.nrows)
        check(load().npdrb_session_copy_radius(self.handle, dptr(out)))
        return out

    def copy_distances(self):
        out = np.empty((self.nrows, self.N), order="C")
        check(load().npdrb_session_copy_distances(self.handle, dptr(out)))
        return out

    def import_pairs(self, dRi, dNN, M):
        check(load().npdrb_session_import_pairs(self.handle, _vp(dRi), _vp(dNN), int(M)))
        self.M = int(M)

    def import_pairs_host(self, Ri, NN):
        Ri = np.ascontiguousarray(Ri, dtype=np.int32); NN = np.ascontiguousarray(NN, dtype=np.int32)
        check(load().npdrb_session_import_pairs_host(self.handle, iptr(Ri), iptr(NN), Ri.size))
        self.M = int(Ri.size)

    def unique_pairs(self, keep_unique_only=False):
        n = _i64()
        check(load().npdrb_session_unique_pairs(self.handle, int(bool(keep_unique_only)), C.byref(n)))
        return n.value

    def regress(self, attr0=0, nattr=None, regression_type="binomial", attr_diff_type="numeric-abs", covar_diff_type=()):
        nattr = self.p - attr0 if nattr is None else nattr
        out = np.empty((nattr, STATS_PER_ATTR))
        cdt = np.asarray([DIFF[t] for t in covar_diff_type] + [DIFF["match-mismatch"]] * MAX_COVARS, dtype=np.int32)[:MAX_COVARS]
        passes = _i32()
        check(load().npdrb_session_regress(self.handle, int(attr0), int(nattr), REG[regression_type], DIFF[attr_diff_type],
                                           iptr(cdt), dptr(out), C.byref(passes)))
        self.irls_passes = passes.value
        return out

    def copy_pairs(self, M):
        Ri = np.empty(M, dtype=np.int32); NN = np.empty(M, dtype=np.int32)
        if M:
            check(load().npdrb_session_copy_pairs(self.handle, iptr(Ri), iptr(NN)))
        return Ri, NN

    def regress_design(self):
        qe, ns, nx = _i32(), _i32(), _i32()
        check(load().npdrb_session_regress_design(self.handle, C.byref(qe), C.byref(ns), C.byref(nx)))
        return {"q_eff": qe.value, "n_streams": ns.value, "n_exact_stops": nx.value}

    def kernel_ms(self):
        out = np.zeros(5)
        check(load().npdrb_session_kernel_ms(self.handle, dptr(out)))
        return dict(zip(("pass_first_ms", "pass_full_ms", "n_pass_full", "evals_pass_full", "records"), out.tolist()))

    def stage_ms(self):
        v = [C.c_double() for _ in range(4)]
        check(load().npdrb_session_stage_ms(self.handle, *[C.byref(x) for x in v]))
        return dict(zip(("distance", "neighbors", "prepare", "regress"), [x.value for x in v]))
