"""ctypes binding of oracle/libnpdr_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (npdr_b200/) never does.
PARITY STATUS: see the header of npdr_oracle.c ("parity unpinned" except npdrDiff).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnpdr_oracle.so")

DIFF = {"numeric-abs": 0, "manhattan": 0, "numeric-sqr": 1, "allele-sharing": 2, "match-mismatch": 3, "raw": 4}
METRIC = {"manhattan": 0, "euclidean": 1, "allele-sharing-manhattan": 2,
          "relief-scaled-manhattan": 3, "relief-scaled-euclidean": 4}
NBD = {"multisurf": 0, "surf": 1, "relieff": 2}
REG = {"lm": 0, "binomial": 1}


def build(force=False):
    src = os.path.join(_HERE, "npdr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_knn_surf.restype = C.c_int64
        L.orc_knn_surf.argtypes = [C.c_int64, C.c_double]
        L.orc_unique_neighbors.restype = C.c_int64
        L.orc_knn_vec.restype = C.c_int64
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP threads of the oracle from now on (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def diff(a, b, diff_type="numeric-abs", norm_fac=1.0):
    a = _f64(np.asarray(a, dtype=np.float64)); b = _f64(np.asarray(b, dtype=np.float64))
    if diff_type == "correlation-data":
        out = np.empty(a.shape[0])
        lib().orc_diff_corr(_d(a), _d(b), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]), C.c_double(norm_fac), _d(out))
        return out
    out = np.empty(a.size)
    rc = lib().orc_diff(_d(a), _d(b), C.c_int64(a.size), C.c_int(DIFF[diff_type]), C.c_double(norm_fac), _d(out))
    assert rc == 0
    return out


def distances(X, metric="manhattan"):
    X = _f64(X)
    N, p = X.shape
    D = np.empty((N, N), order="F")
    if metric == "hamming":
        rc = lib().orc_hamming(_d(X), C.c_int64(N), C.c_int64(p), _d(D))
    else:
        rc = lib().orc_distances(_d(X), C.c_int64(N), C.c_int64(p), C.c_int(METRIC[metric]), _d(D))
    assert rc == 0
    return D


def radius_multisurf(D, sd_frac=0.5, sd_vec=None):
    D = _f64(D); N = D.shape[0]
    r = np.empty(N); sd = np.empty(N)
    sv = _d(_f64(sd_vec)) if sd_vec is not None else None
    rc = lib().orc_radius_multisurf(_d(D), C.c_int64(N), C.c_double(sd_frac), sv, _d(r), _d(sd))
    assert rc == 0
    return r, sd


def radius_surf(D, sd_frac=0.5):
    D = _f64(D); N = D.shape[0]
    r = C.c_double()
    rc = lib().orc_radius_surf(_d(D), C.c_int64(N), C.c_double(sd_frac), C.byref(r))
    assert rc == 0
    return r.value


def knn_surf(m_samples, sd_frac=0.5):
    return int(lib().orc_knn_surf(int(m_samples), float(sd_frac)))


def neighbors(D, method="multisurf", k=0, radius=None):
    """-> (Ri, NN) 1-based int32 arrays, n_empty."""
    D = _f64(D); N = D.shape[0]
    Ri = _ip(); NN = _ip(); M = C.c_int64(); ne = C.c_int64()
    rad = _d(_f64(np.asarray(radius, dtype=np.float64))) if radius is not None else None
    rc = lib().orc_neighbors(_d(D), C.c_int64(N), C.c_int(NBD[method]), C.c_int64(k), rad,
                             C.byref(Ri), C.byref(NN), C.byref(M), C.byref(ne))
    assert rc == 0
    m = M.value
    ri = np.ctypeslib.as_array(Ri, shape=(max(m, 1),))[:m].copy()
    nn = np.ctypeslib.as_array(NN, shape=(max(m, 1),))[:m].copy()
    lib().orc_free(Ri); lib().orc_free(NN)
    return ri, nn, ne.value


def nearest_neighbors(X, nbd_method="multisurf", nbd_metric="manhattan", sd_frac=0.5, k=0, sd_vec=None):
    """nearestNeighbors(), R/nearestNeighbors.R:153-285 -> (Ri, NN, D, radius_or_None)."""
    D = _f64(X) if nbd_metric == "precomputed" else distances(X, nbd_metric)
    N = D.shape[0]
    if nbd_method == "relieff":
        if k == 0:
            k = knn_surf(N, sd_frac)           # R/nearestNeighbors.R:197-203 (same formula as knnSURF)
        ri, nn, _ = neighbors(D, "relieff", k=k)
        return ri, nn, D, None
    if nbd_method == "surf":
        r = np.full(N, radius_surf(D, sd_frac))
    else:
        r, _ = radius_multisurf(D, sd_frac, sd_vec)
    ri, nn, _ = neighbors(D, nbd_method, radius=r)
    return ri, nn, D, r


def nearest_neighbors_hitmiss(X, pheno, nbd_method="relieff", nbd_metric="manhattan", sd_frac=0.5, k=0):
    """nearestNeighborsSeparateHitMiss(), R/nearestNeighbors.R:318-551 -> (Ri, NN, D, r_hit, r_miss)."""
    D = _f64(X) if nbd_metric == "precomputed" else distances(X, nbd_metric)
    N = D.shape[0]
    ph = _f64(np.asarray(pheno, dtype=np.float64))
    rh = np.zeros(N); rm = np.zeros(N)
    if nbd_method == "multisurf":
        rc = lib().orc_radius_hitmiss_multisurf(_d(D), C.c_int64(N), _d(ph), C.c_double(sd_frac), _d(rh), _d(rm))
        assert rc == 0
    elif nbd_method == "surf":
        a = C.c_double(); b = C.c_double()
        rc = lib().orc_radius_hitmiss_surf(_d(D), C.c_int64(N), _d(ph), C.c_double(sd_frac), C.byref(a), C.byref(b))
        assert rc == 0
        rh[:] = a.value; rm[:] = b.value
    Ri = _ip(); NN = _ip(); M = C.c_int64()
    rc = lib().orc_neighbors_hitmiss(_d(D), C.c_int64(N), _d(ph), C.c_int(NBD[nbd_method]), C.c_double(sd_frac), _d(rh), _d(rm),
                                     C.byref(Ri), C.byref(NN), C.byref(M))
    assert rc == 0
    m = M.value
    ri = np.ctypeslib.as_array(Ri, shape=(max(m, 1),))[:m].copy()
    nn = np.ctypeslib.as_array(NN, shape=(max(m, 1),))[:m].copy()
    lib().orc_free(Ri); lib().orc_free(NN)
    return ri, nn, D, rh, rm


def unique_neighbors(Ri, NN, N):
    Ri = np.ascontiguousarray(Ri, dtype=np.int32); NN = np.ascontiguousarray(NN, dtype=np.int32)
    keep = np.zeros(Ri.size, dtype=np.uint8)
    n = lib().orc_unique_neighbors(Ri.ctypes.data_as(_ip), NN.ctypes.data_as(_ip), C.c_int64(Ri.size), C.c_int64(N),
                                   keep.ctypes.data_as(C.POINTER(C.c_uint8)))
    return int(n), keep.astype(bool)


def knn_vec(Ri, N):
    Ri = np.ascontiguousarray(Ri, dtype=np.int32)
    cnt = np.zeros(N, dtype=np.int64)
    n = lib().orc_knn_vec(Ri.ctypes.data_as(_ip), C.c_int64(Ri.size), C.c_int64(N), cnt.ctypes.data_as(C.POINTER(C.c_int64)))
    return cnt[:n]


class _FitInfo(C.Structure):
    _fields_ = [("rank", C.c_int), ("df_resid", C.c_int64), ("rss", C.c_double), ("mss", C.c_double),
                ("r_squared", C.c_double), ("deviance", C.c_double), ("iters", C.c_int),
                ("converged", C.c_int), ("boundary", C.c_int)]


def _fit(fn, Xd, y):
    Xd = _f64(Xd); y = _f64(np.asarray(y, dtype=np.float64))
    n, p = Xd.shape
    coef = np.empty(p); se = np.empty(p); info = _FitInfo()
    rc = fn(_d(Xd), _d(y), C.c_int64(n), C.c_int(p), _d(coef), _d(se), C.byref(info))
    assert rc == 0, rc
    return coef, se, {f[0]: getattr(info, f[0]) for f in _FitInfo._fields_}


def lm(Xd, y):
    return _fit(lib().orc_lm, Xd, y)


def glm_binomial(Xd, y):
    return _fit(lib().orc_glm_binomial, Xd, y)


def pair_diffs(v, Ri, NN, diff_type):
    v = _f64(np.asarray(v, dtype=np.float64))
    Ri = np.ascontiguousarray(Ri, dtype=np.int32); NN = np.ascontiguousarray(NN, dtype=np.int32)
    out = np.empty(Ri.size)
    lib().orc_pair_diffs(_d(v), Ri.ctypes.data_as(_ip), NN.ctypes.data_as(_ip), C.c_int64(Ri.size),
                         C.c_int(DIFF[diff_type]), _d(out))
    return out


def regress_attrs(X, Ri, NN, ydiff, cdiff=None, attr_diff_type="numeric-abs", regression_type="binomial"):
    """Regression loop R/npdr.R:391-430 -> (p, 8) array:
    beta_a, stat_a, beta_0, stat_0, df_resid, r_squared|deviance, iters, flags."""
    X = _f64(X); N, p = X.shape
    Ri = np.ascontiguousarray(Ri, dtype=np.int32); NN = np.ascontiguousarray(NN, dtype=np.int32)
    M = Ri.size
    ydiff = _f64(np.asarray(ydiff, dtype=np.float64))
    q = 0; cd = None
    if cdiff is not None and np.size(cdiff):
        cdiff = _f64(np.asarray(cdiff, dtype=np.float64).reshape(M, -1)); q = cdiff.shape[1]; cd = _d(cdiff)
    out = np.empty((p, 8))
    rc = lib().orc_regress_attrs(_d(X), C.c_int64(N), C.c_int64(p), Ri.ctypes.data_as(_ip), NN.ctypes.data_as(_ip),
                                 C.c_int64(M), _d(ydiff), cd, C.c_int(q), C.c_int(DIFF[attr_diff_type]),
                                 C.c_int(REG[regression_type]), _d(out))
    assert rc == 0, rc
    return out


def distances2(X1, X2, metric="manhattan"):
    """npdrDistances2 (R/npdrLearner.R:323-339): m1 x m2 matrix, rows = instances of X1, columns = instances of X2."""
    X1 = _f64(np.asarray(X1, dtype=np.float64)); X2 = _f64(np.asarray(X2, dtype=np.float64))
    m = metric if metric in ("euclidean", "allele-sharing-manhattan") else "manhattan"
    D = np.empty((X1.shape[0], X2.shape[0]), order="F")
    rc = lib().orc_distances2(_d(X1), C.c_int64(X1.shape[0]), _d(X2), C.c_int64(X2.shape[0]), C.c_int64(X1.shape[1]),
                              C.c_int(METRIC[m]), _d(D))
    assert rc == 0
    return D


def nearest_neighbors2(X1, X2, nbd_method="multisurf", nbd_metric="manhattan", sd_vec=None, sd_frac=0.5, k=0):
    """nearestNeighbors2 (R/npdrLearner.R:378-482) -> (list of m2 int32 arrays of 1-based X1 indices, radius, D)."""
    D = distances2(X1, X2, nbd_metric)
    m1, m2 = D.shape
    radius = np.full(m2, np.nan)
    if nbd_method == "relieff":
        if k == 0:
            k = int(lib().orc_knn_surf(C.c_int64(m1), C.c_double(sd_frac)))       # R/npdrLearner.R:413-418
    else:
        sv = _f64(np.asarray(sd_vec, dtype=np.float64)) if sd_vec is not None else None
        rc = lib().orc_radius2(_d(D), C.c_int64(m1), C.c_int64(m2), C.c_int(NBD[nbd_method]), C.c_double(sd_frac),
                               _d(sv) if sv is not None else None, _d(radius))
        assert rc == 0
    off = np.zeros(m2 + 1, dtype=np.int64)
    idx = _ip()
    rc = lib().orc_neighbors2(_d(D), C.c_int64(m1), C.c_int64(m2), C.c_int(NBD[nbd_method]), C.c_int64(k), _d(radius),
                              off.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(idx))
    assert rc == 0
    M = int(off[-1])
    nn = np.ctypeslib.as_array(idx, shape=(max(M, 1),))[:M].copy()
    lib().orc_free(C.cast(idx, C.c_void_p))
    return [nn[off[i]:off[i + 1]] for i in range(m2)], radius, D
