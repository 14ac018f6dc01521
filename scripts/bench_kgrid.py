"""Batched-k timing: npdrb_npdr_k_grid (one call, one distance matrix, one (d,row) sort) against one npdrb_npdr call per k
(what vwok's foreach loop does, R/adaptive.R:85-111).  usage: python scripts/bench_kgrid.py [N p n_k]  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import npdr_b200 as nb  # noqa: E402

N, p, nk = (int(v) for v in (sys.argv[1:4] + ["2000", "2000", "24"][len(sys.argv) - 1:]))
rng = np.random.default_rng(7)
X = np.asfortranarray(rng.normal(size=(N, p)))
y = (rng.random(N) < 0.5).astype(float)
X[:, :10] += 0.5 * y[:, None]
grid = [int(k) for k in np.unique(np.linspace(1, N // 3, nk).astype(int))]
nb.npdr_k_grid(y, X, grid[:2])                                   # warm-up (contexts, caches)
t0 = time.perf_counter()
g = nb.npdr_k_grid(y, X, grid)
t_grid_py = time.perf_counter() - t0
# the C call alone (the Python wrapper adds pt() / p.adjust() for every k)
import ctypes as C  # noqa: E402
from npdr_b200 import _lib as L_  # noqa: E402
o = L_.Opts(); L_.load().npdrb_opts_default(C.byref(o)); o.nbd_method = 2
kg = np.asarray(grid, dtype=np.int64); st_out = np.empty((len(grid), p, 8)); npairs = np.zeros(len(grid), dtype=np.int64)
t0 = time.perf_counter()
L_.check(L_.load().npdrb_npdr_k_grid(L_.context(0).handle, L_.dptr(X), N, p, L_.dptr(y), None, C.byref(o),
                                     kg.ctypes.data_as(C.POINTER(C.c_int64)), len(grid), L_.dptr(st_out),
                                     npairs.ctypes.data_as(C.POINTER(C.c_int64))))
t_grid = time.perf_counter() - t0
t0 = time.perf_counter()
singles = [nb.npdr_stats(X, y, None, "binomial", "numeric-abs", "relieff", "manhattan", k, 0.5, ())[0] for k in grid]
t_single = time.perf_counter() - t0
# stage breakdown of the batched path through the session API
from npdr_b200 import _lib  # noqa: E402
ses = _lib.Session(_lib.context(0), N, p, 0)
def _t(fn):
    t = time.perf_counter(); r = fn(); _lib.context(0).synchronize(); return time.perf_counter() - t, r
st = {}
st["upload_s"], _ = _t(lambda: ses.upload(X, y))
st["distances_s"], _ = _t(lambda: ses.distances("manhattan", False, 0, N))
st["sort_rows_s"], _ = _t(lambda: ses.sort_rows())
tn = tr = ti = 0.0
for k in grid:
    dt, _ = _t(lambda: ses.neighbors_sorted(k)); tn += dt
    a, b, m = ses.local_pairs()
    dt, _ = _t(lambda: ses.import_pairs(a, b, m)); ti += dt
    dt, _ = _t(lambda: ses.regress(0, p, "binomial", "numeric-abs", ())); tr += dt
st["neighbors_sorted_total_s"] = tn; st["import_total_s"] = ti; st["regress_total_s"] = tr
tsel = 0.0
for k in grid:
    dt, _ = _t(lambda: ses.neighbors("relieff", 0.5, k)); tsel += dt
st["neighbors_select_total_s"] = tsel
ses.close()
same = all(np.array_equal(g["stats"][c], singles[c], equal_nan=True) for c in range(len(grid)))
print(json.dumps({"N": N, "p": p, "k_grid": grid, "pairs_total": int(g["n_pairs"].sum()), "k_grid_call_s": t_grid, "k_grid_python_wrapper_s": t_grid_py,
                  "one_call_per_k_s": t_single, "speedup": t_single / t_grid, "identical_statistics": bool(same), "stages": st}))
