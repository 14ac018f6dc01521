"""One run of the wide-design path (q = 12 covariates, binomial) for profiling wide_fit_kernel:
ncu --set full -k regex:wide_fit_kernel python scripts/wide_case.py [N p q]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npdr_b200 as nb  # noqa: E402

N, p, q = (int(v) for v in (sys.argv[1:4] + ["1500", "1024", "12"][len(sys.argv) - 1:]))
rng = np.random.default_rng(3)
X = np.asfortranarray(rng.normal(size=(N, p)))
y = (rng.random(N) < 0.5).astype(float)
X[:, :8] += 0.6 * y[:, None]
C = np.asfortranarray(np.column_stack([rng.integers(0, 2, N).astype(float) if c % 3 == 0 else rng.normal(40, 12, N).round() for c in range(q)]))
types = ["match-mismatch" if c % 3 == 0 else "numeric-abs" for c in range(q)]
nb.npdr_stats(X[:, :64], y, C, "binomial", "numeric-abs", "multisurf", "manhattan", 0, 0.5, types)      # warm-up
t0 = time.perf_counter()
st, info = nb.npdr_stats(X, y, C, "binomial", "numeric-abs", "multisurf", "manhattan", 0, 0.5, types)
dt = time.perf_counter() - t0
K = q + 3
KT = 8 if K <= 8 else 16 if K <= 16 else 32 if K <= 32 else 64
print({"N": N, "p": p, "q": q, "pairs": info["n_pairs"], "seconds": dt, "regress_ms": info.get("ms_regress"), "iters_max": float(np.nanmax(st[:, 6])),
       "gram_fma_per_s": info["n_pairs"] * p * float(np.nanmean(st[:, 6]) + 1) * KT * KT / 2 / max(dt, 1e-9)})
