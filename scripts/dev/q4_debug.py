"""Hunt for data on which the q = 4 binomial path disagrees with the oracle (NaN pattern / tolerance), with and without
binary-covariate folding.  usage: python scripts/dev/q4_debug.py [n_seeds]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import npdr_b200 as nb
from oracle import oracle as O
O.build()
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for q in (3, 4):
  for seed in range(ns):
    rng = np.random.default_rng(seed)
    N, p = 140, 150
    X = rng.normal(size=(N, p))
    ycls = (rng.random(N) < 0.45).astype(float)
    X[:, 0] += 1.5 * ycls; X[:, 7] = 1.0
    X = np.asfortranarray(X)
    types = ["match-mismatch", "numeric-abs", "numeric-sqr", "numeric-abs"][:q]
    C = np.asfortranarray(np.column_stack([rng.integers(0, 2, N).astype(float), rng.normal(40, 12, N).round(), rng.normal(size=N), rng.normal(size=N)])[:, :q])
    ri, nn, _, _ = O.nearest_neighbors(X)
    yd = O.pair_diffs(ycls, ri, nn, "match-mismatch")
    cd = np.column_stack([O.pair_diffs(C[:, c], ri, nn, types[c]) for c in range(q)])
    exp = O.regress_attrs(X, ri, nn, yd, cd, "numeric-abs", "binomial")
    res = {}
    for name, env in (("fold", {}), ("nofold", {"NPDRB_NO_COVFOLD": "1"}), ("nospec", {"NPDRB_NO_SPECULATE": "1"})):
        os.environ.update(env)
        _, info = nb.npdr(ycls, X, "binomial", "numeric-abs", covars=C, covar_diff_type=types, return_info=True)
        for k in env: os.environ.pop(k)
        st = info["stats"]
        bad = np.where(np.isnan(st[:, 2]) != np.isnan(exp[:, 2]))[0]
        with np.errstate(all="ignore"):
            rel = np.nanmax(np.abs(st[:, 1] - exp[:, 1]) / np.maximum(np.abs(exp[:, 1]), 1e-3))
        res[name] = (bad.tolist(), float(rel))
        if len(bad):
            a = bad[0]
            print(f"  q={q} seed={seed} {name}: attr {a} gpu={st[a]} oracle={exp[a]}")
    if any(r[0] or r[1] > 1e-6 for r in res.values()):
        print(f"q={q} seed={seed}: {res}")
print("done")
