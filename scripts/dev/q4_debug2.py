import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import npdr_b200 as nb
from oracle import oracle as O
O.build()
def run(q, env, seed=35, types_all=("match-mismatch", "numeric-abs", "numeric-sqr", "numeric-abs")):
    rng = np.random.default_rng(seed)
    N, p = 140, 150
    X = rng.normal(size=(N, p)); ycls = (rng.random(N) < 0.45).astype(float)
    X[:, 0] += 1.5 * ycls; X[:, 7] = 1.0
    X = np.asfortranarray(X)
    types = list(types_all[:q])
    C = np.asfortranarray(np.column_stack([rng.integers(0, 2, N).astype(float), rng.normal(40, 12, N).round(), rng.normal(size=N), rng.normal(size=N)])[:, :q])
    ri, nn, _, _ = O.nearest_neighbors(X)
    yd = O.pair_diffs(ycls, ri, nn, "match-mismatch")
    cd = np.column_stack([O.pair_diffs(C[:, c], ri, nn, types[c]) for c in range(q)])
    exp = O.regress_attrs(X, ri, nn, yd, cd, "numeric-abs", "binomial")
    os.environ.update(env)
    _, info = nb.npdr(ycls, X, "binomial", "numeric-abs", covars=C, covar_diff_type=types, return_info=True)
    for k in env: os.environ.pop(k)
    st = info["stats"]
    bad = np.where(np.isnan(st[:, 2]) != np.isnan(exp[:, 2]))[0]
    print(f"q={q} env={env} types={types}: {len(bad)} attrs with NaN-pattern mismatch in beta0; df gpu {st[0,4]} oracle {exp[0,4]}; iters gpu {st[:5,6]} oracle {exp[:5,6]}")
    if len(bad): print("   first:", st[bad[0]], exp[bad[0]])
B = {"NPDRB_NO_COVFOLD": "1"}
run(2, {**B, "NPDRB_NO_SPECULATE": "1"})
run(2, {**B, "NPDRB_NO_SPECULATE": "1", "NPDRB_PASS_KERNEL": "v1"})
run(3, {**B, "NPDRB_NO_SPECULATE": "1"})
run(4, {**B, "NPDRB_NO_SPECULATE": "1"})
run(4, {**B, "NPDRB_NO_SPECULATE": "1"}, types_all=("numeric-abs",) * 4)
run(4, {**B, "NPDRB_NO_SPECULATE": "1", "NPDRB_NO_FOLD": "1"})
run(3, {"NPDRB_NO_SPECULATE": "1"})
run(4, {"NPDRB_NO_SPECULATE": "1"})
