import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import npdr_b200 as nb
from oracle import oracle as O
O.build()
reg, q, diff = "binomial", 0, "allele-sharing"
rng = np.random.default_rng(zlib.crc32(f"{reg}/{q}/{diff}".encode()))
N, p = 140, 150
X = rng.integers(0, 3, size=(N, p)).astype(float)
ycls = (rng.random(N) < 0.45).astype(float)
X[:, 0] += 1.5 * ycls; X[:, 7] = 1.0
X = np.asfortranarray(X)
ri, nn, _, _ = O.nearest_neighbors(X)
yd = O.pair_diffs(ycls, ri, nn, "match-mismatch")
exp = O.regress_attrs(X, ri, nn, yd, None, diff, reg)
np.set_printoptions(precision=12, linewidth=200)
for env in ({}, {"NPDRB_NO_SPECULATE": "1"}, {"NPDRB_NO_FOLD": "1"}, {"NPDRB_NO_SPECULATE": "1", "NPDRB_NO_FOLD": "1"}, {"NPDRB_PASS_KERNEL": "v1"}):
    os.environ.update(env)
    _, info = nb.npdr(ycls, X, reg, diff, return_info=True)
    for k in env: os.environ.pop(k)
    st = info["stats"]
    err = np.abs(st[:, 1] - exp[:, 1]) / np.maximum(np.abs(exp[:, 1]), 1e-300)
    worst = np.argsort(-np.nan_to_num(err))[:3]
    print(env, "worst attrs", worst, err[worst])
    print("  gpu   ", st[worst[0]]); print("  oracle", exp[worst[0]])
