import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
from conftest import load_case
import npdr_b200 as nb
for name, reg in (("cfg1", "binomial"), ("cfg1", "lm"), ("cfg3", "binomial")):
    c = load_case(name)
    out = {}
    for kern in ("v2", "v3", "v3dbg1"):
        os.environ["NPDRB_PASS_KERNEL"] = kern[:2]
        os.environ["NPDRB_V3_DBG"] = "1" if kern.endswith("dbg1") else "0"
        y = c["y"] if reg == "binomial" else c["X"][:, 0] * 0.3 + np.arange(c["X"].shape[0]) % 7
        st, info = nb.npdr_stats(c["X"], y, c["C"], reg, "numeric-abs", c["nbd_method"], c["nbd_metric"], covar_diff_type=c["covar_diff_type"])
        out[kern] = st
    print(name, reg, "v3dbg1 vs v2 max abs diff", np.abs(out["v2"] - out["v3dbg1"]).max(axis=0))
    d = np.abs(out["v2"] - out["v3"])
    print(name, reg, "max abs diff per column", d.max(axis=0), "rows differing", int((d.max(axis=1) > 1e-9).sum()), "of", d.shape[0])
    bad = np.flatnonzero(d.max(axis=1) > 1e-9)[:12]
    print(" first bad attrs", bad)
    if bad.size:
        print(out["v2"][bad[0]], out["v3"][bad[0]])
