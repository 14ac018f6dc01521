import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def load_case(name):
    """Inputs + options of BASELINE configs 1-3 from the committed fixtures."""
    if name == "cfg1":
        g = np.load(os.path.join(GOLDEN, "case_control_train.npz"))
        return dict(X=g["X"], y=(g["class_codes"] == 2).astype(float), C=None, names=list(g["attr_names"]),
                    regression_type="binomial", nbd_method="multisurf", nbd_metric="manhattan", covar_diff_type=())
    if name == "cfg2":
        g = np.load(os.path.join(GOLDEN, "qtrait_train.npz"))
        return dict(X=g["X"], y=g["qtrait"], C=None, names=list(g["attr_names"]),
                    regression_type="lm", nbd_method="multisurf", nbd_metric="euclidean", covar_diff_type=())
    if name == "cfg3":
        g = np.load(os.path.join(GOLDEN, "mdd_rnaseq_small.npz"))
        C = np.asfortranarray(np.column_stack([g["sex_codes"].astype(float), g["age"]]))
        return dict(X=g["X"], y=(g["pheno_codes"] == 2).astype(float), C=C, names=list(g["attr_names"]),
                    regression_type="binomial", nbd_method="relieff", nbd_metric="manhattan",
                    covar_diff_type=("match-mismatch", "numeric-abs"))
    raise KeyError(name)


def oracle_run(case, sd_frac=0.5, k=0):
    """Full path through the CPU oracle -> dict(Ri, NN, D, radius, stats, n_unique)."""
    from oracle import oracle as O
    ri, nn, D, r = O.nearest_neighbors(case["X"], case["nbd_method"], case["nbd_metric"], sd_frac, k)
    ytype = "numeric-abs" if case["regression_type"] == "lm" else "match-mismatch"
    yd = O.pair_diffs(case["y"], ri, nn, ytype)
    cd = None
    if case["C"] is not None:
        cd = np.column_stack([O.pair_diffs(case["C"][:, c], ri, nn, t) for c, t in enumerate(case["covar_diff_type"])])
    stats = O.regress_attrs(case["X"], ri, nn, yd, cd, "numeric-abs", case["regression_type"])
    nu, _ = O.unique_neighbors(ri, nn, case["X"].shape[0])
    return dict(Ri=ri, NN=nn, D=D, radius=r, stats=stats, n_unique=nu)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
