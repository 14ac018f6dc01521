"""Pin the oracle to the REAL reference when an R installation is available.

`Rscript` (with dplyr) is absent from the build container and from the GPU box, so this module skips there and the oracle stays
"parity unpinned" for distances / neighbour lists / statistics.  On a machine with R and a checkout of
insilico/npdr (NPDR_REFERENCE, default /root/reference) it runs scripts/ref_run.R -- the unmodified reference on BASELINE
configs 1-3 -- and compares the oracle with R's own output at the north-star tolerances.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_case, oracle_run

REF = os.environ.get("NPDR_REFERENCE", "/root/reference")


def _r_available():
    exe = shutil.which("Rscript")
    if not exe or not os.path.isdir(os.path.join(REF, "R")):
        return False
    r = subprocess.run([exe, "-e", "suppressPackageStartupMessages({library(dplyr); library(magrittr); library(rlang)})"],
                       capture_output=True)
    return r.returncode == 0


def test_probe_reports_why_the_oracle_is_unpinned():
    """Always runs: records (in the test log) whether the real reference could be executed here."""
    have = _r_available()
    print(f"Rscript: {shutil.which('Rscript')!r}; reference checkout: {os.path.isdir(os.path.join(REF, 'R'))}; runnable: {have}")
    assert have in (True, False)


@pytest.mark.skipif(not _r_available(), reason="no Rscript + dplyr + reference checkout on this machine (parity stays unpinned)")
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_oracle_matches_the_reference_run_in_r(tmp_path_factory, oracle, name):
    import pandas as pd
    out = tmp_path_factory.getbasetemp() / "ref_run"
    if not (out / "cfg3_npdr.csv").exists():
        subprocess.check_call(["Rscript", os.path.join(ROOT, "scripts", "ref_run.R"), REF, str(out)])
    case = load_case(name)
    exp = oracle_run(case)
    D = pd.read_csv(out / f"{name}_dist.csv").to_numpy(dtype=float)
    assert np.max(np.abs(D - exp["D"]) / np.maximum(np.abs(D), 1e-300)) <= 1e-12                     # distances: 1e-12 relative
    pairs = pd.read_csv(out / f"{name}_pairs.csv")
    assert np.array_equal(pairs["Ri_idx"].to_numpy(), exp["Ri"]) and np.array_equal(pairs["NN_idx"].to_numpy(), exp["NN"])
    res = pd.read_csv(out / f"{name}_npdr.csv")
    from npdr_b200 import rstats
    from npdr_b200.api import _pvalues
    pv, _ = _pvalues(exp["stats"], case["regression_type"])
    order = rstats.order(pv)
    assert list(res["att"]) == [case["names"][i] for i in order]                                      # identical ordering
    rtol = 1e-8 if case["regression_type"] == "lm" else 1e-6
    np.testing.assert_allclose(res["beta.raw.att"].to_numpy(), exp["stats"][order, 0], rtol=rtol)
    np.testing.assert_allclose(res["beta.Z.att"].to_numpy(), exp["stats"][order, 1], rtol=rtol)
    np.testing.assert_allclose(res["pval.att"].to_numpy(), pv[order], rtol=max(rtol, 1e-7), atol=1e-300)


@pytest.mark.gpu
@pytest.mark.skipif(not _r_available(), reason="no Rscript + dplyr + reference checkout on this machine")
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_gpu_matches_the_reference_run_in_r(tmp_path_factory, name):
    """The engine itself (through the C ABI) against R's own output -- the north-star bars: pair lists bit-exact, distances
    1e-12, beta / Z 1e-8 (lm) or 1e-6 (binomial), identical attribute ordering."""
    import pandas as pd
    import npdr_b200 as nb
    out = tmp_path_factory.getbasetemp() / "ref_run"
    if not (out / "cfg3_npdr.csv").exists():
        subprocess.check_call(["Rscript", os.path.join(ROOT, "scripts", "ref_run.R"), REF, str(out)])
    case = load_case(name)
    D = pd.read_csv(out / f"{name}_dist.csv").to_numpy(dtype=float)
    got = nb.npdrDistances(case["X"], case["nbd_metric"])
    assert np.max(np.abs(D - got) / np.maximum(np.abs(D), 1e-300)) <= 1e-12
    df = pd.DataFrame(case["X"], columns=case["names"])
    covars = case["C"] if case["C"] is not None else "none"
    res_gpu, info = nb.npdr(case["y"], df, case["regression_type"], "numeric-abs", case["nbd_method"], case["nbd_metric"], covars=covars,
                            covar_diff_type=list(case["covar_diff_type"]) or "match-mismatch", return_info=True)
    pairs = pd.read_csv(out / f"{name}_pairs.csv")
    assert np.array_equal(pairs["Ri_idx"].to_numpy(), info["Ri"]) and np.array_equal(pairs["NN_idx"].to_numpy(), info["NN"])
    res = pd.read_csv(out / f"{name}_npdr.csv")
    assert list(res["att"].astype(str)) == list(res_gpu["att"].astype(str))
    rtol = 1e-8 if case["regression_type"] == "lm" else 1e-6
    np.testing.assert_allclose(res_gpu["beta.raw.att"].to_numpy(), res["beta.raw.att"].to_numpy(), rtol=rtol)
    np.testing.assert_allclose(res_gpu["beta.Z.att"].to_numpy(), res["beta.Z.att"].to_numpy(), rtol=rtol)
