"""CPU tests: the C-ABI library loads and exports everything include/npdr_b200.h declares, fails loudly
without a GPU (no CPU fallback), host-side device helpers (xf80, dmath) are exact, R-side residue."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT, gpu_available


def test_header_symbols_exported():
    from npdr_b200 import _lib
    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "npdr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(npdrb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf" T {name}\b", nm), f"{name} is not an extern \"C\" text symbol"


def test_struct_layouts_match_header():
    from npdr_b200 import _lib
    import ctypes as C
    # sizes and a few offsets as the C compiler sees the header
    prog = r'''#include <stdio.h>
#include <stddef.h>
#include "npdr_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(npdrb_opts), sizeof(npdrb_result), offsetof(npdrb_opts, covar_diff_type),
  offsetof(npdrb_result, Ri), offsetof(npdrb_result, irls_passes)); return 0; }'''
    open("/tmp/nb_sizes.c", "w").write(prog)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "/tmp/nb_sizes.c", "-o", "/tmp/nb_sizes"])
    so, sr, oc, ori, oip = map(int, subprocess.run(["/tmp/nb_sizes"], capture_output=True, text=True).stdout.split())
    assert (C.sizeof(_lib.Opts), C.sizeof(_lib.Result)) == (so, sr)
    assert _lib.Opts.covar_diff_type.offset == oc and _lib.Result.Ri.offset == ori and _lib.Result.irls_passes.offset == oip
    o = _lib.Opts(); _lib.load().npdrb_opts_default(C.byref(o))
    assert (o.regression_type, o.nbd_method, o.nbd_metric, o.msurf_sd_frac, o.knn) == (1, 0, 0, 0.5, 0)
    assert list(o.covar_diff_type) == [3] * _lib.MAX_COVARS and _lib.MAX_COVARS == 60


def test_knn_surf_matches_reference_formula():
    import npdr_b200 as nb
    from scipy.special import ndtr
    for m, f in ((100, 0.5), (157, 0.5), (20000, 0.5), (99, 0.5), (300, 0.25), (50, 1.0)):
        erf = 2 * ndtr((f / np.sqrt(2)) * np.sqrt(2)) - 1
        assert nb.knnSURF(m, f) == int(np.floor((m - 1) * (1 - erf) / 2))
    assert nb.knnSURF(157) == 48 and nb.knnSURF(20000) == 6170


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_cpu_fallback():
    import npdr_b200 as nb
    X = np.random.default_rng(0).normal(size=(10, 4))
    with pytest.raises(nb.NpdrB200Error) as e:
        nb.npdrDistances(X)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(nb.NpdrB200Error):
        nb.npdr(np.arange(10) % 2, X)
    with pytest.raises(nb.NpdrB200Error):
        nb.npdrDiff([1.0], [2.0])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under npdr_b200/ may reference it."""
    for dp, _, files in os.walk(os.path.join(ROOT, "npdr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower().replace("npdr_oracle", "oracle") or f == "__none__", f"{f} mentions the oracle"


def test_unsupported_branches_raise():
    import npdr_b200 as nb
    X = np.zeros((6, 3)); y = np.arange(6) % 2
    with pytest.raises(NotImplementedError):
        nb.npdr(y, X, use_glmnet=True)
    with pytest.raises(NotImplementedError):
        nb.npdr(y, X, attr_diff_type="correlation-data")
    with pytest.raises(NotImplementedError):
        nb.npdr(y, X, separate_hitmiss_nbds=True, rm_attr_from_dist=["1"])
    with pytest.raises(ValueError):
        nb.npdr(y, X, regression_type="poisson")
    with pytest.raises(ValueError):
        nb.npdrDiff([1], [2], "nonsense")


def _compile_run(src, name):
    exe = f"/tmp/{name}"
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "npdr_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpu", src), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_xf80_matches_x87_long_double():
    out = _compile_run("test_xf80.cpp", "test_xf80")
    assert "fails=0" in out


def test_dmath_accuracy():
    out = _compile_run("test_dmath.cpp", "test_dmath")
    assert "max_rel_exp" in out


def test_rstats_p_adjust_and_pt():
    from npdr_b200 import rstats
    p = np.array([0.01, 0.04, 0.03, 0.5, np.nan, 0.2])
    np.testing.assert_allclose(rstats.p_adjust(p, "bonferroni")[[0, 1, 2, 3, 5]], np.minimum(1, 5 * p[[0, 1, 2, 3, 5]]))
    assert np.isnan(rstats.p_adjust(p, "bonferroni")[4])
    q = np.array([0.01, 0.02, 0.03, 0.04, 0.05])
    np.testing.assert_allclose(rstats.p_adjust(q, "BH"), [0.05] * 5)
    np.testing.assert_allclose(rstats.p_adjust(q, "holm"), [0.05, 0.08, 0.09, 0.09, 0.09])
    np.testing.assert_allclose(rstats.p_adjust(q, "hochberg"), [0.05] * 5)
    np.testing.assert_allclose(rstats.p_adjust(q, "BY"), np.minimum(1, 0.05 * np.sum(1 / np.arange(1, 6))) * np.ones(5))
    np.testing.assert_allclose(rstats.p_adjust(q, "hommel"), [0.05] * 5)
    np.testing.assert_allclose(rstats.p_adjust(q, "fdr"), rstats.p_adjust(q, "BH"))
    assert abs(rstats.pt(2.0, 10, lower_tail=False) - stats.t.sf(2.0, 10)) < 1e-15
    assert abs(rstats.pt(3.0, 1e6, lower_tail=False) - stats.norm.sf(3.0)) < 1e-6 * stats.norm.sf(3.0) * 50
    assert list(rstats.order([0.3, np.nan, 0.1, 0.1])) == [2, 3, 0, 1]
    np.testing.assert_allclose(rstats.format_fixed([1.234567, 0.00123456789]), [1.2345670, 0.0012346])
    np.testing.assert_allclose(rstats.format_sci([1.23456789e-5, 0.5]), [1.2346e-5, 0.5])


def test_row_and_attr_partitions():
    from npdr_b200.distributed import partition_attrs, partition_rows
    for N in (100, 157, 4000, 20000, 129):
        for w in (1, 2, 3, 4, 8):
            parts = partition_rows(N, w)
            assert len(parts) == w and parts[0][0] == 0 and sum(n for _, n in parts) == N
            for (r0, n), (r1, _) in zip(parts, parts[1:]):
                assert r0 + n == r1 and (n == 0 or r0 % 128 == 0)
    assert partition_attrs(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]


def test_r_shim_compiles_against_api_stub():
    """R is absent from the image, so the .Call shim cannot be built; it is at least type-checked against a stub of the
    R C API it uses and against include/npdr_b200.h (every C entry point it binds exists with that signature)."""
    import subprocess
    stub = os.path.join(ROOT, "tests", "cpu", "r_api_stub")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", stub, "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "r-package", "src", "npdrb200_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    shim = open(os.path.join(ROOT, "r-package", "src", "npdrb200_shim.c")).read()
    wrapper = open(os.path.join(ROOT, "r-package", "R", "npdr_b200.R")).read()
    import re
    registered = set(re.findall(r'\{"(C_npdrb_\w+)"', shim))
    called = set(re.findall(r'\.Call\("(C_npdrb_\w+)"', wrapper))
    assert called <= registered and len(registered) >= 8, (called - registered)


def test_c_caller_links_and_fails_loudly_without_gpu():
    """examples/npdr_c_abi.c: a plain C program links against the shared library through include/npdr_b200.h alone and,
    on a machine without a B200, stops with NPDRB_ENODEVICE (exit 2) -- on the GPU box it runs the path (exit 0)."""
    import subprocess
    exe = "/tmp/npdr_c_abi_test"
    lib_dir = os.path.join(ROOT, "npdr_b200")
    subprocess.check_call(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "npdr_c_abi.c"),
                           "-L", lib_dir, "-lnpdr_b200", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    if gpu_available():
        assert r.returncode == 0 and "pairs" in r.stdout, r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


def test_npdr_learner_vote_and_level_coding():
    """Host-side logic of npdrLearner (R/npdrLearner.R:208-229, 267-283): level coding and the majority vote."""
    from npdr_b200 import api
    codes, lv, numeric = api._learner_codes(np.array([1.0, -1.0, 1.0, -1.0]))
    assert lv == ["-1", "1"] and numeric and codes.tolist() == [1.0, -1.0, 1.0, -1.0]
    codes, lv, numeric = api._learner_codes(np.array(["MDD", "HC", "HC"]))
    assert lv == ["HC", "MDD"] and not numeric and codes.tolist() == [2.0, 1.0, 1.0]          # levels become 1, 2 in level order
    assert api._r_levels(np.array([10, 9, 9])) == ["9", "10"]                                  # numeric, not lexical
    train = np.array([1.0, -1.0, 1.0, -1.0, -1.0])
    pred = api._knn_vote(train, [np.array([1, 2, 3]), np.array([1, 2]), np.array([2, 4, 5, 1]), np.array([], dtype=np.int32)])
    assert pred[:3].tolist() == [1.0, -1.0, -1.0] and np.isnan(pred[3])                         # tie -> smallest code (which.max of table())
