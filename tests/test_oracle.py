"""CPU tests of the oracle (oracle/npdr_oracle.c): the reference's own known-answer vectors, the
committed golden outputs, and cross-checks against independent numpy/scipy implementations."""
import os

import numpy as np
import pytest
from scipy.spatial.distance import cdist

import sys

from conftest import GOLDEN, ROOT, load_case, oracle_run


# ---- the reference's only pinned vectors: tests/testthat/test-nearestNeighbors.R:1-12 ----
def test_npdrdiff_reference_kats(oracle):
    a = np.array([0, 1, 0, 1.0]); b = np.array([0, 1, 0, 4.0]); d = np.array([2, 1, 0, 1.0])
    a_mat = a.reshape(2, 2, order="F"); b_mat = b.reshape(2, 2, order="F")
    assert np.array_equal(oracle.diff(a, b), [0, 0, 0, 3])
    assert np.array_equal(oracle.diff(a_mat, b_mat, "correlation-data", 2), [0, 1.5])
    assert np.array_equal(oracle.diff(a, b, "numeric-sqr", 3), [0, 0, 0, 3])
    assert np.array_equal(oracle.diff(a, d, "allele-sharing"), [1, 0, 0, 0])
    assert np.array_equal(oracle.diff(a, d, "match-mismatch"), [1, 0, 0, 0])


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_oracle_matches_committed_golden(oracle, name):
    exp = np.load(os.path.join(GOLDEN, f"expected_{name}.npz"))
    r = oracle_run(load_case(name))
    assert np.array_equal(r["Ri"], exp["Ri"]) and np.array_equal(r["NN"], exp["NN"])
    assert r["n_unique"] == int(exp["n_unique"])
    np.testing.assert_allclose(r["stats"], exp["stats"], rtol=1e-12, atol=0)
    assert np.array_equal(r["D"][:, 0], exp["D_row0"])


def test_survey_sanity_values(oracle):
    """SURVEY 8c orientation values (independent survey-time numpy restatement)."""
    r1 = oracle_run(load_case("cfg1"))
    c1 = load_case("cfg1")
    assert r1["Ri"].size == 3804
    i8 = c1["names"].index("simvar8")
    assert abs(r1["stats"][i8, 0] - 0.51583) < 1e-5 and abs(r1["stats"][i8, 1] - 13.520) < 1e-3
    r2 = oracle_run(load_case("cfg2"))
    assert r2["Ri"].size == 3164 and abs(r2["stats"][load_case("cfg2")["names"].index("simvar8"), 1] - 22.060) < 1e-3
    r3 = oracle_run(load_case("cfg3"))
    assert r3["Ri"].size == 7536 and abs(r3["stats"][load_case("cfg3")["names"].index("PTGER2"), 1] - 6.007) < 1e-3
    assert set(np.unique(r3["stats"][:, 6])) <= {3.0, 4.0}


def test_distances_vs_scipy(oracle):
    rng = np.random.default_rng(0)
    X = np.asfortranarray(rng.normal(size=(37, 53)))
    for metric, sp in (("manhattan", "cityblock"), ("euclidean", "euclidean")):
        D = oracle.distances(X, metric)
        np.testing.assert_allclose(D, cdist(X, X, sp), rtol=1e-13, atol=1e-13)
        assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    Xs = (X - X.min(0)) / (X.max(0) - X.min(0))
    np.testing.assert_allclose(oracle.distances(X, "relief-scaled-manhattan"), cdist(Xs, Xs, "cityblock"), rtol=1e-12)
    G = np.asfortranarray(rng.integers(0, 3, size=(20, 30)).astype(float))
    np.testing.assert_allclose(oracle.distances(G, "allele-sharing-manhattan"), cdist(G / 2, G / 2, "cityblock"), rtol=1e-14)


def test_radius_vs_numpy(oracle):
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.normal(size=(60, 40)))
    D = oracle.distances(X)
    r, sd = oracle.radius_multisurf(D, 0.5)
    N = D.shape[0]
    for c in (0, 17, 59):
        col = np.delete(D[:, c], c)
        assert abs(sd[c] - np.std(col, ddof=1)) < 1e-12 * sd[c]
        assert abs(r[c] - (D[:, c].sum() / (N - 1) - 0.5 * np.std(col, ddof=1))) < 1e-12 * abs(r[c])
    rs = oracle.radius_surf(D, 0.5)
    up = D[np.triu_indices(N, 1)]
    assert abs(rs - (D.sum() / (N * (N - 1)) - 0.5 * np.std(up, ddof=1))) < 1e-12 * rs


def test_neighbors_semantics(oracle):
    rng = np.random.default_rng(2)
    X = np.asfortranarray(rng.integers(0, 3, size=(40, 12)).astype(float))      # SNP-like: many exact ties
    X[7] = X[3]                                                                  # a zero-distance duplicate
    D = oracle.distances(X)
    r, _ = oracle.radius_multisurf(D, 0.5)
    ri, nn, _ = oracle.neighbors(D, "multisurf", radius=r)
    for c in range(40):
        want = [j + 1 for j in range(40) if D[j, c] < r[c] and D[j, c] > 0]
        assert list(nn[ri == c + 1]) == want
    k = 5
    ri, nn, _ = oracle.neighbors(D, "relieff", k=k)
    for c in range(40):
        col = D[:, c]
        rank = np.array([1 + np.sum(col < v) for v in col])                      # dplyr::min_rank
        keep = [j for j in range(40) if rank[j] <= k + 1]
        keep.sort(key=lambda j: (col[j], j))                                     # slice_min orders by value, ties by row
        want = [j + 1 for j in keep if col[j] > 0]
        assert list(nn[ri == c + 1]) == want
    assert 8 not in nn[ri == 4] and 4 not in nn[ri == 8]                         # duplicates dropped by filter(> 0)


def test_hitmiss_neighbors_semantics(oracle):
    """nearestNeighborsSeparateHitMiss restated independently in numpy (R/nearestNeighbors.R:318-551)."""
    rng = np.random.default_rng(8)
    N = 60
    X = np.asfortranarray(rng.normal(size=(N, 15)))
    y = (np.arange(N) % 3 == 0).astype(float)                                    # imbalanced 20 / 40
    D = oracle.distances(X)
    ri, nn, _, rh, rm = oracle.nearest_neighbors_hitmiss(X, y, "multisurf")
    for c in (0, 1, 7, 59):
        hits = [j for j in range(N) if j != c and y[j] == y[c]]; miss = [j for j in range(N) if y[j] != y[c]]
        eh = np.mean(D[hits, c]) - 0.5 * np.std(D[hits, c], ddof=1); em = np.mean(D[miss, c]) - 0.5 * np.std(D[miss, c], ddof=1)
        assert abs(rh[c] - eh) < 1e-12 * eh and abs(rm[c] - em) < 1e-12 * em
        want = [j + 1 for j in hits if 0 < D[j, c] < rh[c]] + [j + 1 for j in miss if D[j, c] < rm[c]]
        assert list(nn[ri == c + 1]) == want
    ri, nn, _, _, _ = oracle.nearest_neighbors_hitmiss(X, y, "relieff")
    for c in (0, 1, 59):
        order = sorted(range(N), key=lambda j: (D[j, c], j))
        hits = [j for j in order if y[j] == y[c]]; miss = [j for j in order if y[j] != y[c]]
        kh = oracle.knn_surf(len(hits) - 1, 0.5); km = oracle.knn_surf(len(miss), 0.5)
        assert list(nn[ri == c + 1]) == [j + 1 for j in hits[1:kh + 1]] + [j + 1 for j in miss[:km]]
    yb = (np.arange(N) % 2).astype(float)                                        # balanced: k = floor(knnSURF(N-1)/2) each
    ri, nn, _, _, _ = oracle.nearest_neighbors_hitmiss(X, yb, "relieff")
    k = int(np.floor(0.5 * oracle.knn_surf(N - 1, 0.5)))
    assert np.all(np.bincount(ri)[1:] == 2 * k)


def test_unique_neighbors_and_knnvec(oracle):
    ri = np.array([1, 1, 2, 2, 3, 3], dtype=np.int32); nn = np.array([2, 3, 1, 3, 1, 2], dtype=np.int32)
    n, keep = oracle.unique_neighbors(ri, nn, 3)
    # keys "1,2" "1,3" "1,2" "2,3" "1,3" "2,3" -> !duplicated
    assert n == 3 and list(keep) == [True, True, False, True, False, False]
    assert list(oracle.knn_vec(np.array([1, 1, 3], dtype=np.int32), 4)) == [2, 1]
    assert oracle.knn_surf(100) == 30 and oracle.knn_surf(157) == 48 and oracle.knn_surf(20000) == 6170


def test_lm_vs_lstsq(oracle):
    rng = np.random.default_rng(3)
    n = 500
    Xd = np.asfortranarray(np.column_stack([np.ones(n), rng.normal(size=n), rng.integers(0, 2, n), rng.normal(5, 2, n)]))
    y = Xd @ np.array([0.3, 0.7, -0.2, 0.05]) + rng.normal(size=n)
    coef, se, info = oracle.lm(Xd, y)
    b, *_ = np.linalg.lstsq(Xd, y, rcond=None)
    np.testing.assert_allclose(coef, b, rtol=1e-10)
    res = y - Xd @ b
    cov = np.linalg.inv(Xd.T @ Xd) * (res @ res / (n - 4))
    np.testing.assert_allclose(se, np.sqrt(np.diag(cov)), rtol=1e-10)
    assert info["rank"] == 4 and info["df_resid"] == n - 4
    assert abs(info["r_squared"] - (1 - res @ res / np.sum((y - y.mean()) ** 2))) < 1e-12
    # aliased column: coefficient NA, later columns still estimated (dqrdc2 limited pivoting)
    Xa = Xd.copy(); Xa[:, 1] = 0.0
    coef, se, info = oracle.lm(Xa, y)
    assert np.isnan(coef[1]) and info["rank"] == 3
    b3, *_ = np.linalg.lstsq(Xa[:, [0, 2, 3]], y, rcond=None)
    np.testing.assert_allclose(coef[[0, 2, 3]], b3, rtol=1e-10)


def test_glm_vs_newton(oracle):
    rng = np.random.default_rng(4)
    n = 800
    Xd = np.asfortranarray(np.column_stack([np.ones(n), rng.normal(size=n), rng.integers(0, 2, n)]))
    eta = Xd @ np.array([-0.4, 0.9, 0.5])
    y = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    coef, se, info = oracle.glm_binomial(Xd, y)
    b = np.zeros(3)
    for _ in range(50):                                            # plain Newton-Raphson to the MLE
        mu = 1 / (1 + np.exp(-Xd @ b)); W = mu * (1 - mu)
        b = b + np.linalg.solve(Xd.T @ (Xd * W[:, None]), Xd.T @ (y - mu))
    np.testing.assert_allclose(coef, b, rtol=1e-7)                 # IRLS stops at the 1e-8 deviance criterion
    mu = 1 / (1 + np.exp(-Xd @ coef)); W = mu * (1 - mu)
    np.testing.assert_allclose(se, np.sqrt(np.diag(np.linalg.inv(Xd.T @ (Xd * W[:, None])))), rtol=1e-5)
    assert info["converged"] == 1 and 3 <= info["iters"] <= 8
    dev = -2 * np.sum(y * np.log(mu) + (1 - y) * np.log(1 - mu))
    assert abs(info["deviance"] - dev) < 1e-8 * dev


def test_wide_designs_vs_numpy(oracle):
    """Designs with many covariate columns (up to 64 columns: R/npdr.R:315-345 has no limit): lm against lstsq, glm.fit
    against Newton-Raphson, and an aliased column in the middle of a wide design."""
    rng = np.random.default_rng(40)
    n = 1500
    for k in (8, 23, 41, 62):
        Xd = np.asfortranarray(np.column_stack([np.ones(n)] + [rng.normal(size=n) if c % 3 else rng.integers(0, 2, n) for c in range(k - 1)]))
        bt = rng.normal(scale=0.3, size=k)
        y = Xd @ bt + rng.normal(size=n)
        coef, se, info = oracle.lm(Xd, y)
        b, *_ = np.linalg.lstsq(Xd, y, rcond=None)
        np.testing.assert_allclose(coef, b, rtol=1e-8, atol=1e-11)
        res = y - Xd @ b
        np.testing.assert_allclose(se, np.sqrt(np.diag(np.linalg.inv(Xd.T @ Xd)) * (res @ res / (n - k))), rtol=1e-8)
        assert info["rank"] == k and info["df_resid"] == n - k
        yb = (rng.random(n) < 1 / (1 + np.exp(-Xd @ (bt * 0.5)))).astype(float)
        coef, se, info = oracle.glm_binomial(Xd, yb)
        bn = np.zeros(k)
        for _ in range(60):
            mu = 1 / (1 + np.exp(-Xd @ bn)); W = mu * (1 - mu)
            bn = bn + np.linalg.solve(Xd.T @ (Xd * W[:, None]), Xd.T @ (yb - mu))
        np.testing.assert_allclose(coef, bn, rtol=1e-6, atol=1e-8)
        assert info["converged"] == 1
        Xa = Xd.copy(); Xa[:, k // 2] = Xa[:, 1]                   # duplicate of an earlier column -> aliased
        coef, se, info = oracle.lm(Xa, y)
        assert np.isnan(coef[k // 2]) and info["rank"] == k - 1


def test_diff_regression_quirk3(oracle):
    """Constant attribute: NA without covariates; the first covariate's row with covariates."""
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.normal(size=(30, 3))); X[:, 1] = 2.0
    D = oracle.distances(X); r, _ = oracle.radius_multisurf(D)
    ri, nn, _ = oracle.neighbors(D, "multisurf", radius=r)
    y = rng.normal(size=30); c = rng.normal(size=30)
    yd = oracle.pair_diffs(y, ri, nn, "numeric-abs"); cd = oracle.pair_diffs(c, ri, nn, "numeric-abs")
    out = oracle.regress_attrs(X, ri, nn, yd, None, regression_type="lm")
    assert np.isnan(out[1, 0]) and int(out[1, 7]) & 1 and not np.isnan(out[0, 0])
    out2 = oracle.regress_attrs(X, ri, nn, yd, cd, regression_type="lm")
    Xc = np.column_stack([np.ones(ri.size), cd])
    b, *_ = np.linalg.lstsq(Xc, yd, rcond=None)
    assert int(out2[1, 7]) & 1 and abs(out2[1, 0] - b[1]) < 1e-10 and abs(out2[1, 2] - b[0]) < 1e-10


def test_rectangular_distances_and_neighbors2_against_numpy(oracle):
    """npdrDistances2 / nearestNeighbors2 (R/npdrLearner.R:323-482): oracle vs an independent numpy statement."""
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(11)
    X1 = rng.normal(size=(60, 9)); X2 = rng.normal(size=(17, 9))
    np.testing.assert_allclose(oracle.distances2(X1, X2, "manhattan"), cdist(X1, X2, "cityblock"), rtol=1e-13)
    np.testing.assert_allclose(oracle.distances2(X1, X2, "euclidean"), cdist(X1, X2, "euclidean"), rtol=1e-13)
    np.testing.assert_allclose(oracle.distances2(X1, X2, "allele-sharing-manhattan"), cdist(X1 / 2, X2 / 2, "cityblock"), rtol=1e-13)
    assert np.array_equal(oracle.distances2(X1, X2, "anything-else"), oracle.distances2(X1, X2, "manhattan"))
    D = cdist(X1, X2, "cityblock")
    # multisurf: colMeans - sd.frac * sd(column), all m1 entries; strict `<`, `> 0`, ascending index
    nn, r, _ = oracle.nearest_neighbors2(X1, X2, "multisurf")
    r_np = D.mean(0) - 0.5 * D.std(0, ddof=1)
    np.testing.assert_allclose(r, r_np, rtol=1e-13)
    for i in range(17):
        assert np.array_equal(nn[i], 1 + np.flatnonzero((D[:, i] < r[i]) & (D[:, i] > 0)))
    # surf: one radius from all m1 * m2 entries
    nn, r, _ = oracle.nearest_neighbors2(X1, X2, "surf")
    np.testing.assert_allclose(r, np.full(17, D.mean() - 0.5 * D.std(ddof=1)), rtol=1e-13)
    # relieff: k + 1 nearest (no self distance), nearest first
    k = int(oracle.lib().orc_knn_surf(60, 0.5))
    nn, _, _ = oracle.nearest_neighbors2(X1, X2, "relieff")
    for i in range(17):
        assert np.array_equal(nn[i], 1 + np.argsort(D[:, i], kind="stable")[:k + 1])
    # a test instance that IS a training instance: its zero distance is filtered, leaving k
    X2[0] = X1[5]
    nn, _, _ = oracle.nearest_neighbors2(X1, X2, "relieff")
    assert len(nn[0]) == k and 6 not in nn[0]


def test_lm_glm_restatement_reproduces_published_r_output(oracle):
    """External pin: summary(lm) / summary(glm(binomial)) outputs that R prints for datasets::mtcars
    (tests/golden/r_published_kats.py) -- coefficients, standard errors, R^2 / residual deviance, residual df and the
    number of Fisher scoring iterations."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import r_published_kats as K
    for name, (yv, cols, fam, coef_p, se_p, extra) in K.KATS.items():
        design = np.asfortranarray(np.column_stack([np.ones(yv.size)] + cols))
        coef, se, info = (oracle.lm if fam == "lm" else oracle.glm_binomial)(design, yv)
        for v, pr in zip(coef, coef_p):
            assert K.printed_close(v, pr), (name, "coef", v, pr)
        for v, pr in zip(se, se_p):
            assert K.printed_close(v, pr), (name, "se", v, pr)
        assert info["df_resid"] == extra["df"]
        if fam == "lm":
            assert K.printed_close(info["r_squared"], extra["r_squared"]), name
        else:
            assert K.printed_close(info["deviance"], extra["deviance"]) and info["iters"] == extra["iters"] and info["converged"], name
