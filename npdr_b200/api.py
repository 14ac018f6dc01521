"""Host-side mirror of the reference's operator API for the NPDR hot path.

Same function names, argument meaning (snake_case spellings of the R formals) and error behaviour
as the exported R closures of insilico/npdr; all arithmetic runs in libnpdr_b200.so on the GPU,
the p-value / ordering / p.adjust residue in `rstats` (it stays in R in the R deployment).

  npdr               R/npdr.R:151-634 (default branch; glmnet / correlation-data / separate hit-miss: rejected)
  npdrDiff           R/nearestNeighbors.R:15-40
  npdrDistances      R/nearestNeighbors.R:63-101
  nearestNeighbors   R/nearestNeighbors.R:153-285
  nearestNeighborsSeparateHitMiss   R/nearestNeighbors.R:318-551
  npdrDistances2     R/npdrLearner.R:323-339
  nearestNeighbors2  R/npdrLearner.R:378-482
  npdrLearner        R/npdrLearner.R:204-292 (host-side vote on nearestNeighbors2)
  uniqueNeighbors    R/nearestNeighbors.R:572-583
  knnVec             R/nearestNeighbors.R:603-607
  knnSURF            R/utils.R:12-18
  diffRegression     R/npdr.R:18-71
  uniReg             R/utils.R:66-157
  vwok               R/adaptive.R:39-234 (batched-k caller: one distance matrix for the whole k grid)
"""
import ctypes as C
import sys
import time

import numpy as np

from . import _lib, rstats
from ._lib import DIFF, METRIC, NBD, REG, STATS_PER_ATTR, check, context, dptr, f64, iptr

try:  # pandas is only used for the data-frame shaped inputs/outputs the reference has
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None


def _as_matrix(dataset):
    """-> (float64 Fortran matrix, column names)."""
    if pd is not None and isinstance(dataset, pd.DataFrame):
        names = [str(c) for c in dataset.columns]
        X = np.asfortranarray(dataset.to_numpy(dtype=np.float64))
        return X, names
    X = np.asarray(dataset, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    return np.asfortranarray(X), [str(i + 1) for i in range(X.shape[1])]


def _code_outcome(pheno, regression_type):
    """Phenotype -> float64 vector.  binomial: any two-level labels are mapped to their factor codes
    (only equality matters for the hit/miss diff, R/npdr.R:305); lm: numeric."""
    a = np.asarray(pheno)
    if regression_type == "binomial" or a.dtype.kind in "OUSb":
        _, inv = np.unique(a, return_inverse=True)
        return inv.astype(np.float64)
    return a.astype(np.float64)


def _covars_matrix(covars, covar_diff_type):
    """R/npdr.R:315-345: active iff length(covars) > 1; num.covs = length(covar.diff.type) (quirk 2)."""
    if covars is None or (isinstance(covars, str) and covars == "none"):
        return None, [], []
    if isinstance(covar_diff_type, str):
        covar_diff_type = [covar_diff_type]
    names = None
    if pd is not None and isinstance(covars, (pd.DataFrame, pd.Series)):
        names = [str(c) for c in (covars.columns if isinstance(covars, pd.DataFrame) else [covars.name or "cov1"])]
        covars = covars.to_numpy()
    Cm = np.asarray(covars)
    if Cm.ndim == 1:
        Cm = Cm.reshape(-1, 1)
    if Cm.dtype.kind in "OUS":
        raise TypeError("covars must be numeric (as.matrix() on a mixed data.frame yields characters in R too); "
                        "code factors as integers first")
    ncov = len(covar_diff_type)
    if ncov > Cm.shape[1]:
        raise ValueError("covar_diff_type has more entries than covars has columns")
    if ncov > _lib.MAX_COVARS:
        raise ValueError(f"at most {_lib.MAX_COVARS} covariates are supported")
    Cm = np.asfortranarray(Cm[:, :ncov].astype(np.float64))
    if names is None:
        names = [f"cov{i + 1}" for i in range(ncov)]
    for t in covar_diff_type:
        if t not in DIFF or t == "raw":
            raise ValueError(f"unknown covar.diff.type {t!r}")
    return Cm, list(covar_diff_type), names[:ncov]


def knnSURF(m_samples, sd_frac=0.5):
    """Theoretical expected number of neighbours (R/utils.R:12-18)."""
    return int(_lib.load().npdrb_knn_surf(int(m_samples), float(sd_frac)))


def npdrDiff(a, b, diff_type="manhattan", norm_fac=1, device=0):
    """Projected difference of two value vectors (R/nearestNeighbors.R:15-40)."""
    if diff_type is None:
        diff_type = "numeric-abs"
    if diff_type == "correlation-data":
        raise NotImplementedError("npdrDiff(diff.type='correlation-data') is outside the accelerated path (DESIGN.md)")
    if diff_type not in DIFF or diff_type == "raw":
        raise ValueError(f"'arg' should be one of the npdrDiff diff types, got {diff_type!r}")   # match.arg error
    a = f64(np.ravel(a)); b = f64(np.ravel(b))
    if a.size != b.size:
        raise ValueError("a and b must have the same length")
    out = np.empty(a.size)
    check(_lib.load().npdrb_diff(context(device).handle, dptr(a), dptr(b), a.size, DIFF[diff_type], float(norm_fac), dptr(out)))
    return out


def npdrDistances(attr_mat, metric="manhattan", fast_dist=False, device=0, fast_euclidean=False):
    """m x m distance matrix (R/nearestNeighbors.R:63-101).  fast_dist is accepted for signature
    compatibility (it selects wordspace::dist.matrix in the reference)."""
    X, _ = _as_matrix(attr_mat)
    if metric not in METRIC or metric == "precomputed":
        return None        # the reference falls through and returns NULL for unknown metrics (:97-100)
    N, p = X.shape
    D = np.empty((N, N), order="F")
    check(_lib.load().npdrb_distances(context(device).handle, dptr(X), N, p, METRIC[metric], int(bool(fast_euclidean)), dptr(D)))
    return D


def npdrDistances2(attr_mat1, attr_mat2, metric="manhattan", device=0):
    """R/npdrLearner.R:323-339 -- m1 x m2 distances between the rows of two attribute matrices (flexclust::dist2).
    As in the reference, any metric other than 'euclidean' / 'allele-sharing-manhattan' means manhattan."""
    X1, _ = _as_matrix(attr_mat1)
    X2, _ = _as_matrix(attr_mat2)
    if X1.shape[1] != X2.shape[1]:
        raise ValueError("attr.mat1 and attr.mat2 must have the same attributes (columns)")
    m = metric if metric in ("euclidean", "allele-sharing-manhattan") else "manhattan"
    D = np.empty((X1.shape[0], X2.shape[0]), order="F")
    check(_lib.load().npdrb_distances2(context(device).handle, dptr(X1), X1.shape[0], dptr(X2), X2.shape[0], X1.shape[1], METRIC[m], dptr(D)))
    return D


def nearestNeighbors2(attr_mat1, attr_mat2, nbd_method="multisurf", nbd_metric="manhattan", sd_vec=None, sd_frac=0.5,
                      dopar_nn=False, k=0, device=0, return_info=False):
    """R/npdrLearner.R:378-482 -- for every instance of attr_mat2 (test) its neighbours among the instances of attr_mat1
    (train): a list of m2 integer arrays of 1-based train indices (relieff: nearest first, the k + 1 nearest because a
    test instance has no zero self-distance to drop; surf / multisurf: ascending index)."""
    if nbd_method not in NBD:
        raise ValueError(f"unknown nbd.method {nbd_method!r}")
    X1, _ = _as_matrix(attr_mat1)
    X2, _ = _as_matrix(attr_mat2)
    if X1.shape[1] != X2.shape[1]:
        raise ValueError("attr.mat1 and attr.mat2 must have the same attributes (columns)")
    m1, m2 = X1.shape[0], X2.shape[0]
    m = nbd_metric if nbd_metric in ("euclidean", "allele-sharing-manhattan") else "manhattan"
    sv = f64(sd_vec) if sd_vec is not None else None
    if sv is not None and sv.size != m2:
        raise ValueError("sd.vec must have one entry per instance of attr.mat2")
    off = C.POINTER(C.c_int64)(); idx = C.POINTER(C.c_int32)(); M = C.c_int64()
    radius = np.full(m2, np.nan)
    check(_lib.load().npdrb_nearest_neighbors2(context(device).handle, dptr(X1), m1, dptr(X2), m2, X1.shape[1], NBD[nbd_method], METRIC[m],
                                          dptr(sv), float(sd_frac), int(k), C.byref(off), C.byref(idx), C.byref(M), dptr(radius)))
    offsets = np.ctypeslib.as_array(off, shape=(m2 + 1,)).copy()
    _lib.load().npdrb_free(C.cast(off, C.c_void_p))
    nn = _lib.take_int_array(idx, M.value)
    out = [nn[offsets[i]:offsets[i + 1]] for i in range(m2)]
    return (out, {"radius": radius, "offsets": offsets}) if return_info else out


def _r_levels(values):
    """levels(as.factor(x)): sorted unique values as strings -- numbers sort numerically, everything else as strings."""
    vals = list(dict.fromkeys(np.asarray(values).tolist()))
    try:
        return [str(int(v)) if float(v) == int(float(v)) else repr(float(v)) for v in sorted(vals, key=float)]
    except (TypeError, ValueError):
        return sorted(str(v) for v in vals)


def _learner_codes(values):
    """R/npdrLearner.R:208-229 -- class levels that look numeric are used as numbers, other levels become 1, 2, ... in
    level order.  -> (codes float array, levels as strings, numeric_looking)."""
    lv = _r_levels(values)
    as_str = [str(int(v)) if isinstance(v, (int, np.integer)) or (isinstance(v, (float, np.floating)) and float(v) == int(v))
              else str(v) for v in np.asarray(values).tolist()]
    try:
        num = [float(l) for l in lv]
        lut = dict(zip(lv, num))
        return np.array([lut[a] if a in lut else float(a) for a in as_str]), lv, True
    except ValueError:
        lut = {l: float(i + 1) for i, l in enumerate(lv)}
        return np.array([lut[a] for a in as_str]), lv, False


def _knn_vote(train_codes, neighbor_lists):
    """R/npdrLearner.R:267-276 -- table(train.pheno[nbs]) %>% which.max %>% names %>% as.numeric: the most frequent class
    among the neighbours, ties going to the smallest class code (table() sorts its names, which.max takes the first
    maximum).  An empty neighbourhood predicts NaN (the reference's sapply() would return a ragged list there)."""
    pred = np.full(len(neighbor_lists), np.nan)
    for i, nbs in enumerate(neighbor_lists):
        if len(nbs) == 0:
            continue
        vals, counts = np.unique(train_codes[np.asarray(nbs, dtype=np.int64) - 1], return_counts=True)
        pred[i] = vals[np.argmax(counts)]
    return pred


def npdrLearner(train_outcome, train_data, test_outcome, test_data, nbd_method="relieff", nbd_metric="manhattan",
                msurf_sd_frac=0.5, dopar_nn=False, knn=0, device=0):
    """R/npdrLearner.R:204-292 -- nearest-neighbour classifier on NPDR neighbourhoods: every test instance takes the
    majority class of its nearestNeighbors2() neighbours in the training set (GPU), the vote and the bookkeeping are host
    side.  Returns the reference's list: train_nbs_of_test, actual, confusion (predicted x actual counts), prediction,
    accuracy.  `prediction` is mapped back to the original labels exactly as the reference's replace() loop does it
    (codes equal to a level INDEX are replaced, R/npdrLearner.R:279-283 -- with numeric-looking levels such as -1 / 1 that
    relabels class 1 as the first level; `prediction.codes` holds the untouched numeric prediction).  With separate
    outcome vectors the reference stops at that loop (its level table is undefined); here the codes are returned."""
    Xtr, _, ptr = _split_outcome(train_outcome, train_data)
    Xte, _, pte = _split_outcome(test_outcome, test_data)
    named = np.ndim(train_outcome) == 0 and np.ndim(test_outcome) == 0
    tr_codes, _, _ = _learner_codes(ptr) if np.ndim(train_outcome) == 0 else (np.asarray(ptr, dtype=np.float64), None, True)
    te_codes, te_levels, _ = _learner_codes(pte) if np.ndim(test_outcome) == 0 else (np.asarray(pte, dtype=np.float64), None, True)
    nbs = nearestNeighbors2(Xtr, Xte, nbd_method, nbd_metric, None, msurf_sd_frac, dopar_nn, knn, device)
    codes = _knn_vote(tr_codes, nbs)
    acc = float(np.sum(codes == te_codes) / te_codes.size)
    actual = [str(v) for v in np.asarray(pte).tolist()] if named else [repr(v) for v in te_codes.tolist()]
    pred = [("NA" if np.isnan(c) else (str(int(c)) if c == int(c) else repr(c))) for c in codes]
    if named:
        for i, lev in enumerate(te_levels, start=1):             # replace(test.predict, test.predict %in% i, org_test_levels[i])
            pred = [lev if (q_ == str(i)) else q_ for q_ in pred]
    conf = None
    if pd is not None:
        conf = pd.crosstab(pd.Series(pred, name="Test Predicted"), pd.Series(actual, name="Test Actual"))
    return {"train_nbs_of_test": nbs, "actual": actual, "confusion": conf, "prediction": pred, "prediction.codes": codes,
            "accuracy": acc}


def _nearest(X, nbd_method, nbd_metric, sd_vec, sd_frac, k, unique, device):
    if nbd_method not in NBD:
        raise ValueError(f"unknown nbd.method {nbd_method!r}")
    if nbd_metric not in METRIC:
        raise ValueError(f"unknown nbd.metric {nbd_metric!r}")
    X = f64(X)
    N, p = X.shape
    if nbd_metric == "precomputed" and N != p:
        raise ValueError("precomputed distances must be a square matrix")
    sv = f64(sd_vec) if sd_vec is not None else None
    Ri = C.POINTER(C.c_int32)(); NN = C.POINTER(C.c_int32)()
    M = C.c_int64(); nu = C.c_int64(); ne = C.c_int64()
    radius = np.full(N, np.nan)
    check(_lib.load().npdrb_nearest_neighbors(context(device).handle, dptr(X), N, p, NBD[nbd_method], METRIC[nbd_metric],
                                              dptr(sv), float(sd_frac), int(k), int(bool(unique)), C.byref(Ri), C.byref(NN),
                                              C.byref(M), C.byref(nu), C.byref(ne), dptr(radius)))
    ri = _lib.take_int_array(Ri, M.value)
    nn = _lib.take_int_array(NN, M.value)
    return ri, nn, {"n_unique": nu.value, "n_empty": ne.value, "radius": radius}


def nearestNeighbors(attr_mat, nbd_method="multisurf", nbd_metric="manhattan", sd_vec=None, sd_frac=0.5, k=0,
                     neighbor_sampling="none", att_to_remove=(), fast_dist=False, dopar_nn=False, device=0,
                     return_info=False):
    """(Ri_idx, NN_idx) neighbour-pair list, 1-based (R/nearestNeighbors.R:153-285)."""
    if nbd_metric != "precomputed" and att_to_remove is not None and len(att_to_remove):
        if pd is not None and isinstance(attr_mat, pd.DataFrame):
            try:
                attr_mat = attr_mat.drop(columns=list(att_to_remove))
            except KeyError:
                pass       # the reference swallows this error too (tryCatch, :183-187)
        else:
            Xa = np.asarray(attr_mat)
            keep = [c for c in range(Xa.shape[1]) if c not in set(att_to_remove)]
            attr_mat = Xa[:, keep]
    X, _ = _as_matrix(attr_mat)
    ri, nn, info = _nearest(X, nbd_method, nbd_metric, sd_vec, sd_frac, k, neighbor_sampling == "unique", device)
    if pd is not None:
        out = pd.DataFrame({"Ri_idx": ri, "NN_idx": nn})
    else:  # pragma: no cover
        out = np.column_stack([ri, nn])
    return (out, info) if return_info else out


def nearestNeighborsSeparateHitMiss(attr_mat, pheno_vec, nbd_method="relieff", nbd_metric="manhattan", sd_frac=0.5, k=0,
                                    neighbor_sampling="none", att_to_remove=(), fast_dist=False, dopar_nn=False, device=0,
                                    return_info=False):
    """Hit and miss neighbourhoods formed separately, hits listed first (R/nearestNeighbors.R:318-551).
    pheno_vec: two-class labels (the reference coerces them with as.numeric(as.character()))."""
    if nbd_method not in NBD:
        raise ValueError(f"unknown nbd.method {nbd_method!r}")
    if nbd_metric not in METRIC:
        raise ValueError(f"unknown nbd.metric {nbd_metric!r}")
    if att_to_remove is not None and len(att_to_remove):
        if pd is not None and isinstance(attr_mat, pd.DataFrame):
            try:
                attr_mat = attr_mat.drop(columns=list(att_to_remove))
            except KeyError:
                pass
        else:
            Xa = np.asarray(attr_mat)
            attr_mat = Xa[:, [c for c in range(Xa.shape[1]) if c not in set(att_to_remove)]]
    X, _ = _as_matrix(attr_mat)
    y = _code_outcome(pheno_vec, "binomial")
    if np.unique(y).size != 2:
        raise ValueError("nearestNeighborsSeparateHitMiss needs a two-class outcome")
    N, p = X.shape
    Ri = C.POINTER(C.c_int32)(); NN = C.POINTER(C.c_int32)()
    M = C.c_int64(); nu = C.c_int64()
    rh = np.full(N, np.nan); rm = np.full(N, np.nan)
    check(_lib.load().npdrb_nearest_neighbors_hitmiss(context(device).handle, dptr(X), N, p, dptr(f64(y)), NBD[nbd_method],
                                                      METRIC[nbd_metric], float(sd_frac), int(k), int(neighbor_sampling == "unique"),
                                                      C.byref(Ri), C.byref(NN), C.byref(M), C.byref(nu), dptr(rh), dptr(rm)))
    ri = _lib.take_int_array(Ri, M.value); nn = _lib.take_int_array(NN, M.value)
    out = pd.DataFrame({"Ri_idx": ri, "NN_idx": nn}) if pd is not None else np.column_stack([ri, nn])
    if return_info:
        return out, {"n_unique": nu.value, "radius_hit": rh, "radius_miss": rm}
    return out


def _pairs_arrays(neighbor_pairs):
    if pd is not None and isinstance(neighbor_pairs, pd.DataFrame):
        a = neighbor_pairs.iloc[:, 0].to_numpy(); b = neighbor_pairs.iloc[:, 1].to_numpy()
    else:
        arr = np.asarray(neighbor_pairs)
        a, b = arr[:, 0], arr[:, 1]
    return np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)


def uniqueNeighbors(neighbor_pairs, device=0):
    """Drop the second occurrence of every unordered pair (R/nearestNeighbors.R:572-583).
    The list must be grouped by ascending Ri_idx, as nearestNeighbors() returns it."""
    ri, nn = _pairs_arrays(neighbor_pairs)
    if ri.size and np.any(np.diff(ri) < 0):
        raise ValueError("uniqueNeighbors expects rows grouped by ascending Ri_idx")
    N = int(max(ri.max(initial=1), nn.max(initial=1)))
    s = _lib.Session(context(device), N, 1, 0)
    try:
        s.import_pairs_host(ri, nn)
        nu = s.unique_pairs(keep_unique_only=True)
        r2, n2 = s.copy_pairs(nu)
    finally:
        s.close()
    if pd is not None:
        return pd.DataFrame({"Ri_idx": r2, "NN_idx": n2})
    return np.column_stack([r2, n2])


def knnVec(neighbor_pairs_mat):
    """Number of neighbours of each instance that appears as Ri (R/nearestNeighbors.R:603-607).
    Pure bookkeeping on the pair list (dplyr::count in the reference)."""
    ri, _ = _pairs_arrays(neighbor_pairs_mat)
    if ri.size == 0:
        return np.zeros(0, dtype=np.int64)
    cnt = np.bincount(ri)
    return cnt[cnt > 0]


def _pvalues(stats, regression_type, dof=0):
    """stats (p x 8, see include/npdr_b200.h) -> pval.att, pval.0 as diffRegression forms them (R/npdr.R:55,63)."""
    z_a, z_0, df = stats[:, 1], stats[:, 3], stats[:, 4]
    res_df = df if not dof else np.full_like(df, float(dof))
    pval_att = rstats.pt(z_a, res_df, lower_tail=False)               # one-sided, t even for the logistic z (quirk 5)
    if regression_type == "lm":
        pval_0 = 2.0 * rstats.pt(-np.abs(z_0), df, lower_tail=True)    # summary.lm: two-sided t
    else:
        pval_0 = 2.0 * rstats.pnorm(-np.abs(z_0))                      # summary.glm: two-sided normal
    return pval_att, pval_0


def diffRegression(design_matrix_df, regression_type="binomial", fast_reg=False, dof=0, device=0):
    """One regression of pheno.diff.vec on attr.diff.vec (+ covariate diffs) (R/npdr.R:18-71).
    Returns (pval.att, beta.raw.att, beta.Z.att, beta.0, pval.0[, R.sqr])."""
    if pd is None or not isinstance(design_matrix_df, pd.DataFrame):
        raise TypeError("design_matrix_df must be a pandas DataFrame with a 'pheno.diff.vec' column")
    if "pheno.diff.vec" not in design_matrix_df.columns:
        raise ValueError("design.matrix.df must have a column named 'pheno.diff.vec'")
    if regression_type not in REG:
        raise ValueError(f"unknown regression.type {regression_type!r}")
    y = _code_outcome(design_matrix_df["pheno.diff.vec"].to_numpy(), regression_type)
    preds = design_matrix_df.drop(columns=["pheno.diff.vec"])
    P = np.asfortranarray(preds.to_numpy(dtype=np.float64))
    M, kk = P.shape
    if kk < 1:
        raise ValueError("design.matrix.df needs at least the attribute diff column")
    if kk - 1 > _lib.MAX_COVARS:
        raise ValueError(f"at most {_lib.MAX_COVARS} covariates are supported")
    X = np.asfortranarray(P[:, :1]); Cm = np.asfortranarray(P[:, 1:]) if kk > 1 else None
    out = np.empty((1, STATS_PER_ATTR))
    check(_lib.load().npdrb_unireg(context(device).handle, dptr(X), M, 1, dptr(f64(y)), dptr(Cm), kk - 1,
                                   REG[regression_type], dptr(out)))
    pa, p0 = _pvalues(out, regression_type, dof)
    if int(out[0, 7]) & 1 and kk == 1:
        print("Regression failure. Possible monomorphic SNP or variable with no variation.\n", file=sys.stderr)
    vec = [pa[0], out[0, 0], out[0, 1], out[0, 2], p0[0]]
    if regression_type == "lm":
        vec.append(out[0, 5])
    return np.array(vec)


def uniReg(outcome, dataset, regression_type="lm", padj_method="fdr", covars="none", device=0):
    """Univariate regression of the outcome on every raw attribute (R/utils.R:66-157):
    columns beta, beta.Z.att, pval, p.adj; rounded with format(digits = 5); sorted by pval."""
    X, names, pheno = _split_outcome(outcome, dataset)
    if regression_type not in REG:
        raise ValueError(f"unknown regression.type {regression_type!r}")
    y = _code_outcome(pheno, regression_type)
    Cm = None; q = 0
    if not (isinstance(covars, str) and covars == "none") and covars is not None:
        Cm = np.asarray(covars, dtype=np.float64)
        if Cm.ndim == 1:
            Cm = Cm.reshape(-1, 1)
        Cm = np.asfortranarray(Cm); q = Cm.shape[1]
    N, p = X.shape
    out = np.empty((p, STATS_PER_ATTR))
    check(_lib.load().npdrb_unireg(context(device).handle, dptr(X), N, p, dptr(f64(y)), dptr(Cm), q,
                                   REG[regression_type], dptr(out)))
    beta, stat, df = out[:, 0], out[:, 1], out[:, 4]
    bad = (out[:, 7].astype(int) & 1) == 1                      # aliased attribute -> NA row (R/utils.R:91-95)
    if regression_type == "lm":
        pval = 2.0 * rstats.pt(-np.abs(stat), df)
    else:
        pval = 2.0 * rstats.pnorm(-np.abs(stat))
    beta = np.where(bad, np.nan, beta); stat = np.where(bad, np.nan, stat); pval = np.where(bad, np.nan, pval)
    padj = rstats.format_sci(rstats.p_adjust(pval, padj_method), 5)
    betas = rstats.format_fixed(beta, 5)
    stats_f = rstats.format_fixed(stat, 5)
    pvals = rstats.format_sci(pval, 5)
    o = rstats.order(pvals)
    table = np.column_stack([betas, stats_f, pvals, padj])[o]
    if pd is not None:
        return pd.DataFrame(table, index=[names[i] for i in o], columns=["beta", "beta.Z.att", "pval", "p.adj"])
    return table


def _split_outcome(outcome, dataset):
    """R/npdr.R:166-173 / R/utils.R:68-80: outcome = column name / 1-based index, or a separate vector."""
    scalar = isinstance(outcome, (str, int, np.integer)) or (np.ndim(outcome) == 0)
    if scalar:
        if pd is not None and isinstance(dataset, pd.DataFrame):
            col = outcome if isinstance(outcome, str) else dataset.columns[int(outcome) - 1]
            pheno = dataset[col].to_numpy()
            X, names = _as_matrix(dataset.drop(columns=[col]))
        else:
            arr = np.asarray(dataset)
            j = int(outcome) - 1
            pheno = arr[:, j]
            X, names = _as_matrix(np.delete(arr, j, axis=1))
            names = [str(i + 1) for i in range(arr.shape[1]) if i != j]
    else:
        pheno = np.asarray(outcome)
        X, names = _as_matrix(dataset)
    if len(pheno) != X.shape[0]:
        raise ValueError("outcome and dataset have different numbers of instances")
    return X, names, pheno


def npdr_stats(X, y, Cm=None, regression_type="binomial", attr_diff_type="numeric-abs", nbd_method="multisurf",
               nbd_metric="manhattan", knn=0, msurf_sd_frac=0.5, covar_diff_type=(), neighbor_sampling="none",
               separate_hitmiss_nbds=False, return_pairs=False, device=0, devices=None):
    """The C-ABI call an R `.Call` makes: npdrb_npdr (one device) or npdrb_npdr_multi (`devices`: CUDA ordinals, one host
    thread per device inside the call) on HOST buffers -- column-major float64 X (N x p), coded outcome y, covariates Cm
    (N x q) -- from upload to the p x 8 statistics (include/npdr_b200.h).  -> (stats [p, 8], info dict).  Everything
    between npdr()'s argument parsing and its order()/p.adjust() (R/npdr.R:256-430)."""
    L = _lib.load()
    X = f64(X); N, p = X.shape
    ya = f64(y)
    Ca = f64(Cm) if Cm is not None else None
    cdt = [covar_diff_type] if isinstance(covar_diff_type, str) else list(covar_diff_type)
    o = _lib.Opts()
    L.npdrb_opts_default(C.byref(o))
    o.regression_type = REG[regression_type]; o.attr_diff_type = DIFF[attr_diff_type]
    o.nbd_method = NBD[nbd_method]; o.nbd_metric = METRIC[nbd_metric]
    o.knn = int(knn); o.msurf_sd_frac = float(msurf_sd_frac); o.n_covars = len(cdt) if Ca is not None else 0
    for i, t in enumerate(cdt[:o.n_covars]):
        o.covar_diff_type[i] = DIFF[t]
    o.neighbor_sampling_unique = int(neighbor_sampling == "unique")
    o.return_pairs = int(bool(return_pairs))
    o.separate_hitmiss_nbds = int(bool(separate_hitmiss_nbds))
    res = _lib.Result()
    if devices is not None:
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        check(L.npdrb_npdr_multi(int(devs.size), iptr(devs), dptr(X), N, p, dptr(ya), dptr(Ca), C.byref(o), C.byref(res)))
    else:
        check(L.npdrb_npdr(context(device).handle, dptr(X), N, p, dptr(ya), dptr(Ca), C.byref(o), C.byref(res)))
    stats = np.ctypeslib.as_array(res.stats, shape=(p, STATS_PER_ATTR)).copy()
    info = {f: getattr(res, f) for f in ("n_pairs", "n_unique_pairs", "n_pairs_used", "k_ave_empirical", "n_empty",
                                         "knn_used", "ms_h2d", "ms_distance", "ms_neighbors", "ms_prepare",
                                         "ms_regress", "ms_total", "irls_passes")}
    if return_pairs and res.n_pairs_used > 0:
        info["Ri"] = np.ctypeslib.as_array(res.Ri, shape=(res.n_pairs_used,)).copy()
        info["NN"] = np.ctypeslib.as_array(res.NN, shape=(res.n_pairs_used,)).copy()
    L.npdrb_result_release(C.byref(res))
    return stats, info


def npdr(outcome, dataset, regression_type="binomial", attr_diff_type="numeric-abs", nbd_method="multisurf",
         nbd_metric="manhattan", knn=0, msurf_sd_frac=0.5, covars="none", covar_diff_type="match-mismatch",
         padj_method="bonferroni", verbose=False, use_glmnet=False, glmnet_alpha=1, glmnet_lower=0,
         glmnet_lam="lambda.1se", rm_attr_from_dist=(), neighbor_sampling="none", separate_hitmiss_nbds=False,
         corr_attr_names=None, fast_reg=False, fast_dist=False, dopar_nn=False, dopar_reg=False, unique_dof=False,
         external_dist=None, device=0, return_info=False, k=None, n_gpus=1, devices=None):
    """Nearest-neighbour Projected-Distance Regression (R/npdr.R:151-634), default (non-glmnet) branch.

    Same formals as the reference (dots -> underscores); `k` is accepted as an alias of `knn` because R's
    partial argument matching lets the reference's demo scripts pass it.
    Returns a DataFrame with columns att, pval.adj, pval.att, beta.raw.att, beta.Z.att, beta.0, pval.0
    (+ R.sqr for lm), sorted by pval.att exactly as R/npdr.R:484-508 does.
    fast_reg / fast_dist / dopar_nn / dopar_reg select CPU libraries or fork pools in the reference; they are
    accepted and ignored (the GPU path is always used).  Unsupported branches raise instead of falling back.
    """
    if k is not None:
        knn = k
    if use_glmnet:
        raise NotImplementedError("use.glmnet=TRUE (npdrNET, R/npdr.R:513-632) is outside the accelerated path")
    if attr_diff_type == "correlation-data":
        raise NotImplementedError("attr.diff.type='correlation-data' (R/npdr.R:176-205) is outside the accelerated path")
    if separate_hitmiss_nbds and (nbd_metric == "precomputed" or (rm_attr_from_dist is not None and len(rm_attr_from_dist))):
        raise NotImplementedError("separate.hitmiss.nbds with precomputed distances / rm.attr.from.dist: use "
                                  "nearestNeighborsSeparateHitMiss() and the two-step path")
    if regression_type not in REG:
        raise ValueError(f"unknown regression.type {regression_type!r}")
    if attr_diff_type not in DIFF or attr_diff_type == "raw":
        raise ValueError(f"unknown attr.diff.type {attr_diff_type!r}")
    if nbd_method not in NBD:
        raise ValueError(f"unknown nbd.method {nbd_method!r}")
    if nbd_metric not in METRIC:
        raise ValueError(f"unknown nbd.metric {nbd_metric!r}")

    X, att_names, pheno = _split_outcome(outcome, dataset)
    y = _code_outcome(pheno, regression_type)
    N, p = X.shape
    Cm, cdt, _ = _covars_matrix(covars, covar_diff_type)
    q = len(cdt)
    _, class_counts = np.unique(np.asarray(pheno), return_counts=True)
    balanced_theoretical_k = knnSURF(2 * int(class_counts.min()) - 1, 0.5)

    L = _lib.load()
    ctx = context(device)
    t_start = time.time()
    info = {}
    simple = (nbd_metric != "precomputed") and not (rm_attr_from_dist is not None and len(rm_attr_from_dist))
    if verbose:
        print("Finding nearest neighbor pairs.")
    if simple:
        devs = None
        if (n_gpus and int(n_gpus) > 1) or devices is not None:   # one host process, one thread per device (what R would call)
            devs = np.arange(int(n_gpus)) if devices is None else devices
        stats, info = npdr_stats(X, y, Cm, regression_type, attr_diff_type, nbd_method, nbd_metric, knn, msurf_sd_frac, cdt,
                                 neighbor_sampling, separate_hitmiss_nbds, return_pairs=bool(return_info), device=device, devices=devs)
        num_pairs, num_unique, n_used, k_ave = info["n_pairs"], info["n_unique_pairs"], info["n_pairs_used"], info["k_ave_empirical"]
    else:
        # neighbours from a reduced attribute set or a user distance matrix, regression on all attributes
        if nbd_metric == "precomputed":
            if external_dist is None:
                raise ValueError("nbd.metric='precomputed' needs external.dist")
            src, _ = _as_matrix(external_dist)
        else:
            keep = [i for i, nm in enumerate(att_names) if nm not in set(map(str, rm_attr_from_dist))]
            src = np.asfortranarray(X[:, keep])
        ri, nn, ninfo = _nearest(src, nbd_method, nbd_metric, None, msurf_sd_frac, knn, False, device)
        num_pairs = ri.size
        num_unique = ninfo["n_unique"]
        k_ave = float(np.mean(knnVec(np.column_stack([ri, nn])))) if ri.size else 0.0
        if neighbor_sampling == "unique":
            u = uniqueNeighbors(np.column_stack([ri, nn]), device=device)
            ri, nn = _pairs_arrays(u)
        n_used = ri.size
        stats = np.empty((p, STATS_PER_ATTR))
        cdt_arr = np.asarray([DIFF[t] for t in cdt] + [DIFF["match-mismatch"]] * _lib.MAX_COVARS, dtype=np.int32)
        check(L.npdrb_regress(ctx.handle, dptr(X), N, p, iptr(ri), iptr(nn), n_used, dptr(f64(y)), dptr(Cm), q,
                              iptr(cdt_arr), DIFF[attr_diff_type], REG[regression_type], dptr(stats)))
        info = {"n_pairs": num_pairs, "n_unique_pairs": num_unique, "n_pairs_used": n_used, "k_ave_empirical": k_ave,
                "n_empty": ninfo["n_empty"]}
        if return_info:
            info["Ri"], info["NN"] = ri, nn
    if verbose:
        print(f"Neighborhood calculation: {time.time() - t_start:.3f} secs")
        print(f"{num_pairs} total neighbor pairs (possible repeats).")
        print(f"{num_unique} unique neighbor pairs.")
        print(f"Theoretical multiSURF average neighbors: {knnSURF(N, msurf_sd_frac)}.")
        print(f"Theoretical best average neighbors for class imbalance: {balanced_theoretical_k}.")
        print(f"Empirical (computed from neighborhood) average neighbors: {k_ave}.")
        if neighbor_sampling == "unique":
            print(f"{n_used} unique neighbor pairs.\n\nPerforming projected distance regression.")

    dof = (num_unique - 2) if unique_dof else 0                       # R/npdr.R:419-423
    pval_att, pval_0 = _pvalues(stats, regression_type, dof)
    flags = stats[:, 7].astype(int)
    if np.any((flags & 1) == 1) and q == 0:
        print("Regression failure. Possible monomorphic SNP or variable with no variation.\n", file=sys.stderr)
    o = rstats.order(pval_att)                                        # R/npdr.R:486
    pval_adj = rstats.p_adjust(pval_att[o], padj_method)              # R/npdr.R:488
    cols = {"att": [att_names[i] for i in o], "pval.adj": pval_adj, "pval.att": pval_att[o],
            "beta.raw.att": stats[o, 0], "beta.Z.att": stats[o, 1], "beta.0": stats[o, 2], "pval.0": pval_0[o]}
    if regression_type == "lm":
        cols["R.sqr"] = stats[o, 5]
    out = pd.DataFrame(cols) if pd is not None else cols
    if return_info:
        info["stats"] = stats
        info["order"] = o
        return out, info
    return out


def npdr_k_grid(outcome, dataset, k_grid=None, regression_type="binomial", attr_diff_type="numeric-abs",
                nbd_metric="manhattan", msurf_sd_frac=0.5, separate_hitmiss_nbds=False, padj_method="bonferroni",
                covars="none", covar_diff_type="match-mismatch", neighbor_sampling="none", device=0):
    """npdr(nbd.method = "relieff", knn = k) for every k of a grid through ONE C-ABI call (npdrb_npdr_k_grid): the attribute
    matrix is uploaded once, the distance matrix, the per-row (d, row) sort and the instance-major attribute copy are
    computed once and reused for all k (what vwok's foreach loop recomputes per k, R/adaptive.R:85-111).
    -> dict(att, k_grid, beta_Z [p, len(k_grid)], pval_adj, beta_raw, pval_att, n_pairs), attributes in input order."""
    X, att_names, pheno = _split_outcome(outcome, dataset)
    y = _code_outcome(pheno, regression_type)
    Cm, cdt, _ = _covars_matrix(covars, covar_diff_type)
    N, p = X.shape
    ks = list(range(1, N)) if k_grid is None else [int(k) for k in k_grid]       # seq.int(m - 1), R/adaptive.R:73
    o = _lib.Opts()
    L = _lib.load()
    L.npdrb_opts_default(C.byref(o))
    o.regression_type = REG[regression_type]; o.attr_diff_type = DIFF[attr_diff_type]; o.nbd_method = _lib.NBD["relieff"]
    o.nbd_metric = METRIC[nbd_metric]; o.msurf_sd_frac = float(msurf_sd_frac)
    o.n_covars = len(cdt) if Cm is not None else 0
    for i, t in enumerate(cdt[:o.n_covars]):
        o.covar_diff_type[i] = DIFF[t]
    o.neighbor_sampling_unique = int(neighbor_sampling == "unique")
    o.separate_hitmiss_nbds = int(bool(separate_hitmiss_nbds))
    kg = np.asarray(ks, dtype=np.int64)
    stats = np.empty((len(ks), p, STATS_PER_ATTR))
    npairs = np.zeros(len(ks), dtype=np.int64)
    check(L.npdrb_npdr_k_grid(context(device).handle, dptr(X), N, p, dptr(f64(y)), dptr(Cm), C.byref(o),
                              kg.ctypes.data_as(C.POINTER(C.c_int64)), len(ks), dptr(stats),
                              npairs.ctypes.data_as(C.POINTER(C.c_int64))))
    bz = np.empty((p, len(ks))); br = np.empty((p, len(ks))); pa = np.empty((p, len(ks))); padj = np.empty((p, len(ks)))
    for c in range(len(ks)):
        pv, _ = _pvalues(stats[c], regression_type)
        bz[:, c] = stats[c][:, 1]; br[:, c] = stats[c][:, 0]; pa[:, c] = pv
        padj[:, c] = rstats.p_adjust(pv, padj_method)
    return {"att": att_names, "k_grid": ks, "beta_Z": bz, "beta_raw": br, "pval_att": pa, "pval_adj": padj,
            "n_pairs": npairs, "stats": stats}


def vwok(dats, k_grid=None, verbose=False, attr_diff_type="numeric-abs", corr_attr_names=None, signal_names=None,
         separate_hitmiss_nbds=False, label="class", device=0):
    """Variable-Wise Optimized k (R/adaptive.R:39-234): NPDR for every k of the grid, then per attribute the k with the
    largest standardized beta.  Returns {"vwok.out": DataFrame(att, best.ks, betas, pval.att) sorted by decreasing beta,
    "betas": p x |k| frame, "pvals": p x |k| frame}; rows of the per-k matrices are ordered by attribute name as the
    reference's arrange(att) does.  The auPRC part (PRROC, signal.names) and correlation-data are not accelerated."""
    if attr_diff_type == "correlation-data":
        raise NotImplementedError("vwok(attr.diff.type='correlation-data') is outside the accelerated path")
    if signal_names is not None:
        raise NotImplementedError("vwok(signal.names=...): the PRROC auPRC scoring is outside the accelerated path")
    if pd is None or not isinstance(dats, pd.DataFrame):
        raise TypeError("dats must be a pandas DataFrame that contains the outcome column `label`")
    g = npdr_k_grid(label, dats, k_grid, "binomial", "numeric-abs", "manhattan", 0.5, separate_hitmiss_nbds, "bonferroni",
                    device=device)
    order = sorted(range(len(g["att"])), key=lambda i: g["att"][i])              # arrange(att)
    atts = [g["att"][i] for i in order]
    cols = [f"k.{k}" for k in g["k_grid"]]
    betas = pd.DataFrame(g["beta_Z"][order], index=atts, columns=cols)
    pvals = pd.DataFrame(g["pval_adj"][order], index=atts, columns=cols)
    best = np.nanargmax(betas.to_numpy(), axis=1)                                # which.max: first maximum
    out = pd.DataFrame({"att": atts, "best.ks": best + 1,
                        "betas": betas.to_numpy()[np.arange(len(atts)), best],
                        "pval.att": pvals.to_numpy()[np.arange(len(atts)), best]})
    out = out.sort_values("betas", ascending=False, kind="stable").reset_index(drop=True)
    return {"vwok.out": out, "betas": betas, "pvals": pvals}

