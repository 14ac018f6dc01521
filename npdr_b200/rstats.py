"""The R-side residue of npdr(): pt(), pnorm(), order(), p.adjust() (R/npdr.R:55,63,486-488).

In the R deployment these stay in R (BASELINE north_star); this module is their host-side stand-in
for the Python harness so that outputs have the same columns and ordering as the reference.
"""
import numpy as np
from scipy import special, stats


def pnorm(x, lower_tail=True):
    x = np.asarray(x, dtype=np.float64)
    return special.ndtr(x) if lower_tail else special.ndtr(-x)


def pt(x, df, lower_tail=True):
    """R's pt(): for df > 4e5 R switches to the normal approximation
    pnorm(x*(1 - 1/(4n))/sqrt(1 + x^2/(2n))) (nmath/pt.c); otherwise the incomplete-beta form."""
    x = np.asarray(x, dtype=np.float64)
    df = np.broadcast_to(np.asarray(df, dtype=np.float64), x.shape)
    out = np.full(x.shape, np.nan)
    ok = np.isfinite(x) & np.isfinite(df) & (df > 0)
    big = ok & (df > 4e5)
    if big.any():
        val = 1.0 / (4.0 * df[big])
        z = x[big] * (1.0 - val) / np.sqrt(1.0 + x[big] * x[big] * 2.0 * val)
        out[big] = special.ndtr(z) if lower_tail else special.ndtr(-z)
    small = ok & ~big
    if small.any():
        out[small] = stats.t.cdf(x[small], df[small]) if lower_tail else stats.t.sf(x[small], df[small])
    return out


def order(x):
    """R's order(x, decreasing = FALSE): stable, NA last, ties keep the original order."""
    return np.argsort(np.asarray(x, dtype=np.float64), kind="stable")


def p_adjust(p, method="bonferroni"):
    """stats::p.adjust: NAs are dropped, n = number of non-NA p-values."""
    p = np.asarray(p, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    nna = ~np.isnan(p)
    q = p[nna]
    n = q.size
    method = {"fdr": "BH"}.get(method, method)
    if n <= 1 or method == "none":
        out[nna] = q
        return out
    if method == "bonferroni":
        r = np.minimum(1.0, n * q)
    elif method == "holm":
        o = np.argsort(q, kind="stable"); ro = np.argsort(o, kind="stable")
        i = np.arange(1, n + 1)
        r = np.minimum(1.0, np.maximum.accumulate((n + 1 - i) * q[o]))[ro]
    elif method == "hochberg":
        o = np.argsort(-q, kind="stable"); ro = np.argsort(o, kind="stable")
        i = np.arange(n, 0, -1)
        r = np.minimum(1.0, np.minimum.accumulate((n + 1 - i) * q[o]))[ro]
    elif method in ("BH", "BY"):
        o = np.argsort(-q, kind="stable"); ro = np.argsort(o, kind="stable")
        i = np.arange(n, 0, -1)
        c = np.sum(1.0 / np.arange(1, n + 1)) if method == "BY" else 1.0
        r = np.minimum(1.0, np.minimum.accumulate(c * n / i * q[o]))[ro]
    elif method == "hommel":
        o = np.argsort(q, kind="stable"); ro = np.argsort(o, kind="stable")
        ps = q[o]
        i = np.arange(1, n + 1)
        qq = np.full(n, np.min(n * ps / i)); pa = qq.copy()
        for m in range(n - 1, 1, -1):
            i1 = np.arange(0, n - m + 1)
            i2 = np.arange(n - m + 1, n)
            q1 = np.min(m * ps[i2] / np.arange(2, m + 1))
            qq[i1] = np.minimum(m * ps[i1], q1)
            qq[i2] = qq[n - m]
            pa = np.maximum(pa, qq)
        r = np.maximum(pa, ps)[ro]
    else:
        raise ValueError(f"unknown p.adjust method {method!r}")
    out[nna] = r
    return out


def _sig_needed(x, digits):
    """Per element: decimal exponent e and the minimal number of significant digits (<= digits)
    that represents x rounded to `digits` significant digits (R's formatReal/scientific())."""
    x = np.asarray(x, dtype=np.float64)
    e = np.zeros(x.shape, dtype=np.int64)
    sig = np.ones(x.shape, dtype=np.int64)
    for idx in np.ndindex(x.shape):
        v = x[idx]
        if not np.isfinite(v) or v == 0:
            continue
        mant, ex = f"{abs(v):.{digits - 1}e}".split("e")
        e[idx] = int(ex)
        sig[idx] = max(1, len(mant.replace(".", "").rstrip("0")))
    return e, sig


def format_fixed(x, digits=5):
    """as.numeric(format(x, scientific = FALSE, digits = d)) as uniReg applies it to the beta and
    statistic columns (R/utils.R:144-145): fixed notation with ONE common number of decimals, enough
    for the element that needs the most to show `digits` significant digits."""
    x = np.asarray(x, dtype=np.float64)
    e, sig = _sig_needed(x, digits)
    fin = np.isfinite(x)
    rgt = int(np.max(np.maximum(0, sig - 1 - e)[fin])) if fin.any() else 0
    out = x.copy()
    out[fin] = np.round(x[fin], rgt)
    return out


def format_sci(x, digits=5):
    """as.numeric(format(x, scientific = TRUE, digits = d)) (R/utils.R:143,146): scientific notation
    with one common number of significant digits (the most any element needs, <= digits)."""
    x = np.asarray(x, dtype=np.float64)
    e, sig = _sig_needed(x, digits)
    fin = np.isfinite(x) & (x != 0)
    ns = int(np.max(sig[fin])) if fin.any() else 1
    out = x.copy()
    out[fin] = np.array([float(f"{v:.{ns - 1}e}") for v in x[fin]])
    return out
