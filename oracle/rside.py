"""numpy restatements of the R-side callers of the path (fastGLOBETROTTER.R), TEST INFRASTRUCTURE ONLY -- the
checker for the opt-in entry points of include/fgt_companion.h part 3.  R is not installed in this image, so these follow the R
source line by line; where R calls BLAS / LINPACK (%*%, lm) the tests compare at a stated tolerance.  "Parity unpinned by R":
no R here to run the original; the date fit is additionally checked against the published tutorial result (main.txt)."""
from __future__ import annotations

import numpy as np


def null_normalise(tempweights, null_mat, S, K, G, means):
    """fastGLOBETROTTER.R:1216-1249.  tempweights: data_read layout [(k*K+l)*S*S + mn]; null_mat [K*K, G]; means [S*S, G].
    Returns (means / nullcurves, nullcurves)."""
    tw = np.asarray(tempweights, dtype=np.float64).reshape(K * K, S * S).T          # R's tempweights matrix: S^2 x K^2
    results = tw @ np.asarray(null_mat, dtype=np.float64).reshape(K * K, G)          # R:1222
    for k in range(G):                                                               # R:1223
        m = results[:, k].reshape(S, S)
        results[:, k] = ((m + m.T) / 2.0).ravel()
    sums = np.zeros(G)
    sums2 = np.zeros((S, G))
    sums3 = np.zeros((S, G))
    for k in range(S):                                                               # R:1229-1237
        for l in range(S):
            sums = sums + results[k * S + l]
            sums2[k] = sums2[k] + results[k * S + l]
            sums3[l] = sums3[l] + results[k * S + l]
    exp = np.zeros((S * S, G))
    for k in range(S):                                                               # R:1238-1241
        for l in range(S):
            exp[k * S + l] = sums2[k] * sums3[l] / sums
    for k in range(G):                                                               # R:1242-1246
        m = exp[:, k].reshape(S, S)
        exp[:, k] = ((m + m.T) / 2.0).ravel()
    results = results / exp                                                          # R:1247
    return np.asarray(means, dtype=np.float64).reshape(S * S, G) / results, results  # R:1248


def _lm(y, pred):
    """lm(y ~ pred): coefficients (NaN where lm reports NA for a rank-deficient column, tolerance 1e-7 like dqrdc2) and residuals"""
    T = y.size
    X = np.column_stack([np.ones(T), pred])
    keep = [0]
    Q = [X[:, 0] / np.linalg.norm(X[:, 0])]
    for j in range(1, X.shape[1]):
        v = X[:, j].copy()
        for q in Q:
            v -= (q @ v) * q
        if np.linalg.norm(v) >= 1e-7 * np.linalg.norm(X[:, j]) and np.linalg.norm(X[:, j]) > 0:
            keep.append(j)
            Q.append(v / np.linalg.norm(v))
    coef = np.full(X.shape[1], np.nan)
    sol, *_ = np.linalg.lstsq(X[:, keep], y, rcond=None)
    coef[keep] = sol
    res = y - X[:, keep] @ sol
    return coef, res


def getlhood(means, times, admixtimes, popfracs, likpenalty):
    """fastGLOBETROTTER.R:441-489"""
    means = np.asarray(means, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    admixtimes = np.atleast_1d(np.asarray(admixtimes, dtype=np.float64))
    pred = np.column_stack([np.exp(-times * a / 100.0) for a in admixtimes])
    R = means.shape[0]
    S = int(round(np.sqrt(R)))
    lhood, neg = 0.0, False
    to_sum = np.zeros(admixtimes.size)
    capture = set((np.arange(0, R, S) + np.arange(S)).tolist())                      # 0-based seq.to.capture
    for i in range(R):
        coef, res = _lm(means[i], pred)
        cur = -res.size / 2.0 * np.log(np.mean(res ** 2))
        if i in capture:
            T = res.size
            tail = np.abs(res[max(T // 2 - 1, 0):])
            sd3 = 3.0 * np.std(tail, ddof=1)
            pf = popfracs[int(i / S)]
            if admixtimes.size == 2:
                if admixtimes[0] != admixtimes[1] and not np.isnan(coef[2]):
                    if np.sign(coef[1]) != np.sign(coef[2]) and min(coef[1], coef[2]) < -sd3:
                        neg = True
                    to_sum += pf * np.abs(coef[1:3])
                if np.isnan(coef[2]):
                    neg = True
            else:
                to_sum += pf * np.abs(coef[1:2])
        lhood += cur
    if np.isnan(np.max(to_sum)):
        to_sum = np.zeros(1)
        neg = True
    if neg or np.max(to_sum) > 2.0:
        return lhood / likpenalty
    return lhood


def getmax(means, times, nparam, popfracs):
    """fastGLOBETROTTER.R:491-520 (returnall = F): returns (dates, value, likpenalty).  scipy's Nelder-Mead stands in for R's optim."""
    from scipy.optimize import minimize
    x = 10.0 ** np.arange(0, 3.0001, 0.25)
    if nparam == 1:
        grid = x[:, None]
    else:
        grid = np.array([(x[i], x[j]) for i in range(x.size) for j in range(i, x.size)])
    pen = getlhood(means, times, grid[0], popfracs, 1.0)
    maxlik, init = 0.0, np.sqrt(grid[0])
    for g in grid:
        v = getlhood(means, times, g, popfracs, pen)
        if v > maxlik:
            maxlik, init = v, np.sqrt(g)
    r = minimize(lambda p: -getlhood(means, times, 1.0 + p ** 2, popfracs, pen), init, method="Nelder-Mead",
                 options={"xatol": 1e-10, "fatol": 1e-12, "maxiter": 2000})
    return 1.0 + r.x ** 2, r.fun, pen


def project(raw_full, tw, K, S, G):
    """reference C:578-595 on one [K*K, G] raw tensor: results[mn] = sum over k<=h of tw[(K*k+h)*S*S + mn] * raw[k*K+h]"""
    out = np.zeros((S * S, G))
    tw = np.asarray(tw, dtype=np.float64)
    for k in range(K):
        for h in range(k, K):
            out += np.outer(tw[(K * k + h) * S * S:(K * k + h + 1) * S * S], raw_full[k * K + h])
    return out


def normalise_mean(results_all, S, G):
    """reference C:599-648: results_all [N, S*S, G] -> means [S*S, G]"""
    N = results_all.shape[0]
    means = np.zeros((S * S, G))
    for n in range(N):
        r = results_all[n].reshape(S, S, G)
        sym = (r + r.transpose(1, 0, 2)) / 2.0
        rs = sym.sum(axis=1)                       # [S, G]
        tot = rs.sum(axis=0)
        ex = rs[:, None, :] * rs[None, :, :] / tot
        ex = (ex + ex.transpose(1, 0, 2)) / 2.0
        means += (sym / ex / N).reshape(S * S, G)
    return means
