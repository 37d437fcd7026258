"""shared fixtures for the parity tests (seeded inputs, the reference's tolerance form)"""
import numpy as np


def close(x, y, tol=1e-5):
    """|x-y| <= tol * (1 + max(|x|,|y|)) — data-beans-alg/src/collapse_data/stats_tests.rs:20-30"""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    return bool(np.all(np.abs(x - y) <= tol * (1.0 + np.maximum(np.abs(x), np.abs(y)))))


def max_err(x, y):
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    return float(np.max(np.abs(x - y) / (1.0 + np.maximum(np.abs(x), np.abs(y))))) if x.size else 0.0


def random_csc(rng, D, N, density=0.05, max_count=6, empty_every=0):
    """integer-valued counts, rows ascending inside a column (canonical CSC)"""
    indptr, idx, val = [0], [], []
    for j in range(N):
        if empty_every and j % empty_every == 0:
            indptr.append(len(idx))
            continue
        n = rng.binomial(D, density)
        rows = np.sort(rng.choice(D, size=n, replace=False))
        idx.extend(rows.tolist())
        val.extend(np.minimum(rng.geometric(0.7, size=n), max_count).astype(np.float32).tolist())
        indptr.append(len(idx))
    return np.array(indptr, np.uint64), np.array(idx, np.uint64), np.array(val, np.float32)


def tiny_fixture():
    """data-beans-alg/src/random_projection.rs:578-605: 8 genes x 12 cells"""
    d, n = 8, 12
    indptr, idx, val = [0], [], []
    for j in range(n):
        for i in range(d):
            if (i * 7 + j * 3) % 5 < 3:
                idx.append(i)
                val.append(1.0 + ((i + j) % 4))
        indptr.append(len(idx))
    return d, n, np.array(indptr, np.uint64), np.array(idx, np.uint64), np.array(val, np.float32)


def toy_stat(G, S, B):
    """data-beans-alg/src/collapse_data/stats_tests.rs:6-18, in (S, G) / (B, G) / (S, B) layout"""
    f = lambda a, b: 1.0 + ((a * 7 + b * 13) % 11)
    obs = np.array([[f(g, c) for g in range(G)] for c in range(S)], np.float32)
    imp = np.array([[0.5 * f(g + 1, c + 2) for g in range(G)] for c in range(S)], np.float32)
    res = np.array([[0.3 * f(g + 2, c + 1) for g in range(G)] for c in range(S)], np.float32)
    size = np.array([2.0 + (c % 3) for c in range(S)], np.float32)
    obs_db = np.array([[f(g, b) + 0.7 for g in range(G)] for b in range(B)], np.float32)
    n_bs = np.array([[1.0 + ((b + c) % 4) for b in range(B)] for c in range(S)], np.float32)
    return obs, imp, res, size, obs_db, n_bs


def dense_columns(ip, ix, v, D):
    n = len(ip) - 1
    out = np.zeros((n, D))
    for j in range(n):
        sl = slice(int(ip[j]), int(ip[j + 1]))
        out[j, ix[sl].astype(np.int64)] = v[sl]
    return out


def matched_stat_f64(ip, ix, v, D, grp, S, idx, dist):
    """float64 restatement of collect_matched_stat_visitor (stats.rs:26-108): (imputed (S, D), residual (S, D))"""
    Y = dense_columns(ip, ix, v, D)
    imp, res = np.zeros((S, D)), np.zeros((S, D))
    for j in range(len(ip) - 1):
        live = idx[j] != 0xFFFFFFFF
        y1 = Y[j].copy()
        if live.any():
            m, d = idx[j][live].astype(np.int64), dist[j][live].astype(np.float64)
            w = np.exp(-d - (-d).min())
            w /= w.sum()
            yhat = (w[:, None] * Y[m]).sum(0)
            scale = Y[j].sum() / yhat.sum() if yhat.sum() > 0 else 1.0
            pos = (yhat > 0) & (y1 > 0)
            y1[pos] = y1[pos] / (yhat[pos] * scale)
            imp[grp[j]] += yhat
        res[grp[j]] += y1
    return imp, res


def nystrom_f64(ip, ix, v, D, basis, delta, pb, csn):
    ip = ip.astype(np.int64)
    out = np.zeros((len(ip) - 1, basis.shape[0]))
    for j in range(len(ip) - 1):
        rows = ix[ip[j]:ip[j + 1]].astype(np.int64)
        x = v[ip[j]:ip[j + 1]].astype(np.float64)
        if len(x) == 0:
            continue
        x = x / max(np.sqrt((x * x).sum()), 1e-8) * csn
        if delta is not None and pb[j] < delta.shape[0]:
            d = delta[pb[j], rows].astype(np.float64)
            scale = x.sum() / d.sum() if d.sum() > 0 else 1.0
            x = np.where(d > 0, x / np.where(d > 0, d * scale, 1.0), x)
        z = np.log1p(x)
        mu, sig = z.mean(), z.std()
        z = (z - mu) / sig if sig > 0 else z - mu
        out[j] = basis[:, rows].astype(np.float64) @ z
    return out
