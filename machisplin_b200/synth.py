"""Synthetic inputs of the BASELINE.json configs: grids, knots, residual field,
covariate planes and seeded ensemble descriptors.  Used by bench.py and the tests; pure numpy /
scikit-learn on the host (this is input generation, not the hot path)."""
from __future__ import annotations

import numpy as np

from .engine import Geom

CONFIGS = {
    # name: (nrow, ncol, knots, covariates, models, responses)
    "c2": dict(nrow=2048, ncol=2048, knots=1000, C=3, kept="", L=1, seed=1236),
    "c3": dict(nrow=8192, ncol=8192, knots=5000, C=6, kept="bgnmrv", L=1, seed=1237),
    "c4": dict(nrow=16384, ncol=16384, knots=10000, C=4, kept="bgnmrv", L=1, seed=1238),
    "c5": dict(nrow=8192, ncol=8192, knots=5000, C=6, kept="gnmv", L=20, seed=1239),
}
ENSEMBLE_WEIGHTS = {"b": 0.31, "g": 0.22, "n": 0.17, "m": 0.12, "r": 0.10, "v": 0.08}


def make_geom(nrow: int, ncol: int, cell: float = 0.0) -> Geom:
    """Square cells of size 1/max(nrow, ncol) (or ``cell``); extent [0, ncol r] x [0, nrow r], NW origin."""
    r = cell if cell > 0 else 1.0 / max(nrow, ncol)
    return Geom(0.0, ncol * r, 0.0, nrow * r, nrow, ncol)


def make_knots(geom: Geom, n: int, seed: int):
    """n points at distinct cell centres (uniform), returned as (xy, row, col)."""
    rng = np.random.default_rng(seed)
    cells = rng.choice(geom.nrow * geom.ncol, size=n, replace=False)
    row, col = cells // geom.ncol, cells % geom.ncol
    x = geom.xmin + (col + 0.5) * geom.rx
    y = geom.ymax - (row + 0.5) * geom.ry
    return np.column_stack([x, y]), row.astype(np.int32), col.astype(np.int32)


def residual_field(xy: np.ndarray, seed: int, L: int = 1, noise: float = 0.1) -> np.ndarray:
    """y = sin(6x) cos(4y) + 0.1 N(0,1): interior GCV minimum (lambda ~ 1.7e-3 @1k, 5e-3 @5k)."""
    rng = np.random.default_rng(seed + 7919)
    base = np.sin(6 * xy[:, 0]) * np.cos(4 * xy[:, 1])
    if L == 1:
        return base + noise * rng.standard_normal(xy.shape[0])
    k = np.arange(L)
    out = np.sin((6 + 0.3 * k)[None, :] * xy[:, :1]) * np.cos((4 - 0.1 * k)[None, :] * xy[:, 1:2])
    return out + noise * rng.standard_normal((xy.shape[0], L))


def covariate_planes(geom: Geom, C: int, nan_frac: float = 0.02, seed: int = 99, dtype=np.float32,
                     row0: int = 0, row1: int | None = None) -> np.ndarray:
    """C smooth planes (sum of 3 random sinusoids each, seed 99 + k); plane 0 carries a NaN mask made
    of a few discs covering ~nan_frac of the grid ("sea").  Rows [row0, row1) only, to bound memory."""
    row1 = geom.nrow if row1 is None else row1
    r = np.arange(row0, row1)
    x = geom.xmin + (np.arange(geom.ncol) + 0.5) * geom.rx
    y = geom.ymax - (r + 0.5) * geom.ry
    out = np.empty((C, r.size, geom.ncol), dtype=dtype)
    for k in range(C):
        rng = np.random.default_rng(seed + k)
        acc = np.zeros((r.size, geom.ncol))
        for _ in range(3):
            fx, fy = rng.uniform(2, 12, 2)
            ph = rng.uniform(0, 2 * np.pi)
            amp = rng.uniform(0.5, 2.0)
            acc += amp * np.sin(fx * x[None, :] + fy * y[:, None] + ph)
        out[k] = (100.0 * (k + 1) + 50.0 * acc).astype(dtype)
    if nan_frac > 0:
        rng = np.random.default_rng(seed + 1000)
        ndisc = 8
        rad = np.sqrt(nan_frac * (geom.xmax - geom.xmin) * (geom.ymax - geom.ymin) / (ndisc * np.pi))
        for _ in range(ndisc):
            cx = rng.uniform(geom.xmin, geom.xmax)
            cy = rng.uniform(geom.ymin, geom.ymax)
            m = (x[None, :] - cx) ** 2 + (y[:, None] - cy) ** 2 < rad * rad
            out[0][m] = np.nan
    return out


# ---------------------------------------------------------------------------------------------------
# seeded ensemble descriptors (flat layout of include/machisplin_b200.h)
# ---------------------------------------------------------------------------------------------------
def _training_table(geom: Geom, C: int, n: int, seed: int):
    """Training rows as the reference builds them (V73:145-154): covariates at the knot cells + LONG, LAT."""
    xy, row, col = make_knots(geom, n, seed)
    X = np.empty((n, C + 2))
    for k in range(C):
        rng = np.random.default_rng(99 + k)
        acc = np.zeros(n)
        for _ in range(3):
            fx, fy = rng.uniform(2, 12, 2)
            ph = rng.uniform(0, 2 * np.pi)
            amp = rng.uniform(0.5, 2.0)
            acc += amp * np.sin(fx * xy[:, 0] + fy * xy[:, 1] + ph)
        X[:, k] = np.float32(100.0 * (k + 1) + 50.0 * acc)
    X[:, C] = xy[:, 0]
    X[:, C + 1] = xy[:, 1]
    rng = np.random.default_rng(seed + 31)
    resp = 200.0 + 0.05 * X[:, 0] - 0.03 * X[:, min(1, C - 1)] + 30 * np.sin(5 * xy[:, 0]) * np.cos(3 * xy[:, 1]) \
        + 2.0 * rng.standard_normal(n)
    return X, resp


def make_models(geom: Geom, C: int, n_train: int, seed: int, kept: str = "bgnmrv", rf_trees: int = 500,
                gbm_trees: int = 1000, svm_frac: float = 0.5, mars_terms: int = 21,
                keep_estimators: bool = False, table=None) -> dict:
    """Six fitted-model descriptors with the reference's structural sizes: gam (linear), nnet(10),
    earth (<= 21 hinge terms), ksvm (~0.5 n SVs), randomForest (500 trees), gbm (1000 trees, 5 splits).
    table = (X, resp): fit on that training table (rows = points, columns = cov_1..cov_C, LONG, LAT as at V73:145-154)
    instead of the synthetic one - the tests on the reference's bundled data use this."""
    P = C + 2
    X, resp = _training_table(geom, C, n_train, seed) if table is None else (np.asarray(table[0], float), np.asarray(table[1], float))
    assert X.shape[1] == P
    rng = np.random.default_rng(seed + 77)
    out = {}
    sk = {"X": X, "resp": resp}     # fitted scikit-learn objects (tests: independent check of the descriptors)
    if "g" in kept:
        A = np.column_stack([np.ones(len(X)), X])
        out["g"] = {"coef": np.linalg.lstsq(A, resp, rcond=None)[0]}
    if "n" in kept:
        H = 10
        mn, mx = resp.min(), resp.max()
        mu, sd = X.mean(0), X.std(0)
        Wh = rng.standard_normal((H, P)) / sd[None, :]           # inputs are unscaled in the reference
        bh = -(Wh * mu[None, :]).sum(1) + 0.3 * rng.standard_normal(H)
        wo = np.concatenate([[0.4], rng.standard_normal(H) * 0.2])
        wts = np.concatenate([np.column_stack([bh, Wh]).ravel(), wo])
        out["n"] = {"wts": wts, "H": H, "max2": mx - mn, "min": mn}
    if "m" in kept:
        T = mars_terms
        dirs = np.zeros((T, P), dtype=np.int8)
        cuts = np.zeros((T, P))
        basis = [np.ones(len(X))]
        for t in range(1, T):
            v = int(rng.integers(0, P))
            dirs[t, v] = 1 if (t % 2) else -1
            cuts[t, v] = X[int(rng.integers(0, len(X))), v]
            basis.append(np.maximum(0.0, dirs[t, v] * (X[:, v] - cuts[t, v])))
        coef = np.linalg.lstsq(np.column_stack(basis), resp, rcond=None)[0]
        out["m"] = {"dirs": dirs, "cuts": cuts, "coef": coef}
    if "v" in kept:
        from sklearn.svm import SVR
        xc, xs = X.mean(0), X.std(0, ddof=1)
        yc, ys = resp.mean(), resp.std(ddof=1)
        Xs, ysc = (X - xc) / xs, (resp - yc) / ys
        sigma = 1.0 / (2.0 * P) * 1.5
        # epsilon tuned so that about svm_frac of the rows become support vectors
        svr = SVR(kernel="rbf", gamma=sigma, C=1.0, epsilon=0.1 if svm_frac >= 0.5 else 0.3).fit(Xs, ysc)
        sk["v"] = svr
        out["v"] = {"sv": svr.support_vectors_.copy(), "alpha": svr.dual_coef_.ravel().copy(),
                    "b": -float(svr.intercept_[0]), "sigma": sigma, "x_center": xc, "x_scale": xs,
                    "y_center": yc, "y_scale": ys}
    if "r" in kept:
        from sklearn.ensemble import RandomForestRegressor
        rf = RandomForestRegressor(n_estimators=rf_trees, min_samples_leaf=5, max_features=max(1, P // 3),
                                   random_state=seed, n_jobs=-1).fit(X, resp)
        sk["r"] = rf
        nrn = max(e.tree_.node_count for e in rf.estimators_)
        nt = len(rf.estimators_)
        left = np.zeros((nt, nrn), np.int32); right = np.zeros((nt, nrn), np.int32)
        status = np.zeros((nt, nrn), np.int8); bestvar = np.zeros((nt, nrn), np.int32)
        split = np.zeros((nt, nrn)); pred = np.zeros((nt, nrn))
        for t, e in enumerate(rf.estimators_):
            tr = e.tree_
            k = tr.node_count
            leaf = tr.children_left == -1
            left[t, :k] = np.where(leaf, 0, tr.children_left + 1)
            right[t, :k] = np.where(leaf, 0, tr.children_right + 1)
            status[t, :k] = np.where(leaf, -1, -3)
            bestvar[t, :k] = np.where(leaf, 0, tr.feature + 1)
            split[t, :k] = np.where(leaf, 0.0, tr.threshold)
            pred[t, :k] = tr.value[:, 0, 0]
        out["r"] = {"ntree": nt, "nrnodes": nrn, "left": left, "right": right, "status": status,
                    "bestvar": bestvar, "split": split, "nodepred": pred}
    if "b" in kept:
        from sklearn.ensemble import GradientBoostingRegressor
        lr = 0.01
        gb = GradientBoostingRegressor(n_estimators=gbm_trees, learning_rate=lr, max_leaf_nodes=6, max_depth=None,
                                       subsample=0.5, random_state=seed).fit(X, resp)
        sk["b"] = gb
        off = [0]
        sv, sc, ln, rn, mn_ = [], [], [], [], []
        for e in gb.estimators_[:, 0]:
            tr = e.tree_
            k = tr.node_count
            leaf = tr.children_left == -1
            sv.append(np.where(leaf, -1, tr.feature))
            sc.append(np.where(leaf, lr * tr.value[:, 0, 0], tr.threshold))
            ln.append(np.where(leaf, 0, tr.children_left))
            rn.append(np.where(leaf, 0, tr.children_right))
            # gbm's MissingNode: a terminal holding the node mean; emulate by routing NA to the larger child
            big = np.where(tr.n_node_samples[np.maximum(tr.children_left, 0)] >=
                           tr.n_node_samples[np.maximum(tr.children_right, 0)], tr.children_left, tr.children_right)
            mn_.append(np.where(leaf, 0, big))
            off.append(off[-1] + k)
        # sklearn splits on x <= thr, gbm on x < thr: nudge thresholds up one ulp so both agree on float inputs
        scat = np.concatenate(sc)
        svat = np.concatenate(sv)
        scat = np.where(svat >= 0, np.nextafter(scat, np.inf), scat)
        out["b"] = {"initF": float(gb.init_.constant_[0, 0]), "tree_off": np.array(off, np.int32),
                    "splitvar": svat.astype(np.int32), "splitcode": scat,
                    "left": np.concatenate(ln).astype(np.int32), "right": np.concatenate(rn).astype(np.int32),
                    "missing": np.concatenate(mn_).astype(np.int32)}
    if keep_estimators:
        out["_sk"] = sk
    return out


def ensemble_weights(kept: str):
    """(kept letters, rounded weights, unrounded total over all six candidates) - fixed synthetic weights."""
    w = np.array([ENSEMBLE_WEIGHTS[k] for k in kept])
    return kept, w, float(sum(ENSEMBLE_WEIGHTS.values())) if set(kept) - set("gnmv") else \
        float(sum(ENSEMBLE_WEIGHTS[k] for k in "gnmv"))
