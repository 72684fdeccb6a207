"""Oracle: per-cell predictors of the six ensemble members + the weight rule + RSS objective.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the
model packages (mgcv, nnet, earth, kernlab, randomForest, gbm) are CRAN
dependencies without version pins (DESCRIPTION:12) and absent from
``/root/reference``; the predictors below restate the documented structure of
their fitted objects.  Only *prediction* is in scope -
fitting stays in R.

Reference call sites:
  V73:468-475  nnet    terra::predict(rast_stack, mod.nn)  * max2.resp.f + min.resp.f
  V73:497-499  gbm     terra::predict(rast_stack, mod.brt, n.trees = best.trees, type="response")
  V73:521-523  rf      terra::predict(rast_stack, mod.rf, type="response")
  V73:543-545  earth   terra::predict(rast_stack, mod.MARS)
  V73:582-584  ksvm    terra::predict(rast_stack, mod.SVM, na.rm=TRUE)
  V73:604-606  gam     terra::predict(rast_stack, mod.GAM)
  V73:619      pred.elev <- pred.elev / OptX.mfit.wt.tot
  V73:329-331 / 369-371   the RSS objective handed to optimx
  V73:336-393  the 5 % cut and round(., 2)

Feature order everywhere is that of ``xnam`` (V73:194): cov_1..cov_C, LONG, LAT.
A model descriptor is a dict of numpy arrays (the same flat layout the C ABI takes,
see include/machisplin_b200.h):

  gam  : coef[P+1]                                     (intercept first)
  nnet : wts[(P+1)*H + H+1], H, max2, min              (nnet ordering: per hidden unit bias,
                                                        inputs 1..P; then output bias, hidden 1..H)
  mars : dirs[T,P] (0, +-1, 2), cuts[T,P], coef[T]
  svm  : sv[S,P] (scaled), alpha[S], b, sigma, x_center[P], x_scale[P], y_center, y_scale
  rf   : ntree, nrnodes, left[ntree,nrnodes], right[..] (1-based, 0 = none),
         status[..] (-1 terminal), bestvar[..] (1-based), split[..], nodepred[..]
  gbm  : initF, tree_off[n_trees+1], splitvar[] (-1 terminal), splitcode[], left[], right[], missing[]
"""
from __future__ import annotations

import numpy as np

MODEL_LETTERS_ALL = "bgnmrv"       # order of k1..k6 in the objective, V73:329
MODEL_LETTERS_SMOOTH = "gnmv"      # smooth.outputs.only, V73:369


# ---------------------------------------------------------------------------------------
# predictors: X is (n, P) float64, NaN = NA.  Each returns (n,) float64.
# ---------------------------------------------------------------------------------------
def predict_gam(m, X):
    """mgcv::gam with a purely parametric formula (V73:194-195, 600): X beta."""
    coef = np.asarray(m["coef"], dtype=np.float64)
    return coef[0] + X @ coef[1:]


def predict_nnet(m, X):
    """nnet(size=H, linout=TRUE) (V73:463): logistic hidden layer, linear output; then the
    response un-scaling of V73:469-470."""
    H = int(m["H"])
    P = X.shape[1]
    w = np.asarray(m["wts"], dtype=np.float64)
    Wh = w[: (P + 1) * H].reshape(H, P + 1)          # row h: bias, inputs
    wo = w[(P + 1) * H:]
    z = Wh[:, 0][None, :] + X @ Wh[:, 1:].T
    with np.errstate(over="ignore"):
        a = 1.0 / (1.0 + np.exp(-z))
    out = wo[0] + a @ wo[1:]
    return out * float(m["max2"]) + float(m["min"])


def predict_mars(m, X):
    """earth (degree-1 default, V73:539): sum_t beta_t prod_v h(dir, x_v - cut)."""
    dirs = np.asarray(m["dirs"])
    cuts = np.asarray(m["cuts"], dtype=np.float64)
    coef = np.asarray(m["coef"], dtype=np.float64)
    out = np.zeros(X.shape[0])
    for t in range(dirs.shape[0]):
        b = np.full(X.shape[0], coef[t])
        for v in np.nonzero(dirs[t])[0]:
            if dirs[t, v] == 2:
                b = b * X[:, v]
            else:
                b = b * np.maximum(0.0, dirs[t, v] * (X[:, v] - cuts[t, v]))
        out += b
    return out


def predict_svm(m, X, chunk=2048):
    """kernlab::ksvm eps-svr, rbfdot, scaled=TRUE (V73:560):
    (sum_i alpha_i exp(-sigma |xs - sv_i|^2) - b) * y_scale + y_center."""
    sv = np.asarray(m["sv"], dtype=np.float64)
    alpha = np.asarray(m["alpha"], dtype=np.float64)
    out = np.empty(X.shape[0])
    sv2 = np.sum(sv * sv, axis=1)
    for a in range(0, X.shape[0], chunk):
        xs = (X[a:a + chunk] - m["x_center"]) / m["x_scale"]
        d2 = np.sum(xs * xs, axis=1)[:, None] + sv2[None, :] - 2.0 * xs @ sv.T
        # the expansion above is only used for speed; guard tiny negatives
        k = np.exp(-float(m["sigma"]) * np.maximum(d2, 0.0))
        out[a:a + chunk] = (k @ alpha - float(m["b"])) * float(m["y_scale"]) + float(m["y_center"])
    out[np.isnan(X).any(axis=1)] = np.nan
    return out


def predict_svm_exact(m, X):
    """Same as predict_svm with |xs - sv|^2 formed by explicit differences (slow, small n)."""
    sv = np.asarray(m["sv"], dtype=np.float64)
    xs = (X - m["x_center"]) / m["x_scale"]
    d2 = ((xs[:, None, :] - sv[None, :, :]) ** 2).sum(-1)
    k = np.exp(-float(m["sigma"]) * d2)
    return (k @ np.asarray(m["alpha"], dtype=np.float64) - float(m["b"])) * float(m["y_scale"]) + float(m["y_center"])


def predict_rf(m, X):
    """randomForest regression: mean over trees of nodepred at the reached leaf;
    left iff x[bestvar] <= xbestsplit."""
    ntree = int(m["ntree"])
    n = X.shape[0]
    acc = np.zeros(n)
    bad = np.isnan(X).any(axis=1)
    Xs = np.where(np.isnan(X), 0.0, X)
    rows = np.arange(n)
    for t in range(ntree):
        left, right = m["left"][t], m["right"][t]
        status, bestvar = m["status"][t], m["bestvar"][t]
        split, nodepred = m["split"][t], m["nodepred"][t]
        k = np.zeros(n, dtype=np.int64)          # 0-based node index (R: k = 1)
        active = status[k] != -1
        while active.any():
            ka = k[active]
            xv = Xs[rows[active], bestvar[ka] - 1]
            goleft = xv <= split[ka]
            k[active] = np.where(goleft, left[ka], right[ka]) - 1
            active = status[k] != -1
        acc += nodepred[k]
    out = acc / ntree
    out[bad] = np.nan
    return out


def predict_gbm(m, X):
    """gbm gaussian, type="response": initF + sum_t leaf_t(x); left iff x < split;
    NA goes to MissingNode (so gbm returns a number where the others return NA)."""
    n = X.shape[0]
    acc = np.full(n, float(m["initF"]))
    off = np.asarray(m["tree_off"])
    sv, sc = np.asarray(m["splitvar"]), np.asarray(m["splitcode"], dtype=np.float64)
    ln, rn, mn = np.asarray(m["left"]), np.asarray(m["right"]), np.asarray(m["missing"])
    rows = np.arange(n)
    for t in range(off.shape[0] - 1):
        o = off[t]
        k = np.zeros(n, dtype=np.int64)
        active = sv[o + k] != -1
        while active.any():
            ka = o + k[active]
            xv = X[rows[active], sv[ka]]
            nxt = np.where(np.isnan(xv), mn[ka], np.where(xv < sc[ka], ln[ka], rn[ka]))
            k[active] = nxt
            active = sv[o + k] != -1
        acc += sc[o + k]
    return acc


PREDICTORS = {"b": predict_gbm, "g": predict_gam, "n": predict_nnet,
              "m": predict_mars, "r": predict_rf, "v": predict_svm}


def _na_to_nan(letter, X, out):
    """terra::predict: any NA predictor -> NA, except gbm which routes through MissingNode."""
    if letter != "b":
        out = np.array(out, copy=True)
        out[np.isnan(X).any(axis=1)] = np.nan
    return out


def predict_model(letter, m, X):
    return _na_to_nan(letter, X, PREDICTORS[letter](m, X))


# ---------------------------------------------------------------------------------------
# weights (V73:336-393) and the ensemble combine (V73:471-475 ... 619)
# ---------------------------------------------------------------------------------------
def r_round2(x: float) -> float:
    """R's round(x, 2) (IEEE round-half-even on the decimal representation is a good model
    of R >= 4.0.0's 'long double' algorithm for 2 digits)."""
    return float(np.round(x, 2))


def select_models(p: np.ndarray, letters: str = MODEL_LETTERS_ALL):
    """Keep model k iff round(p_k, 2) > 0.05 * sum(p); kept weight = round(p_k, 2);
    the divisor stays the *unrounded* total over all candidates (V73:337-362, 619)."""
    p = np.asarray(p, dtype=np.float64)
    tot = float(np.sum(p))
    cut = 0.05 * tot
    kept, w = "", []
    for k, letter in enumerate(letters):
        rk = r_round2(p[k])
        if rk > cut:
            kept += letter
            w.append(rk)
    return kept, np.array(w), tot


def ensemble_predict(models: dict, kept: str, w: np.ndarray, w_total: float, X: np.ndarray):
    """pred.elev = sum_k w_k f_k(x) / w_total  (V73:471-475, 498-499, ..., 619)."""
    acc = np.zeros(X.shape[0])
    for letter, wk in zip(kept, w):
        acc = acc + predict_model(letter, models[letter], X) * wk
    return acc / w_total


# ---------------------------------------------------------------------------------------
# RSS objective (V73:329-331, 369-371) and its Gram form
# ---------------------------------------------------------------------------------------
def rss_objective(k: np.ndarray, R: np.ndarray) -> float:
    """sum_j ( sum_k k_k r_kj / sum_k k_k )^2 with R = n x K cross-validation residuals."""
    k = np.asarray(k, dtype=np.float64)
    return float(np.sum((R @ (k / np.sum(k))) ** 2))


def gram(R: np.ndarray) -> np.ndarray:
    return R.T @ R


def rss_from_gram(k: np.ndarray, G: np.ndarray) -> float:
    k = np.asarray(k, dtype=np.float64)
    return float(k @ G @ k / np.sum(k) ** 2)
