"""Oracle: float64 restatement of ``fields::Tps`` + ``predict.Krig``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the
``fields`` sources are not in ``/root/reference`` and R is not installed; this
restates the published algorithm (Nychka et al., ``fields`` >= 9, functions
``Tps``, ``Krig``, ``Krig.engine.default``, ``gcv.Krig``, ``Krig.df.to.lambda``,
``bisection.search``, ``golden.section.search``, ``Krig.find.gcvmin``,
``Krig.coef``, ``predict.Krig``, ``Rad.cov``/``radfun``) as laid out in
SURVEY.md 3.2, 3.3 and Appendix A.

Reference call sites this stands in for:
  V73:722  mod.tps.elev <- fields::Tps(MyTPSdata[2:3], MyTPSdata[1])
  V73:751  mod.tps.elev <- fields::Tps(dat[,c(n.covars,n.covars+1)], res.FINAL)
  V73:726  terra::interpolate(terra::rast(rb), mod.tps.elev)
  V73:753  terra::interpolate(terra::rast(rast_stack), mod.tps.elev)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# radbas.constant(m=2, d=2) = (-1)^(1+m+d/2) 2^(1-2m) pi^(-d/2) / (Gamma(m) Gamma(m-d/2+1))
RBF_CONSTANT = 1.0 / (8.0 * math.pi)
D2_CLAMP = 1e-20  # Fortran radfun: if (d2 .lt. 1e-20) d2 = 1e-20
NT = 3  # null space of m=2, d=2: [1, s1, s2]


def radfun(d2: np.ndarray) -> np.ndarray:
    """``radfun(d2, par1 = p/2 = 1, par2 = 1)`` = log(d2)/2 * d2 with the 1e-20 clamp."""
    d2 = np.maximum(d2, D2_CLAMP)
    return 0.5 * np.log(d2) * d2


def rad_cov(s1: np.ndarray, s2: np.ndarray) -> np.ndarray:
    """``Rad.cov(x1, x2, m=2)`` for d=2: (1/8pi) * radfun(|s1_i - s2_j|^2)."""
    dx = s1[:, None, 0] - s2[None, :, 0]
    dy = s1[:, None, 1] - s2[None, :, 1]
    return RBF_CONSTANT * radfun(dx * dx + dy * dy)


@dataclass
class TpsFit:
    """What the B200 engine must reproduce of a ``Krig`` object."""

    center: np.ndarray        # x.center (2)  = column minima   (scale.type="range")
    scale: np.ndarray         # x.scale  (2)  = column ranges
    knots_s: np.ndarray       # np x 2 scaled unique locations
    knots_xy: np.ndarray      # np x 2 unscaled
    c: np.ndarray             # np
    d: np.ndarray             # 3
    lam: float
    eff_df: float
    weights: np.ndarray       # np replicate counts (weightsM)
    yM: np.ndarray            # np replicate means
    eta: np.ndarray = field(repr=False, default=None)   # eigenvalues of Q2' W K W Q2 (np-3)
    u: np.ndarray = field(repr=False, default=None)     # c(0,0,0, V' Q2' sqrt(w) yM)
    gcv_grid: np.ndarray = field(repr=False, default=None)  # 200 x (lambda, trA, GCV)
    pure_ss: float = 0.0
    n_obs: int = 0
    gcv_at_endpoint: bool = False


# ----------------------------------------------------------------------------------------
# Householder QR of the N x 3 polynomial block (R: qr(), qr.q2ty, qr.qy, qr.coef)
# ----------------------------------------------------------------------------------------
class _QRT:
    def __init__(self, T: np.ndarray):
        A = np.array(T, dtype=np.float64, copy=True)
        n, k = A.shape
        self.v = []
        self.tau = []
        for j in range(k):
            x = A[j:, j].copy()
            alpha = -math.copysign(np.linalg.norm(x), x[0] if x[0] != 0 else 1.0)
            v = x
            v[0] -= alpha
            vn = np.linalg.norm(v)
            if vn == 0.0:
                raise ValueError("Regression matrix for fixed part of model is colinear")
            v /= vn
            A[j:, j:] -= 2.0 * np.outer(v, v @ A[j:, j:])
            self.v.append(v)
            self.tau.append(2.0)
        self.R = np.triu(A[:k, :k])
        if np.min(np.abs(np.diag(self.R))) < 1e-12 * np.max(np.abs(np.diag(self.R))):
            raise ValueError("Regression matrix for fixed part of model is colinear")
        self.n, self.k = n, k

    def qty(self, B: np.ndarray) -> np.ndarray:
        B = np.array(B, dtype=np.float64, copy=True)
        for j, v in enumerate(self.v):
            B[j:] -= 2.0 * np.multiply.outer(v, v @ B[j:]) if B.ndim > 1 else 2.0 * v * (v @ B[j:])
        return B

    def qy(self, B: np.ndarray) -> np.ndarray:
        B = np.array(B, dtype=np.float64, copy=True)
        for j in reversed(range(self.k)):
            v = self.v[j]
            B[j:] -= 2.0 * np.multiply.outer(v, v @ B[j:]) if B.ndim > 1 else 2.0 * v * (v @ B[j:])
        return B

    def q2ty(self, B):           # qr.q2ty: rows nt+1..n of Q'B
        return self.qty(B)[self.k:]

    def coef(self, b):           # qr.coef: R^-1 Q1' b
        return np.linalg.solve(self.R, self.qty(b)[: self.k])


# ----------------------------------------------------------------------------------------
# GCV machinery
# ----------------------------------------------------------------------------------------
def tr_a(lam: float, D: np.ndarray) -> float:
    return float(np.sum(1.0 / (1.0 + lam * D)))


def gcv_value(lam: float, D: np.ndarray, u: np.ndarray, n_obs: int, pure_ss: float,
              nt: int = NT, cost: float = 1.0, offset: float = 0.0) -> float:
    """``Krig.fgcv``."""
    lD = D * lam
    rss = float(np.sum(((u * lD) / (1.0 + lD)) ** 2))
    npts = lD.shape[0]
    mse = rss / npts
    if n_obs - npts > 0:
        mse += pure_ss / (n_obs - npts)
    tra = float(np.sum(1.0 / (1.0 + lD)))
    den = 1.0 - (cost * (tra - nt - offset) + nt) / npts
    return mse / den ** 2 if den > 0 else float("nan")


def _bisection_search(x1, x2, f, tol=1e-7, niter=25):
    f1, f2 = f(x1), f(x2)
    if f1 > f2:
        raise ValueError(" f1 must be < f2 ")
    for _ in range(niter):
        xm = (x1 + x2) / 2.0
        fm = f(xm)
        if fm < 0:
            x1, f1 = xm, fm
        else:
            x2, f2 = xm, fm
        if abs(fm) < tol:
            break
    return (x1 + x2) / 2.0


def df_to_lambda(df: float, D: np.ndarray, guess: float = 1.0, tolerance: float = 1e-5) -> float:
    """``Krig.df.to.lambda``: x4 bracketing then bisection on log(lambda)."""
    l1 = guess
    for _ in range(25):
        if tr_a(l1, D) <= df:
            break
        l1 *= 4.0
    l2 = guess
    for _ in range(25):
        if tr_a(l2, D) >= df:
            break
        l2 /= 4.0
    out = _bisection_search(math.log(l1), math.log(l2),
                            lambda ll: tr_a(math.exp(ll), D) - df, tol=tolerance)
    return math.exp(out)


def golden_section_search(ax, bx, cx, f, niter=25, tol=1e-5):
    r = 0.61803399
    con = 1.0 - r
    x0, x3 = ax, cx
    if abs(cx - bx) > abs(bx - ax):
        x1 = bx
        x2 = bx + con * (bx - ax)
    else:
        x2 = bx
        x1 = bx - con * (bx - ax)
    f1, f2 = f(x1), f(x2)
    for _ in range(niter):
        if f2 < f1:
            x0, x1 = x1, x2
            x2 = r * x1 + con * x3
            f1 = f2
            f2 = f(x2)
        else:
            x3, x2 = x2, x1
            x1 = r * x2 + con * x0
            f2 = f1
            f1 = f(x1)
        if abs(f2 - f1) < tol:
            break
    return (x1, f1) if f1 < f2 else (x2, f2)


def gcv_search(D: np.ndarray, u: np.ndarray, n_obs: int, pure_ss: float,
               nstep: int = 200, tol: float = 1e-5):
    """``gcv.Krig`` + ``Krig.find.gcvmin`` for method="GCV": returns (lambda, grid, endpoint?)."""
    npts = D.shape[0]
    df = np.linspace(NT, 0.95 * npts, nstep)
    df[0] += 0.001
    lam_grid = np.sort(np.array([df_to_lambda(x, D) for x in df]))
    g = np.array([gcv_value(l, D, u, n_obs, pure_ss) for l in lam_grid])
    tra = np.array([tr_a(l, D) for l in lam_grid])
    grid = np.stack([lam_grid, tra, g], axis=1)
    ok = ~np.isnan(g)
    lg, gg = lam_grid[ok], g[ok]
    il = int(np.argmin(gg))
    if 0 < il < lg.shape[0] - 1:
        lam, _ = golden_section_search(lg[il - 1], lg[il], lg[il + 1],
                                       lambda l: gcv_value(l, D, u, n_obs, pure_ss),
                                       tol=tol * gg[il])
        return float(lam), grid, False
    return float(lg[il]), grid, True


# ----------------------------------------------------------------------------------------
# Tps fit
# ----------------------------------------------------------------------------------------
def pool_replicates(xy: np.ndarray, y: np.ndarray):
    """``Krig.replicates``: collapse duplicated locations to (mean, count) + pure-error SS."""
    xy = np.asarray(xy, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    uniq, inv, cnt = np.unique(xy, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    if uniq.shape[0] == xy.shape[0]:
        return xy, y, np.ones(xy.shape[0]), 0.0
    sums = np.zeros(uniq.shape[0])
    np.add.at(sums, inv, y)
    yM = sums / cnt
    pure_ss = float(np.sum((y - yM[inv]) ** 2))
    return uniq, yM, cnt.astype(np.float64), pure_ss


def tps_fit(xy: np.ndarray, y: np.ndarray, lam: float | None = None) -> TpsFit:
    """``fields::Tps(x, Y)`` with the defaults used at V73:722 / V73:751.

    m=2, p=2, scale.type="range", method="GCV" (``lam=None``) or a fixed lambda.
    """
    xy = np.asarray(xy, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n_obs = xy.shape[0]
    center = xy.min(axis=0)
    scale = xy.max(axis=0) - center
    if np.any(scale <= 0):
        raise ValueError("degenerate knot cloud (zero range)")
    xyM, yM, wM, pure_ss = pool_replicates(xy, y)
    s = (xyM - center) / scale
    npts = s.shape[0]
    if npts <= NT:
        raise ValueError("need more than 3 unique locations")
    w2 = np.sqrt(wM)
    T = np.column_stack([np.ones(npts), s[:, 0], s[:, 1]])
    qr = _QRT(w2[:, None] * T)
    K = rad_cov(s, s)
    M = w2[:, None] * K * w2[None, :]
    M = qr.q2ty(M)            # Q2' (W K W)
    M = qr.q2ty(M.T)          # Q2' (..)' = Q2' W K W Q2   (symmetric)
    M = 0.5 * (M + M.T)
    eta, V = np.linalg.eigh(M)
    eta, V = eta[::-1], V[:, ::-1]        # R eigen(): decreasing
    D = np.concatenate([np.zeros(NT), 1.0 / eta])
    z = qr.q2ty(w2 * yM)
    u = np.concatenate([np.zeros(NT), V.T @ z])
    grid = None
    endpoint = False
    if lam is None:
        lam, grid, endpoint = gcv_search(D, u, n_obs, pure_ss)
    # Krig.coef (WBW branch)
    beta = V @ (u[NT:] / (eta + lam))
    tmp = np.concatenate([np.zeros(NT), beta])
    c = w2 * qr.qy(tmp)
    d = qr.coef(w2 * (yM - K @ c))
    return TpsFit(center=center, scale=scale, knots_s=s, knots_xy=xyM, c=c, d=d,
                  lam=float(lam), eff_df=tr_a(lam, D), weights=wM, yM=yM, eta=eta, u=u,
                  gcv_grid=grid, pure_ss=pure_ss, n_obs=n_obs, gcv_at_endpoint=endpoint)


# ----------------------------------------------------------------------------------------
# predict.Krig / terra::interpolate
# ----------------------------------------------------------------------------------------
def tps_predict_points(fit: TpsFit, xy: np.ndarray, chunk: int = 4096) -> np.ndarray:
    """``predict.Krig(model, x)``: [1, s] d + Rad.cov(s, knots, C = c)  (Fortran multrb)."""
    xy = np.asarray(xy, dtype=np.float64)
    out = np.empty(xy.shape[0])
    ks = fit.knots_s
    for a in range(0, xy.shape[0], chunk):
        s = (xy[a:a + chunk] - fit.center) / fit.scale
        dx = s[:, None, 0] - ks[None, :, 0]
        dy = s[:, None, 1] - ks[None, :, 1]
        k = radfun(dx * dx + dy * dy)
        out[a:a + chunk] = fit.d[0] + fit.d[1] * s[:, 0] + fit.d[2] * s[:, 1] + RBF_CONSTANT * (k @ fit.c)
    return out


def cell_centres(geom, rows: np.ndarray, cols: np.ndarray):
    """terra xFromCol / yFromRow: x = xmin + (col + 0.5) rx ; y = ymax - (row + 0.5) ry."""
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    rx = (xmax - xmin) / ncol
    ry = (ymax - ymin) / nrow
    return xmin + (np.asarray(cols) + 0.5) * rx, ymax - (np.asarray(rows) + 0.5) * ry


def tps_interpolate(fit: TpsFit, geom, row0=0, row1=None, col0=0, col1=None) -> np.ndarray:
    """``terra::interpolate(rast(template), model)`` on the window rows [row0,row1) x cols [col0,col1).

    The template carries no values, so every cell - land or not - is predicted (V73:726,753);
    output is row-major from the NW corner (terra cell order).
    """
    _, _, _, _, nrow, ncol = geom
    row1 = nrow if row1 is None else row1
    col1 = ncol if col1 is None else col1
    rows = np.arange(row0, row1)
    cols = np.arange(col0, col1)
    x, y = cell_centres(geom, rows, cols)
    out = np.empty((rows.size, cols.size))
    sx = (x - fit.center[0]) / fit.scale[0]
    ks = fit.knots_s
    dx2 = (sx[:, None] - ks[None, :, 0]) ** 2          # cols x np, reused by every row
    for i, yy in enumerate(y):
        sy = (yy - fit.center[1]) / fit.scale[1]
        dy2 = (sy - ks[:, 1]) ** 2
        k = radfun(dx2 + dy2[None, :])
        out[i] = fit.d[0] + fit.d[1] * sx + fit.d[2] * sy + RBF_CONSTANT * (k @ fit.c)
    return out
