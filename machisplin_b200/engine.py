"""Thin Python objects over the C ABI: Engine (context), Spline, Ensemble.

Host arrays are numpy; device buffers are whatever owns a raw CUDA pointer (torch tensors in
bench.py / parallel.py: ``t.data_ptr()``).  All compute happens in libmachisplin_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Grid, Window, Models, check

EVAL_DIRECT = 0
EVAL_FAST = 1
_METHODS = {"direct": EVAL_DIRECT, "fast": EVAL_FAST, 0: 0, 1: 1}


def _pd(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_lib.PD)


def _f64(a, order="C") -> np.ndarray:
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O", order[0]])


@dataclass(frozen=True)
class Geom:
    """Extent + dimensions of a terra SpatRaster."""
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    nrow: int
    ncol: int

    def as_tuple(self):
        return (self.xmin, self.xmax, self.ymin, self.ymax, self.nrow, self.ncol)

    def c(self) -> Grid:
        return Grid(self.xmin, self.xmax, self.ymin, self.ymax, self.nrow, self.ncol)

    @property
    def rx(self):
        return (self.xmax - self.xmin) / self.ncol

    @property
    def ry(self):
        return (self.ymax - self.ymin) / self.nrow

    def full_window(self):
        return (0, self.nrow, 0, self.ncol)


def as_geom(g) -> Geom:
    return g if isinstance(g, Geom) else Geom(*g)


def _win(geom: Geom, window) -> Window:
    w = geom.full_window() if window is None else window
    return Window(int(w[0]), int(w[1]), int(w[2]), int(w[3]))


class Spline:
    """Handle of a fitted thin-plate spline (the part of a ``Krig`` object the hot path needs)."""

    def __init__(self, engine: "Engine", handle: int):
        self.engine = engine
        self._h = C.c_void_p(handle)
        self.np = engine.lib.mb_spline_np(self._h)
        c = np.empty(self.np)
        d = np.empty(3)
        center = np.empty(2)
        scale = np.empty(2)
        knots = np.empty((2, self.np))
        lam, edf, gcv = C.c_double(), C.c_double(), C.c_double()
        check(engine.lib.mb_spline_get(self._h, _pd(c), _pd(d), _pd(center), _pd(scale), _pd(knots),
                                       C.byref(lam), C.byref(edf), C.byref(gcv)))
        self.c, self.d, self.center, self.scale = c, d, center, scale
        self.knots_xy = knots.T.copy()
        self.lam, self.eff_df, self.gcv = lam.value, edf.value, gcv.value

    def decomposition(self):
        """GCV fits: eigenvalues of Q2'KQ2 (decreasing) and the tridiagonal form (diag, off, zhat) the lambda
        search ran on."""
        m = self.np - 3
        eta, dg, of, zh = np.empty(m), np.empty(m), np.empty(m - 1), np.empty(m)
        check(self.engine.lib.mb_spline_get_decomp(self._h, _pd(eta), _pd(dg), _pd(of), _pd(zh)))
        return eta, (dg, of, zh)

    def free(self):
        if self._h:
            self.engine.lib.mb_spline_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Ensemble:
    def __init__(self, engine: "Engine", handle: int, geom: Geom, keepalive):
        self.engine = engine
        self._h = C.c_void_p(handle)
        self.geom = geom
        self._keepalive = keepalive

    def free(self):
        if self._h:
            self.engine.lib.mb_ensemble_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pack_models(models: dict, P: int):
    """dict of flat numpy descriptors (layout documented in include/machisplin_b200.h and
    DESIGN.md section 3) -> (Models struct, keepalive list)."""
    m = Models()
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(np.asarray(a, dtype=dt))
        keep.append(a)
        return a

    m.P = P
    if "g" in models:
        a = arr(models["g"]["coef"], np.float64)
        assert a.size == P + 1, "gam: coef must have P+1 entries"
        m.gam_coef = a.ctypes.data_as(_lib.PD)
    if "n" in models:
        d = models["n"]
        H = int(d["H"])
        a = arr(d["wts"], np.float64)
        assert a.size == (P + 1) * H + H + 1, "nnet: wts has the wrong length"
        m.nn_wts, m.nn_H = a.ctypes.data_as(_lib.PD), H
        m.nn_max2, m.nn_min = float(d["max2"]), float(d["min"])
    if "m" in models:
        d = models["m"]
        dirs = arr(d["dirs"], np.int8)
        cuts = arr(d["cuts"], np.float64)
        coef = arr(d["coef"], np.float64)
        assert dirs.shape == cuts.shape == (coef.size, P)
        m.mars_T = coef.size
        m.mars_dirs = dirs.ctypes.data_as(_lib.PI8)
        m.mars_cuts = cuts.ctypes.data_as(_lib.PD)
        m.mars_coef = coef.ctypes.data_as(_lib.PD)
    if "v" in models:
        d = models["v"]
        sv = arr(d["sv"], np.float64)
        al = arr(d["alpha"], np.float64)
        xc = arr(d["x_center"], np.float64)
        xs = arr(d["x_scale"], np.float64)
        assert sv.shape == (al.size, P) and xc.size == P and xs.size == P
        m.svm_S = al.size
        m.svm_sv, m.svm_alpha = sv.ctypes.data_as(_lib.PD), al.ctypes.data_as(_lib.PD)
        m.svm_b, m.svm_sigma = float(d["b"]), float(d["sigma"])
        m.svm_x_center, m.svm_x_scale = xc.ctypes.data_as(_lib.PD), xs.ctypes.data_as(_lib.PD)
        m.svm_y_center, m.svm_y_scale = float(d["y_center"]), float(d["y_scale"])
    if "r" in models:
        d = models["r"]
        nt, nn = int(d["ntree"]), int(d["nrnodes"])
        left, right = arr(d["left"], np.int32), arr(d["right"], np.int32)
        status, bestvar = arr(d["status"], np.int8), arr(d["bestvar"], np.int32)
        split, pred = arr(d["split"], np.float64), arr(d["nodepred"], np.float64)
        for a in (left, right, status, bestvar, split, pred):
            assert a.shape == (nt, nn)
        m.rf_ntree, m.rf_nrnodes = nt, nn
        m.rf_left, m.rf_right = left.ctypes.data_as(_lib.PI32), right.ctypes.data_as(_lib.PI32)
        m.rf_status, m.rf_bestvar = status.ctypes.data_as(_lib.PI8), bestvar.ctypes.data_as(_lib.PI32)
        m.rf_split, m.rf_nodepred = split.ctypes.data_as(_lib.PD), pred.ctypes.data_as(_lib.PD)
    if "b" in models:
        d = models["b"]
        off = arr(d["tree_off"], np.int32)
        sv_, sc = arr(d["splitvar"], np.int32), arr(d["splitcode"], np.float64)
        ln, rn, mn = arr(d["left"], np.int32), arr(d["right"], np.int32), arr(d["missing"], np.int32)
        m.gbm_ntrees, m.gbm_initF = off.size - 1, float(d["initF"])
        m.gbm_tree_off = off.ctypes.data_as(_lib.PI32)
        m.gbm_splitvar, m.gbm_splitcode = sv_.ctypes.data_as(_lib.PI32), sc.ctypes.data_as(_lib.PD)
        m.gbm_left, m.gbm_right = ln.ctypes.data_as(_lib.PI32), rn.ctypes.data_as(_lib.PI32)
        m.gbm_missing = mn.ctypes.data_as(_lib.PI32)
    return m, keep


def tiles_owned_window(lib, geom, wins, ncol: int, nrow: int, t: int) -> tuple:
    """``mb_tiles_owned_window``: needs the library but neither a context nor a GPU."""
    geom = as_geom(geom)
    nt = ncol * nrow
    wa = (Window * nt)(*[Window(*map(int, w)) for w in wins])
    own = Window()
    g = geom.c()
    check(lib.mb_tiles_owned_window(C.byref(g), ncol, nrow, wa, int(t), C.byref(own)))
    return (own.r0, own.r1, own.c0, own.c1)


class Engine:
    """One context = one GPU (one process per GPU in multi-GPU runs)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if self.lib.mb_device_count() <= 0:
            raise RuntimeError("machisplin_b200: no CUDA device visible; the engine has no CPU fallback")
        h = C.c_void_p()
        check(self.lib.mb_init(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.mb_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.mb_sync(self._h))

    @property
    def launches(self) -> int:
        return int(self.lib.mb_launch_count(self._h))

    def timing(self, on: bool):
        check(self.lib.mb_timing_enable(self._h, 1 if on else 0))

    def timing_collect(self) -> dict:
        """{kernel name: (total device ms, launches)} since the last collect (CUDA events on the launching stream)."""
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = self.lib.mb_timing_collect(self._h, cap, names, ms, cnt)
        if n < 0:
            check(n)
        return {names[i].decode(): (ms[i], int(cnt[i])) for i in range(n)}

    def set_fast_eval_params(self, cheb_p=0, leaf_cols=0, leaf_rows=0):
        check(self.lib.mb_set_fast_eval_params(self._h, cheb_p, leaf_cols, leaf_rows))

    def debug_values(self, name: str, cap: int = 16):
        out = np.zeros(cap)
        n = self.lib.mb_debug_values(self._h, name.encode(), _pd(out), cap)
        if n < 0:
            check(n)
        return out[:n]

    def set_param(self, name: str, value: int):
        check(self.lib.mb_set_param(self._h, name.encode(), int(value)))

    # -- a1: fields::Tps ---------------------------------------------------------------------------
    def tps_fit(self, xy, y, lam: Optional[float] = None):
        """``fields::Tps(xy, y)`` (V73:722, 751).  y may be (n,) or (n, L); returns Spline or list."""
        xy = np.asarray(xy, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        single = y.ndim == 1
        Y = y.reshape(len(y), -1)
        n, L = Y.shape
        assert xy.shape == (n, 2)
        xy_f = np.asfortranarray(xy)
        Y_f = np.asfortranarray(Y)
        hs = (C.c_void_p * L)()
        check(self.lib.mb_tps_fit(self._h, _pd(xy_f), _pd(Y_f), n, L, -1.0 if lam is None else float(lam), hs))
        out = [Spline(self, h) for h in hs]
        return out[0] if single else out

    def spline_create(self, knots_xy, c, d, center, scale) -> Spline:
        k = np.asfortranarray(np.asarray(knots_xy, dtype=np.float64))
        c, d, center, scale = _f64(c), _f64(d), _f64(center), _f64(scale)
        h = C.c_void_p()
        check(self.lib.mb_spline_create(self._h, _pd(k), k.shape[0], _pd(c), _pd(d), _pd(center), _pd(scale),
                                        C.byref(h)))
        return Spline(self, h.value)

    # -- a2: terra::interpolate ------------------------------------------------------------------
    def tps_eval(self, spline: Spline, geom, window=None, method="fast") -> np.ndarray:
        geom = as_geom(geom)
        w = _win(geom, window)
        out = np.empty((w.r1 - w.r0, w.c1 - w.c0))
        g = geom.c()
        check(self.lib.mb_tps_eval(self._h, spline._h, C.byref(g), C.byref(w), _METHODS[method], _pd(out)))
        return out

    def tps_eval_dev(self, spline: Spline, geom, out_ptr: int, row_stride: int, window=None, method="fast",
                     stream: int = 0):
        geom = as_geom(geom)
        w = _win(geom, window)
        g = geom.c()
        check(self.lib.mb_tps_eval_dev(self._h, spline._h, C.byref(g), C.byref(w), _METHODS[method],
                                       C.c_void_p(out_ptr), row_stride, C.c_void_p(stream)))

    def tps_predict_points(self, spline: Spline, xy) -> np.ndarray:
        xy = np.asfortranarray(np.asarray(xy, dtype=np.float64))
        out = np.empty(xy.shape[0])
        check(self.lib.mb_tps_predict_points(self._h, spline._h, _pd(xy), xy.shape[0], _pd(out)))
        return out

    # -- a5: ensemble ----------------------------------------------------------------------------------
    def ensemble_create(self, geom, models: dict, kept: str, w: Sequence[float], w_total: float, P: int) -> Ensemble:
        geom = as_geom(geom)
        m, keep = pack_models({k: models[k] for k in kept}, P)
        wv = _f64(w)
        assert wv.size == len(kept)
        h = C.c_void_p()
        g = geom.c()
        check(self.lib.mb_ensemble_create(self._h, C.byref(g), C.byref(m), kept.encode(), _pd(wv), float(w_total),
                                          C.byref(h)))
        return Ensemble(self, h.value, geom, keep)

    def ensemble_eval(self, ens: Ensemble, cov: Optional[np.ndarray], spline: Optional[Spline] = None,
                      tps_surface: Optional[np.ndarray] = None, window=None) -> np.ndarray:
        """pred.elev (+ TPS) on a window; cov is (C, nrow, ncol) float32 of the full grid."""
        geom = ens.geom
        w = _win(geom, window)
        Cn = 0 if cov is None else cov.shape[0]
        if cov is not None:
            cov = np.ascontiguousarray(cov, dtype=np.float32)
            assert cov.shape == (Cn, geom.nrow, geom.ncol)
        out = np.empty((w.r1 - w.r0, w.c1 - w.c0))
        ts = None if tps_surface is None else np.ascontiguousarray(tps_surface, dtype=np.float64)
        check(self.lib.mb_ensemble_eval(self._h, ens._h, None if cov is None else cov.ctypes.data_as(_lib.PF), Cn,
                                        spline._h if spline is not None else None, _pd(ts), C.byref(w), _pd(out)))
        return out

    def ensemble_eval_dev(self, ens: Ensemble, cov_ptr: int, Cn: int, out_ptr: int, spline: Optional[Spline] = None,
                          tps_surface_ptr: int = 0, window=None, stream: int = 0):
        w = _win(ens.geom, window)
        check(self.lib.mb_ensemble_eval_dev(self._h, ens._h, C.c_void_p(cov_ptr), Cn,
                                            spline._h if spline is not None else None,
                                            C.c_void_p(tps_surface_ptr) if tps_surface_ptr else None, C.byref(w),
                                            C.c_void_p(out_ptr), C.c_void_p(stream)))

    def ensemble_predict_points(self, ens: Ensemble, X) -> np.ndarray:
        X = np.asfortranarray(np.asarray(X, dtype=np.float64))
        out = np.empty(X.shape[0])
        check(self.lib.mb_ensemble_predict_points(self._h, ens._h, _pd(X), X.shape[0], _pd(out)))
        return out

    # -- a3 + a4 -----------------------------------------------------------------------------------------
    def tiles_tps(self, geom, knots_xy, resid, tile_px=1500, fit_halo=0.2, keep_halo=0.025, min_pts=10,
                  lam: Optional[float] = None, method="fast") -> np.ndarray:
        geom = as_geom(geom)
        k = np.asfortranarray(np.asarray(knots_xy, dtype=np.float64))
        r = _f64(resid)
        out = np.empty((geom.nrow, geom.ncol))
        g = geom.c()
        check(self.lib.mb_tiles_tps(self._h, C.byref(g), _pd(k), _pd(r), k.shape[0], tile_px, fit_halo, keep_halo,
                                    min_pts, -1.0 if lam is None else float(lam), _METHODS[method], _pd(out)))
        return out

    def tiles_tps_dev(self, geom, knots_xy, resid, out_ptr: int, tile_px=1500, fit_halo=0.2, keep_halo=0.025,
                      min_pts=10, lam: Optional[float] = None, method="fast", stream: int = 0):
        geom = as_geom(geom)
        k = np.asfortranarray(np.asarray(knots_xy, dtype=np.float64))
        r = _f64(resid)
        g = geom.c()
        check(self.lib.mb_tiles_tps_dev(self._h, C.byref(g), _pd(k), _pd(r), k.shape[0], tile_px, fit_halo, keep_halo,
                                        min_pts, -1.0 if lam is None else float(lam), _METHODS[method],
                                        C.c_void_p(out_ptr), C.c_void_p(stream)))

    def tiles_merge(self, geom, wins, rasters, ncol: int, nrow: int) -> np.ndarray:
        geom = as_geom(geom)
        nt = ncol * nrow
        assert len(wins) == nt and len(rasters) == nt
        wa = (Window * nt)(*[Window(*map(int, w)) for w in wins])
        rs = [np.ascontiguousarray(r, dtype=np.float64) for r in rasters]
        for w, r in zip(wins, rs):
            assert r.shape == (w[1] - w[0], w[3] - w[2])
        pa = (_lib.PD * nt)(*[_pd(r) for r in rs])
        out = np.empty((geom.nrow, geom.ncol))
        g = geom.c()
        check(self.lib.mb_tiles_merge(self._h, C.byref(g), ncol, nrow, wa, pa, _pd(out)))
        return out

    def tiles_merge_dev(self, geom, wins, tile_ptrs: Sequence[int], ncol: int, nrow: int, out_ptr: int, stream: int = 0):
        """``machisplin.tiles.merge`` on device-resident tile rasters (row-major, window-shaped); asynchronous."""
        geom = as_geom(geom)
        nt = ncol * nrow
        assert len(wins) == nt and len(tile_ptrs) == nt
        wa = (Window * nt)(*[Window(*map(int, w)) for w in wins])
        pa = (C.c_void_p * nt)(*[C.c_void_p(int(p)) for p in tile_ptrs])
        g = geom.c()
        check(self.lib.mb_tiles_merge_dev(self._h, C.byref(g), ncol, nrow, wa, pa, C.c_void_p(out_ptr), C.c_void_p(stream)))

    def tiles_owned_window(self, geom, wins, ncol: int, nrow: int, t: int) -> tuple:
        """(r0, r1, c0, c1) of the cells tile ``t`` owns in the sharded merge (host arithmetic only)."""
        return tiles_owned_window(self.lib, geom, wins, ncol, nrow, t)

    def tiles_merge_shard_dev(self, geom, wins, my_tile_ptrs: dict, ncol: int, nrow: int, out_ptrs: dict, stream: int = 0):
        """``machisplin.tiles.merge`` with the tiles spread over the ranks of this engine's communicator (tile t on rank
        t % world): ``my_tile_ptrs`` / ``out_ptrs`` map the tile indices of THIS rank to device pointers of the tile raster
        (window-shaped) and of the owned-window-shaped output.  Seam strips travel over NCCL inside the library."""
        geom = as_geom(geom)
        nt = ncol * nrow
        assert len(wins) == nt
        wa = (Window * nt)(*[Window(*map(int, w)) for w in wins])
        pa = (C.c_void_p * nt)(*[C.c_void_p(int(my_tile_ptrs[t])) if t in my_tile_ptrs else None for t in range(nt)])
        po = (C.c_void_p * nt)(*[C.c_void_p(int(out_ptrs[t])) if t in out_ptrs else None for t in range(nt)])
        g = geom.c()
        check(self.lib.mb_tiles_merge_shard_dev(self._h, C.byref(g), ncol, nrow, wa, pa, po, C.c_void_p(stream)))

    # -- a6 / a7 --------------------------------------------------------------------------------------------
    def gram(self, R) -> np.ndarray:
        R = np.asfortranarray(np.asarray(R, dtype=np.float64))
        n, K = R.shape
        G = np.empty((K, K))
        check(self.lib.mb_gram(self._h, _pd(R), n, K, _pd(G)))
        return G

    def gather_cells_dev(self, raster_ptr: int, row_stride: int, nrow: int, ncol: int, row, col, stream: int = 0) -> np.ndarray:
        """``f.actual <- extract(final, points)`` (V73:910) on a device raster; NaN for cells outside it."""
        row = np.ascontiguousarray(row, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        out = np.empty(row.size)
        check(self.lib.mb_gather_cells_dev(self._h, C.c_void_p(raster_ptr), row_stride, int(nrow), int(ncol),
                                           row.ctypes.data_as(_lib.PI32), col.ctypes.data_as(_lib.PI32), row.size,
                                           _pd(out), C.c_void_p(stream)))
        return out

    # -- multi-GPU: NCCL communicator of the context (one process per GPU) ------------------------------------
    def comm_unique_id(self) -> bytes:
        """Rank 0 creates the id; the host ships its 128 bytes to the other processes, each calls ``comm_init``."""
        buf = C.create_string_buffer(128)
        check(self.lib.mb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        assert len(uid) == 128
        check(self.lib.mb_comm_init(self._h, int(nranks), int(rank), C.c_char_p(uid)))

    def comm_destroy(self):
        check(self.lib.mb_comm_destroy(self._h))

    @property
    def rank(self) -> int:
        return int(self.lib.mb_comm_rank(self._h))

    @property
    def world(self) -> int:
        return int(self.lib.mb_comm_size(self._h))

    def comm_backend(self) -> str:
        return self.lib.mb_comm_backend().decode()

    def allreduce(self, values, op: str = "sum") -> np.ndarray:
        v = np.ascontiguousarray(np.atleast_1d(np.asarray(values, dtype=np.float64))).copy()
        check(self.lib.mb_comm_allreduce_f64(self._h, _pd(v), v.size, {"sum": 0, "max": 1}[op]))
        return v

    def gram_allreduce(self, R_local) -> np.ndarray:
        """a6 with the k-fold residual rows sharded over the ranks (V73:329-333): every rank gets G = R'R."""
        R = np.asfortranarray(np.asarray(R_local, dtype=np.float64))
        n, K = R.shape
        G = np.empty((K, K))
        check(self.lib.mb_gram_allreduce(self._h, _pd(R) if n else None, n, K, _pd(G)))
        return G

    def spline_bcast(self, spline: Optional[Spline], max_knots: int, root: int = 0) -> Spline:
        h = C.c_void_p(spline._h.value if spline is not None else None)
        check(self.lib.mb_spline_bcast(self._h, C.byref(h), int(max_knots), int(root)))
        if self.rank == root:
            return spline
        return Spline(self, h.value)

    def mltps_predict_shard_dev(self, geom_block, ens: Optional[Ensemble], cov_ptr: int, Cn: int, knots_xy, resid, n: int,
                                out_ptr: int, lam: Optional[float] = None, root: int = 0, stream: int = 0):
        """Parts 2-5 for this rank's row block of a raster sharded over the communicator; one fit (on ``root``), broadcast."""
        geom = as_geom(geom_block)
        k = r = None
        if knots_xy is not None:
            k = np.asfortranarray(np.asarray(knots_xy, dtype=np.float64))
            r = _f64(resid)
            assert k.shape == (n, 2) and r.shape == (n,)
        g = geom.c()
        h = C.c_void_p()
        check(self.lib.mb_mltps_predict_shard_dev(self._h, C.byref(g), ens._h if ens is not None else None,
                                                  C.c_void_p(cov_ptr) if cov_ptr else None, Cn, _pd(k), _pd(r), int(n),
                                                  -1.0 if lam is None else float(lam), int(root), C.c_void_p(out_ptr),
                                                  C.byref(h), C.c_void_p(stream)))
        return Spline(self, h.value) if h.value else None

    def mltps_predict_shard(self, geom_block, ens: Optional[Ensemble], cov: Optional[np.ndarray], knots_xy, resid, n: int,
                            lam: Optional[float] = None, root: int = 0, out: Optional[np.ndarray] = None):
        """Host-buffer twin of ``mltps_predict_shard_dev``.  Returns (block raster, Spline)."""
        geom = as_geom(geom_block)
        k = r = None
        if knots_xy is not None:
            k = np.asfortranarray(np.asarray(knots_xy, dtype=np.float64))
            r = _f64(resid)
            assert k.shape == (n, 2) and r.shape == (n,)
        Cn = 0 if cov is None else cov.shape[0]
        if cov is not None:
            cov = np.ascontiguousarray(cov, dtype=np.float32)
            assert cov.shape == (Cn, geom.nrow, geom.ncol)
        if out is None:
            out = np.empty((geom.nrow, geom.ncol))
        g = geom.c()
        h = C.c_void_p()
        check(self.lib.mb_mltps_predict_shard(self._h, C.byref(g), ens._h if ens is not None else None,
                                              None if cov is None else cov.ctypes.data_as(_lib.PF), Cn, _pd(k), _pd(r), int(n),
                                              -1.0 if lam is None else float(lam), int(root), _pd(out), C.byref(h)))
        return out, (Spline(self, h.value) if h.value else None)

    # -- mltps parts 2-5 in one call (V73:442-932) ---------------------------------------------------------
    def _mltps_args(self, geom, ens, knots_xy, resid, lam, tile_px):
        geom = as_geom(geom)
        k = r = None
        n = 0
        if knots_xy is not None:
            k = np.asfortranarray(np.asarray(knots_xy, dtype=np.float64))
            r = _f64(resid)
            n = k.shape[0]
            assert k.shape == (n, 2) and r.shape == (n,)
        return geom, k, r, n, (-1.0 if lam is None else float(lam)), int(tile_px)

    def mltps_predict_dev(self, geom, ens: Optional[Ensemble], cov_ptr: int, Cn: int, knots_xy, resid, out_ptr: int,
                          lam: Optional[float] = None, tile_px: int = 0, stream: int = 0, want_spline: bool = True):
        """Ensemble raster prediction (part 2) overlapped with the TPS fit of the residuals (part 3), then the
        fused surface + combine pass (parts 3-5).  Device pointers; asynchronous on ``stream`` after the fit."""
        geom, k, r, n, lamv, tile_px = self._mltps_args(geom, ens, knots_xy, resid, lam, tile_px)
        g = geom.c()
        h = C.c_void_p()
        check(self.lib.mb_mltps_predict_dev(self._h, C.byref(g), ens._h if ens is not None else None,
                                            C.c_void_p(cov_ptr) if cov_ptr else None, Cn, _pd(k), _pd(r), n, lamv,
                                            tile_px, C.c_void_p(out_ptr), C.byref(h) if want_spline else None,
                                            C.c_void_p(stream)))
        return Spline(self, h.value) if (want_spline and h.value) else None

    def mltps_predict(self, geom, ens: Optional[Ensemble], cov: Optional[np.ndarray], knots_xy, resid,
                      lam: Optional[float] = None, tile_px: int = 0, out: Optional[np.ndarray] = None):
        """Host-buffer twin: cov (C, nrow, ncol) float32 in, (nrow, ncol) float64 out.  Returns (raster, Spline|None)."""
        geom, k, r, n, lamv, tile_px = self._mltps_args(geom, ens, knots_xy, resid, lam, tile_px)
        Cn = 0 if cov is None else cov.shape[0]
        if cov is not None:
            cov = np.ascontiguousarray(cov, dtype=np.float32)
            assert cov.shape == (Cn, geom.nrow, geom.ncol)
        if out is None:
            out = np.empty((geom.nrow, geom.ncol))
        assert out.shape == (geom.nrow, geom.ncol) and out.dtype == np.float64 and out.flags.c_contiguous
        g = geom.c()
        h = C.c_void_p()
        check(self.lib.mb_mltps_predict(self._h, C.byref(g), ens._h if ens is not None else None,
                                        None if cov is None else cov.ctypes.data_as(_lib.PF), Cn, _pd(k), _pd(r), n,
                                        lamv, tile_px, _pd(out), C.byref(h)))
        return out, (Spline(self, h.value) if h.value else None)
