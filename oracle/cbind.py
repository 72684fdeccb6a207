"""ctypes binding of oracle/c/liboracle.so (the C restatement used as checker on larger grids and as
the CPU baseline of bench.py).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent / "c"
LIB = HERE / "liboracle.so"
_lib = None
PD, PF, PI32, PI8 = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int8)


class OrcModels(C.Structure):
    _fields_ = [
        ("P", C.c_int),
        ("gam_coef", PD),
        ("nn_wts", PD), ("nn_H", C.c_int), ("nn_max2", C.c_double), ("nn_min", C.c_double),
        ("mars_T", C.c_int), ("mars_dirs", PI8), ("mars_cuts", PD), ("mars_coef", PD),
        ("svm_S", C.c_int), ("svm_sv", PD), ("svm_alpha", PD), ("svm_b", C.c_double), ("svm_sigma", C.c_double),
        ("svm_x_center", PD), ("svm_x_scale", PD), ("svm_y_center", C.c_double), ("svm_y_scale", C.c_double),
        ("rf_ntree", C.c_int), ("rf_nrnodes", C.c_int),
        ("rf_left", PI32), ("rf_right", PI32), ("rf_status", PI8), ("rf_bestvar", PI32),
        ("rf_split", PD), ("rf_nodepred", PD),
        ("gbm_ntrees", C.c_int), ("gbm_initF", C.c_double), ("gbm_tree_off", PI32),
        ("gbm_splitvar", PI32), ("gbm_splitcode", PD), ("gbm_left", PI32), ("gbm_right", PI32),
        ("gbm_missing", PI32),
    ]


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
        _lib = C.CDLL(str(LIB))
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(load().orc_max_threads())


def _p(a, t):
    return a.ctypes.data_as(t)


def tps_eval(fit, geom, window=None, threads=0) -> np.ndarray:
    """multrb-style double loop over the window; fit is an oracle TpsFit (or anything with the same fields)."""
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    r0, r1, c0, c1 = window if window is not None else (0, nrow, 0, ncol)
    ks = np.ascontiguousarray(fit.knots_s)
    sx, sy = np.ascontiguousarray(ks[:, 0]), np.ascontiguousarray(ks[:, 1])
    c = np.ascontiguousarray(fit.c, dtype=np.float64)
    d = np.ascontiguousarray(fit.d, dtype=np.float64)
    cen = np.ascontiguousarray(fit.center, dtype=np.float64)
    sc = np.ascontiguousarray(fit.scale, dtype=np.float64)
    out = np.empty((r1 - r0, c1 - c0))
    load().orc_tps_eval(_p(sx, PD), _p(sy, PD), _p(c, PD), C.c_int(c.size), _p(d, PD), _p(cen, PD), _p(sc, PD),
                        C.c_double(xmin), C.c_double(ymax), C.c_double((xmax - xmin) / ncol),
                        C.c_double((ymax - ymin) / nrow), C.c_int(r0), C.c_int(r1), C.c_int(c0), C.c_int(c1),
                        _p(out, PD), C.c_int(threads))
    return out


def pack_models(models: dict, P: int):
    m = OrcModels()
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(np.asarray(a, dtype=dt))
        keep.append(a)
        return a

    m.P = P
    if "g" in models:
        m.gam_coef = _p(arr(models["g"]["coef"], np.float64), PD)
    if "n" in models:
        d = models["n"]
        m.nn_wts, m.nn_H = _p(arr(d["wts"], np.float64), PD), int(d["H"])
        m.nn_max2, m.nn_min = float(d["max2"]), float(d["min"])
    if "m" in models:
        d = models["m"]
        m.mars_T = len(d["coef"])
        m.mars_dirs, m.mars_cuts = _p(arr(d["dirs"], np.int8), PI8), _p(arr(d["cuts"], np.float64), PD)
        m.mars_coef = _p(arr(d["coef"], np.float64), PD)
    if "v" in models:
        d = models["v"]
        m.svm_S = len(d["alpha"])
        m.svm_sv, m.svm_alpha = _p(arr(d["sv"], np.float64), PD), _p(arr(d["alpha"], np.float64), PD)
        m.svm_b, m.svm_sigma = float(d["b"]), float(d["sigma"])
        m.svm_x_center, m.svm_x_scale = _p(arr(d["x_center"], np.float64), PD), _p(arr(d["x_scale"], np.float64), PD)
        m.svm_y_center, m.svm_y_scale = float(d["y_center"]), float(d["y_scale"])
    if "r" in models:
        d = models["r"]
        m.rf_ntree, m.rf_nrnodes = int(d["ntree"]), int(d["nrnodes"])
        m.rf_left, m.rf_right = _p(arr(d["left"], np.int32), PI32), _p(arr(d["right"], np.int32), PI32)
        m.rf_status, m.rf_bestvar = _p(arr(d["status"], np.int8), PI8), _p(arr(d["bestvar"], np.int32), PI32)
        m.rf_split, m.rf_nodepred = _p(arr(d["split"], np.float64), PD), _p(arr(d["nodepred"], np.float64), PD)
    if "b" in models:
        d = models["b"]
        m.gbm_ntrees, m.gbm_initF = len(d["tree_off"]) - 1, float(d["initF"])
        m.gbm_tree_off = _p(arr(d["tree_off"], np.int32), PI32)
        m.gbm_splitvar, m.gbm_splitcode = _p(arr(d["splitvar"], np.int32), PI32), _p(arr(d["splitcode"], np.float64), PD)
        m.gbm_left, m.gbm_right = _p(arr(d["left"], np.int32), PI32), _p(arr(d["right"], np.int32), PI32)
        m.gbm_missing = _p(arr(d["missing"], np.int32), PI32)
    return m, keep


def ensemble_eval(models: dict, kept: str, w, w_total, cov: np.ndarray, geom, window=None, tps=None, threads=0):
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    r0, r1, c0, c1 = window if window is not None else (0, nrow, 0, ncol)
    Cn = cov.shape[0]
    cov = np.ascontiguousarray(cov, dtype=np.float32)
    m, keep = pack_models({k: models[k] for k in kept}, Cn + 2)
    wv = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty((r1 - r0, c1 - c0))
    tp = None if tps is None else _p(np.ascontiguousarray(tps, dtype=np.float64), PD)
    load().orc_ensemble_eval(C.byref(m), kept.encode(), _p(wv, PD), C.c_double(w_total), _p(cov, PF), C.c_int(Cn),
                             C.c_int(nrow), C.c_int(ncol), C.c_double(xmin), C.c_double(ymax),
                             C.c_double((xmax - xmin) / ncol), C.c_double((ymax - ymin) / nrow), C.c_int(r0),
                             C.c_int(r1), C.c_int(c0), C.c_int(c1), tp, _p(out, PD), C.c_int(threads))
    return out


def interpolate_c(fit, geom, r0=0, r1=None, c0=0, c1=None):
    """Drop-in for ``oracle.tps.tps_interpolate`` (same arithmetic, the C + OpenMP loop): the ``interpolate`` hook of
    ``oracle.tiles.tps_tiled_surface`` / ``oracle.mltps.mltps_one_response`` on config-sized rasters."""
    nrow, ncol = int(geom[4]), int(geom[5])
    return tps_eval(fit, geom, (r0, nrow if r1 is None else r1, c0, ncol if c1 is None else c1))


def ensemble_raster_c(geom, cov, models, kept, w, w_total):
    """Drop-in for ``oracle.mltps.ensemble_raster`` (pred.elev, V73:447-619) through the C loops."""
    return ensemble_eval(models, kept, w, w_total, cov, geom)
