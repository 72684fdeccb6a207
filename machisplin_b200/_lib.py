"""ctypes binding of libmachisplin_b200.so (include/machisplin_b200.h).

The product path has no CPU fallback: if the shared library is missing or does not export a
declared symbol this module raises at import of the symbol, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# MB_LIB: another build of the SAME library (tools/build_variants.sh: kernels compiled with other tuning constants for A/B runs)
LIB_PATH = Path(os.environ["MB_LIB"]) if os.environ.get("MB_LIB") else HERE / "libmachisplin_b200.so"


def _pin_nccl():
    """One NCCL per process.  The library binds NCCL with dlopen at the first collective; a Python host usually also runs torch,
    whose libtorch_cuda.so NEEDS the libnccl.so.2 bundled with it (nvidia/nccl).  If the library mapped the SYSTEM copy first
    (older, same SONAME), a later ``import torch`` would resolve against that copy and fail on a missing symbol - so, unless the
    user chose a library, point MB_NCCL_LIB at the copy torch will load (found without importing torch)."""
    if os.environ.get("MB_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["MB_NCCL_LIB"] = cand
                return
    except Exception:
        pass


_pin_nccl()


class MbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"machisplin_b200 error {code}: {msg}")
        self.code = code


class Grid(C.Structure):
    _fields_ = [("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
                ("nrow", C.c_int32), ("ncol", C.c_int32)]


class Window(C.Structure):
    _fields_ = [("r0", C.c_int32), ("r1", C.c_int32), ("c0", C.c_int32), ("c1", C.c_int32)]


PD = C.POINTER(C.c_double)
PF = C.POINTER(C.c_float)
PI32 = C.POINTER(C.c_int32)
PI8 = C.POINTER(C.c_int8)


class Models(C.Structure):
    _fields_ = [
        ("P", C.c_int32),
        ("gam_coef", PD),
        ("nn_wts", PD), ("nn_H", C.c_int32), ("nn_max2", C.c_double), ("nn_min", C.c_double),
        ("mars_T", C.c_int32), ("mars_dirs", PI8), ("mars_cuts", PD), ("mars_coef", PD),
        ("svm_S", C.c_int32), ("svm_sv", PD), ("svm_alpha", PD), ("svm_b", C.c_double), ("svm_sigma", C.c_double),
        ("svm_x_center", PD), ("svm_x_scale", PD), ("svm_y_center", C.c_double), ("svm_y_scale", C.c_double),
        ("rf_ntree", C.c_int32), ("rf_nrnodes", C.c_int32),
        ("rf_left", PI32), ("rf_right", PI32), ("rf_status", PI8), ("rf_bestvar", PI32),
        ("rf_split", PD), ("rf_nodepred", PD),
        ("gbm_ntrees", C.c_int32), ("gbm_initF", C.c_double), ("gbm_tree_off", PI32),
        ("gbm_splitvar", PI32), ("gbm_splitcode", PD), ("gbm_left", PI32), ("gbm_right", PI32),
        ("gbm_missing", PI32),
    ]


# name -> (restype, argtypes); every symbol include/machisplin_b200.h declares
VP = C.c_void_p
PVP = C.POINTER(C.c_void_p)
PG = C.POINTER(Grid)
PW = C.POINTER(Window)
SIGNATURES = {
    "mb_version": (C.c_int, []),
    "mb_device_count": (C.c_int, []),
    "mb_last_error": (C.c_char_p, []),
    "mb_init": (C.c_int, [C.c_int, PVP]),
    "mb_shutdown": (None, [VP]),
    "mb_sync": (C.c_int, [VP]),
    "mb_launch_count": (C.c_int64, [VP]),
    "mb_tps_fit": (C.c_int, [VP, PD, PD, C.c_int, C.c_int, C.c_double, PVP]),
    "mb_spline_create": (C.c_int, [VP, PD, C.c_int, PD, PD, PD, PD, PVP]),
    "mb_spline_np": (C.c_int, [VP]),
    "mb_spline_get": (C.c_int, [VP, PD, PD, PD, PD, PD, PD, PD, PD]),
    "mb_spline_get_decomp": (C.c_int, [VP, PD, PD, PD, PD]),
    "mb_spline_free": (None, [VP]),
    "mb_debug_values": (C.c_int, [VP, C.c_char_p, PD, C.c_int]),
    "mb_tps_eval": (C.c_int, [VP, VP, PG, PW, C.c_int, PD]),
    "mb_tps_eval_dev": (C.c_int, [VP, VP, PG, PW, C.c_int, VP, C.c_int64, VP]),
    "mb_tps_predict_points": (C.c_int, [VP, VP, PD, C.c_int, PD]),
    "mb_ensemble_create": (C.c_int, [VP, PG, C.POINTER(Models), C.c_char_p, PD, C.c_double, PVP]),
    "mb_ensemble_free": (None, [VP]),
    "mb_ensemble_eval": (C.c_int, [VP, VP, PF, C.c_int, VP, PD, PW, PD]),
    "mb_ensemble_eval_dev": (C.c_int, [VP, VP, VP, C.c_int, VP, VP, PW, VP, VP]),
    "mb_ensemble_predict_points": (C.c_int, [VP, VP, PD, C.c_int, PD]),
    "mb_tiles_tps": (C.c_int, [VP, PG, PD, PD, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                               C.c_int, PD]),
    "mb_tiles_tps_dev": (C.c_int, [VP, PG, PD, PD, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                   C.c_int, VP, VP]),
    "mb_tiles_merge": (C.c_int, [VP, PG, C.c_int, C.c_int, PW, C.POINTER(PD), PD]),
    "mb_tiles_merge_dev": (C.c_int, [VP, PG, C.c_int, C.c_int, PW, PVP, VP, VP]),
    "mb_tiles_owned_window": (C.c_int, [PG, C.c_int, C.c_int, PW, C.c_int, PW]),
    "mb_tiles_merge_shard_dev": (C.c_int, [VP, PG, C.c_int, C.c_int, PW, PVP, PVP, VP]),
    "mb_gram": (C.c_int, [VP, PD, C.c_int, C.c_int, PD]),
    "mb_gram_dev": (C.c_int, [VP, VP, C.c_int, C.c_int, VP, VP]),
    "mb_gather_cells_dev": (C.c_int, [VP, VP, C.c_int64, C.c_int, C.c_int, PI32, PI32, C.c_int, PD, VP]),
    "mb_mltps_predict_dev": (C.c_int, [VP, PG, VP, VP, C.c_int, PD, PD, C.c_int, C.c_double, C.c_int, VP, PVP, VP]),
    "mb_mltps_predict": (C.c_int, [VP, PG, VP, PF, C.c_int, PD, PD, C.c_int, C.c_double, C.c_int, PD, PVP]),
    "mb_comm_unique_id": (C.c_int, [VP]),
    "mb_comm_init": (C.c_int, [VP, C.c_int, C.c_int, VP]),
    "mb_comm_destroy": (C.c_int, [VP]),
    "mb_comm_rank": (C.c_int, [VP]),
    "mb_comm_size": (C.c_int, [VP]),
    "mb_comm_backend": (C.c_char_p, []),
    "mb_comm_allreduce_f64": (C.c_int, [VP, PD, C.c_int, C.c_int]),
    "mb_gram_allreduce": (C.c_int, [VP, PD, C.c_int, C.c_int, PD]),
    "mb_spline_bcast": (C.c_int, [VP, PVP, C.c_int, C.c_int]),
    "mb_mltps_predict_shard_dev": (C.c_int, [VP, PG, VP, VP, C.c_int, PD, PD, C.c_int, C.c_double, C.c_int, VP, PVP, VP]),
    "mb_mltps_predict_shard": (C.c_int, [VP, PG, VP, PF, C.c_int, PD, PD, C.c_int, C.c_double, C.c_int, PD, PVP]),
    "mbC_gram": (None, [PD, PI32, PI32, PD, PI32]),
    "mbC_tps_surface": (None, [PD, PD, PI32, PD, PD, PI32, PD, PD, PI32]),
    "mbC_tiles_merge": (None, [PD, PI32, PI32, PI32, PD, PD, PI32]),
    "mbC_last_error": (None, [C.POINTER(C.c_char_p)]),
    "mbC_shutdown": (None, []),
    "mb_dev_alloc": (C.c_int, [VP, C.c_size_t, PVP]),
    "mb_dev_free": (C.c_int, [VP, VP]),
    "mb_h2d": (C.c_int, [VP, VP, VP, C.c_size_t]),
    "mb_d2h": (C.c_int, [VP, VP, VP, C.c_size_t]),
    "mb_timing_enable": (C.c_int, [VP, C.c_int]),
    "mb_timing_collect": (C.c_int, [VP, C.c_int, C.POINTER(C.c_char_p), PD, C.POINTER(C.c_int64)]),
    "mb_set_fast_eval_params": (C.c_int, [VP, C.c_int, C.c_int, C.c_int]),
    "mb_set_param": (C.c_int, [VP, C.c_char_p, C.c_int]),
    "mb_tiff_info": (C.c_int, [C.c_char_p, VP]),
    "mb_tiff_read_f32": (C.c_int, [C.c_char_p, C.c_int, PF, C.c_int]),
    "mb_tiff_read_f32_dev": (C.c_int, [VP, C.c_char_p, C.c_int, VP, C.c_int, VP, VP]),
    "mb_tiff_write_f32": (C.c_int, [C.c_char_p, PG, PF, C.c_int, C.c_int, C.c_int]),
    "mb_tiff_write_f64": (C.c_int, [C.c_char_p, PG, PD, C.c_int, C.c_int, C.c_int]),
}

_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m machisplin_b200.build` "
            "(__graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.DEFAULT_MODE)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        raise MbError(code, load().mb_last_error().decode("utf-8", "replace"))
