"""Builds r/mb_shim.c against the stub R API of tests/r_stub/ and drives it through ctypes the way R's .Call would -
TEST INFRASTRUCTURE (the build image has no R)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")
OUT = os.path.join(STUB, "build", "mb_shim_stub.so")


def build() -> str:
    src = [os.path.join(ROOT, "r", "mb_shim.c"), os.path.join(STUB, "r_stub.c")]
    deps = src + [os.path.join(STUB, "Rinternals.h"), os.path.join(ROOT, "include", "machisplin_b200.h"),
                  os.path.join(ROOT, "machisplin_b200", "libmachisplin_b200.so")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.join(ROOT, "machisplin_b200")
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I" + STUB, "-I" + os.path.join(ROOT, "include"),
                    *src, "-L" + libdir, "-lmachisplin_b200", "-Wl,-rpath," + libdir, "-o", OUT], check=True)
    return OUT


class RStub:
    """A pretend R session: make SEXPs, .Call() wrappers of the shim, read results back as numpy."""

    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        S = C.c_void_p
        for name, res, args in [
            ("stub_nil", S, []), ("stub_real", S, [C.POINTER(C.c_double), C.c_ssize_t, C.c_int, C.c_int]),
            ("stub_int", S, [C.POINTER(C.c_int), C.c_ssize_t]), ("stub_raw", S, [C.c_void_p, C.c_ssize_t]),
            ("stub_string", S, [C.c_char_p]), ("stub_strings", S, [C.POINTER(C.c_char_p), C.c_int]), ("stub_list", S, [C.c_int]),
            ("stub_list_set", None, [S, C.c_int, C.c_char_p, S]), ("stub_type", C.c_int, [S]),
            ("stub_len", C.c_ssize_t, [S]), ("stub_data", C.c_void_p, [S]), ("stub_nrow", C.c_int, [S]),
            ("stub_ncol", C.c_int, [S]), ("stub_prot", S, [S]), ("stub_last_error", C.c_char_p, []),
            ("stub_protect_depth", C.c_int, []), ("stub_finalize", None, [S]),
            ("stub_call", S, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
        ]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.nil = L.stub_nil()

    # ---- R values ----------------------------------------------------------------------------------------------------
    def real(self, v, matrix=False):
        a = np.asfortranarray(np.asarray(v, dtype=np.float64))
        nrow, ncol = (a.shape if (matrix and a.ndim == 2) else (0, 0))
        flat = np.ascontiguousarray(a.ravel(order="F"))
        return self.lib.stub_real(flat.ctypes.data_as(C.POINTER(C.c_double)), flat.size, nrow, ncol)

    def integer(self, v):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.int32)))
        return self.lib.stub_int(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)

    def raw(self, b: bytes):
        return self.lib.stub_raw(C.c_char_p(b), len(b))

    def string(self, s: str):
        return self.lib.stub_string(s.encode())

    def strings(self, v):
        arr = (C.c_char_p * len(v))(*[x.encode() for x in v])
        return self.lib.stub_strings(arr, len(v))

    def list_elt(self, lst, i: int):
        return C.cast(self.lib.stub_data(lst), C.POINTER(C.c_void_p))[i]

    def named_list(self, d: dict):
        lst = self.lib.stub_list(len(d))
        for i, (k, v) in enumerate(d.items()):
            self.lib.stub_list_set(lst, i, k.encode(), v)
        return lst

    # ---- .Call -------------------------------------------------------------------------------------------------------
    def call(self, name: str, *args):
        fn = C.cast(getattr(self.lib, name), C.c_void_p)
        arr = (C.c_void_p * len(args))(*args)
        r = self.lib.stub_call(fn, len(args), arr)
        if not r:
            raise RuntimeError("R error: " + self.lib.stub_last_error().decode())
        assert self.lib.stub_protect_depth() == 0, "PROTECT / UNPROTECT imbalance in " + name
        return r

    def as_numpy(self, s):
        t, n = self.lib.stub_type(s), self.lib.stub_len(s)
        ct = {14: C.c_double, 13: C.c_int, 24: C.c_ubyte}[t]
        a = np.ctypeslib.as_array(C.cast(self.lib.stub_data(s), C.POINTER(ct)), shape=(n,)).copy()
        nr, nc = self.lib.stub_nrow(s), self.lib.stub_ncol(s)
        return a.reshape((nr, nc), order="F") if nr else a

    def finalize(self, s):
        self.lib.stub_finalize(s)


def models_to_r(rs: RStub, models: dict, P: int):
    """What mb_export_models() of r/machisplin_b200.R returns, built from the flat descriptors of synth.make_models."""
    d = {"P": rs.real([P])}
    if "g" in models:
        d["gam_coef"] = rs.real(models["g"]["coef"])
    if "n" in models:
        m = models["n"]
        d.update(nn_wts=rs.real(m["wts"]), nn_H=rs.integer(m["H"]), nn_max2=rs.real([m["max2"]]), nn_min=rs.real([m["min"]]))
    if "m" in models:
        m = models["m"]
        d.update(mars_T=rs.integer(len(m["coef"])), mars_dirs=rs.raw(np.ascontiguousarray(m["dirs"], dtype=np.int8).tobytes()),
                 mars_cuts=rs.real(np.ascontiguousarray(m["cuts"]).ravel()), mars_coef=rs.real(m["coef"]))
    if "v" in models:
        m = models["v"]
        d.update(svm_S=rs.integer(len(m["alpha"])), svm_sv=rs.real(np.ascontiguousarray(m["sv"]).ravel()), svm_alpha=rs.real(m["alpha"]),
                 svm_b=rs.real([m["b"]]), svm_sigma=rs.real([m["sigma"]]), svm_x_center=rs.real(m["x_center"]),
                 svm_x_scale=rs.real(m["x_scale"]), svm_y_center=rs.real([m["y_center"]]), svm_y_scale=rs.real([m["y_scale"]]))
    return rs.named_list(d)
