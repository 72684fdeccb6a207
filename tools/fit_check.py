#!/usr/bin/env python
"""GPU check of the GCV fit paths: two-stage default (mode `default`), one-stage persistent tridiagonalisation (sytrd_mode 1), one kernel per phase
(sytrd_mode 2), cuSOLVER validation path (eigen_impl 1).  Usage: fit_check.py <mode: default|persistent|phases|cusolver|chol> [n ...]"""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import machisplin_b200 as mb
from machisplin_b200 import synth

mode = sys.argv[1]
sizes = [int(a) for a in sys.argv[2:]] or [35, 200, 1100, 5000]
eng = mb.Engine(0)
if mode.startswith("persistent"):
    eng.set_param("sytrd_mode", 1)
if mode == "phases":
    eng.set_param("sytrd_mode", 2)
elif mode == "cusolver":
    eng.set_param("eigen_impl", 1)
elif mode == "persistent1":
    eng.set_param("sytrd_ctas_per_sm", 1)
geom = synth.make_geom(8192, 8192)
for n in sizes:
    xy, _, _ = synth.make_knots(geom, n, 300 + n)
    y = synth.residual_field(xy, 300 + n)
    lam = 1e-3 if mode == "chol" else None
    sp = eng.tps_fit(xy, y, lam=lam)          # warm-up (allocations, attributes)
    eng.timing(True); eng.timing_collect()
    t0 = time.perf_counter()
    sp = eng.tps_fit(xy, y, lam=lam)
    dt = time.perf_counter() - t0
    kt = eng.timing_collect(); eng.timing(False)
    top = sorted(kt.items(), key=lambda kv: -kv[1][0])[:6]
    eta = sp.decomposition()[0] if lam is None else np.zeros(1)
    f = eng.tps_predict_points(sp, xy)
    resid = np.max(np.abs(f - (y - sp.lam * sp.c))) / max(1.0, np.abs(y).max())
    print(f"{mode} n={n} wall={dt*1e3:.1f} ms lam={sp.lam!r} edf={sp.eff_df:.6f} eta[0]={eta[0]!r} eta[-1]={eta[-1]!r} "
          f"knot-identity={resid:.2e} sum|c|={np.abs(sp.c).sum():.6e}", flush=True)
    if mode.startswith("persistent"):
        print("   k_sytrd phases [P1 bar P2 bar P3 bar upd bar] ms:", np.round(eng.debug_values("sytrd_phase_ms"), 2), flush=True)
    print("   kernels:", ", ".join(f"{k} {v[0]:.2f}ms x{v[1]}" for k, v in top), flush=True)
eng.close()
