#!/usr/bin/env python
"""GPU check of the two-stage tridiagonalisation (sytrd_mode 3, csrc/sbr.cu) against the one-stage kernel (sytrd_mode 1):
eigenvalues of the stage-1 band matrix, eigenvalues of the final tridiagonal, the transformed right-hand side (through
RSS(lambda)), the selected lambda, and per-kernel times.  Usage: sbr_check.py [n ...]"""
import sys, time, traceback
import numpy as np
from scipy.linalg import eig_banded, eigvalsh_tridiagonal, solveh_banded
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import machisplin_b200 as mb
from machisplin_b200 import synth


def rss(d, e, zh, lam):
    ab = np.zeros((2, len(d)))
    ab[0] = d + lam
    ab[1, :-1] = e
    x = solveh_banded(ab, zh, lower=True)
    return lam * lam * float(x @ x)


sizes = [int(a) for a in sys.argv[1:]] or [20, 36, 70, 200, 600, 1100, 5000]
eng = mb.Engine(0)
geom = synth.make_geom(8192, 8192)
for n in sizes:
    try:
        xy, _, _ = synth.make_knots(geom, n, 300 + n)
        y = synth.residual_field(xy, 300 + n)
        eng.set_param("sytrd_mode", 1)
        sp0 = eng.tps_fit(xy, y)
        eta0, (d0, e0, z0) = sp0.decomposition()
        eng.set_param("sytrd_mode", 3)
        eng.set_param("sbr_debug", 1)
        sp3 = eng.tps_fit(xy, y)
        m = n - 3
        band = eng.debug_values("sbr_band", 64 * m).reshape(m, 64).T      # [off, j]
        eng.set_param("sbr_debug", 0)
        eta3, (d3, e3, z3) = sp3.decomposition()
        scale = np.abs(eta0).max()
        msg = f"n={n} m={m}"
        if band.size:
            kb = min(32, m - 1)
            evb = eig_banded(band[:kb + 1], lower=True, eigvals_only=True)[::-1]
            msg += f" | stage1: band-eig err {np.abs(evb - eta0).max() / scale:.2e} bulge rows max {np.abs(band[33:]).max():.1e}"
        evt = eigvalsh_tridiagonal(d3, e3)[::-1]
        msg += f" | stage2: tri-eig err {np.abs(evt - eta0).max() / scale:.2e}, device eta err {np.abs(eta3 - eta0).max() / scale:.2e}"
        msg += f" | |z| {abs(np.linalg.norm(z3) - np.linalg.norm(z0)) / np.linalg.norm(z0):.1e}"
        lam = sp0.lam
        msg += f" rss rel {abs(rss(d3, e3, z3, lam) - rss(d0, e0, z0, lam)) / rss(d0, e0, z0, lam):.2e}"
        msg += f" | lam {sp3.lam!r} vs {sp0.lam!r} rel {abs(sp3.lam - sp0.lam) / sp0.lam:.2e} | c err {np.abs(sp3.c - sp0.c).max() / np.abs(sp0.c).max():.2e}"
        print(msg, flush=True)
        # three responses at once
        if n in (200, 1100):
            Y = np.stack([y, y[::-1].copy(), y * y], axis=1)
            eng.set_param("sytrd_mode", 1); a = eng.tps_fit(xy, Y)
            eng.set_param("sytrd_mode", 3); b = eng.tps_fit(xy, Y)
            print("   L=3 lambda rel diff", [f"{abs(p.lam - q.lam) / q.lam:.1e}" for p, q in zip(b, a)], flush=True)
        if n in (600, 1100):
            eng.set_param("sytrd_mode", 3); eng.set_param("sbr_qr_grid", 1)
            g = eng.tps_fit(xy, y)
            eng.set_param("sbr_qr_grid", 0)
            print(f"   grid-barrier QR: lambda rel diff {abs(g.lam - sp0.lam) / sp0.lam:.1e}", flush=True)
        try:                                   # the three-warp chase kernel (default: + watcher / publisher warps)
            eng.set_param("sytrd_mode", 3); eng.set_param("sbr_chase_impl", 2)
            eng.tps_fit(xy, y)
            eng.timing(True); eng.timing_collect()
            b = eng.tps_fit(xy, y)
            kt = eng.timing_collect(); eng.timing(False)
            print(f"   sbr_chase_impl 2 (three warps, no watcher): lambda rel diff {abs(b.lam - sp0.lam) / sp0.lam:.1e}, c err {np.abs(b.c - sp0.c).max() / np.abs(sp0.c).max():.2e}, "
                  f"k_sbr_chase {kt.get('k_sbr_chase', (0, 0))[0]:.2f} ms", flush=True)
        except Exception:
            traceback.print_exc()
        finally:
            eng.timing(False); eng.set_param("sbr_chase_impl", 0)
        try:                                   # dense Cholesky instead of the band-form coefficient solve (default)
            eng.set_param("sytrd_mode", 3); eng.set_param("coef_impl", 2)
            eng.tps_fit(xy, y)
            eng.timing(True); eng.timing_collect()
            t0 = time.perf_counter()
            b = eng.tps_fit(xy, y)
            dt = time.perf_counter() - t0
            kt = eng.timing_collect(); eng.timing(False)
            print(f"   coef_impl 2 (dense Cholesky): wall {dt * 1e3:.1f} ms, c err {np.abs(b.c - sp0.c).max() / np.abs(sp0.c).max():.2e}, "
                  f"Cholesky kernels {sum(kt.get(k, (0, 0))[0] for k in ('k_potrf_diag', 'k_trsm_panel', 'k_syrk_dmma', 'k_chol_sweep')):.2f} ms", flush=True)
        except Exception:
            traceback.print_exc()
        finally:
            eng.timing(False); eng.set_param("coef_impl", 0)
        for mode in (1, 3):
            eng.set_param("sytrd_mode", mode)
            eng.timing(True); eng.timing_collect()
            t0 = time.perf_counter()
            eng.tps_fit(xy, y)
            dt = time.perf_counter() - t0
            kt = eng.timing_collect(); eng.timing(False)
            top = sorted(kt.items(), key=lambda kv: -kv[1][0])[:8]
            print(f"   mode {mode} wall {dt * 1e3:.1f} ms:", ", ".join(f"{k} {v[0]:.2f}ms x{v[1]}" for k, v in top), flush=True)
    except Exception:
        traceback.print_exc()
        print(f"n={n} FAILED", flush=True)
eng.close()
