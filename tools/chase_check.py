#!/usr/bin/env python
"""Bulge-chase variants against each other: bit-identity of lambda / c / eigenvalues and kernel time.

    python tools/chase_check.py [--sizes 36,70,200,1100,5000] [--impls 2,1,3] [--reps 3]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import machisplin_b200 as mb                                     # noqa: E402
from machisplin_b200 import synth                                # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="36,70,200,1100,5000")
ap.add_argument("--impls", default="2,1,3")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
eng = mb.Engine(0)
geom = synth.make_geom(2048, 2048)
names = {1: "k_sbr_chase_dec", 2: "k_sbr_chase", 3: "k_sbr_chase_ll"}
for n in [int(v) for v in args.sizes.split(",")]:
    xy, _, _ = synth.make_knots(geom, n, 800 + n)
    y = synth.residual_field(xy, 800 + n)
    Y = np.stack([y, y[::-1].copy(), y * y], axis=1)
    ref = None
    for impl in [int(v) for v in args.impls.split(",")]:
        eng.set_param("sbr_chase_impl", impl)
        try:
            eng.tps_fit(xy, y)
            eng.timing(True); eng.timing_collect()
            t0 = time.perf_counter()
            fits = [eng.tps_fit(xy, y) for _ in range(args.reps)]
            dt = (time.perf_counter() - t0) / args.reps
            kt = eng.timing_collect(); eng.timing(False)
            f3 = eng.tps_fit(xy, Y)
            sig = (fits[0].lam, fits[0].c.copy(), fits[0].decomposition()[0].copy(), [f.lam for f in f3], np.concatenate([f.c for f in f3]))
            same_runs = all(f.lam == fits[0].lam and np.array_equal(f.c, fits[0].c) for f in fits)
            if ref is None:
                ref = sig
                cmp = "reference"
            else:
                cmp = "identical" if (sig[0] == ref[0] and np.array_equal(sig[1], ref[1]) and np.array_equal(sig[2], ref[2]) and
                                      sig[3] == ref[3] and np.array_equal(sig[4], ref[4])) else \
                    f"DIFFERENT: lam {sig[0]!r} vs {ref[0]!r}, c err {np.abs(sig[1] - ref[1]).max() / np.abs(ref[1]).max():.2e}, L=3 lam {sig[3]} vs {ref[3]}"
            k = names[impl]
            print(f"n={n:5d} impl {impl}: fit wall {dt * 1e3:7.2f} ms, {k} {kt.get(k, (0, 0))[0] / args.reps:7.3f} ms, repeated runs identical {same_runs}, vs impl "
                  f"{args.impls.split(',')[0]}: {cmp}", flush=True)
        except Exception as ex:
            eng.timing(False)
            print(f"n={n} impl {impl} FAILED: {str(ex)[:300]}", flush=True)
            eng = mb.Engine(0)
    eng.set_param("sbr_chase_impl", 0)
