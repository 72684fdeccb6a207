#!/usr/bin/env python
"""A/B of the grid-evaluation kernel (k_leaf_fused / k_leaf) on an 8192 x 8192 raster with a 5 000-knot spline (fixed lambda: the
fit is not what is measured here) and a gam-only ensemble as the accumulator's producer.

    python tools/leaf_check.py [--param name=v1,v2 ...] [--reps 20]
prints per variant the device time of the kernel, GB/s at 16 B / cell (fused) or 8 B / cell (TPS only) and the fraction of the
measured HBM peak; the outputs of all variants are compared bit for bit.
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402  (device buffers only)
import machisplin_b200 as mb                                     # noqa: E402
from machisplin_b200 import synth                                # noqa: E402
import bench                                                     # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--param", action="append", default=[], help="name=v1,v2,... (mb_set_param); the product of all lists is run")
ap.add_argument("--nrow", type=int, default=8192)
ap.add_argument("--ncol", type=int, default=8192)
ap.add_argument("--knots", type=int, default=5000)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0)
eng = mb.Engine(0)
peak = bench.measured_peak()[0]
geom = synth.make_geom(args.nrow, args.ncol)
xy, krow, kcol = synth.make_knots(geom, args.knots, 1237)
resid = synth.residual_field(xy, 1237)
sp = eng.tps_fit(xy, resid, lam=5e-3)
C = 3
models = synth.make_models(geom, C, 500, 5, kept="g")
kept, w, wt = synth.ensemble_weights("g")
ens = eng.ensemble_create(geom, models, kept, w, wt, C + 2)
cov = bench.device_covariates(geom, C, dev)
out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device=dev)
names = [kv.split("=")[0] for kv in args.param]
lists = [[int(v) for v in kv.split("=")[1].split(",")] for kv in args.param]
ref = {}
cells = geom.nrow * geom.ncol
for combo in itertools.product(*lists) if lists else [()]:
    for n, v in zip(names, combo):
        eng.set_param(n, v)
    for mode, kname, bpc in (("fused", "k_leaf_fused", 16), ("tps", "k_leaf", 8)):
        def run():
            if mode == "fused":
                eng.ensemble_eval_dev(ens, cov.data_ptr(), C, out.data_ptr(), spline=sp)
            else:
                eng.tps_eval_dev(sp, geom, out.data_ptr(), geom.ncol)
        run(); run(); run()
        torch.cuda.synchronize()
        eng.timing(True); eng.timing_collect()
        for _ in range(args.reps):
            run()
        torch.cuda.synchronize()
        kt = eng.timing_collect(); eng.timing(False)
        ms = kt[kname][0] / kt[kname][1]
        chk = out.clone()
        same = None
        if mode in ref:
            same = bool(torch.equal(torch.nan_to_num(chk, nan=-7.0), torch.nan_to_num(ref[mode], nan=-7.0)))
        else:
            ref[mode] = chk
        gbs = cells * bpc / (ms * 1e-3) / 1e9
        print(json.dumps({"params": dict(zip(names, combo)), "kernel": kname, "ms": round(ms, 4), "GBs": round(gbs, 1),
                          "frac_of_measured_hbm": round(gbs / peak, 4), "identical_to_first_variant": same,
                          "other": {k: round(v[0] / v[1], 4) for k, v in kt.items() if k != kname and v[0] / args.reps > 0.02}}), flush=True)
for n in names:
    eng.set_param(n, 0)
