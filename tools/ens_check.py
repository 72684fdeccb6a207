#!/usr/bin/env python
"""Per-kernel device time of the per-cell ensemble kernels (a5) on an 8192 x 8192 raster - the synthetic smooth planes of BASELINE
config 3, and the reference's REAL rasters (alt / slope / TWI tiled up to 8192^2, forests fitted on the bundled points) if a copy
of inst/extdata sits under baseline/_ref/extdata (git-ignored; `tools/copy_extdata.sh` makes it in the build container).

    python tools/ens_check.py [synthetic|real|both] [--levels 1,2] [--svm 1,2] [--nrow N --ncol N]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402  (device buffers only)
import machisplin_b200 as mb                                     # noqa: E402
from machisplin_b200 import geotiff, synth                       # noqa: E402
from machisplin_b200.engine import Geom                          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="?", default="both")
ap.add_argument("--levels", default="1,2")
ap.add_argument("--svm", default="1,2")
ap.add_argument("--fuse", default="1,2", help="ens_overlap values for the full ensemble (1 forest and ksvm kernels side by side, 2 one after the other)")
ap.add_argument("--per-sm", default="0", help="svm_ctas_per_sm values in overlap mode")
ap.add_argument("--kept", default="rb,v,bgnmrv")
ap.add_argument("--nrow", type=int, default=8192)
ap.add_argument("--ncol", type=int, default=8192)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
eng = mb.Engine(0)


def run(tag, geom, cov_t, models, kept_sets):
    C = cov_t.shape[0]
    out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device=dev)
    for kept in kept_sets:
        kk, w, wt = synth.ensemble_weights(kept)
        full = bool(set(kept) & set("rb")) and "v" in kept
        for fz, psm in ([(int(x), int(q)) for x in args.fuse.split(",") for q in (args.per_sm.split(",") if int(x) == 1 else ["0"])] if full else [(0, 0)]):
          for lv in [int(x) for x in args.levels.split(",")] if (set(kept) & set("rb") and not full) else [0]:
           for sv in [int(x) for x in args.svm.split(",")] if "v" in kept else [0]:
            eng.set_param("tree_levels", lv)
            eng.set_param("svm_impl", sv)
            eng.set_param("ens_overlap", fz); eng.set_param("svm_ctas_per_sm", psm)
            ens = eng.ensemble_create(geom, models, kk, w, wt, C + 2)
            eng.ensemble_eval_dev(ens, cov_t.data_ptr(), C, out.data_ptr())
            torch.cuda.synchronize()
            eng.timing(True); eng.timing_collect()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                eng.ensemble_eval_dev(ens, cov_t.data_ptr(), C, out.data_ptr())
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.reps
            kt = eng.timing_collect(); eng.timing(False)
            ks = ", ".join(f"{k} {v[0] / args.reps:.2f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0]) if v[0] / args.reps > 0.05)
            nanf = float(torch.isnan(out).float().mean())
            if "v" in kept:
                ref_out = globals().setdefault("_ref_out", {})
                key = (tag, kept)
                if key not in ref_out:
                    ref_out[key] = out[::7, ::5].clone()
                else:
                    d = (out[::7, ::5] - ref_out[key]).abs()
                    print(f"   svm_impl={sv}: max |diff to first variant| / max|ref| = {float(d[~torch.isnan(d)].max() / ref_out[key][~torch.isnan(ref_out[key])].abs().max()):.3e}")
            print(f"{tag} kept={kept:7s} tree_levels={lv} svm_impl={sv} ens_overlap={fz} svm_ctas_per_sm={psm}: wall {dt * 1e3:7.2f} ms | {ks} | NA {nanf:.4f}", flush=True)
            ens.free()
    eng.set_param("tree_levels", 0); eng.set_param("svm_impl", 0); eng.set_param("ens_overlap", 0); eng.set_param("svm_ctas_per_sm", 0)


if args.what in ("synthetic", "both"):
    sys.path.insert(0, ROOT)
    import bench
    cfg = dict(synth.CONFIGS["c3"]); cfg["nrow"], cfg["ncol"] = args.nrow, args.ncol
    geom, xy, krow, kcol, resid, models, kept, w, wt = bench.build_inputs(cfg, 0)
    cov = bench.device_covariates(geom, cfg["C"], dev)
    run("synthetic", geom, cov, models, args.kept.split(","))
    del cov

ext = os.path.join(ROOT, "baseline", "_ref", "extdata")
if args.what in ("real", "both"):
    if not os.path.isdir(ext):
        print("real rasters: baseline/_ref/extdata not present, skipped")
    else:
        planes = []
        for name in ("alt", "slope", "TWI"):
            g, a = geotiff.read_raster(os.path.join(ext, name + ".tif"))
            planes.append(a)
        small = np.stack(planes)                                 # 3 x 2476 x 3264
        g0 = Geom(g.xmin, g.xmax, g.ymin, g.ymax, g.nrow, g.ncol)
        z = np.load(os.path.join(ROOT, "tests", "golden", "bundled_c1.npz"))
        pts = z["points"]
        krow, kcol = z["krow"], z["kcol"]
        X = np.column_stack([small[:, krow, kcol].T.astype(np.float64), z["knots_xy"]])
        ok = ~np.isnan(X).any(axis=1)
        models = synth.make_models(g0, 3, 0, 5, kept="bgnmrv", table=(X[ok], pts[ok, 2]))
        # tile the real raster up to nrow x ncol cells (mirror tiling keeps the fields continuous); same cell size, larger extent
        ry = -(-args.nrow // g.nrow); rx = -(-args.ncol // g.ncol)
        rows = np.concatenate([np.arange(g.nrow) if k % 2 == 0 else np.arange(g.nrow)[::-1] for k in range(ry)])[:args.nrow]
        cols = np.concatenate([np.arange(g.ncol) if k % 2 == 0 else np.arange(g.ncol)[::-1] for k in range(rx)])[:args.ncol]
        big = torch.from_numpy(small).to(dev)[:, torch.from_numpy(rows.copy()).to(dev)][:, :, torch.from_numpy(cols.copy()).to(dev)].contiguous()
        geom = Geom(g.xmin, g.xmin + args.ncol * g0.rx, g.ymax - args.nrow * g0.ry, g.ymax, args.nrow, args.ncol)
        run("real     ", geom, big, models, args.kept.split(","))
eng.close()
