#!/usr/bin/env python
"""bench.py - Mcells/s interpolated (TPS + ensemble), BASELINE.json metric, on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3|c2]

A "step" is one pass of the hot path over one synthetic raster: fields::Tps fit of the residuals
(GCV), per-cell six-model ensemble prediction + TPS surface + NA-propagating sum (mltps parts 2-5,
V73:442-932), and the point extraction of part 5.  Default workload = BASELINE config 3 (the one the
north-star target is quoted on): 8192 x 8192 cells, 5 000 knots, 6 covariates, 6 models + TPS, global
spline.  N > 1 is weak scaling over machisplin.tiles-style partitions (one 8192^2 tile per GPU).

  value  device-resident inputs, CUDA events, max over ranks
  e2e    same step through the host-buffer path: pinned host covariates -> device, result -> host
  roofline / kernels   per-kernel device time from CUDA events recorded by the library on the
         launching stream inside the timed region (mb_timing_*), achieved = algorithmic bytes / time
  cpu_baseline   the oracle's C restatement (oracle/c) on a bounded row sample, all host threads

--impl reference times that CPU restatement alone (the reference itself is R + CRAN packages and
cannot run in this image; oracle/ is the stand-in).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcells/s interpolated (TPS+ensemble)"


# ------------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--nrow", type=int, default=0, help="override the grid size (debug)")
    ap.add_argument("--ncol", type=int, default=0)
    ap.add_argument("--knots", type=int, default=0)
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="rows of the CPU sample (0 = auto, ~10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--lam", type=float, default=None, help="fixed lambda (Cholesky path) instead of GCV")
    ap.add_argument("--tree-rows", type=int, default=0, help="forest tile: cells per thread (1, 2, 4; 0 = auto)")
    ap.add_argument("--eval-precision", type=int, default=0, help="leaf kernel path: 0 auto, 1 float64, 2 mixed")
    ap.add_argument("--param", action="append", default=[], help="engine tunable name=value (mb_set_param), repeatable")
    return ap.parse_args()


def workload(args):
    from machisplin_b200 import synth
    cfg = dict(synth.CONFIGS[args.config])
    if args.nrow:
        cfg["nrow"] = args.nrow
    if args.ncol:
        cfg["ncol"] = args.ncol
    if args.knots:
        cfg["knots"] = args.knots
    return cfg


def build_inputs(cfg, rank):
    """Host-side synthetic inputs of one rank's tile."""
    from machisplin_b200 import synth
    geom = synth.make_geom(cfg["nrow"], cfg["ncol"])
    seed = cfg["seed"] + 101 * rank
    xy, krow, kcol = synth.make_knots(geom, cfg["knots"], seed)
    resid = synth.residual_field(xy, seed)
    kept = cfg["kept"]
    models, w, wt = {}, np.zeros(0), 1.0
    if kept:
        cache = f"/tmp/mb_models_{cfg['nrow']}x{cfg['ncol']}_{cfg['knots']}_{cfg['C']}_{kept}_{seed}.npz"
        if os.path.exists(cache):
            z = np.load(cache, allow_pickle=True)
            models = z["models"].item()
        else:
            models = synth.make_models(geom, cfg["C"], cfg["knots"], seed, kept=kept)
            try:
                np.savez(cache, models=np.array(models, dtype=object))
            except Exception:
                pass
        kept, w, wt = synth.ensemble_weights(kept)
    return geom, xy, krow, kcol, resid, models, kept, w, wt


def device_covariates(geom, C, device, seed=99, disc_geom=None):
    """Same generator as synth.covariate_planes, evaluated on the GPU with torch (plumbing only).  disc_geom: the
    grid whose extent places the NaN discs (the full raster when geom is one tile of it)."""
    import torch
    dg = disc_geom or geom
    x = torch.tensor(geom.xmin, dtype=torch.float64, device=device) + \
        (torch.arange(geom.ncol, device=device, dtype=torch.float64) + 0.5) * geom.rx
    y = torch.tensor(geom.ymax, dtype=torch.float64, device=device) - \
        (torch.arange(geom.nrow, device=device, dtype=torch.float64) + 0.5) * geom.ry
    out = torch.empty((C, geom.nrow, geom.ncol), dtype=torch.float32, device=device)
    blk = 1024
    for k in range(C):
        rng = np.random.default_rng(seed + k)
        pars = [(rng.uniform(2, 12, 2), rng.uniform(0, 2 * np.pi), rng.uniform(0.5, 2.0)) for _ in range(3)]
        for r0 in range(0, geom.nrow, blk):
            r1 = min(geom.nrow, r0 + blk)
            acc = torch.zeros((r1 - r0, geom.ncol), dtype=torch.float64, device=device)
            for (f, ph, amp) in pars:
                acc += amp * torch.sin(f[0] * x[None, :] + f[1] * y[r0:r1, None] + ph)
            out[k, r0:r1] = (100.0 * (k + 1) + 50.0 * acc).to(torch.float32)
    rng = np.random.default_rng(seed + 1000)
    ndisc, nan_frac = 8, 0.02
    rad = np.sqrt(nan_frac * (dg.xmax - dg.xmin) * (dg.ymax - dg.ymin) / (ndisc * np.pi))
    for _ in range(ndisc):
        cx, cy = rng.uniform(dg.xmin, dg.xmax), rng.uniform(dg.ymin, dg.ymax)
        for r0 in range(0, geom.nrow, blk):
            r1 = min(geom.nrow, r0 + blk)
            m = (x[None, :] - cx) ** 2 + (y[r0:r1, None] - cy) ** 2 < rad * rad
            out[0, r0:r1][m] = float("nan")
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = upper half of the observed clocks
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on a bounded row sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(geom, fit_like, models, kept, w, wt, cov_rows, r0, threads=0):
    """TPS surface + ensemble on rows [r0, r0 + S) with the oracle's C loops.  Returns (raster, seconds)."""
    from oracle import cbind
    S = cov_rows.shape[1]
    sub = (geom.xmin, geom.xmax, geom.ymax - (r0 + S) * geom.ry, geom.ymax - r0 * geom.ry, S, geom.ncol)
    t0 = time.perf_counter()
    surf = cbind.tps_eval(fit_like, sub, threads=threads)
    if kept:
        out = cbind.ensemble_eval(models, kept, w, wt, cov_rows, sub, tps=surf, threads=threads)
    else:
        out = surf
    return out, time.perf_counter() - t0


class FitLike:
    def __init__(self, sp):
        self.knots_s = (sp.knots_xy - sp.center) / sp.scale
        self.c, self.d, self.center, self.scale = sp.c, sp.d, sp.center, sp.scale


def run_reference(args):
    """--impl reference: the CPU restatement of the path on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cbind, tps as otps
    cfg = workload(args)
    geom, xy, krow, kcol, resid, models, kept, w, wt = build_inputs(cfg, 0)
    from machisplin_b200 import synth
    threads = cbind.max_threads()
    # fit once with a fixed lambda close to the GCV optimum (a 5k-knot GCV fit alone takes ~10 s of LAPACK)
    nfit = min(cfg["knots"], 1500)
    t0 = time.perf_counter()
    fit = otps.tps_fit(xy[:nfit], resid[:nfit])
    fit_s = time.perf_counter() - t0
    # evaluation sample uses a spline with the configured knot count: pad coefficients by refitting at fixed lambda
    if nfit < cfg["knots"]:
        fit = otps.tps_fit(xy, resid, lam=fit.lam) if cfg["knots"] <= 3000 else _cheap_full_spline(xy, resid, fit)
    S = args.cpu_sample_rows or 4
    r0 = geom.nrow // 2
    cov_rows = synth.covariate_planes(geom, cfg["C"], row0=r0, row1=r0 + S) if cfg["C"] else np.zeros((0, S, geom.ncol), np.float32)
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = cpu_sample(geom, fit, models, kept, w, wt, cov_rows, r0, threads)
        if i >= args.warmup:
            times.append(dt)
        if i == 0 and not args.cpu_sample_rows:   # size the sample to ~10 s per step after the first probe
            S = int(max(1, min(geom.nrow, round(S * 10.0 / max(dt, 1e-3)))))
            cov_rows = synth.covariate_planes(geom, cfg["C"], row0=r0, row1=min(geom.nrow, r0 + S)) if cfg["C"] else \
                np.zeros((0, S, geom.ncol), np.float32)
            S = cov_rows.shape[1]
    cells = S * geom.ncol
    ms = 1e3 * float(np.mean(times))
    value = cells / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args, "global"),
        "cpu_baseline": {"value": value, "unit": "Mcells/s", "cores": threads, "kind": "port",
                         "sample": f"{S} full-width rows ({cells} cells) per step: TPS surface ({cfg['knots']} knots, "
                                   f"float64 pair loop) + {len(kept)}-model ensemble; fit excluded (oracle GCV fit of "
                                   f"{nfit} knots took {fit_s:.1f} s)"},
        "e2e": {"value": value, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _cheap_full_spline(xy, resid, small_fit):
    """A spline with the full knot count for CPU *timing* only (the pair loop costs the same whatever
    the coefficients are): coefficients of the small fit tiled over all knots."""
    class F:
        pass
    f = F()
    f.center, f.scale = xy.min(0), xy.max(0) - xy.min(0)
    f.knots_s = (xy - f.center) / f.scale
    reps = int(np.ceil(len(xy) / small_fit.c.size))
    f.c = np.tile(small_fit.c, reps)[:len(xy)]
    f.d = small_fit.d
    return f


def workload_config(cfg, args, tps_mode):
    return {"workload": f"BASELINE config {args.config[1]}: {cfg['nrow']}x{cfg['ncol']} cells, {cfg['knots']} knots, "
                        f"{cfg['C']} covariates, models '{cfg['kept']}' + TPS ({tps_mode} spline, "
                        f"{'GCV' if args.lam is None else 'fixed lambda'}), {cfg['L']} response",
            "grid": [cfg["nrow"], cfg["ncol"]], "knots": cfg["knots"], "covariates": cfg["C"], "models": cfg["kept"],
            "tps_mode": tps_mode, "l2": "inputs_exceed_l2", "parallelism": f"tiles{args.gpus}"}


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import machisplin_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload(args)
    geom, xy, krow, kcol, resid, models, kept, w, wt = build_inputs(cfg, rank)
    eng = mb.Engine(local)
    eng.set_param("tree_rows", args.tree_rows)
    eng.set_param("eval_precision", args.eval_precision)
    for kv in args.param:
        name, val = kv.split("=")
        eng.set_param(name, int(val))
    C = cfg["C"]
    P = C + 2
    cov = device_covariates(geom, C, dev) if C else torch.zeros((0,), device=dev)
    out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device=dev)
    ens = eng.ensemble_create(geom, models, kept, w, wt, P) if kept else None
    # cross-validation residual matrix of this rank's points for the Gram reduction (a6)
    rng = np.random.default_rng(5 + rank)
    Rcv = rng.standard_normal((cfg["knots"], 6 if "b" in kept or not kept else 4))
    stream = torch.cuda.current_stream().cuda_stream
    cells = geom.nrow * geom.ncol
    state = {}

    def step_device():
        G = eng.gram(Rcv)
        if world > 1:
            g = torch.from_numpy(G).to(dev)
            dist.all_reduce(g)
        # parts 2-5 (V73:442-932): ensemble kernels || fields::Tps fit, then the fused per-cell pass
        sp = eng.mltps_predict_dev(geom, ens, cov.data_ptr() if C else 0, C, xy, resid, out.data_ptr(), lam=args.lam,
                                   stream=stream)
        f_actual = eng.gather_cells_dev(out.data_ptr(), geom.ncol, geom.nrow, geom.ncol, krow, kcol, stream=stream)
        state["sp"], state["f_actual"] = sp, f_actual
        return f_actual

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing(True)
    eng.timing_collect()
    l0 = eng.launches
    ms = timed(step_device, args.steps)
    launches = eng.launches - l0
    ktimes = eng.timing_collect()
    eng.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    value = world * cells / (ms * 1e-3) / 1e6

    # ---- e2e: pinned host inputs -> device -> result back to pinned host ---------------------------
    e2e = None
    if not args.no_e2e:
        # the call a user (the R shim) makes: host buffers in, host raster out, through the C ABI
        h_cov = torch.empty(cov.shape, dtype=torch.float32, pin_memory=True)
        h_cov.copy_(cov)
        h_out = torch.empty(out.shape, dtype=torch.float64, pin_memory=True)
        cov_np = h_cov.numpy() if C else None
        out_np = h_out.numpy()

        host_s = {"ensemble_create": 0.0, "mltps_predict": 0.0}

        def step_e2e():
            t0 = time.perf_counter()
            ens2 = eng.ensemble_create(geom, models, kept, w, wt, P) if kept else None   # descriptor upload + tree packing
            t1 = time.perf_counter()
            eng.mltps_predict(geom, ens2, cov_np, xy, resid, lam=args.lam, out=out_np)
            host_s["ensemble_create"] += t1 - t0
            host_s["mltps_predict"] += time.perf_counter() - t1
            return out_np[krow, kcol]

        step_e2e()
        for k in host_s:
            host_s[k] = 0.0
        ms_e2e = timed(step_e2e, max(1, args.steps))
        e2e = {"value": world * cells / (ms_e2e * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h_cov.numel() * 4 + xy.nbytes + resid.nbytes),
               "d2h_bytes_per_step": int(h_out.numel() * 8),
               "host_ms_per_step": {k: v * 1e3 / max(1, args.steps) for k, v in host_s.items()}}
        del h_cov

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the grid-evaluation kernel (north-star kernel) + per-kernel table -----------------
    peak, peak_src = measured_peak()
    kern = {}
    tot_ms = sum(v[0] for v in ktimes.values()) or 1.0
    heavy = bool(kept) and bool(set(kept) & set("brv"))
    # algorithmic HBM bytes per cell: k_leaf writes the float64 surface; k_leaf_fused reads the
    # float64 ensemble accumulator and writes the final raster; the ensemble kernels read the C float32 planes
    # and write (k_ens_trees) or read-modify-write (k_ens_svm, k_ens_smooth) the accumulator
    bytes_per_cell = {"k_leaf": 8.0, "k_leaf_f64": 8.0, "k_leaf_fused": 16.0, "k_leaf_f64_fused": 16.0,
                      "k_ens_final": 24.0, "k_ens_trees": 4.0 * C + 8.0,
                      "k_ens_svm": 4.0 * C + (16.0 if set(kept) & set("br") else 8.0),
                      "k_ens_smooth": 4.0 * C + (16.0 if heavy else 8.0)}
    for name, (tms, cnt) in sorted(ktimes.items(), key=lambda kv: -kv[1][0]):
        per_launch = tms / max(cnt, 1)
        ent = {"ms_per_step": tms / args.steps, "launches_per_step": cnt / args.steps, "share": tms / tot_ms}
        twin = {"k_leaf": "k_leaf_f64", "k_leaf_f64": "k_leaf", "k_leaf_fused": "k_leaf_f64_fused",
                "k_leaf_f64_fused": "k_leaf_fused"}.get(name)
        if twin in ktimes and ktimes[twin][0] > tms:
            ent["note"] = "twin of the selected leaf kernel: exits at its first instruction (DESIGN.md 4.1)"
        elif name in bytes_per_cell:
            gbs = cells * bytes_per_cell[name] / (per_launch * 1e-3) / 1e9
            ent.update({"algorithmic_bytes_per_cell": bytes_per_cell[name], "achieved_gbs": gbs, "hbm_frac": gbs / peak})
        kern[name] = ent
    leaf_names = ("k_leaf_fused", "k_leaf", "k_leaf_f64_fused", "k_leaf_f64")
    lname = next((k for k in leaf_names if "hbm_frac" in kern.get(k, {})), next((k for k in leaf_names if k in kern), "k_leaf"))
    leaf = kern.get(lname, {})
    # DRAM traffic of the kernel per launch from the committed ncu --set full capture of this workload (profiles/)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        key = f"{lname}@{cfg['nrow']}x{cfg['ncol']}"
        traffic = tr.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": f"{lname} (grid-evaluation kernel: per-cell TPS surface" +
                          (" + ensemble combine, mltps part 5)" if "fused" in lname else ")"),
                "bound": "hbm", "achieved": leaf.get("achieved_gbs"),
                "peak": peak, "unit": "GB/s", "frac": leaf.get("hbm_frac"), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_cell": bytes_per_cell[lname],
                "note": "roofline of the north-star kernel; `kernels` lists every kernel of the step with its share"}

    # ---- parity + CPU baseline on a bounded row sample ---------------------------------------------------
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import cbind
        sp = state["sp"]
        S = args.cpu_sample_rows or 2
        r0 = geom.nrow // 2
        fl = FitLike(sp)
        cov_rows = cov[:, r0:r0 + S].cpu().numpy() if C else np.zeros((0, S, geom.ncol), np.float32)
        ref, dt = cpu_sample(geom, fl, models, kept, w, wt, cov_rows, r0)
        if not args.cpu_sample_rows and dt < 8.0:       # grow the sample to ~10-15 s of CPU work
            S2 = int(max(S, min(geom.nrow - r0, round(S * 12.0 / max(dt, 1e-3)))))
            if S2 > S:
                S = S2
                cov_rows = cov[:, r0:r0 + S].cpu().numpy() if C else np.zeros((0, S, geom.ncol), np.float32)
                ref, dt = cpu_sample(geom, fl, models, kept, w, wt, cov_rows, r0)
        got = out[r0:r0 + S].cpu().numpy()
        m = ~np.isnan(ref)
        parity = {"rows": [r0, r0 + S], "max_abs_err": float(np.max(np.abs(got[m] - ref[m]))),
                  "max_rel_err": float(np.max(np.abs(got[m] - ref[m])) / np.max(np.abs(ref[m]))),
                  "na_mask_equal": bool(np.array_equal(np.isnan(got), np.isnan(ref))), "tolerance": 1e-5}
        cpu = {"value": S * geom.ncol / dt / 1e6, "unit": "Mcells/s", "cores": cbind.max_threads(), "kind": "port",
               "sample": f"{S} full-width rows ({S * geom.ncol} cells), TPS surface + ensemble with the GPU-fitted "
                         f"coefficients, {dt:.1f} s; fit excluded"}

    line = {
        "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, args, "global"),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kern,
        "cpu_baseline": cpu, "parity": parity,
        "fit": {"lambda": state["sp"].lam, "eff_df": state["sp"].eff_df, "knots": state["sp"].np},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# config 4: one raster tiled across the GPUs with machisplin.tiles.* (strong scaling)
# ------------------------------------------------------------------------------------------------
def run_tiled(args):
    """BASELINE config 4: machisplin.tiles.create cuts the raster into world-size tiles with a feather halo
    (V73:1165-1256); every rank runs mltps parts 2-5 on its tile (internal 1500-px tiling, V73:649-895); the
    tiles travel to rank 0 over NCCL and machisplin.tiles.merge blends the seams there (V73:1392-1548); the
    Gram of the cross-validation residuals is all-reduced (V73:329-333)."""
    import torch
    import torch.distributed as dist
    import machisplin_b200 as mb
    from machisplin_b200 import synth, tiles as mtiles, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload(args)
    geom = synth.make_geom(cfg["nrow"], cfg["ncol"])
    xy, krow, kcol = synth.make_knots(geom, cfg["knots"], cfg["seed"])
    resid = synth.residual_field(xy, cfg["seed"])
    nC, nR = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
    ts = mtiles.tiles_create(geom, xy, nC, nR, feather_d=50)
    tile = ts.tiles[rank]
    tg, pts = tile.geom, tile.points
    kept, C = cfg["kept"], cfg["C"]
    cache = f"/tmp/mb_models_{cfg['nrow']}x{cfg['ncol']}_{cfg['knots']}_{C}_{kept}_{cfg['seed']}.npz"
    if os.path.exists(cache):
        models = np.load(cache, allow_pickle=True)["models"].item()
    else:
        models = synth.make_models(geom, C, min(cfg["knots"], 5000), cfg["seed"], kept=kept)
        if rank == 0:
            try:
                np.savez(cache, models=np.array(models, dtype=object))
            except Exception:
                pass
    kept, w, wt = synth.ensemble_weights(kept)
    eng = mb.Engine(local)
    for kv in args.param:
        name, val = kv.split("=")
        eng.set_param(name, int(val))
    cov = device_covariates(tg, C, dev, disc_geom=geom)
    ens = eng.ensemble_create(tg, models, kept, w, wt, C + 2)
    out_tile = torch.empty((tg.nrow, tg.ncol), dtype=torch.float64, device=dev)
    out_full = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device=dev) if rank == 0 else None
    shapes = [(t.geom.nrow, t.geom.ncol) for t in ts.tiles]
    wins = [t.win for t in ts.tiles]
    Rcv = np.random.default_rng(5).standard_normal((cfg["knots"], 6))[parallel.shard_rows(cfg["knots"], world, rank)]
    stream = torch.cuda.current_stream().cuda_stream
    tile_px = 1500

    def step():
        G = eng.gram(Rcv)
        if world > 1:
            g = torch.from_numpy(G).to(dev)
            dist.all_reduce(g)
        eng.mltps_predict_dev(tg, ens, cov.data_ptr(), C, xy[pts], resid[pts], out_tile.data_ptr(), lam=args.lam,
                              tile_px=tile_px, stream=stream)
        tl = parallel.gather_tiles_device(out_tile, shapes, dst=0)
        if rank == 0:
            eng.tiles_merge_dev(geom, wins, [t.data_ptr() for t in tl], nC, nR, out_full.data_ptr(), stream=stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing(True)
    eng.timing_collect()
    l0 = eng.launches
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = eng.launches - l0
    ktimes = eng.timing_collect()
    eng.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        cells = geom.nrow * geom.ncol
        tot = sum(v[0] for v in ktimes.values()) or 1.0
        kern = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps, "share": v[0] / tot}
                for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:12]}
        conf = workload_config(cfg, args, f"mltps tiling {tile_px} px inside {nC}x{nR} machisplin.tiles")
        conf["parallelism"] = f"tiles{nC}x{nR}"
        nan_frac = float(torch.isnan(out_full).float().mean().item())
        line = {"metric": METRIC, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
                "e2e": None, "gpu_launches": int(launches), "clocks": clocks, "kernels_rank0": kern,
                "roofline": None, "cpu_baseline": None, "na_fraction": nan_frac,
                "collectives": {"gram": "all_reduce 36 doubles", "tile_gather_bytes": int(sum(a * b for a, b in shapes[1:]) * 8)}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# config 5: 20 response layers sharing the knots (smooth.outputs.only = TRUE: gam, nnet, earth, ksvm)
# ------------------------------------------------------------------------------------------------
def run_batch(args):
    """BASELINE config 5: the NA filter runs on the joined table (V73:154), so all L response layers share the knots:
    ONE tridiagonalisation serves the L GCV searches (mb_tps_fit with L responses), then every layer gets its own
    ensemble + TPS raster (V73:203 loop).  Multi-GPU: the layers are dealt round-robin to the ranks (the deleted
    snowfall path of the reference parallelised exactly this loop, old/...V69.R:937-968)."""
    import torch
    import torch.distributed as dist
    import machisplin_b200 as mb
    from machisplin_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload(args)
    L = cfg["L"]
    geom, xy, krow, kcol, _, models, kept, w, wt = build_inputs(cfg, 0)
    Y = synth.residual_field(xy, cfg["seed"], L=L)
    mine = [r for r in range(L) if r % world == rank]
    eng = mb.Engine(local)
    for kv in args.param:
        name, val = kv.split("=")
        eng.set_param(name, int(val))
    C = cfg["C"]
    cov = device_covariates(geom, C, dev)
    out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device=dev)
    ens = eng.ensemble_create(geom, models, kept, w, wt, C + 2)     # same descriptors for every layer (synthetic)
    stream = torch.cuda.current_stream().cuda_stream
    state = {}

    def step():
        sps = eng.tps_fit(xy, Y[:, mine]) if len(mine) > 1 else [eng.tps_fit(xy, Y[:, mine[0]])]
        for sp in sps:
            eng.ensemble_eval_dev(ens, cov.data_ptr(), C, out.data_ptr(), spline=sp, stream=stream)
            state["f"] = eng.gather_cells_dev(out.data_ptr(), geom.ncol, geom.nrow, geom.ncol, krow, kcol, stream=stream)
        state["lam"] = [sp.lam for sp in sps]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing(True)
    eng.timing_collect()
    l0 = eng.launches
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = eng.launches - l0
    ktimes = eng.timing_collect()
    eng.timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        cells = geom.nrow * geom.ncol * L
        tot = sum(v[0] for v in ktimes.values()) or 1.0
        kern = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps, "share": v[0] / tot}
                for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:12]}
        conf = workload_config(cfg, args, "global")
        conf["parallelism"] = f"layers{world}"
        line = {"metric": METRIC, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
                "e2e": None, "gpu_launches": int(launches), "clocks": clocks, "kernels_rank0": kern,
                "roofline": None, "cpu_baseline": None, "layers": L, "lambda_rank0": state["lam"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c4":
        run_tiled(args)
    elif args.config == "c5":
        run_batch(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
