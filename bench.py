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
    ap.add_argument("--no-tiled", action="store_true", help="skip the secondary mltps-tiled measurement (N = 1)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one raster of N x nrow rows (a config-sized block per GPU); strong = the nrow x ncol raster cut in N")
    ap.add_argument("--lam", type=float, default=None, help="fixed lambda (Cholesky path) instead of GCV")
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


def build_inputs(cfg, rank, model_cfg=None):
    """Host-side synthetic inputs of one rank's tile.  model_cfg: the config whose geometry the synthetic model
    descriptors are fitted on (weak scaling keeps the descriptors of the base config - the number of support vectors of the
    scikit-learn SVR they are exported from depends on the training table - so that the work per cell is the same at every N)."""
    from machisplin_b200 import synth
    geom = synth.make_geom(cfg["nrow"], cfg["ncol"], cfg.get("cell", 0.0))
    mcfg = model_cfg or cfg
    mgeom = synth.make_geom(mcfg["nrow"], mcfg["ncol"])
    seed = cfg["seed"] + 101 * rank
    xy, krow, kcol = synth.make_knots(geom, cfg["knots"], seed)
    resid = synth.residual_field(xy, seed)
    kept = cfg["kept"]
    models, w, wt = {}, np.zeros(0), 1.0
    if kept:
        cache = f"/tmp/mb_models_{mcfg['nrow']}x{mcfg['ncol']}_{cfg['knots']}_{cfg['C']}_{kept}_{seed}.npz"
        if os.path.exists(cache):
            z = np.load(cache, allow_pickle=True)
            models = z["models"].item()
        else:
            models = synth.make_models(mgeom, cfg["C"], cfg["knots"], seed, kept=kept)
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
    threads = host_threads()       # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    lapack_threads(threads)
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
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args, "global"),
        "note": "CPU throughput of the whole box (all host cores), independent of --gpus: the per-cell work is the same for every row",
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
def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def lapack_threads(n: int):
    """Give the BLAS / OpenMP pools behind numpy / scipy n threads (OMP_NUM_THREADS=1 from torchrun was read at import)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass


# what bounds each kernel according to its committed ncu --set full capture (profiles/): the pipe that saturates first
TRUE_BOUND = {
    "k_ens_svm": "mufu (ex2) / issue balanced: XU pipe 53 %, FMA pipe 67 %, issue 64 % (profiles/r1s_ncu_full_c3.md)",
    "k_ens_svm_mma": "mufu (ex2): dot products on the tensor pipe (3 x TF32)",
    "k_ens_svm_tma": "mufu (ex2): XU pipe 85 %, tensor pipe 42 %, issue 39 % (profiles/r3d_ncu_full_svm_f16.md); 2 500 exponentials per cell, "
                     "floor 36 ms at 16 / clk / SM, 38.0 ms alone; dot products as two HMMA.16816 with FP16 split operands, covariate tiles by "
                     "TMA tensor copies",
    "k_ens_trees": "issue / L2 latency: issue-active 65 %, L2 hit 97 %, L1 hit 29 % (profiles/r3z_ncu_full_trees.md); inside the step it runs on the "
                   "84-SM ensemble partition beside stage 1 of the fit (30.5 ms on all SMs)",
    "k_ens_fused": "issue + mufu: forest warps and support-vector warps share the SM",
    "k_sbr_chase": "latency: dependent L2 round trips between consecutive sweeps",
    "k_leaf_fused": "hbm / issue: DRAM traffic 1.077 GB for 1.074 GB algorithmic, issue 65 % (profiles/r3z_ncu_full_leaf.md; before the 2-D "
                    "tensor copy of the accumulator tile it was bound by the box barrier: 51 % of the stall samples, r2h_ncu_full_leaf_before_tma.md)",
    "k_leaf": "hbm write / issue",
}


def run_b200(args):
    """Default arm.  One raster, one fitted spline, one ensemble: N = 1 is BASELINE config 3 as is; at N > 1 the SAME problem is
    sharded by row blocks over the library's own NCCL communicator - `--scaling weak` (default): the raster grows to
    (N x nrow) x ncol cells so that every GPU owns one config-3-sized block (per-GPU work fixed); `--scaling strong`: the
    nrow x ncol raster itself is cut into N blocks.  Either way fields::Tps runs ONCE (on rank 0) beside the per-cell ensemble
    kernels of every rank, its descriptor is broadcast, the Gram and the R^2 sums are all-reduced - inside the library."""
    import torch
    import torch.distributed as dist
    import machisplin_b200 as mb
    from machisplin_b200 import parallel as par, synth

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
    weak = args.scaling == "weak"
    base_cfg = dict(cfg)
    if world > 1 and weak:
        # one shared raster, N times as tall, SAME cell size (every block looks like the base raster: the covariates are functions of
        # the coordinates, so a smaller cell would make them smoother and the forest pruning cheaper); the knots spread over all of it
        cfg["cell"] = 1.0 / max(cfg["nrow"], cfg["ncol"])
        cfg["nrow"] = cfg["nrow"] * world
    geom, xy, krow, kcol, resid, models, kept, w, wt = build_inputs(cfg, 0, model_cfg=base_cfg)
    r0, r1 = par.row_blocks(geom.nrow, world)[rank]
    bgeom = par.block_geom(geom, r0, r1)
    eng = mb.Engine(local)
    par.comm_init(eng)                              # the library's own communicator; torch only ships the 128-byte id
    eng.set_param("eval_precision", args.eval_precision)
    for kv in args.param:
        name, val = kv.split("=")
        eng.set_param(name, int(val))
    C = cfg["C"]
    P = C + 2
    n = len(resid)
    cov = device_covariates(bgeom, C, dev, disc_geom=geom) if C else torch.zeros((0,), device=dev)
    out = torch.empty((bgeom.nrow, bgeom.ncol), dtype=torch.float64, device=dev)
    ens = eng.ensemble_create(bgeom, models, kept, w, wt, P) if kept else None
    # k-fold residual matrix for the Gram reduction (a6): the same matrix everywhere, every rank owns a slice of its rows
    Rcv_all = np.random.default_rng(5).standard_normal((cfg["knots"], 6 if "b" in kept or not kept else 4))
    Rcv = Rcv_all[par.shard_rows(cfg["knots"], world, rank)]
    mine = (krow >= r0) & (krow < r1)               # the points whose cells this rank owns (part 5 extract)
    krow_b, kcol_b, resid_b = (krow[mine] - r0).astype(np.int32), kcol[mine], resid[mine]
    stream = torch.cuda.current_stream().cuda_stream
    cells_total = geom.nrow * geom.ncol
    state = {}

    def step_device():
        G = eng.gram_allreduce(Rcv)                                           # V73:329-333 (ncclAllReduce of K x K doubles)
        # parts 2-5 (V73:442-932): ensemble kernels of this block || fields::Tps fit on rank 0, broadcast, fused per-cell pass
        sp = eng.mltps_predict_shard_dev(bgeom, ens, cov.data_ptr() if C else 0, C, xy, resid, n, out.data_ptr(), lam=args.lam,
                                         root=0, stream=stream)
        f_actual = eng.gather_cells_dev(out.data_ptr(), bgeom.ncol, bgeom.nrow, bgeom.ncol, krow_b, kcol_b, stream=stream)
        rss = eng.allreduce([float(np.nansum((resid_b - f_actual) ** 2))])    # V73:912 summed over the ranks
        state.update(sp=sp, f_actual=f_actual, rss=float(rss[0]), G=G)
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
    value = cells_total / (ms * 1e-3) / 1e6

    # ---- the same raster the way machisplin.mltps() itself would cut it: 1500-px TPS sub-tiles (V73:649-895), N = 1 only -----
    tiled = None
    if world == 1 and not args.no_tiled and kept:
        def step_tiled():
            eng.gram_allreduce(Rcv)
            eng.mltps_predict_dev(geom, ens, cov.data_ptr() if C else 0, C, xy, resid, out.data_ptr(), lam=args.lam, tile_px=1500,
                                  stream=stream, want_spline=False)
            return eng.gather_cells_dev(out.data_ptr(), geom.ncol, geom.nrow, geom.ncol, krow, kcol, stream=stream)
        step_tiled()
        ms_t = timed(step_tiled, max(1, args.steps))
        nt = -(-geom.nrow // 1500) * -(-geom.ncol // 1500)
        tiled = {"tps_mode": f"mltps tiling, 1500 px: {nt} sub-tiles with own GCV fits, mean mosaic + seam feather",
                 "value": cells_total / (ms_t * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms_t}
        step_device()                                  # leave the global-spline raster in `out` for the parity check below

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
            eng.gram_allreduce(Rcv)
            ens2 = eng.ensemble_create(bgeom, models, kept, w, wt, P) if kept else None   # descriptor upload + tree packing
            t1 = time.perf_counter()
            eng.mltps_predict_shard(bgeom, ens2, cov_np, xy, resid, n, lam=args.lam, root=0, out=out_np)
            host_s["ensemble_create"] += t1 - t0
            host_s["mltps_predict"] += time.perf_counter() - t1
            f = out_np[krow_b, kcol_b]
            eng.allreduce([float(np.nansum((resid_b - f) ** 2))])
            return f

        step_e2e()
        for k in host_s:
            host_s[k] = 0.0
        ms_e2e = timed(step_e2e, max(1, args.steps))
        e2e = {"value": cells_total / (ms_e2e * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(world * (h_cov.numel() * 4) + xy.nbytes + resid.nbytes),
               "d2h_bytes_per_step": int(world * h_out.numel() * 8),
               "host_ms_per_step": {k: v * 1e3 / max(1, args.steps) for k, v in host_s.items()}}
        del h_cov

    # ---- parity sample: a few rows of EVERY rank's block travel to rank 0 (outside the timed region) --------------------
    S_par = 2
    pr0 = bgeom.nrow // 2
    sample = {"rank": rank, "row0": r0 + pr0, "rows": out[pr0:pr0 + S_par].cpu().numpy(),
              "cov": cov[:, pr0:pr0 + S_par].cpu().numpy() if C else np.zeros((0, S_par, bgeom.ncol), np.float32),
              "lam": state["sp"].lam}
    samples = [sample]
    if world > 1:
        box = [None] * world if rank == 0 else None
        dist.gather_object(sample, box, dst=0)
        samples = box

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: dominant kernel by share, the north-star kernel, and the whole step against section 8(d)'s bytes -------
    peak, peak_src = measured_peak()
    kern = {}
    tot_ms = sum(v[0] for v in ktimes.values()) or 1.0
    heavy = bool(kept) and bool(set(kept) & set("brv"))
    cells_rank = bgeom.nrow * bgeom.ncol
    # algorithmic HBM bytes per cell: k_leaf writes the float64 surface; k_leaf_fused reads the
    # float64 ensemble accumulator and writes the final raster; the ensemble kernels read the C float32 planes
    # and write (k_ens_trees, k_ens_fused) or read-modify-write (k_ens_svm, k_ens_smooth) the accumulator
    bytes_per_cell = {"k_leaf": 8.0, "k_leaf_f64": 8.0, "k_leaf_fused": 16.0, "k_leaf_f64_fused": 16.0,
                      "k_ens_final": 24.0, "k_ens_trees": 4.0 * C + 8.0, "k_ens_fused": 4.0 * C + 8.0,
                      "k_ens_svm": 4.0 * C + (16.0 if set(kept) & set("br") else 8.0),
                      "k_ens_svm_mma": 4.0 * C + (16.0 if set(kept) & set("br") else 8.0),
                      "k_ens_svm_tma": 4.0 * C + (16.0 if set(kept) & set("br") else 8.0),
                      "k_ens_smooth": 4.0 * C + (16.0 if heavy else 8.0)}
    for name, (tms, cnt) in sorted(ktimes.items(), key=lambda kv: -kv[1][0]):
        per_launch = tms / max(cnt, 1)
        ent = {"ms_per_step": tms / args.steps, "launches_per_step": cnt / args.steps, "share_of_kernel_time": tms / tot_ms,
               "share_of_step": tms / args.steps / ms}
        twin = {"k_leaf": "k_leaf_f64", "k_leaf_f64": "k_leaf", "k_leaf_fused": "k_leaf_f64_fused",
                "k_leaf_f64_fused": "k_leaf_fused"}.get(name)
        if twin in ktimes and ktimes[twin][0] > tms:
            ent["note"] = "twin of the selected leaf kernel: exits at its first instruction (DESIGN.md 4.1)"
        elif name in bytes_per_cell:
            gbs = cells_rank * bytes_per_cell[name] / (per_launch * 1e-3) / 1e9
            ent.update({"algorithmic_bytes_per_cell": bytes_per_cell[name], "achieved_gbs": gbs, "hbm_frac": gbs / peak})
        if name in TRUE_BOUND:
            ent["true_bound"] = TRUE_BOUND[name]
        kern[name] = ent
    leaf_names = ("k_leaf_fused", "k_leaf", "k_leaf_f64_fused", "k_leaf_f64")
    lname = next((k for k in leaf_names if "hbm_frac" in kern.get(k, {})), next((k for k in leaf_names if k in kern), "k_leaf"))
    leaf = kern.get(lname, {})
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        tr = {}

    def traffic_of(name):
        return tr.get(f"{name}@{bgeom.nrow}x{bgeom.ncol}", {}).get("dram_bytes_per_launch")

    dom = next((k for k in kern if "hbm_frac" in kern[k]), lname)          # kern is sorted by time: the dominant per-cell kernel
    step_bytes = (4.0 * C + 8.0) * cells_total                             # C float32 planes in, one float64 raster out
    step_gbs = step_bytes / (ms * 1e-3) / 1e9
    roofline = {
        "kernel": dom, "bound": "hbm", "achieved": kern.get(dom, {}).get("achieved_gbs"), "peak": peak, "unit": "GB/s",
        "frac": kern.get(dom, {}).get("hbm_frac"), "traffic": traffic_of(dom), "peak_source": peak_src,
        "algorithmic_bytes_per_cell": bytes_per_cell.get(dom), "share_of_step": kern.get(dom, {}).get("share_of_step"),
        "true_bound": TRUE_BOUND.get(dom, "see profiles/"),
        "note": "dominant per-cell kernel of the step by device time; its HBM fraction is small because it is bound by the pipe named "
                "in true_bound, not by memory - the memory-bound kernel of the path is north_star_kernel",
        "north_star_kernel": {"kernel": f"{lname} (grid evaluation: per-cell TPS surface" +
                                        (" + ensemble combine, mltps part 5)" if "fused" in lname else ")"),
                              "bound": "hbm", "achieved": leaf.get("achieved_gbs"), "peak": peak, "unit": "GB/s",
                              "frac": leaf.get("hbm_frac"), "traffic": traffic_of(lname),
                              "algorithmic_bytes_per_cell": bytes_per_cell[lname], "share_of_step": leaf.get("share_of_step")},
        "step": {"algorithmic_bytes": step_bytes, "bytes_per_cell": 4.0 * C + 8.0, "achieved": step_gbs, "peak": peak * world,
                 "unit": "GB/s", "frac": step_gbs / (peak * world),
                 "note": "whole step against SURVEY.md 8(d): (4 C + 8) B/cell over ms_per_step; the step is bound by the serial "
                         "GCV fit and the MUFU / issue-bound ensemble kernels, not by HBM"}}

    # ---- parity + CPU baseline on a bounded row sample ---------------------------------------------------
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        from oracle import cbind, tps as otps
        threads = host_threads()
        lapack_threads(threads)                                # LAPACK of the oracle fit: torchrun pinned it to 1
        sp = state["sp"]
        # parity: fields::Tps restated by the ORACLE (LAPACK eigendecomposition + fields' lambda search, ~10 s at 5 000 knots),
        # then the oracle's C loops on the sampled rows - nothing of the GPU fit enters the reference values
        t0 = time.perf_counter()
        ofit = otps.tps_fit(xy, resid, lam=args.lam)
        fit_s = time.perf_counter() - t0
        worst_abs, worst_rel, na_ok, rows_checked = 0.0, 0.0, True, []
        for smp in samples:
            ref, _ = cpu_sample(geom, ofit, models, kept, w, wt, smp["cov"], smp["row0"], threads)
            got = smp["rows"]
            m = ~np.isnan(ref)
            na_ok = na_ok and bool(np.array_equal(np.isnan(got), np.isnan(ref)))
            worst_abs = max(worst_abs, float(np.max(np.abs(got[m] - ref[m]))))
            worst_rel = max(worst_rel, float(np.max(np.abs(got[m] - ref[m])) / np.max(np.abs(ref[m]))))
            rows_checked.append([int(smp["row0"]), int(smp["row0"]) + S_par])
        parity = {"rows": rows_checked, "reference": "oracle GCV fit (LAPACK) + oracle C loops", "max_abs_err": worst_abs,
                  "max_rel_err": worst_rel, "na_mask_equal": na_ok, "tolerance": 1e-5,
                  "lambda_gpu": sp.lam, "lambda_oracle": ofit.lam, "lambda_rel_diff": abs(sp.lam - ofit.lam) / ofit.lam,
                  "lambda_equal_on_all_ranks": bool(all(s["lam"] == sp.lam for s in samples)),
                  "gram_max_rel_err": float(np.max(np.abs(state["G"] - Rcv_all.T @ Rcv_all)) / np.max(np.abs(Rcv_all.T @ Rcv_all)))}
        # CPU baseline: the same rows of work per cell, timed on a sample sized to ~10-15 s
        S = args.cpu_sample_rows or 2
        pr = min(pr0, bgeom.nrow - S)
        cov_rows = cov[:, pr:pr + S].cpu().numpy() if C else np.zeros((0, S, geom.ncol), np.float32)
        _, dt = cpu_sample(geom, ofit, models, kept, w, wt, cov_rows, r0 + pr, threads)
        if not args.cpu_sample_rows and dt < 8.0:
            S2 = int(max(S, min(bgeom.nrow - pr, round(S * 12.0 / max(dt, 1e-3)))))
            if S2 > S:
                S = S2
                cov_rows = cov[:, pr:pr + S].cpu().numpy() if C else np.zeros((0, S, geom.ncol), np.float32)
                _, dt = cpu_sample(geom, ofit, models, kept, w, wt, cov_rows, r0 + pr, threads)
        cpu = {"value": S * geom.ncol / dt / 1e6, "unit": "Mcells/s", "cores": threads, "kind": "port",
               "sample": f"{S} full-width rows ({S * geom.ncol} cells), TPS surface + ensemble, {dt:.1f} s; fit excluded "
                         f"(the oracle's GCV fit took {fit_s:.1f} s on the same cores)"}

    par_desc = "1 GPU" if world == 1 else \
        (f"{world} row blocks of {bgeom.nrow} rows of ONE {geom.nrow}x{geom.ncol} raster ({args.scaling} scaling); one fields::Tps fit on "
         f"rank 0, spline descriptor by ncclBroadcast, Gram + R^2 sums by ncclAllReduce, all inside the library; no raster exchange")
    conf = workload_config(cfg, args, "global")
    conf["parallelism"] = par_desc
    conf["comm"] = eng.comm_backend() if world > 1 else None
    line = {
        "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": conf,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kern,
        "cpu_baseline": cpu, "parity": parity, "mltps_tiled": tiled,
        "limiter": "serial part: the GCV fit of fields::Tps on rank 0 (two-stage tridiagonalisation, bulge chase) - every rank waits "
                   "for its broadcast; per-cell kernels shard perfectly" if world > 1 else
                   "SM-time: ensemble kernels (84 ms of the whole GPU) + full-GPU kernels of the fit (17 ms) + the SMs the bulge chase holds "
                   "(19 ms) packed into 134 ms by SM partitions (DESIGN.md section 6)",
        "fit": {"lambda": state["sp"].lam, "eff_df": state["sp"].eff_df, "knots": state["sp"].np, "rss_final": state["rss"]},
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
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
    from machisplin_b200 import synth, tiles as mtiles, parallel as par

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
    wins = [t.win for t in ts.tiles]
    par.comm_init(eng)                                  # the library's own communicator: Gram all-reduce + seam-strip exchange
    own = eng.tiles_owned_window(geom, wins, nC, nR, rank)
    out_own = torch.empty((own[1] - own[0], own[3] - own[2]), dtype=torch.float64, device=dev)   # this rank's part of the merged raster
    Rcv = np.random.default_rng(5).standard_normal((cfg["knots"], 6))[par.shard_rows(cfg["knots"], world, rank)]
    stream = torch.cuda.current_stream().cuda_stream
    tile_px = 1500

    def step():
        eng.gram_allreduce(Rcv)                                                # V73:329-333, ncclAllReduce inside the library
        eng.mltps_predict_dev(tg, ens, cov.data_ptr(), C, xy[pts], resid[pts], out_tile.data_ptr(), lam=args.lam,
                              tile_px=tile_px, stream=stream, want_spline=False)
        # machisplin.tiles.merge across the GPUs: seam strips point to point, every rank blends the cells it owns (no gather)
        eng.tiles_merge_shard_dev(geom, wins, {rank: out_tile.data_ptr()}, nC, nR, {rank: out_own.data_ptr()}, stream=stream)

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

    # ---- parity sample of the MERGED raster against the oracle: a patch across the seam between tiles 0 and 1 (two GPUs) -------
    parity = None
    if world > 1 and not args.no_cpu_baseline:
        parity = tiled_parity_patch(eng, dist, dev, rank, geom, ts, own, out_own, cov, models, kept, w, wt, xy, resid, tile_px, args.lam)
    nan_own = torch.isnan(out_own).sum().to(torch.float64)
    if world > 1:
        dist.all_reduce(nan_own)
    if rank == 0:
        cells = geom.nrow * geom.ncol
        tot = sum(v[0] for v in ktimes.values()) or 1.0
        kern = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps, "share": v[0] / tot}
                for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])[:12]}
        conf = workload_config(cfg, args, f"mltps tiling {tile_px} px inside {nC}x{nR} machisplin.tiles")
        conf["parallelism"] = (f"tiles{nC}x{nR}: one machisplin.tiles.create tile per GPU (own knots, own GCV fits), tile-border blend by "
                               f"seam-strip exchange (ncclSend / ncclRecv) + local blend of the owned cells, Gram by ncclAllReduce; no gather")
        conf["comm"] = eng.comm_backend() if world > 1 else None
        strip_bytes = 0
        for a in range(len(wins)):
            for b in range(len(wins)):
                if a % world != b % world:
                    ob = eng.tiles_owned_window(geom, wins, nC, nR, b)
                    rr, cc = min(wins[a][1], ob[1]) - max(wins[a][0], ob[0]), min(wins[a][3], ob[3]) - max(wins[a][2], ob[2])
                    if rr > 0 and cc > 0:
                        strip_bytes += rr * cc * 8
        line = {"metric": METRIC, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
                "e2e": None, "gpu_launches": int(launches), "clocks": clocks, "kernels_rank0": kern,
                "roofline": None, "cpu_baseline": None, "parity": parity, "na_fraction": float(nan_own.item()) / cells,
                "collectives": {"gram": "ncclAllReduce 36 doubles", "seam_boxes": "ncclAllReduce(min) 4 ints per seam",
                                "seam_strip_bytes_per_step": int(strip_bytes),
                                "whole_tile_gather_bytes_avoided": int(sum((t.win[1] - t.win[0]) * (t.win[3] - t.win[2]) for t in ts.tiles[1:]) * 8)}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        eng.comm_destroy()
        dist.destroy_process_group()


def tiled_parity_patch(eng, dist, dev, rank, geom, ts, own, out_own, cov, models, kept, w, wt, xy, resid, tile_px, lam):
    """Oracle values of the merged raster on a patch that straddles the seam between tiles 0 and 1 (east-west neighbours on two
    GPUs): per tile, the oracle's fields::Tps fit of the internal 1500-px sub-tile that covers the patch + its C loops for the
    surface and the ensemble (V73:649-753, 468-619), then oracle.tiles.tiles_merge on the two patch-high tiles (V73:1392-1548:
    the cross-fade weights depend on x only and the patch spans the whole overlap).  Ranks 0 and 1 compute their tile's side;
    rank 0 compares with the cells the two ranks own."""
    import torch
    from oracle import cbind, tiles as otl, tps as otps
    A, B = ts.tiles[0], ts.tiles[1]
    assert A.win[0] == B.win[0] and A.win[1] == B.win[1]
    lay = otl.mltps_tile_layout(A.geom.as_tuple(), tile_px)
    share = A.geom.nrow / lay.nRx
    tr0 = int((lay.nRx // 2 + 0.5) * share) - 2                     # 4 rows at the centre of an internal sub-tile row (tile coordinates)
    pr0, pr1 = A.win[0] + tr0, A.win[0] + tr0 + 4                   # ... in raster coordinates
    pc0, pc1 = B.win[2] - 16, A.win[3] + 16                         # the whole overlap zone + 16 cells either side
    res = None
    if rank in (0, 1):
        T = A if rank == 0 else B
        tgeom = T.geom.as_tuple()
        c0, c1 = max(pc0, T.win[2]) - T.win[2], min(pc1, T.win[3]) - T.win[2]        # patch columns inside this tile (tile coordinates)
        r0, r1 = pr0 - T.win[0], pr1 - T.win[0]
        layT = otl.mltps_tile_layout(tgeom, tile_px)
        hit = [k for k, kw in enumerate(layT.keep_win) if kw[0] < r1 and kw[1] > r0 and kw[2] < c1 and kw[3] > c0]
        ok = len(hit) == 1 and layT.keep_win[hit[0]][0] <= r0 and layT.keep_win[hit[0]][1] >= r1 and \
            layT.keep_win[hit[0]][2] <= c0 and layT.keep_win[hit[0]][3] >= c1
        if ok:
            fw = layT.fit_win[hit[0]]
            kxy = xy[T.points]
            krow, kcol = otl.cell_of_points(tgeom, kxy)
            inside = (krow >= fw[0]) & (krow < fw[1]) & (kcol >= fw[2]) & (kcol < fw[3])
            fit = otps.tps_fit(kxy[inside], resid[T.points][inside], lam=lam)
            tps = cbind.interpolate_c(fit, tgeom, r0, r1, c0, c1)
            Cn = cov.shape[0]
            cv = np.zeros((Cn, T.geom.nrow, T.geom.ncol), dtype=np.float32)           # only the patch rows are read
            cv[:, r0:r1, c0:c1] = cov[:, r0:r1, c0:c1].cpu().numpy()
            ens = cbind.ensemble_eval(models, kept, w, wt, cv, tgeom, window=(r0, r1, c0, c1))
            res = {"win": (pr0, pr1, T.win[2] + c0, T.win[2] + c1), "val": ens + tps, "knots": int(inside.sum()), "lam": fit.lam}
        # the cells of the patch this rank owns
        oc0, oc1 = max(pc0, own[2]), min(pc1, own[3])
        mine = out_own[pr0 - own[0]:pr1 - own[0], oc0 - own[2]:oc1 - own[2]].cpu().numpy()
        res = {"oracle": res, "own_cols": (oc0, oc1), "got": mine}
    box = [None, None]
    if rank == 0:
        box[0] = res
        other = [None]
        dist.recv_object_list(other, src=1)
        box[1] = other[0]
    elif rank == 1:
        dist.send_object_list([res], dst=0)
    if rank != 0:
        return None
    if box[0]["oracle"] is None or box[1]["oracle"] is None:
        return {"skipped": "the patch touches an internal sub-tile seam of a tile"}
    oa, ob = box[0]["oracle"], box[1]["oracle"]
    ref = otl.tiles_merge((geom.xmin, geom.xmax, geom.ymax - pr1 * geom.ry, geom.ymax - pr0 * geom.ry, pr1 - pr0, geom.ncol),
                          [(0, pr1 - pr0, oa["win"][2], oa["win"][3]), (0, pr1 - pr0, ob["win"][2], ob["win"][3])],
                          [oa["val"], ob["val"]], 2, 1)[:, pc0:pc1]
    got = np.full((pr1 - pr0, pc1 - pc0), np.nan)
    for b in box:
        got[:, b["own_cols"][0] - pc0:b["own_cols"][1] - pc0] = b["got"]
    m = ~np.isnan(ref)
    err = float(np.max(np.abs(got[m] - ref[m])))
    return {"patch": {"rows": [int(pr0), int(pr1)], "cols": [int(pc0), int(pc1)], "overlap_cols": [int(B.win[2]), int(A.win[3])]},
            "reference": "oracle sub-tile GCV fits + oracle C loops per tile, oracle.tiles.tiles_merge across the seam of tiles 0 | 1",
            "max_abs_err": err, "max_rel_err": err / float(np.max(np.abs(ref[m]))), "na_mask_equal": bool(np.array_equal(np.isnan(got), np.isnan(ref))),
            "tolerance": 1e-5, "knots_of_the_two_sub_tiles": [oa["knots"], ob["knots"]], "lambda_oracle": [oa["lam"], ob["lam"]]}


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
        conf["parallelism"] = f"layers{world}: the {L} response layers dealt round-robin to the GPUs, one shared tridiagonalisation per GPU for its layers"
        parity = None
        if not args.no_cpu_baseline:
            # the raster of the LAST layer of rank 0 (still in `out`) against the oracle: own LAPACK GCV fit of that layer + C loops
            from oracle import tps as otps
            threads = host_threads()
            lapack_threads(threads)
            lay = mine[-1]
            ofit = otps.tps_fit(xy, Y[:, lay], lam=args.lam)
            pr = geom.nrow // 2
            cov_rows = cov[:, pr:pr + 2].cpu().numpy()
            ref, _ = cpu_sample(geom, ofit, models, kept, w, wt, cov_rows, pr, threads)
            got = out[pr:pr + 2].cpu().numpy()
            m = ~np.isnan(ref)
            err = float(np.max(np.abs(got[m] - ref[m])))
            parity = {"layer": int(lay), "rows": [pr, pr + 2], "reference": "oracle GCV fit (LAPACK) + oracle C loops",
                      "max_abs_err": err, "max_rel_err": err / float(np.max(np.abs(ref[m]))),
                      "na_mask_equal": bool(np.array_equal(np.isnan(got), np.isnan(ref))), "tolerance": 1e-5,
                      "lambda_gpu": state["lam"][-1], "lambda_oracle": ofit.lam}
        line = {"metric": METRIC, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
                "e2e": None, "gpu_launches": int(launches), "clocks": clocks, "kernels_rank0": kern,
                "roofline": None, "cpu_baseline": None, "parity": parity, "layers": L, "lambda_rank0": state["lam"]}
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
