#!/usr/bin/env python
"""Headline benchmark: column*wavelength*layer flux updates per second of the fused RCE step.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (default)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU code on the host cores

Workload (BASELINE.json configs[3]): 65,536-column synthetic perturbed-profile ensemble PER GPU,
repwvl-100 table (100 wavelengths x 20 layers x 30 angles), one "step" = one full reference loop
iteration (main.cpp:531-583) for every column.  Columns shard over the ranks with no data-path
collective (weak scaling); the only inter-GPU traffic is the per-step allreduce of four scalars.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the roofline definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

NLAY = 20
METRIC = "column*wavelength*layer flux updates/s (repwvl-100 RCE step)"
UNIT = "updates/s"

# FP64-pipe work per unit (one column x wavelength x layer through one step), counted on the REFERENCE's
# expression tree: 31 exp, 33 divides, ~520 add/mul/fma.
ALG_EXP, ALG_DIV, ALG_FMA = 31, 33, 520
# FP64-pipe instructions this repo's kernel really executes per unit, and DRAM bytes per column-step:
# from the ncu capture committed under profiles/ (r1f_ncu_step_kernel_regions.md: 1,377,349,632 FP64 warp
# instructions and 85.9 + 24.7 MB of DRAM traffic for one launch of 65,536 columns x 100 wavelengths x 20 layers).
EXEC_FP64_PER_UNIT = 336.3
DRAM_BYTES_PER_COLUMN_STEP = 1688.0


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def build_ensemble(rcm, ncol, seed):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    Tlev, vlev = rcm.make_ensemble(ncol, seed, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    st["plevel"] = pl
    st["Tsurf"] = np.full(ncol, 288.2)  # main.cpp:357
    return st


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the reference's own code (oracle/_ref, built from /root/reference) on all
# host cores, one process per core, or the plain-C port when _ref is not there.
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, table, seed, ncol, nsteps = args
    import our_first_climate_model_b200 as rcm
    st = build_ensemble(rcm, ncol, seed)
    solar = rcm.solar_setup()["solar_irr"]
    t0 = time.perf_counter()
    if kind == "reference":
        from oracle import refcpu as R
        R.advance(table, st["plevel"], st["rel_hum"], solar, st["Tlayer"], st["Tsurf"], st["vmr9"], nsteps)
    else:
        from oracle import port as P
        P.advance(P.load_rcmtab(table), st["plevel"], st["rel_hum"], solar, st["Tlayer"], st["Tsurf"], st["vmr9"],
                  nsteps)
    return time.perf_counter() - t0


def cpu_throughput(nwvl, cols_per_core, nsteps, repeats=1):
    """-> (units/s over all cores, cores, kind, per-repeat wall times)"""
    import multiprocessing as mp
    from oracle import refcpu as R
    kind = "reference" if R.available() else "port"
    if kind == "port":
        from oracle import port as P
        P.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    table = os.path.join(GOLDEN, f"Reduced{nwvl}Forcing.rcmtab")
    walls = []
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(kind, table, 1, 2, 1)] * cores)  # load libraries, page in the table
        for rep in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(kind, table, 1000 + i, cols_per_core, nsteps) for i in range(cores)])
            walls.append(time.perf_counter() - t0)
    units = cores * cols_per_core * nsteps * nwvl * NLAY
    return units / statistics.median(walls), cores, kind, walls


def workload_config(ncol, nwvl, world):
    """`config` of the JSON line - the same for this repo's arm and the reference arm."""
    return {"workload": f"{ncol}-column synthetic perturbed-profile ensemble per GPU, repwvl-{nwvl} "
                        "RCE step (BASELINE configs[3])", "columns_per_gpu": ncol, "nwvl": nwvl,
            "nlayer": NLAY, "nangle": 30, "parallelism": f"columns sharded over {world} GPU(s)",
            "l2": "no flush: per-step working set (state + fluxes, ~130 MB at 65,536 columns) exceeds the "
                  "126 MB L2 and the kernel is FP64-pipe bound (DRAM < 1% of peak)"}


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    cols = args.cpu_cols
    # every "step" is one bounded sample: cols columns x 1 reference iteration per core, all cores busy
    vals = []
    cores = kind = None
    for i in range(args.warmup + args.steps):
        v, cores, kind, walls = cpu_throughput(args.nwvl, cols, 1)
        if i >= args.warmup:
            vals.append((v, walls[0]))
    value = statistics.median(v for v, _ in vals)
    ms = 1e3 * statistics.median(w for _, w in vals)
    sample = f"{cols} columns x 1 step per core on {cores} cores per bench step, table cached in RAM"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.ncol, args.nwvl, args.gpus),
                           sample=f"each bench step = {cols} columns x 1 reference iteration (main.cpp:531-583) per "
                                  f"host core, a bounded sample of that workload"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import our_first_climate_model_b200 as rcm
    from our_first_climate_model_b200 import distributed as rdist

    if not torch.cuda.is_available() or rcm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device - the solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        rdist.init("nccl")
    ncol = args.ncol
    st = build_ensemble(rcm, ncol, 12345 + rank)
    solver = rcm.Solver(local_rank)
    stream = torch.cuda.Stream()  # the solver launches on this torch stream: events and NCCL see its work
    torch.cuda.set_stream(stream)
    solver.set_stream(stream.cuda_stream)
    solver.set_repwvl_table_from(rcm.Table(os.path.join(GOLDEN, f"Reduced{args.nwvl}Forcing.rcmtab")))
    solver.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
    nwvl = solver.nwvl
    units_per_step = ncol * nwvl * NLAY

    peaks = {}
    if rank == 0:
        for i, nm in enumerate(("dfma", "exp", "div", "exp_solver")):
            peaks[nm] = solver.fp64_microbench(i)  # 1e9 ops/s

    exch = rdist.StepScalarExchange(torch.device("cuda", local_rank))

    def one_step():
        ptr = solver.advance_async(1)  # fused K1-K5 kernel + scalar reduction, all on `stream`
        # the per-step collective: one asynchronous 32-byte all_gather per step (no stall, see StepScalarExchange)
        exch.submit(rdist.device_view(ptr, 4))
        return ptr

    for _ in range(args.warmup):
        one_step()
    exch.latest()  # also loads the handful of torch kernels the read-out uses before the clock starts
    torch.cuda.synchronize()
    solver.kernel_time_ms(reset=True)
    l0 = solver.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        rdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    scal = exch.latest()  # waits for the last step's collective: inside the timed region
    e1.record(stream)
    if world > 1:
        rdist.barrier()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    toa_mean = float(scal[0]) / (world * ncol)
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.launch_count() - l0
    k_ms, k_n = solver.kernel_time_ms(reset=True)
    ms_total = rdist.max_over_ranks(ms_total) if world > 1 else ms_total
    value = world * units_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in and out through rcm_step_host, copies inside the timed region ------------
    nact = solver.nactive
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory()
    T_in, Ts_in, v_in = pin(ncol, NLAY), pin(ncol), pin(ncol, nact, NLAY)
    Ed, Eu, dE, T_out, Ts_out = pin(ncol, 21), pin(ncol, 21), pin(ncol, NLAY), pin(ncol, NLAY), pin(ncol)
    T_in.copy_(torch.from_numpy(st["Tlayer"]))
    Ts_in.copy_(torch.from_numpy(st["Tsurf"]))
    active = [k for k in range(9) if solver.params.species_mask >> k & 1]
    v_in.copy_(torch.from_numpy(np.ascontiguousarray(st["vmr9"][:, active, :])))
    ptrs = [t.data_ptr() for t in (T_in, Ts_in, v_in, Ed, Eu, dE, T_out, Ts_out)]
    h2d = (T_in.numel() + Ts_in.numel() + v_in.numel()) * 8
    d2h = (Ed.numel() + Eu.numel() + dE.numel() + T_out.numel() + Ts_out.numel()) * 8
    for _ in range(max(1, args.warmup // 2)):
        solver.step_host_ptrs(*ptrs)
    if world > 1:
        rdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.step_host_ptrs(*ptrs)
        if world > 1:
            pass  # scalars stay per rank in this leg; the copies dominate
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = rdist.max_over_ranks(e2e_s) if world > 1 else e2e_s
    e2e_value = world * units_per_step * args.steps / e2e_s
    olr_check = float(Eu[0, 0])

    if rank != 0:
        return 0
    # ---- roofline of the dominant (only) kernel: FP64 pipe -----------------------------------------------
    r_fma, r_exp, r_div = peaks["dfma"], peaks["exp"], peaks["div"]
    alg_per_unit = ALG_FMA + ALG_EXP * (r_fma / r_exp) + ALG_DIV * (r_fma / r_div)
    k_units_per_s = units_per_step / (k_ms * 1e-3) if k_ms > 0 else 0.0
    achieved = k_units_per_s * EXEC_FP64_PER_UNIT / 1e9  # FP64-pipe instructions really executed per second
    roofline = {"bound": "fp64", "achieved": achieved, "peak": r_fma, "unit": "G FP64-pipe instr/s",
                "frac": achieved / r_fma if r_fma else None,
                "traffic": DRAM_BYTES_PER_COLUMN_STEP * ncol if DRAM_BYTES_PER_COLUMN_STEP else None,
                "kernel": "rcm_step_kernel<MODE_STEP>", "kernel_ms": k_ms, "kernel_launches": k_n,
                "algorithmic_fp64_instr_per_unit": alg_per_unit,
                "executed_fp64_instr_per_unit": EXEC_FP64_PER_UNIT,
                "reference_tree_frac": (k_units_per_s * alg_per_unit / 1e9 / r_fma) if r_fma else None,
                "peak_source": "measured in this run: DFMA, exp(), divide microbenchmarks "
                               f"({r_fma:.0f}/{r_exp:.0f}/{r_div:.0f} Gop/s); MEASURED_PEAKS.json has no FP64 figure",
                "hbm_GBps": (DRAM_BYTES_PER_COLUMN_STEP * ncol / (k_ms * 1e-3) / 1e9) if k_ms > 0 else None}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, kind, walls = cpu_throughput(nwvl, args.cpu_cols, args.cpu_steps)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{args.cpu_cols} columns x {args.cpu_steps} steps per core, one process per core, "
                         f"table cached in RAM ({walls[0]:.1f} s wall)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(ncol, nwvl, world),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": "rcm_step_host (pinned host buffers; columns travel in 8 chunks through 3 streams, copies overlap the step)",
                    "check_olr_col0": olr_check},
            "gpu_launches": launches, "clocks": clocks, "fp64_peaks_Gops": peaks,
            "ensemble": {"toa_net_mean_Wm2": toa_mean, "collective": "1 async all_gather of 4 doubles per step" if world > 1
                         else "none (1 rank)"}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------
# Optional second workload (--workload lbl): BASELINE configs[4], a 4,096-column line-by-line ensemble with 2xCO2
# forcing, columns sharded over the ranks (total fixed: strong scaling).  The reference's LBL tables are not
# distributed: synthetic tables in its format (rcm_make_lbl_tables), --lbl-nwvl wavelengths.
# ------------------------------------------------------------------------------------------------------
LBL_EXEC_FP64_PER_UNIT = 400.1  # ncu, rcm_lbl_rt_kernel, 512 columns x 20,000 wavelengths (profiles/r1f_lbl_kernel.md)


def build_lbl_case(rcm, ncol, nwvl, seed):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.lbl.atm"))
    full = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    Tlev, vlev = rcm.make_ensemble(ncol, seed, pl, atm[:, 2].copy(), full[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev, 1.0)  # 2xCO2 enters through co2_factor of the LBL tables
    base = rcm.init_columns(pl, atm[None, :, 2].copy(), full[None, :, 4:9].transpose(0, 2, 1).copy(), 1.0)
    h2o_ref, o3_ref = base["vmr9"][0, 0].copy(), base["vmr9"][0, 2].copy()
    wvl, tau5 = rcm.make_lbl_tables(nwvl, 777, pl, h2o_ref, o3_ref)
    return dict(pl=pl, st=st, Tsurf=Tlev[:, 20].copy(), wvl=wvl, tau5=tau5, h2o_ref=h2o_ref, o3_ref=o3_ref)


def _lbl_cpu_worker(args):
    seed, ncol, nwvl = args
    import our_first_climate_model_b200 as rcm
    from oracle import port as P
    c = build_lbl_case(rcm, ncol, nwvl, seed)
    solar = rcm.solar_setup()["solar_irr"]
    t0 = time.perf_counter()
    P.lbl_advance(c["wvl"], c["tau5"], c["pl"], c["st"]["rel_hum"], c["h2o_ref"], c["st"]["vmr9"][:, 2] / c["o3_ref"], 2.0,
                  solar, c["st"]["Tlayer"], c["Tsurf"], c["st"]["vmr9"][:, 0], 1)
    return time.perf_counter() - t0


def run_b200_lbl(args, rank, world, local_rank):
    import torch
    import our_first_climate_model_b200 as rcm
    from our_first_climate_model_b200 import distributed as rdist
    if not torch.cuda.is_available() or rcm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device - the solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        rdist.init("nccl")
    total, nwvl = args.lbl_ncol, args.lbl_nwvl
    lo, hi = rdist.shard_range(total, rank, world)
    ncol = hi - lo
    c = build_lbl_case(rcm, total, nwvl, 4242)
    sl = slice(lo, hi)
    solver = rcm.Solver(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    solver.set_stream(stream.cuda_stream)
    solver.set_lbl_tables(c["wvl"], c["tau5"], c["h2o_ref"], c["o3_ref"], 2.0)
    solver.set_columns(c["pl"], c["st"]["Tlayer"][sl], c["Tsurf"][sl], c["st"]["vmr9"][sl], c["st"]["rel_hum"][sl])
    exch = rdist.StepScalarExchange(torch.device("cuda", local_rank))
    peak = solver.fp64_microbench(0) if rank == 0 else 0.0

    def one_step():
        exch.submit(rdist.device_view(solver.advance_async(1), 4))

    for _ in range(args.warmup):
        one_step()
    exch.latest()
    torch.cuda.synchronize()
    l0 = solver.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        rdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    scal = exch.latest()
    e1.record(stream)
    if world > 1:
        rdist.barrier()
    torch.cuda.synchronize()
    ms_total = rdist.max_over_ranks(e0.elapsed_time(e1)) if world > 1 else e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.launch_count() - l0
    units_per_step = total * nwvl * NLAY
    value = units_per_step * args.steps / (ms_total * 1e-3)
    # e2e: host buffers through rcm_step_host (T, Tsurf, active VMRs up; fluxes, heating rates, T down)
    nact = solver.nactive
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory()
    T_in, Ts_in, v_in = pin(ncol, NLAY), pin(ncol), pin(ncol, nact, NLAY)
    Ed, Eu, dE, T_out, Ts_out = pin(ncol, 21), pin(ncol, 21), pin(ncol, NLAY), pin(ncol, NLAY), pin(ncol)
    T_in.copy_(torch.from_numpy(c["st"]["Tlayer"][sl]))
    Ts_in.copy_(torch.from_numpy(c["Tsurf"][sl]))
    active = [k for k in range(9) if solver.params.species_mask >> k & 1]
    v_in.copy_(torch.from_numpy(np.ascontiguousarray(c["st"]["vmr9"][sl][:, active, :])))
    ptrs = [t.data_ptr() for t in (T_in, Ts_in, v_in, Ed, Eu, dE, T_out, Ts_out)]
    h2d = (T_in.numel() + Ts_in.numel() + v_in.numel()) * 8
    d2h = (Ed.numel() + Eu.numel() + dE.numel() + T_out.numel() + Ts_out.numel()) * 8
    solver.step_host_ptrs(*ptrs)
    if world > 1:
        rdist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.step_host_ptrs(*ptrs)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = rdist.max_over_ranks(e2e_s) if world > 1 else e2e_s
    if rank != 0:
        return 0
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        cols = 32
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_lbl_cpu_worker, [(1, 1, 200)] * cores)
            t0 = time.perf_counter()
            pool.map(_lbl_cpu_worker, [(100 + i, cols, nwvl) for i in range(cores)])
            wall = time.perf_counter() - t0
        cpu = {"value": cores * cols * nwvl * NLAY / wall, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cols} columns x 1 step per core of the oracle's LBL composition (the reference has no LBL driver), "
                         f"{wall:.1f} s wall"}
    achieved = value / world * LBL_EXEC_FP64_PER_UNIT / 1e9
    line = {"metric": "column*wavelength*layer flux updates/s (line-by-line RCE step)", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{total}-column line-by-line ensemble, 2xCO2, {nwvl} synthetic wavelengths "
                                   "(BASELINE configs[4]); columns sharded over the ranks", "columns_total": total,
                       "nwvl": nwvl, "nlayer": NLAY, "nangle": 30, "parallelism": f"{world} GPU(s)",
                       "l2": "tables (5 x nwvl x 20 doubles = 16 MB at 20,000 wavelengths) are re-read by every tile: L2-resident"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "G FP64-pipe instr/s",
                         "frac": achieved / peak if peak else None, "traffic": None, "kernel": "rcm_lbl_rt_kernel",
                         "executed_fp64_instr_per_unit": LBL_EXEC_FP64_PER_UNIT},
            "cpu_baseline": cpu,
            "e2e": {"value": units_per_step * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps, "api": "rcm_step_host"},
            "gpu_launches": launches, "clocks": clocks,
            "ensemble": {"toa_net_mean_Wm2": float(scal[0]) / total}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncol", type=int, default=65536, help="columns per GPU")
    ap.add_argument("--nwvl", type=int, default=100, choices=[10, 20, 100])
    ap.add_argument("--cpu-cols", type=int, default=None, help="columns per core in a CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="repwvl", choices=["repwvl", "lbl"],
                    help="repwvl = the headline metric (BASELINE configs[3]); lbl = configs[4], optional")
    ap.add_argument("--lbl-ncol", type=int, default=4096, help="columns of the LBL ensemble (all ranks together)")
    ap.add_argument("--lbl-nwvl", type=int, default=20000)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank, world, local_rank = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        if args.cpu_cols is None:
            args.cpu_cols = 256
        return run_reference(args, rank, world)
    if args.cpu_cols is None:
        args.cpu_cols = 512
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.workload == "lbl":
        return run_b200_lbl(args, rank, world, local_rank)
    return run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
