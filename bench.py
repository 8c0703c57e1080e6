#!/usr/bin/env python
"""Benchmark of the pyxrd_b200 hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

One JSON line on stdout (rank 0).  Primary metric: pattern-pts/s on BASELINE.json configs[1]
(c2: two-component illite-smectite R1, 2-theta 3-70 deg step 0.01, X = 6700), a population of
candidates per GPU; a "step" = all phase patterns of the population (+ mixture residuals).
The second half of BASELINE's metric, refinement trial evaluations/s, is measured on configs[3]
(c4: AD + EG specimens x 6 shared-parameter phases, population 256 per GPU) and reported under
the ``trials`` key of the same line.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyxrd_b200 import synthetic as syn  # noqa: E402

SEED = 2024
DEFAULT_CANDIDATES = {"c1": 4096, "c2": 2048, "c3": 64, "c3s": 256, "c4": 256, "c5": 1024}


# ---------------------------------------------------------------------------------------
# algorithmic work: flops per phase-pattern point
# ---------------------------------------------------------------------------------------
def structured_transitions(P, G):
    """1 if P (r x r, r > G) is zero outside the columns (I mod r/G) G + j of row I: the transition structure of
    every Reichweite >= 2 model (base_models.py:243-261), for which the kernels take G instead of r MACs per state;
    2 if r = G >= 2 and every row of P is the same (Reichweite 0): scalar series; else 0"""
    P = np.asarray(P)
    r = P.shape[0]
    if G < 2 or r % G:
        return 0
    if r == G:
        return 2 if bool(np.all(P == P[0])) else 0
    tails = r // G
    I, J = np.indices(P.shape)
    return 1 if bool(np.all(P[(J // G) != (I % tails)] == 0.0)) else 0


def work_per_point(G, r, n_terms, atoms, structured=0):
    """flops of the algorithm the kernels execute (fma = 2, mul / add = 1), DESIGN.md §3:
    series per term: rank 1 scalar Horner 7; rank 2 Clenshaw on (trace, det) 15; otherwise the vector
    Horner step with the REAL probability matrix, r*r real-x-complex MACs (4 flop; r*G for a matrix with the
    Reichweite >= 2 transition structure) + r complex updates (10 flop), Reichweite 0 (equal rows): the rank-1 count;
    transcendental equivalents as
    SURVEY.md §8(d): sincos 32, exp 20, erf 30."""
    per_term = 7 if (r == 1 or structured == 2) else 15 if r == 2 else 4 * r * (G if structured == 1 else r) + 10 * r
    series = n_terms * per_term + 16 * r
    factors = atoms * (32 + 8 + 2) + G * (20 + 32 + 12)
    return series + factors + 120


def measured_traffic(workload, candidates):
    """dram__bytes_read.sum + dram__bytes_write.sum of the phase kernel from the committed ncu --set full
    capture of this workload at this population size (profiles/r1_traffic.json), else None"""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[workload]
        return rec["dram_bytes"] if rec.get("candidates") == candidates else None
    except Exception:
        return None


def mixture_work(mix):
    """(total flops, total points, per-(G,r) breakdown) of all phase patterns of one candidate"""
    flops = 0.0
    points = 0
    for spec in mix.specimens:
        X = spec.range_theta.size
        for row in spec.phases:
            for ph in row:
                if ph is None or not ph.valid_probs:
                    continue
                atoms = sum(len(c.layer_atoms) + len(c.interlayer_atoms) for c in ph.components)
                n_terms = max(int(ph.CSDS.maximum) - 1, 0)
                flops += X * work_per_point(ph.G, ph.P.shape[0], n_terms, atoms, structured_transitions(ph.P, ph.G))
                points += X
    return flops, points


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def build_population(name, count, rank, ctx_calc):
    cfg = syn.CONFIGS[name]
    truth = cfg.truth()
    mixes = cfg.population(SEED + rank, count)

    def calc_total(mix):
        ctx_calc.mixture.calculate_mixture(mix)
        return [np.array(s.total_intensity) for s in mix.specimens]

    syn.attach_observed(mixes, truth, calc_total)
    return cfg, mixes


# ---------------------------------------------------------------------------------------
def time_workload(name, count, args, rank, world, torch, dist, ctx, stream, flush, optimise=False):
    """-> dict with resident-in-HBM and end-to-end timings of one workload on this rank.
    optimise=False: a step = all phase patterns + residuals at the candidates' solutions;
    optimise=True:  a step = a refinement trial per candidate: all phase patterns + the mixture
    optimiser (get_optimized_residual semantics)"""
    import pyxrd_b200.calculations as calc
    from pyxrd_b200.batched import pack_population
    cfg, mixes = build_population(name, count, rank, calc)
    t0 = time.perf_counter()
    static, batch = pack_population(mixes)
    pack_s = time.perf_counter() - t0
    flops_1, points_1 = 0.0, 0
    for m in mixes:
        f, p = mixture_work(m)
        flops_1 += f
        points_1 += p
    with ctx.lock:
        ctx.set_static(static, force=True)
        ctx.set_batch(batch)
        n_int = batch.K * batch.M * static.Z * static.sumX
        pinned_int = torch.empty(n_int, dtype=torch.float64).pin_memory().numpy()
        pinned_res = torch.empty(batch.K * (batch.S + 1), dtype=torch.float64).pin_memory().numpy()

        def sync_all():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- resident: parameters already in HBM ------------------------------------------
        def step_resident():
            if optimise:
                ctx.compute_intensities()
                ctx.optimize(refine_bg=True)
            else:
                ctx.compute_all(False)

        for _ in range(args.warmup):
            step_resident()
        sync_all()
        launches0 = ctx.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        stage = dict(phase=0.0, wavelength=0.0, residual=0.0)
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()                      # > L2: every step starts cold
            a.record(stream)
            step_resident()
            b.record(stream)
            st = ctx.stage_times()             # synchronises the stream
            for k in stage:
                stage[k] += st[k]
        sync_all()
        wall = time.perf_counter() - wall0
        launches = ctx.launch_count - launches0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # ---- end to end: host buffers in, host buffers out -------------------------------
        want_int = not optimise
        pinned_sol = torch.empty(batch.K * (batch.M + 2 * batch.S), dtype=torch.float64).pin_memory().numpy()

        def step_e2e():
            if optimise:
                ctx.optimize_host(batch, refine_bg=True, out_residuals=pinned_res, out_solutions=pinned_sol)
            else:
                ctx.evaluate_host(batch, out_residuals=pinned_res, out_intensities=pinned_int if want_int else None)

        for _ in range(min(args.warmup, 2)):
            step_e2e()
        sync_all()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        sync_all()
        e2e_s = time.perf_counter() - e0
        h2d = sum(a.nbytes for a in (batch.phase_i, batch.phase_d, batch.phase_comp, batch.matrices, batch.comp_d,
                                     batch.comp_i, batch.atom_d, batch.atom_type, batch.job_phase, batch.fractions,
                                     batch.scales, batch.bgshifts)) + 20 * batch.job_phase.size
        d2h = pinned_res.nbytes + (pinned_int.nbytes if want_int else pinned_sol.nbytes)
        # ---- population from a template: solution vectors x[K, d] in (host), residuals out (host) ----------
        pop = None
        if name in ("c1", "c2", "c4", "c5"):
            from pyxrd_b200.population import PopulationTemplate
            tmix, refs = syn.template(name)
            for spec, src in zip(tmix.specimens, mixes[0].specimens):
                spec.observed_intensity = np.array(src.observed_intensity)
            tmpl = PopulationTemplate(tmix, refs)
            x = syn.sample_solutions(refs, SEED + rank, batch.K)
            ctx.set_static(tmpl.static)

            def step_pop():
                if optimise:
                    return ctx.optimize_population(tmpl.batch, x, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor, refine_bg=True)[0]
                return ctx.evaluate_population(tmpl.batch, x, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor)

            for _ in range(min(args.warmup, 2)):
                step_pop()
            sync_all()
            p0 = time.perf_counter()
            for _ in range(args.steps):
                r_pop = step_pop()
            sync_all()
            pop = dict(ms=(time.perf_counter() - p0) * 1e3 / args.steps, d=x.shape[1], h2d=x.nbytes, d2h=r_pop.nbytes,
                       finite=bool(np.isfinite(r_pop).all()))
    t = torch.tensor([dev_ms, e2e_s * 1e3, wall * 1e3, pop["ms"] * args.steps if pop else 0.0], dtype=torch.float64,
                     device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms, pop_ms = [float(v) for v in t.tolist()]
    if pop:
        pop["ms"] = pop_ms / args.steps
    return dict(cfg=cfg, K=batch.K, points=points_1, flops=flops_1, dev_ms=dev_ms / args.steps,
                e2e_ms=e2e_ms / args.steps, wall_ms=wall_ms / args.steps, stage={k: v / args.steps for k, v in stage.items()},
                launches=launches, h2d=h2d, d2h=d2h, pack_s=pack_s, residual_sample=float(pinned_res[0]), pop=pop,
                n_specimens=batch.S, n_phases=batch.M, X=[int(v) for v in static.npts])


def cpu_sample(name, budget_s=12.0, processes=1):
    """oracle (numpy port of the reference algorithm) on a bounded sample of the same workload"""
    from oracle import pyxrd_oracle as orc
    cfg = syn.CONFIGS[name]
    mixes = cfg.population(SEED, 4096 if processes > 1 else 64)
    rng = np.random.default_rng(1)
    for m in mixes:
        for spec in m.specimens:
            spec.observed_intensity = rng.uniform(20.0, 60.0, (spec.range_theta.size, 1))
    _, pts = mixture_work(mixes[0])
    done = 0
    t0 = time.perf_counter()
    if processes <= 1:
        for m in mixes:
            orc.calculate_mixture(m)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(processes) as pool:
            t0 = time.perf_counter()
            chunk = processes * 2
            pos = 0
            while time.perf_counter() - t0 < budget_s and pos < len(mixes):
                pool.map(_cpu_one, [(name, SEED, pos + i) for i in range(chunk)])
                pos += chunk
                done += chunk
    dt = time.perf_counter() - t0
    return dict(patterns=done, seconds=dt, pts_per_s=done * pts / dt, trials_per_s=done / dt)


def cpu_trials(budget_s=10.0, name="c4"):
    """oracle get_optimized_residual (= the reference's per-trial callback) on a few candidates"""
    from oracle import pyxrd_oracle as orc
    cfg = syn.CONFIGS[name]
    truth = cfg.truth()
    mixes = cfg.population(SEED, 64)

    def calc_total(mix):
        orc.calculate_mixture(mix)
        return [np.array(s.total_intensity) for s in mix.specimens]

    syn.attach_observed(mixes, truth, calc_total)
    for m in mixes:
        m.parsed = m.calculated = m.optimized = False
    t0 = time.perf_counter()
    n = 0
    for m in mixes:
        orc.get_optimized_residual(m)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(n=n, seconds=dt, trials_per_s=n / dt)


def _cpu_one(arg):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import pyxrd_oracle as orc
    name, seed, idx = arg
    m = syn.CONFIGS[name].candidate(seed, idx)
    rng = np.random.default_rng(1)
    for spec in m.specimens:
        spec.observed_intensity = rng.uniform(20.0, 60.0, (spec.range_theta.size, 1))
    orc.calculate_mixture(m)
    return float(m.residuals[0])


# ---------------------------------------------------------------------------------------
def run_reference(args, rank):
    """the reference arm: the CPU algorithm (oracle port; the reference itself is pure Python and
    cannot travel to the GPU box) on all host cores, same workload / metric / unit"""
    if rank != 0:
        return
    name = args.workload
    cores = os.cpu_count() or 1
    steps = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(name, budget_s=max(4.0, 40.0 / max(args.steps, 1)), processes=cores)
        if i >= args.warmup:
            steps.append(r)
    pts = sum(r["patterns"] for r in steps) * mixture_work(syn.CONFIGS[name].truth())[1]
    secs = sum(r["seconds"] for r in steps)
    value = pts / secs
    line = {"impl": "reference", "metric": "pattern-pts/s", "value": value, "unit": "pattern-pts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / max(len(steps), 1) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (name, syn.CONFIGS[name].description),
                       "candidates_per_gpu": args.candidates or DEFAULT_CANDIDATES[name],
                       "points_per_pattern": [int(sp.range_theta.size) for sp in syn.CONFIGS[name].truth().specimens],
                       "specimens": syn.CONFIGS[name].n_specimens, "phases": syn.CONFIGS[name].n_phases,
                       "l2": "256 MiB flush write between timed steps", "population_seed": SEED},
            "cpu_baseline": {"value": value, "unit": "pattern-pts/s", "cores": cores, "kind": "port",
                             "sample": "%d candidates per step in a %d-process pool (multiprocessing, as "
                                       "pyxrd/server/pyxrd_server.py:57-76 does)" % (steps[0]["patterns"], cores)},
            "e2e": {"value": value, "unit": "pattern-pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "trials": {"value": sum(r["patterns"] for r in steps) / secs, "unit": "calculate_mixture evals/s"}}
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyxrd_b200 has no CPU fallback")
    from pyxrd_b200.distributed import bind_process_to_device
    numa_bound = bind_process_to_device(local_rank)          # before any pinned allocation
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from pyxrd_b200.runtime import default_context
    ctx = default_context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        peak_tf, _ = ctx.dfma_peak(8192)
        sampler = ClockSampler(local_rank)
        sampler.start()
        main = time_workload(args.workload, args.candidates or DEFAULT_CANDIDATES[args.workload], args, rank, world,
                             torch, dist, ctx, stream, flush)
        clocks = sampler.stop()
        trials = None
        if args.workload != args.trials_workload and not args.no_trials:
            trials = time_workload(args.trials_workload, args.trials_candidates or DEFAULT_CANDIDATES[args.trials_workload],
                                   args, rank, world, torch, dist, ctx, stream, flush, optimise=True)
        # residual gather (the only collective of the path): [K, 1+S] per rank over NCCL
        gather_us = None
        if world > 1:
            local = ctx.residuals_tensor()
            out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device="cuda")
            dist.all_gather_into_tensor(out, local)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                dist.all_gather_into_tensor(out, local)
            torch.cuda.synchronize()
            gather_us = (time.perf_counter() - t0) / 10 * 1e6
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    total_pts = main["points"] * world
    value = total_pts / (main["dev_ms"] * 1e-3)
    e2e_value = total_pts / (main["e2e_ms"] * 1e-3)
    phase_ms = main["stage"]["phase"]
    achieved_tf = main["flops"] / (phase_ms * 1e-3) * 1e-12
    line = {
        "metric": "pattern-pts/s", "value": value, "unit": "pattern-pts/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["dev_ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, main["cfg"].description),
                   "candidates_per_gpu": main["K"], "points_per_pattern": main["X"], "specimens": main["n_specimens"],
                   "phases": main["n_phases"], "l2": "256 MiB flush write between timed steps",
                   "population_seed": SEED},
        "e2e": {"value": e2e_value, "unit": "pattern-pts/s", "h2d_bytes_per_step": main["h2d"],
                "d2h_bytes_per_step": main["d2h"], "ms_per_step": main["e2e_ms"],
                "call": "pyxrd_evaluate_host (packed host parameters in, pinned host patterns + residuals out)"},
        "gpu_launches": main["launches"],
        "roofline": {"bound": "fp64", "kernel": "k_phase_* (phase-intensity kernels)", "achieved": achieved_tf,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                     "peak_source": "DFMA microbenchmark pyxrd_dfma_peak measured in this run (MEASURED_PEAKS.json "
                                    "has no fp64 entry)",
                     "kernel_ms": phase_ms, "algorithmic_flops_per_launch": main["flops"],
                     "traffic": measured_traffic(args.workload, main["K"]),
                     "hbm": {"achieved_gbs": 16.0 * main["points"] / (phase_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                             "frac": 16.0 * main["points"] / (phase_ms * 1e-3) * 1e-9 / hbm_peak,
                             "peak_source": "of measured" if peaks else "of fallback"}},
        "stage_ms": main["stage"], "host_pack_s": main["pack_s"], "clocks": clocks,
        "cpu_affinity": "NVML ideal affinity of the rank's GPU" if numa_bound else "unchanged",
    }
    if main["pop"]:
        p = main["pop"]
        line["population"] = {"call": "pyxrd_evaluate_population: solution vectors x[K, d] on the host -> native expansion "
                                      "of the packed template + binding table -> phase patterns -> residuals on the host",
                              "value": total_pts / (p["ms"] * 1e-3), "unit": "pattern-pts/s", "ms_per_step": p["ms"],
                              "evals_per_s": main["K"] * world / (p["ms"] * 1e-3), "refinables": p["d"],
                              "h2d_bytes_per_step": p["h2d"], "d2h_bytes_per_step": p["d2h"]}
    if trials is not None:
        tv = trials["K"] * world / (trials["dev_ms"] * 1e-3)
        line["trials"] = {"metric": "refinement trial evals/s (get_optimized_residual semantics: all phase patterns "
                                    "of the trial + optimisation of fractions / scales / background + residual)",
                          "value": tv, "unit": "trials/s", "ms_per_step": trials["dev_ms"],
                          "e2e": {"value": trials["K"] * world / (trials["e2e_ms"] * 1e-3), "unit": "trials/s",
                                  "h2d_bytes_per_step": trials["h2d"], "d2h_bytes_per_step": trials["d2h"]},
                          "config": {"workload": "%s: %s" % (args.trials_workload, trials["cfg"].description),
                                     "candidates_per_gpu": trials["K"]},
                          "stage_ms": {"phase": trials["stage"]["phase"], "wavelength": trials["stage"]["wavelength"],
                                       "optimiser": trials["stage"]["residual"]},
                          "optimiser": "device majorise-minimise (pyxrd_optimize), 32 reweightings x <=2 Newton steps; "
                                       "contract: <= reference L-BFGS-B optimum * (1 + 5e-3)",
                          "population": None if not trials["pop"] else {
                              "call": "pyxrd_optimize_population: x[K, d] on the host -> optimised residuals + solutions on the host",
                              "value": trials["K"] * world / (trials["pop"]["ms"] * 1e-3), "unit": "trials/s",
                              "ms_per_step": trials["pop"]["ms"], "refinables": trials["pop"]["d"]},
                          "gpu_launches": trials["launches"],
                          "pattern_pts_per_s": trials["points"] * world / (trials["dev_ms"] * 1e-3)}
    if gather_us is not None:
        line["residual_gather_us"] = gather_us
    if world == 1 and not args.no_cpu:
        c = cpu_sample(args.workload, budget_s=12.0, processes=1)
        line["cpu_baseline"] = {"value": c["pts_per_s"], "unit": "pattern-pts/s", "cores": 1, "kind": "port",
                                "sample": "%d candidates of %s through oracle.calculate_mixture, single process, "
                                          "%.1f s" % (c["patterns"], args.workload, c["seconds"])}
        if trials is not None:
            ct = cpu_trials(budget_s=10.0, name=args.trials_workload)
            line["trials"]["cpu_baseline"] = {"value": ct["trials_per_s"], "unit": "trials/s", "cores": 1, "kind": "port",
                                              "sample": "%d c4 candidates through oracle.get_optimized_residual (scipy "
                                                        "L-BFGS-B), single process, %.1f s" % (ct["n"], ct["seconds"])}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(syn.CONFIGS))
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU (default per workload)")
    ap.add_argument("--no-trials", action="store_true")
    ap.add_argument("--trials-workload", default="c4", choices=sorted(syn.CONFIGS),
                    help="configuration of the trial-evaluation measurement (BASELINE configs[3] by default)")
    ap.add_argument("--trials-candidates", type=int, default=0, help="trials per GPU of that measurement")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
