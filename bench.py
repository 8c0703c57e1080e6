#!/usr/bin/env python
"""Benchmark of the pyxrd_b200 hot path (contract: the task statement / DESIGN.md section 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on stdout (rank 0).  What a "step" is, at every N: one refinement GENERATION of BASELINE.json configs[1]
(c2: two-component illite-smectite R1, 2-theta 3-70 deg step 0.01, X = 6700) -- a solution matrix x[K, d] of 2048
candidates per GPU -> every rank's block expanded on the device from the packed template + binding table -> per-candidate
prologue (probability matrices, CSDS, absolute scale, z-stretch) -> all phase patterns -> wavelength distribution ->
mixture residuals -> (N > 1) NCCL all-gather of the residual rows f64[K, 1+S] straight from the device buffer.

  value  pattern-pts/s of that step with x already in HBM, CUDA events on the launching stream (gather inside);
  e2e    the same through the product's sharded call (pyxrd_b200.distributed.evaluate_solutions_sharded): x on the
         host in, the gathered residual table on the host out, wall clock between device synchronisations;
  trials refinement trial evaluations/s (get_optimized_residual semantics: patterns + mixture optimiser) on configs[3]
         (c4: AD + EG specimens x 6 shared-parameter phases, CMA-ES population 256 per GPU), same two figures;
  sweep  configs[4]: c5 populations of 1k / 4k / 16k / 64k trials in TOTAL, split over the N ranks (strong scaling),
         with a bitwise comparison of the gathered table against rank 0 evaluating all of it alone;
  configs (N = 1) device-resident pattern-pts/s and roofline fraction of c1, c3, c3s, c4, c5;
  patterns_e2e (N = 1) the pattern-download call (pyxrd_evaluate_host: packed parameters in, all patterns out).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyxrd_b200 import synthetic as syn  # noqa: E402

SEED = 2024
DEFAULT_CANDIDATES = {"c1": 4096, "c2": 2048, "c3": 64, "c3s": 256, "c4": 256, "c5": 1024}
SWEEP_TOTALS = (1024, 4096, 16384, 65536)
L2_NOTE = "256 MiB flush write between timed steps"


# ---------------------------------------------------------------------------------------
# algorithmic work ("Roofline bookkeeping"): flops per phase-pattern point
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


def series_flops(G, r, n_terms, structured=0):
    """flops of the series the kernels execute (fma = 2): rank 1 scalar Horner 7 per term; rank 2 Clenshaw on (trace,
    det) 15; otherwise the vector Horner step with the REAL probability matrix, r*r real-x-complex MACs (4 flop; r*G
    for a matrix with the Reichweite >= 2 transition structure) + r complex updates (10 flop); Reichweite 0 (equal
    rows): the rank-1 count"""
    per_term = 7 if (r == 1 or structured == 2) else 15 if r == 2 else 4 * r * (G if structured == 1 else r) + 10 * r
    return n_terms * per_term + 16 * r


def work_per_point(G, r, n_terms, atoms, structured=0, exponentials=None):
    """ALGORITHMIC flops per phase-pattern point (SURVEY.md 8(d) conventions: sincos 32, exp 20, erf 30): series +
    per atom (phase factor 32 + MAC 8 + ASF MAC 2) + per component (exp 20 + sincos 32 + 12) + LPF / corrections 120.
    ``exponentials``: the EXECUTED count instead -- the kernels evaluate one complex exponential per distinct plane of
    the phase (atoms on one plane, and planes several components share, are merged; candidate-invariant layer stacks
    come from a table), so 32 + 8 is charged per executed exponential and 2 per atom"""
    if exponentials is None:
        factors = atoms * (32 + 8 + 2)
    else:
        factors = exponentials * (32 + 8) + atoms * 2
    return series_flops(G, r, n_terms, structured) + factors + G * (20 + 32 + 12) + 120


def mixture_work(mix):
    """(algorithmic flops, phase-pattern points) of all phase patterns of one candidate graph"""
    flops, points = 0.0, 0
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


_MODEL_STRUCTURE = {3: 1, 4: 1, 7: 1, 1: 2}        # PYXRD_PROB_R2G2 / R3G2 / R2G3: structured; R0: equal rows


def population_work(tmpl, x):
    """(algorithmic flops, executed flops, phase-pattern points) of the K candidates ``x`` describes, from the natively
    expanded arrays (no Python graphs): per computed pattern X_s * work_per_point(phase)"""
    e = tmpl.expand(x)
    st = tmpl.static
    K, M = e["K"], e["M"]
    pi, ci = e["phase_i"], e["comp_i"]
    NP = pi.shape[0]
    atoms_of_comp = ci[:, 1] + ci[:, 2]
    per_phase_alg, per_phase_exe = np.zeros(NP), np.zeros(NP)
    t_pi = tmpl.batch.phase_i
    npt = t_pi.shape[0]
    t_struct = np.zeros(npt, dtype=int)
    t_exp = np.zeros(npt, dtype=int)
    bound_layers = set()          # components with a refinable on one of their layer atoms (PYXRD_BIND_ATOM_D = 2)
    for row in tmpl.bindings:
        if int(row["target"]) == 2:
            atom = int(row["index"]) // 2
            for c in range(tmpl.batch.comp_i.shape[0]):
                a0, nl = int(tmpl.batch.comp_i[c, 0]), int(tmpl.batch.comp_i[c, 1])
                if a0 <= atom < a0 + nl:
                    bound_layers.add(c)
    for p in range(npt):
        G, r = int(t_pi[p, 0]), int(t_pi[p, 1])
        if not (t_pi[p, 4] & 1) or (t_pi[p, 4] & 8) or G < 1:
            continue
        model = int(tmpl.batch.prob_model[p]) if tmpl.batch.prob_model is not None else 0
        if model:
            t_struct[p] = _MODEL_STRUCTURE.get(model, 0) if G >= 2 else 0
        else:
            off = int(t_pi[p, 6])
            t_struct[p] = structured_transitions(tmpl.batch.matrices[off + r * r: off + 2 * r * r].reshape(r, r), G)
        # complex exponentials the kernels evaluate for the phase: one per distinct plane (z after the interlayer stretch
        # of the TEMPLATE; the count does not depend on x).  The layer planes of a component no refinable drives come from
        # the layer table of the batch (k_layer_table: populations of >= 3 candidates) and cost no exponential here.
        zs = set()
        for c in tmpl.batch.phase_comp[t_pi[p, 5]: t_pi[p, 5] + G]:
            a0, nl, ni = (int(v) for v in tmpl.batch.comp_i[c, :3])
            d001, default_c, _, lattice_d = tmpl.batch.comp_d[c, :4]
            zf = (d001 - lattice_d) / (default_c - lattice_d)
            tabulated = K >= 3 and c not in bound_layers
            for a in range(nl + ni):
                z0 = tmpl.batch.atom_d[a0 + a, 1]
                if a < nl and tabulated:
                    continue
                zs.add(float(z0) if a < nl else float(lattice_d + zf * (z0 - lattice_d)))
        t_exp[p] = len(zs)
    for p in range(NP):
        G, r = int(pi[p, 0]), int(pi[p, 1])
        if not (pi[p, 4] & 1) or (pi[p, 4] & 8) or G < 1:
            continue
        comps = e["phase_comp"][pi[p, 5]: pi[p, 5] + G]
        atoms = int(atoms_of_comp[comps].sum())
        n_terms = max(int(pi[p, 3]) - 1, 0)
        tp = p % npt
        per_phase_alg[p] = work_per_point(G, r, n_terms, atoms, t_struct[tp])
        per_phase_exe[p] = work_per_point(G, r, n_terms, atoms, t_struct[tp], exponentials=int(t_exp[tp]))
    jobs = e["job_phase"].reshape(K, st.S, st.Z, M)
    alg = exe = 0.0
    points = 0
    for s in range(st.S):
        X = int(st.npts[s])
        pb = jobs[:, s].ravel()
        pb = pb[pb >= 0]
        live = pb[per_phase_alg[pb] > 0]
        alg += X * float(per_phase_alg[live].sum())
        exe += X * float(per_phase_exe[live].sum())
        points += X * int(live.size)
    return alg, exe, points


def optimiser_flops(M, selected_points, passes):
    """flops of the Gram passes of k_mixture_optimize for one candidate (fma = 2): per selected point and pass the
    model value (2 MB), the residual and its weight (3), the upper triangle of the weighted Gram matrix (2 NG), the
    right-hand side (2 MB), sum |r| (1); MB = M + 1 columns (phases + background), NG = MB (MB + 1) / 2.  The small
    Newton systems are not counted."""
    MB = M + 1
    NG = MB * (MB + 1) // 2
    return float(passes) * float(selected_points) * (2 * NG + 4 * MB + 4)


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
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


class Env(object):
    """what every measurement needs: torch, the process group, the library context on one stream, the L2 flush"""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world, self.local_rank = rank, world, local_rank
        from pyxrd_b200.runtime import default_context
        self.ctx = default_context(local_rank)
        for item in getattr(args, "option", []):
            key, _, value = item.partition("=")
            self.ctx.set_option(key, float(value))
        self.stream = torch.cuda.Stream()
        self.ctx.set_stream(self.stream.cuda_stream)
        self.device = torch.device("cuda", local_rank)
        with torch.cuda.stream(self.stream):
            self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def sync_all(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def time_device(self, step, steps, warmup, stages=None):
        """EXACTLY ``steps`` steps, each bracketed by CUDA events on the launching stream, L2 flushed before each,
        barrier + synchronize on both sides; -> (sum of event ms, stage sums, launches)"""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.sync_all()
        launches0 = self.ctx.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        acc = {k: 0.0 for k in (stages or ())}
        for a, b in ev:
            self.flush.zero_()                     # > L2: every step starts cold
            a.record(self.stream)
            step()
            b.record(self.stream)
            if stages:
                st = self.ctx.stage_times()        # synchronises the stream
                for k in acc:
                    acc[k] += st[k]
        self.sync_all()
        return sum(a.elapsed_time(b) for a, b in ev), acc, self.ctx.launch_count - launches0

    def time_wall(self, step, steps, warmup):
        for _ in range(warmup):
            step()
        self.sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.sync_all()
        return (time.perf_counter() - t0) * 1e3


def optimiser_threads(candidates, sm_count=148):
    """block size k_mixture_optimize picks for a population of that size (pyxrd_b200.cu: launch_optimize)"""
    return 256 if candidates <= 2 * sm_count else 128


def observed_for(name, calc):
    """the 'observed' patterns of a configuration: its truth evaluated by the device path + noise (seed 1234)"""
    cfg = syn.CONFIGS[name]
    truth = cfg.truth()

    def calc_total(mix):
        calc.mixture.calculate_mixture(mix)
        return [np.array(s.total_intensity) for s in mix.specimens]

    syn.attach_observed([], truth, calc_total)
    return [np.array(s.observed_intensity) for s in truth.specimens]


def make_template(name, calc):
    from pyxrd_b200.population import PopulationTemplate
    observed = observed_for(name, calc)
    tmix, refs = syn.template(name)
    for spec, obs in zip(tmix.specimens, observed):
        spec.observed_intensity = obs.copy()
    return PopulationTemplate(tmix, refs), refs


# ---------------------------------------------------------------------------------------
def measure_generation(env, name, per_gpu, optimise, verify=True):
    """one configuration through the template path, K = per_gpu * world solutions in total (weak scaling)"""
    import pyxrd_b200.calculations as calc
    from pyxrd_b200 import distributed
    torch, dist, ctx, args = env.torch, env.dist, env.ctx, env.args
    tmpl, refs = make_template(name, calc)
    S, M = tmpl.batch.S, tmpl.batch.M
    x_total = syn.sample_solutions(refs, SEED, per_gpu * env.world)
    lo, hi = distributed.shard_bounds(len(x_total), env.world)[env.rank]
    x_mine = np.ascontiguousarray(x_total[lo:hi])
    alg, exe, points = population_work(tmpl, x_mine)
    refine_bg = tmpl.mixture_refines_background()
    with torch.cuda.stream(env.stream), ctx.lock:
        tmpl._load(ctx, x_mine)
        gathered = torch.empty((env.world * per_gpu, S + 1), dtype=torch.float64, device="cuda")
        # the exchange: one NCCL all_gather_into_tensor of the rows; --gather peer: this rank's rows stored into every
        # rank's table over NVLink by the library's kernel + one barrier (distributed.PeerGather)
        peer = None
        if env.world > 1 and args.gather == "peer":
            peer = distributed.peer_gather_for(ctx, per_gpu, S + 1, env.device)
        gather_ctx = ctx if peer is not None else None

        def step_resident():
            ctx.update_population(None)            # expansion + per-candidate prologue from the resident x
            if optimise:
                ctx.compute_intensities()
                ctx.optimize(refine_bg=refine_bg)
            else:
                ctx.compute_all(residuals_only=True)
            if peer is not None:
                peer(ctx.residuals_tensor(optimised=optimise), per_gpu * env.world)
            elif env.world > 1:
                dist.all_gather_into_tensor(gathered, ctx.residuals_tensor(optimised=optimise))

        dev_ms, stage, launches = env.time_device(step_resident, args.steps, args.warmup,
                                                  stages=("phase", "wavelength", "residual", "prologue"))
        table = [None]

        def step_e2e():
            if optimise:
                table[0] = distributed.evaluate_solutions_sharded(
                    x_total, lambda xb: tmpl.optimized_residuals(xb, ctx=ctx, device=True), device=env.device, columns=S + 1,
                    ctx=gather_ctx)
            else:
                table[0] = distributed.evaluate_solutions_sharded(
                    x_total, lambda xb: tmpl.residuals(xb, ctx=ctx, device=True), device=env.device, columns=S + 1,
                    ctx=gather_ctx)

        e2e_ms = env.time_wall(step_e2e, args.steps, min(args.warmup, 2))
        check = None
        if verify and env.world > 1 and env.rank == 0:
            # the gathered table equals a single-GPU evaluation of all K solutions, bit for bit (same
            # kernels: the optimiser's block size, which it otherwise picks by population size, is held at the shards')
            ctx.set_option("optimizer_threads", optimiser_threads(per_gpu))
            alone = (tmpl.optimized_residuals(x_total, ctx=ctx, device=True) if optimise
                     else tmpl.residuals(x_total, ctx=ctx, device=True)).cpu().numpy()
            ctx.set_option("optimizer_threads", 0)
            check = bool(np.array_equal(alone, table[0]))
            if not check:
                raise AssertionError("%s: the gathered residual table differs from the single-GPU table" % name)
    dev_ms, e2e_ms = env.max_over_ranks([dev_ms, e2e_ms])
    sel_points = sum(int(np.count_nonzero(tmpl.static.selected[tmpl.static.offsets[s]:tmpl.static.offsets[s + 1]])) * tmpl.static.Z
                     for s in range(S))
    return dict(name=name, cfg=syn.CONFIGS[name], K=per_gpu, d=x_total.shape[1], S=S, M=M, points=points, flops=alg,
                flops_executed=exe, dev_ms=dev_ms / args.steps, e2e_ms=e2e_ms / args.steps,
                stage={k: v / args.steps for k, v in stage.items()}, launches=launches,
                h2d=int(x_mine.nbytes), d2h=int(per_gpu * env.world * (S + 1) * 8), finite=bool(np.isfinite(table[0]).all()),
                gather_matches_single_gpu=check, selected_points=sel_points, X=[int(v) for v in tmpl.static.npts],
                residual_sample=float(table[0][0, 0]), tmpl=tmpl, x_mine=x_mine,
                gather=("none (one rank)" if env.world == 1 else "rows stored into every rank's table over NVLink by "
                        "k_put_rows_to_peers + one symmetric-memory barrier" if peer is not None else "NCCL all_gather_into_tensor"))


def measure_patterns_download(env, gen):
    """N = 1 only: the pattern-download call -- packed parameter blocks of the K candidates on the host in, all phase
    patterns + residuals in pinned host memory out (pyxrd_evaluate_host: 8 candidate chunks, copies overlap compute)"""
    torch, ctx, args = env.torch, env.ctx, env.args
    tmpl = gen["tmpl"]
    e = tmpl.expand(gen["x_mine"])
    batch = types.SimpleNamespace(S=tmpl.batch.S, Z=tmpl.batch.Z, **e)
    n_int = batch.K * batch.M * tmpl.static.Z * tmpl.static.sumX
    with torch.cuda.stream(env.stream), ctx.lock:
        pinned_int = torch.empty(n_int, dtype=torch.float64).pin_memory().numpy()
        pinned_res = torch.empty(batch.K * (batch.S + 1), dtype=torch.float64).pin_memory().numpy()
        ctx.set_static(tmpl.static)
        ms = env.time_wall(lambda: ctx.evaluate_host(batch, out_residuals=pinned_res, out_intensities=pinned_int),
                           args.steps, min(args.warmup, 2))
    h2d = sum(v.nbytes for v in e.values() if isinstance(v, np.ndarray)) + 20 * e["job_phase"].size
    return dict(ms=ms / args.steps, h2d=int(h2d), d2h=int(pinned_int.nbytes + pinned_res.nbytes))


def measure_sweep(env, name="c5", totals=SWEEP_TOTALS, steps=2):
    """BASELINE configs[4]: populations of ``totals`` trials in TOTAL split over the ranks (strong scaling), through the
    product's sharded call, optimised residuals gathered on every rank"""
    import pyxrd_b200.calculations as calc
    from pyxrd_b200 import distributed
    torch, ctx = env.torch, env.ctx
    tmpl, refs = make_template(name, calc)
    S = tmpl.batch.S
    rows = []
    with torch.cuda.stream(env.stream), ctx.lock:
        for total in totals:
            x = syn.sample_solutions(refs, SEED + 1, total)
            table = [None]

            def step():
                table[0] = distributed.evaluate_solutions_sharded(
                    x, lambda xb: tmpl.optimized_residuals(xb, ctx=ctx, device=True), device=env.device, columns=S + 1,
                    ctx=ctx if (env.world > 1 and env.args.gather == "peer") else None)

            ms = env.time_wall(step, steps, 1)
            check = None
            if env.world > 1 and env.rank == 0:
                ctx.set_option("optimizer_threads", optimiser_threads(-(-total // env.world)))
                alone = tmpl.optimized_residuals(x, ctx=ctx, device=True).cpu().numpy()
                ctx.set_option("optimizer_threads", 0)
                check = bool(np.array_equal(alone, table[0]))
                if not check:
                    raise AssertionError("c5 sweep K=%d: gathered table differs from the single-GPU table" % total)
            ms, = env.max_over_ranks([ms])
            rows.append({"trials_total": int(total), "ms_per_generation": ms / steps,
                         "trials_per_s": total / (ms / steps * 1e-3), "finite": bool(np.isfinite(table[0]).all()),
                         "gather_matches_single_gpu": check})
    return {"workload": "%s: %s" % (name, syn.CONFIGS[name].description), "scaling": "strong", "refinables": int(x.shape[1]),
            "call": "distributed.evaluate_solutions_sharded(x[K, d] on the host -> optimised residual table on the host of "
                    "every rank), NCCL gather inside", "rows": rows}


def measure_packed_config(env, name, count, peak_tf):
    """device-resident pattern throughput of one configuration from packed candidate graphs (c3 / c3s: matrices no
    probability model produces, so no template)"""
    import pyxrd_b200.calculations as calc
    from pyxrd_b200.batched import pack_population
    torch, ctx, args = env.torch, env.ctx, env.args
    cfg = syn.CONFIGS[name]
    observed = observed_for(name, calc)
    mixes = cfg.population(SEED, count)
    for m in mixes:
        for spec, obs in zip(m.specimens, observed):
            spec.observed_intensity = obs.copy()
    static, batch = pack_population(mixes)
    flops = points = 0
    for m in mixes:
        f, p = mixture_work(m)
        flops += f
        points += p
    with torch.cuda.stream(env.stream), ctx.lock:
        ctx.set_static(static, force=True)
        ctx.set_batch(batch)
        dev_ms, stage, _ = env.time_device(lambda: ctx.compute_all(False), max(2, min(args.steps, 4)), 2, stages=("phase",))
    steps = max(2, min(args.steps, 4))
    return config_row(name, count, points, flops, None, dev_ms / steps, stage["phase"] / steps, peak_tf, "packed graphs")


def config_row(name, count, points, flops, flops_exe, dev_ms, phase_ms, peak_tf, path):
    alg = flops / (phase_ms * 1e-3) * 1e-12 / peak_tf if phase_ms > 0 else None
    exe = flops_exe / (phase_ms * 1e-3) * 1e-12 / peak_tf if (phase_ms > 0 and flops_exe is not None) else None
    # frac: executed count where the template path knows it; packed graphs (c3 / c3s: one plane per atom group, no shared
    # planes worth mentioning next to a rank-27 series): the algorithmic count
    return {"workload": "%s: %s" % (name, syn.CONFIGS[name].description), "candidates": count, "path": path,
            "value": points / (dev_ms * 1e-3), "unit": "pattern-pts/s", "ms_per_step": dev_ms, "kernel_ms": phase_ms,
            "frac": exe if exe is not None else alg, "frac_algorithmic": alg}


# ---------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------
def cpu_functions():
    """(calculate_mixture, get_optimized_residual, kind): the unmodified reference from oracle/_ref when it travelled
    with the snapshot (kind "reference"), else the numpy port of the oracle (kind "port")"""
    from oracle import reference_arm
    ref = reference_arm.load()
    if ref is not None:
        return ref.calculate_mixture, ref.get_optimized_residual, "reference"
    from oracle import pyxrd_oracle as orc
    return orc.calculate_mixture, orc.get_optimized_residual, "port"


_CPU_OBSERVED = {}


def _cpu_candidate(name, seed, idx, optimised):
    """candidate idx of the population with 'observed' patterns; for trials the observed pattern is the truth's own
    (CPU-calculated once per process) pattern + noise, as on the GPU arm"""
    m = syn.CONFIGS[name].candidate(seed, idx)
    if optimised:
        obs = _CPU_OBSERVED.get(name)
        if obs is None:
            calc_mix = cpu_functions()[0]
            truth = syn.CONFIGS[name].truth()

            def calc_total(mix):
                calc_mix(mix)
                return [np.array(s.total_intensity) for s in mix.specimens]

            for spec in truth.specimens:
                spec.observed_intensity = np.zeros((spec.range_theta.size, 1))
            syn.attach_observed([], truth, calc_total)
            obs = _CPU_OBSERVED[name] = [np.array(s.observed_intensity) for s in truth.specimens]
        for spec, o in zip(m.specimens, obs):
            spec.observed_intensity = o.copy()
    else:
        rng = np.random.default_rng(1)
        for spec in m.specimens:
            spec.observed_intensity = rng.uniform(20.0, 60.0, (spec.range_theta.size, 1))
    return m


def _single_threaded_blas():
    """one BLAS / OpenMP thread per process: the reference's Pool runs one trial per worker (pyxrd_server.py:57-76), and
    scipy's L-BFGS-B would otherwise start a thread team in every worker (16 workers x 16 threads on 16 cores: measured
    no faster than ONE process)"""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass


def _cpu_one(arg):
    name, seed, idx, optimised = arg
    calc_mix, opt_res, _ = cpu_functions()
    m = _cpu_candidate(name, seed, idx, optimised)
    if optimised:
        return float(np.ravel(opt_res(m))[0])
    calc_mix(m)
    return float(m.residuals[0])


def cpu_sample(name, budget_s, processes, optimised=False):
    """the CPU implementation on a bounded sample of the workload; processes > 1: a multiprocessing Pool mapping the
    per-trial callback over candidates, as pyxrd/server/pyxrd_server.py:57-76 does"""
    kind = cpu_functions()[2]
    pts = mixture_work(syn.CONFIGS[name].truth())[1]
    done = 0
    _single_threaded_blas()
    if processes <= 1:
        _cpu_one((name, SEED, 0, optimised))            # imports, observed pattern
        t0 = time.perf_counter()
        while True:
            _cpu_one((name, SEED, done, optimised))
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp
        _cpu_candidate(name, SEED, 0, optimised)          # the observed pattern is computed once, before the fork
        with mp.get_context("fork").Pool(processes, initializer=_single_threaded_blas) as pool:
            pool.map(_cpu_one, [(name, SEED, i, optimised) for i in range(processes)])      # warm every worker
            t0 = time.perf_counter()
            pos = processes
            while time.perf_counter() - t0 < budget_s:
                chunk = processes * (1 if optimised else 2)
                pool.map(_cpu_one, [(name, SEED, pos + i, optimised) for i in range(chunk)])
                pos += chunk
                done += chunk
            dt = time.perf_counter() - t0
    return dict(candidates=done, seconds=dt, pts_per_s=done * pts / dt, trials_per_s=done / dt, kind=kind)


def run_reference(args, rank):
    """the reference arm: the reference's own CPU implementation of the path on all host cores (oracle/_ref = the
    unmodified pyxrd.calculations when it travelled with the snapshot, else the oracle port), same workload, metric,
    unit; rank 0 alone runs it"""
    if rank != 0:
        return
    name = "c2"
    cores = os.cpu_count() or 1
    steps = []
    for i in range(args.warmup + args.steps):
        budget = max(3.0, 30.0 / max(args.steps, 1)) if i >= args.warmup else 1.0
        r = cpu_sample(name, budget_s=budget, processes=cores)
        if i >= args.warmup:
            steps.append(r)
    pts1 = mixture_work(syn.CONFIGS[name].truth())[1]
    secs = sum(r["seconds"] for r in steps)
    done = sum(r["candidates"] for r in steps)
    value = done * pts1 / secs
    kind = steps[0]["kind"]
    tr = cpu_sample("c4", budget_s=25.0, processes=cores, optimised=True)
    what = ("the unmodified pyxrd.calculations (oracle/_ref)" if kind == "reference" else "the oracle port (numpy)")
    line = {"impl": "reference", "metric": "pattern-pts/s", "value": value, "unit": "pattern-pts/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / max(len(steps), 1) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main_config(name, DEFAULT_CANDIDATES[name]),
            "cpu_baseline": {"value": value, "unit": "pattern-pts/s", "cores": cores, "kind": kind,
                             "sample": "%d candidates per step through calculate_mixture of %s in a %d-process pool "
                                       "(multiprocessing, as pyxrd/server/pyxrd_server.py:57-76 does)"
                                       % (steps[0]["candidates"], what, cores)},
            "e2e": {"value": value, "unit": "pattern-pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "trials": {"metric": "refinement trial evals/s (get_optimized_residual)", "value": tr["trials_per_s"],
                       "unit": "trials/s", "cores": cores, "kind": tr["kind"],
                       "config": {"workload": "c4: %s" % syn.CONFIGS["c4"].description},
                       "sample": "%d c4 trials through get_optimized_residual of %s (scipy L-BFGS-B per trial) in a "
                                 "%d-process pool, %.1f s" % (tr["candidates"], what, cores, tr["seconds"])}}
    print(json.dumps(line), flush=True)


def main_config(name, per_gpu):
    cfg = syn.CONFIGS[name]
    return {"workload": "%s: %s" % (name, cfg.description), "candidates_per_gpu": per_gpu,
            "points_per_pattern": [int(sp.range_theta.size) for sp in cfg.truth().specimens],
            "specimens": cfg.n_specimens, "phases": cfg.n_phases, "l2": L2_NOTE, "population_seed": SEED,
            "step": "one generation: x[K, d] -> device expansion + prologue -> phase patterns -> wavelength distribution "
                    "-> residuals (-> NCCL all-gather of f64[K, 1+S] at N > 1)"}


# ---------------------------------------------------------------------------------------
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
    env = Env(args, rank, world, local_rank)
    with torch.cuda.stream(env.stream):
        peak_tf, _ = env.ctx.dfma_peak(8192)
    sampler = ClockSampler(local_rank)
    sampler.start()
    name = args.workload
    main = measure_generation(env, name, args.candidates or DEFAULT_CANDIDATES[name], optimise=False)
    trials = None
    if not args.no_trials:
        tw = args.trials_workload
        trials = measure_generation(env, tw, args.trials_candidates or DEFAULT_CANDIDATES[tw], optimise=True)
    sweep = None if args.no_sweep else measure_sweep(env)
    clocks = sampler.stop()            # sampled over every timed region above (headline, trials, population sweep)
    download = configs = None
    if world == 1:
        download = measure_patterns_download(env, main)
        if not args.no_configs:
            configs = []
            for cname in ("c1", "c3", "c3s", "c4", "c5"):
                if cname in ("c3", "c3s"):
                    configs.append(measure_packed_config(env, cname, DEFAULT_CANDIDATES[cname], peak_tf))
                else:
                    g = measure_generation(env, cname, DEFAULT_CANDIDATES[cname], optimise=False, verify=False)
                    configs.append(config_row(cname, g["K"], g["points"], g["flops"], g["flops_executed"], g["dev_ms"],
                                              g["stage"]["phase"], peak_tf, "template + binding table"))
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
    phase_ms = main["stage"]["phase"]
    achieved_tf = main["flops"] / (phase_ms * 1e-3) * 1e-12
    executed_tf = main["flops_executed"] / (phase_ms * 1e-3) * 1e-12
    traffic = None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[name]
        traffic = rec["dram_bytes"] if rec.get("candidates") == main["K"] else None
    except Exception:
        pass
    call = ("pyxrd_b200.distributed.evaluate_solutions_sharded: x[K_total, d] on the host -> every rank's block through "
            "PopulationTemplate.residuals (pyxrd_update_population + pyxrd_compute_all) -> all-gather from the device "
            "buffer -> residual table f64[K_total, 1+S] on the host of every rank")
    line = {
        "metric": "pattern-pts/s", "value": total_pts / (main["dev_ms"] * 1e-3), "unit": "pattern-pts/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["dev_ms"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": main_config(name, main["K"]),
        "e2e": {"value": total_pts / (main["e2e_ms"] * 1e-3), "unit": "pattern-pts/s", "h2d_bytes_per_step": main["h2d"],
                "d2h_bytes_per_step": main["d2h"], "ms_per_step": main["e2e_ms"], "call": call,
                "evals_per_s": main["K"] * world / (main["e2e_ms"] * 1e-3), "refinables": main["d"],
                "gather_matches_single_gpu": main["gather_matches_single_gpu"], "gather": main["gather"]},
        "gpu_launches": main["launches"],
        "roofline": {"bound": "fp64", "kernel": "k_phase_* (phase-intensity kernels)", "achieved": executed_tf,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": executed_tf / peak_tf if peak_tf else None,
                     "count": "flops the kernels execute (DESIGN.md section 3): series + (32 + 8) per complex exponential "
                              "evaluated (one per distinct interlayer plane of a phase; candidate-invariant layer stacks "
                              "are read from the layer table of the batch) + 2 per atom + 64 per component + 120 (LPF, "
                              "corrections)",
                     "algorithmic": {"achieved": achieved_tf, "frac": achieved_tf / peak_tf if peak_tf else None,
                                     "count": "SURVEY 8(d): one complex exponential per ATOM (32 + 8 + 2 each); exceeds 1 "
                                              "where the kernels no longer evaluate them"},
                     "peak_source": "DFMA microbenchmark pyxrd_dfma_peak measured in this run (MEASURED_PEAKS.json "
                                    "has no fp64 entry)",
                     "kernel_ms": phase_ms, "flops_per_launch": main["flops_executed"],
                     "algorithmic_flops_per_launch": main["flops"], "traffic": traffic,
                     "hbm": {"achieved_gbs": 16.0 * main["points"] / (phase_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                             "frac": 16.0 * main["points"] / (phase_ms * 1e-3) * 1e-9 / hbm_peak,
                             "peak_source": "of measured" if peaks else "of fallback"}},
        "stage_ms": main["stage"], "clocks": clocks,
        "cpu_affinity": "NVML ideal affinity of the rank's GPU" if numa_bound else "unchanged",
    }
    if download is not None:
        line["patterns_e2e"] = {"call": "pyxrd_evaluate_host (N = 1 only): packed parameter blocks of the K candidates on "
                                        "the host in, ALL phase patterns + residuals in pinned host memory out",
                                "value": total_pts / (download["ms"] * 1e-3), "unit": "pattern-pts/s",
                                "ms_per_step": download["ms"], "h2d_bytes_per_step": download["h2d"],
                                "d2h_bytes_per_step": download["d2h"]}
    if trials is not None:
        t = trials
        passes = 33
        opt_ms = t["stage"]["residual"]
        opt_flops = t["K"] * optimiser_flops(t["M"], t["selected_points"], passes)
        opt_tf = opt_flops / (opt_ms * 1e-3) * 1e-12 if opt_ms > 0 else 0.0
        ph_tf = t["flops_executed"] / (t["stage"]["phase"] * 1e-3) * 1e-12 if t["stage"]["phase"] > 0 else 0.0
        line["trials"] = {
            "metric": "refinement trial evals/s (get_optimized_residual semantics: all phase patterns of the trial + "
                      "optimisation of fractions / scales / background + residual)",
            "value": t["K"] * world / (t["dev_ms"] * 1e-3), "unit": "trials/s", "ms_per_step": t["dev_ms"],
            "e2e": {"value": t["K"] * world / (t["e2e_ms"] * 1e-3), "unit": "trials/s", "ms_per_step": t["e2e_ms"],
                    "h2d_bytes_per_step": t["h2d"], "d2h_bytes_per_step": t["d2h"],
                    "gather_matches_single_gpu": t["gather_matches_single_gpu"],
                    "call": "distributed.evaluate_solutions_sharded over PopulationTemplate.optimized_residuals"},
            "config": {"workload": "%s: %s" % (t["name"], t["cfg"].description), "candidates_per_gpu": t["K"],
                       "refinables": t["d"], "l2": L2_NOTE},
            "stage_ms": {"prologue": t["stage"]["prologue"], "phase": t["stage"]["phase"],
                         "wavelength": t["stage"]["wavelength"], "optimiser": opt_ms},
            "roofline": {"bound": "fp64", "kernel": "k_mixture_optimize", "achieved": opt_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": opt_tf / peak_tf if peak_tf else None, "kernel_ms": opt_ms,
                         "algorithmic_flops_per_launch": opt_flops,
                         "count": "%d Gram passes x selected points x (2 NG + 4 MB + 4), MB = M + 1, NG = MB (MB + 1) / 2; "
                                  "the kernel is one latency chain per candidate (DESIGN.md section 4)" % passes,
                         "phase": {"kernel": "k_phase_*", "achieved": ph_tf, "frac": ph_tf / peak_tf if peak_tf else None,
                                   "kernel_ms": t["stage"]["phase"], "count": "executed flops, as the headline roofline"}},
            "optimiser": "device majorise-minimise (pyxrd_optimize), 32 reweightings x <= 2 Newton steps; contract: <= "
                         "reference L-BFGS-B optimum * (1 + 5e-3)",
            "gpu_launches": t["launches"], "pattern_pts_per_s": t["points"] * world / (t["dev_ms"] * 1e-3)}
    if sweep is not None:
        line["sweep"] = sweep
    if configs is not None:
        line["configs"] = configs
    if world == 1 and not args.no_cpu:
        c = cpu_sample(name, budget_s=12.0, processes=1)
        what = "the unmodified pyxrd.calculations (oracle/_ref)" if c["kind"] == "reference" else "the oracle port"
        line["cpu_baseline"] = {"value": c["pts_per_s"], "unit": "pattern-pts/s", "cores": 1, "kind": c["kind"],
                                "sample": "%d candidates of %s through calculate_mixture of %s, single process, %.1f s"
                                          % (c["candidates"], name, what, c["seconds"])}
        if trials is not None:
            ct = cpu_sample(args.trials_workload, budget_s=10.0, processes=1, optimised=True)
            line["trials"]["cpu_baseline"] = {"value": ct["trials_per_s"], "unit": "trials/s", "cores": 1, "kind": ct["kind"],
                                              "sample": "%d %s trials through get_optimized_residual of %s (scipy L-BFGS-B), "
                                                        "single process, %.1f s" % (ct["candidates"], args.trials_workload,
                                                                                     what, ct["seconds"])}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c4", "c5"],
                    help="configuration of the headline measurement (BASELINE configs[1] by default)")
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU (default per workload)")
    ap.add_argument("--no-trials", action="store_true")
    ap.add_argument("--trials-workload", default="c4", choices=["c1", "c2", "c4", "c5"],
                    help="configuration of the trial-evaluation measurement (BASELINE configs[3] by default)")
    ap.add_argument("--trials-candidates", type=int, default=0, help="trials per GPU of that measurement")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] population sweep")
    ap.add_argument("--no-configs", action="store_true", help="skip the table of the other configurations (N = 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE",
                    help="pyxrd_set_option applied to the context before the measurements (experiments; repeatable)")
    ap.add_argument("--gather", default="nccl", choices=["nccl", "peer"],
                    help="exchange of the residual rows at N > 1: one NCCL all-gather (default), or the library's stores into "
                         "peer memory + a symmetric-memory barrier (measured equal at 2 GPUs, 2 %% slower at 8)")
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
