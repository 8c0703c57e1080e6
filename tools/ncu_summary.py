#!/usr/bin/env python
"""Summarise an .ncu-rep for profiles/: headline metrics, pipe mix, stall reasons, hot source lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--lines 25]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors.sum",
    "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units = rows[0], rows[1]
    for data in rows[2:]:
        rec = dict(zip(head, zip(units, data)))
        print("## %s" % rec.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in rec:
                print("%-86s %14s %s" % (k, rec[k][1], rec[k][0]))
        stalls = [(float(v[1]), k) for k, v in rec.items()
                  if k.startswith("smsp__average_warp") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and v[1] not in ("", "n/a")]
        stalls.sort(reverse=True)
        print("-- warp latency per issued instruction by stall reason (cycles):")
        for v, k in stalls[:8]:
            print("   %6.2f  %s" % (v, k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    hdr = [i for i, r in enumerate(src) if r and r[0] == "Line No"]
    if not hdr:
        return
    h = src[hdr[0]]
    sec = src[hdr[0] + 1: (hdr[1] - 2) if len(hdr) > 1 else None]
    ii, isamp = h.index("Instructions Executed"), h.index("# Samples")

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    by_line = collections.defaultdict(lambda: [0.0, 0.0, ""])
    ops = collections.Counter()
    for r in sec:
        if len(r) <= ii:
            continue
        e = by_line[r[0]]
        e[0] += f(r[ii]); e[1] += f(r[isamp]); e[2] = r[1][:100]
        toks = r[3].split()
        if toks:
            name = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[name.split(".")[0]] += f(r[ii])
    ti = sum(v[0] for v in by_line.values()) or 1.0
    ts = sum(v[1] for v in by_line.values()) or 1.0
    print("-- hot source lines (share of executed warp instructions / of stall samples):")
    for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:nlines]:
        print("   %5.1f%% %5.1f%%  L%-4s %s" % (100 * v[0] / ti, 100 * v[1] / ts, k, v[2]))
    print("-- executed warp instructions by opcode:")
    print("   " + "  ".join("%s %.1f%%" % (k, 100 * v / ti) for k, v in ops.most_common(16)))


if __name__ == "__main__":
    main()
