#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + opcode mix + hot SASS regions.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--regions]
Writes plain text to stdout; the per-round copies live under profiles/.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        name = vals[hdr.index("Kernel Name")]
        print(f"== launch {k}: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print(f"  {h:75s} {v:>16s} {u}")
    rows = ncu(rep, "source")
    hdr = rows[1]
    data = rows[2:]
    ia, ie, it, isamp = (hdr.index(x) for x in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    byop, byop_t, samp = collections.Counter(), collections.Counter(), collections.Counter()
    stalls = collections.Counter()
    recs = []
    tot = 0
    for r in data:
        try:
            n, t, s = int(r[ie]), int(r[it]), int(r[isamp])
        except (ValueError, IndexError):
            continue
        toks = r[ia].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += n
        byop_t[op] += t
        samp[op] += s
        tot += n
        recs.append((n, t, s, r[ia]))
        for i in stall_cols:
            try:
                stalls[hdr[i]] += int(r[i])
            except ValueError:
                pass
    print(f"\n== opcode mix (warp instructions executed, total {tot/1e6:.1f} M)")
    for op, n in byop.most_common(22):
        print(f"  {op:10s} {n/1e6:9.1f} M {100*n/tot:5.1f} %  avg lanes {byop_t[op]/max(n,1):5.1f}  samples {samp[op]}")
    tt = sum(byop_t.values())
    print(f"  average active lanes per instruction: {tt/max(tot,1):.2f}")
    fp64 = sum(byop[o] for o in ("DFMA", "DMUL", "DADD"))
    print(f"  fp64 arithmetic share of issued instructions: {100*fp64/tot:.1f} %")
    print("\n== warp stall samples")
    ts = sum(stalls.values())
    for k, v in stalls.most_common(10):
        print(f"  {k:28s} {v:9d} {100*v/max(ts,1):5.1f} %")
    if "--regions" in sys.argv:
        print("\n== SASS regions with equal execution count (share of issued instructions > 0.4 %)")
        prev, start, acc_t, acc_s = None, 0, 0, 0
        out = []
        for i, (n, t, s, _) in enumerate(recs):
            if prev is None:
                prev, start = n, i
            if n != prev:
                out.append((start, i - 1, prev, acc_t, acc_s))
                prev, start, acc_t, acc_s = n, i, 0, 0
            acc_t += t
            acc_s += s
        out.append((start, len(recs) - 1, prev, acc_t, acc_s))
        for a, b, n, t, s in out:
            if (b - a + 1) * n > 0.004 * tot:
                print(f"  inst[{a:4d}-{b:4d}] len {b-a+1:4d} exec {n/1e6:8.2f} M share {100*(b-a+1)*n/tot:5.1f} % "
                      f"lanes {t/max((b-a+1)*n,1):5.1f} samples {s}")


if __name__ == "__main__":
    main()
