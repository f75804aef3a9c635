#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full --import-source on) into a small committed text summary:
key metrics, per-segment instruction / stall-sample shares, top stalled SASS lines.
usage: ncu_summary.py report.ncu-rep out.md [title] [kernel-name-regex]   (regex: pick one kernel of a multi-kernel report)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__maximum_warps_avg_per_active_cycle", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_xu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


FILTER = []


def run(args):
    return subprocess.run(["ncu", "-i"] + args + FILTER, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    if len(sys.argv) > 4:
        FILTER.extend(["--kernel-name", "regex:" + sys.argv[4]])
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, val = raw[0], raw[1], raw[2]
    lines = [f"# {title}", "", f"kernel: `{val[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'}`", "",
             "| metric | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"| {w} | {val[i]} | {units[i]} |")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    h = src[1]
    iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    data = [r for r in src[2:] if len(r) > max(iS, iN, iE) and r[iE].strip().lstrip('-').isdigit() and r[iN].strip().isdigit()]
    segs, cur = [], None
    for k, r in enumerate(data):
        e, s = int(r[iE]), int(r[iN])
        if cur is None or not (0.7 * cur["e0"] <= e <= 1.4 * cur["e0"]):
            cur = dict(start=k, e0=max(e, 1), inst=0, samp=0, n=0)
            segs.append(cur)
        cur["inst"] += e; cur["samp"] += s; cur["n"] += 1; cur["end"] = k
    te, ts = sum(s["inst"] for s in segs), max(1, sum(s["samp"] for s in segs))
    lines += ["", f"SASS segments by execution-count plateau (total {te} warp instructions, {ts} stall samples):", "",
              "| SASS lines | #instr | executions each | % of instructions | % of stall samples | first instruction |", "|---|---|---|---|---|---|"]
    for s in segs:
        if s["inst"] > 0.01 * te or s["samp"] > 0.01 * ts:
            lines.append(f"| {s['start']}-{s['end']} | {s['n']} | {s['e0']} | {100 * s['inst'] / te:.1f} | {100 * s['samp'] / ts:.1f} | `{data[s['start']][iS].strip()[:48]}` |")
    lines += ["", "Top stalled instructions:", "", "| line | samples | % | executions | SASS |", "|---|---|---|---|---|"]
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][iN]))[:12]:
        lines.append(f"| {k} | {r[iN]} | {100 * int(r[iN]) / ts:.1f} | {r[iE]} | `{r[iS].strip()[:64]}` |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


main()
