#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full capture of sweep_kernel) into the text summary committed under
profiles/: headline counters, stall reasons and the executed instruction mix by SASS opcode.
    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    print(f"# ncu summary of {rep}")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}")
        for k in keys:
            if k in idx:
                print(f"{k:70s} {r[idx[k]]} {units[idx[k]]}")
        stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        vals = sorted(((float(r[idx[h]].replace(",", "")), h) for h in stall), reverse=True)
        print("stall reasons (warps stalled per issue-active cycle): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in vals[:8]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
            continue
        if cur is not None:
            cur.append(r)
    if blocks:
        b = blocks[0]
        hdr, data = b[0], b[1:]
        ix = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        samp = sum(int(r[ix["# Samples"]]) for r in data) or 1
        op, ops = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            name = m.group(2).split(".")[0] if m else "?"
            if name == "IMAD" and "IMAD.MOV" in r[ix["Source"]]:
                name = "IMAD.MOV"
            op[name] += int(r[ix["Instructions Executed"]])
            ops[name] += int(r[ix["# Samples"]])
        print(f"\n## executed instruction mix of launch 0 ({tot} warp instructions, {len(data)} SASS instructions in the kernel)")
        for k, v in op.most_common(20):
            print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}% of instructions   {100 * ops[k] / samp:5.1f}% of stall samples")


if __name__ == "__main__":
    main()
