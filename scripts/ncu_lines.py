#!/usr/bin/env python
"""Per-source-line view of an ncu --set full --import-source on capture: stall samples and executed
warp instructions aggregated by CUDA-C line of the first kernel block (needs -lineinfo).
    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    h = rows[hi[0]]
    end = hi[1] - 2 if len(hi) > 1 else len(rows)
    ix = {n: i for i, n in enumerate(h)}
    num = lambda r, k: int(r[ix[k]]) if r[ix[k]].isdigit() else 0
    data = [r for r in rows[hi[0] + 1:end] if r and r[0] != ""]
    tot = sum(num(r, "# Samples") for r in data) or 1
    toti = sum(num(r, "Instructions Executed") for r in data) or 1
    print(f"# {rep}: {len(data)} source lines, {tot} stall samples, {toti} warp instructions (launch 0)")
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    data.sort(key=lambda r: -num(r, "# Samples"))
    for r in data[:top_n]:
        top = sorted(((num(r, k), k[6:]) for k in stalls), reverse=True)[:3]
        tops = " ".join(f"{k}={100 * v / tot:.1f}" for v, k in top if v)
        print(f"{r[0]:>4} {100 * num(r, '# Samples') / tot:5.1f}% samples {100 * num(r, 'Instructions Executed') / toti:5.1f}% instr  [{tops}]  {r[1].strip()[:100]}")


if __name__ == "__main__":
    main()
