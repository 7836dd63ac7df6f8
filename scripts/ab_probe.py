"""A/B timing of builds of libqsx.so on the same box: each variant in its own process, interleaved.
    python scripts/ab_probe.py name=path/to/libqsx.so ...      (needs a B200)"""
import json
import os
import subprocess
import sys

CHILD = r'''
import json, sys
sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W
out = {}
for name, n, text in (("c2", 12, W.random_clifford_t_qasm(12, 2000, 20261019)), ("ct14", 14, W.random_clifford_t_qasm(14, 1000, 20261019)),
                      ("qft14", 14, W.qft_qasm(14))):
    qc = host.Circuit(qasm_text=text)
    plan = Q.Plan(n, 2**n, qc.gate_array())
    u = Q.Unitary(n)
    plan.run(u); plan.run(u)
    ms = sorted(plan.run(u).device_ms for _ in range(9))
    out[name] = [round(ms[0], 4), round(ms[4], 4)]
    u.close()
print(json.dumps(out))
'''


def main():
    variants = [a.split("=", 1) for a in sys.argv[1:]]
    for rep in range(3):
        for name, path in variants:
            env = dict(os.environ, QSX_LIB=path)
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
            print(json.dumps({"variant": name, "rep": rep, "min_median_ms": json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-300:]}), flush=True)


if __name__ == "__main__":
    main()
