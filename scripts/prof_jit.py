"""Fixed workload for ncu: one BASELINE workload (default C4), library-default schedule, specialised sweeps compiled
before the first run (jit = 2), operand A built twice.
    python scripts/prof_jit.py c4 && ncu --set full --clock-control none --import-source on -k regex:qsx_sweep -s <sweeps> -c 2 -o gpurun_out/prof python scripts/prof_jit.py c4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsyn_b200 as Q
from qsyn_b200 import host, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
jit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, desc, text = workloads.workload_qasm(name)
qc = host.Circuit(qasm_text=text)
plan = Q.Plan(n, 2**n, qc.gate_array(), Q.PlanOptions(jit=jit))
u = Q.Unitary(n)
for _ in range(2):
    r = plan.run(u)
print("ok", desc, "sweeps", plan.stats().n_sweeps, "launches", r.launches, "specialised", r.jit_launches, "ms", round(r.device_ms, 3))
