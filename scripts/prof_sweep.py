"""Small fixed workload for ncu: n=12, 2000-gate Clifford+T (BASELINE config C2), default schedule, run twice."""
import sys

sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
g = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
fuse = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dense = int(sys.argv[4]) if len(sys.argv) > 4 else 0
qc = host.Circuit(qasm_text=workloads.random_clifford_t_qasm(n, g, 20261019))
plan = Q.Plan(n, 2**n, qc.gate_array(), Q.PlanOptions(fuse=fuse, dense_block=dense))
u = Q.Unitary(n)
for _ in range(2):
    r = plan.run(u)
print("ok", plan.stats().n_sweeps, r.launches, round(r.device_ms, 3))
