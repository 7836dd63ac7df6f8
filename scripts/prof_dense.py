"""Fixed workload for ncu: n = 12 random Clifford+T folded into dense k-qubit blocks (dense_kernel<k>), run twice.
    python scripts/prof_dense.py 5 && ncu --set full ... -k regex:dense_kernel -s 20 -c 2 -o gpurun_out/prof python scripts/prof_dense.py 5"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsyn_b200 as Q
from qsyn_b200 import host, workloads

k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = 12
qc = host.Circuit(qasm_text=workloads.random_clifford_t_qasm(n, 400, 20261019))
plan = Q.Plan(n, 2**n, qc.gate_array(), Q.PlanOptions(dense_block=k))
u = Q.Unitary(n)
for _ in range(2):
    r = plan.run(u)
st = plan.stats()
print("ok dense", k, "sweeps", st.n_sweeps, "ms", round(r.device_ms, 3), "us/sweep", round(r.device_ms * 1e3 / st.n_sweeps, 2),
      "TFLOP/s", round(8 * 2**k * 4.0**n * st.n_sweeps / (r.device_ms * 1e-3) / 1e12, 2))
