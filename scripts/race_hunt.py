"""Determinism check: runs one plan repeatedly from identity and reports how many distinct
results appear (must be 1), for a few option sets."""
import hashlib
import json
import sys

sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = int(sys.argv[2]) if len(sys.argv) > 2 else 600
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 123
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 12
qc = host.Circuit(qasm_text=W.random_clifford_t_qasm(n, g, seed))
ga = qc.gate_array()
variants = [{}, {"tile_qubits": 9}, {"no_perm_folding": 1}, {"no_diag_conjugation": 1}, {"min_tail_ops": -1}, {"fuse": 0},
            {"max_ops_per_sweep": 8}, {"log2_tile_cols": 4}, {"tile_qubits": 7}, {"reg_qubits": 3}]
for v in variants:
    plan = Q.Plan(n, 2**n, ga, Q.PlanOptions(**v))
    seen = {}
    for _ in range(reps):
        u = Q.Unitary(n)
        plan.run(u)
        h = hashlib.sha1(u.download().tobytes()).hexdigest()[:10]
        seen[h] = seen.get(h, 0) + 1
        u.close()
    st = plan.stats()
    print(json.dumps({"opt": v, "sweeps": st.n_sweeps, "stages": st.n_stages, "distinct": len(seen), "hist": sorted(seen.values(), reverse=True)}), flush=True)
