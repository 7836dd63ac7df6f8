"""Offline (no GPU) evaluation of scheduler options with the measured cost model of the sweep
kernel at n = 12 (profiles/r1_microbench_n12.jsonl): sweep 95 us, extra stage 34 us, micro-op
3.3 us, folded permutation 0.37 us; all scale with 4^(n-12).
    python scripts/plan_cost.py '{"min_tail_ops": 8}' '{"search": 2}' ..."""
import json
import sys

sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W


def model_us(st, n):
    us = 95 * st.n_sweeps + 34 * (st.n_stages - st.n_sweeps) + 3.3 * st.n_micro_ops + 0.37 * st.n_folded_perms
    return us * 4.0 ** (n - 12)


def main():
    variants = [json.loads(a) for a in sys.argv[1:]] or [{}]
    cases = [("c2", W.random_clifford_t_qasm(12, 2000, 20261019), 12), ("rot12", W.random_rotation_qasm(12, 2000, 20261019), 12),
             ("c3", W.random_clifford_t_qasm(14, 1000, 20261019), 14), ("ct16", W.random_clifford_t_qasm(16, 1000, 7), 16),
             ("qft14", W.qft_qasm(14), 14)]
    for name, text, n in cases:
        qc = host.Circuit(qasm_text=text)
        ga = qc.gate_array()
        for v in variants:
            plan = Q.Plan(n, 2**n, ga, Q.PlanOptions(**v))
            st = plan.stats()
            print(f"{name:6s} {json.dumps(v):40s} sweeps {st.n_sweeps:3d} stages {st.n_stages:3d} uops {st.n_micro_ops:4d} folded {st.n_folded_perms:3d}"
                  f" plan {st.plan_host_ms:6.2f} ms  model {model_us(st, n) / 1e3:8.3f} ms")


if __name__ == "__main__":
    main()
