"""Offline (no GPU) evaluation of scheduler options with the measured cost model of the sweep
kernel at n = 12 (profiles/r1_microbench_n12.jsonl); everything scales with 4^(n-12).
    python scripts/plan_cost.py '{"min_tail_ops": 8}' '{"search": 2}' ..."""
import json
import sys

sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W


def model_us(plan, st, n):
    """max(memory floor, fixed + micro-ops + stage transitions) per sweep (light instantiation)."""
    cost = {"H": 3.8, "U1": 8.1, "X": 3.6, "XC": 2.7, "PHJ": 3.4, "DTABH": 3.4, "DTABF": 4.5, "SCALN": 3.0, "SWAP": 3.6, "AD": 6.0, "U2": 20.0}
    us = 0.0
    for sw in plan.dump()["sweeps"]:
        c = 32.0 + 23.0 * (len(sw["stages"]) - 1)
        for stg in sw["stages"]:
            c += 0.13 * (len(stg["pre"]) + len(stg["post"]))
            for u in stg["uops"]:
                k = cost.get(u["kind"], 4.0)
                if u["kind"] == "X" and (u["cother"] or u["czero"]):
                    k = 2.4
                c += k
        us += max(90.0, c + 12.0)  # +12: the part of the memory time that never hides
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
                  f" plan {st.plan_host_ms:6.2f} ms  model {model_us(plan, st, n) / 1e3:8.3f} ms")


if __name__ == "__main__":
    main()
