"""Column blocking through the L2 (qsx_plan_options.l2_block_mib) on memory-bound and compute-bound plans.
    python scripts/l2_block_probe.py > gpurun_out/l2_block_probe.jsonl      (needs a B200)"""
import json
import sys

sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W


def main():
    cases, keep = [], []
    for n in (12, 14):
        text = W.random_clifford_t_qasm(n, 2000 if n == 12 else 400, 20261019)
        full, part = host.Circuit(qasm_text=text), host.Circuit(qasm_text="\n".join(text.splitlines()[:3 + 120]) + "\n")
        keep += [full, part]  # the gate arrays point into the circuits
        ga, head = full.gate_array(), part.gate_array()
        cases += [(f"n{n}_one_gate_per_sweep_120_gates", n, head, dict(fuse=0)),
                  (f"n{n}_cap6", n, ga, dict(max_ops_per_sweep=6, min_tail_ops=-1)),
                  (f"n{n}_cap12", n, ga, dict(max_ops_per_sweep=12)),
                  (f"n{n}_default", n, ga, dict())]
    for name, n, ga, kw in cases:
        u = Q.Unitary(n) if "--dry" not in sys.argv else None
        row = {"case": name}
        for mib in (-1, 0, 32, 64):
            plan = Q.Plan(n, 2**n, ga, Q.PlanOptions(l2_block_mib=mib, **kw))
            st = plan.stats()
            if "--dry" in sys.argv:
                print(name, mib, st.n_sweeps, st.n_micro_ops, plan.dump()["block_cols"])
                continue
            u.set_identity()
            plan.run(u)
            ms = min(plan.run(u).device_ms for _ in range(3))
            row.update(sweeps=st.n_sweeps, uops=st.n_micro_ops)
            row["auto_block_cols" if mib == 0 else f"ms_{'off' if mib < 0 else mib}"] = plan.dump()["block_cols"] if mib == 0 else round(ms, 4)
            if mib == 0:
                row["ms_auto"] = round(ms, 4)
            plan.close()
        if u is not None:
            u.close()
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
