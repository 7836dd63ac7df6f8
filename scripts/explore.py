"""Exploratory timing of plan options on one GPU (not the bench; numbers guide tuning)."""
import json
import sys
import time

sys.path.insert(0, ".")
import qsyn_b200 as Q
from qsyn_b200 import host, workloads as W


def time_plan(n, gs, opt, reps=3):
    ncols = 2**n
    plan = Q.Plan(n, ncols, gs, opt)
    st = plan.stats()
    u = Q.Unitary(n)
    plan.run(u)  # warm
    best = 1e30
    for _ in range(reps):
        r = plan.run(u)
        best = min(best, r.device_ms)
    u.close()
    return dict(sweeps=st.n_sweeps, stages=st.n_stages, ms=round(best, 3),
                alg_GBps=round(st.alg_bytes_per_run / best / 1e6, 1), eff_GBps=round(st.gate_bytes_per_run / best / 1e6, 1),
                us_per_sweep=round(best * 1e3 / max(st.n_sweeps, 1), 1), plan_ms=round(st.plan_host_ms, 2))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    g = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    configs = json.loads(sys.argv[3]) if len(sys.argv) > 3 else [
        (0, 4, 10, 3, 48), (0, 4, 10, 5, 48),
        (1, 4, 10, 3, 8), (1, 4, 10, 3, 16), (1, 4, 10, 3, 48), (1, 4, 10, 3, 200),
        (1, 4, 9, 3, 48), (1, 4, 8, 3, 48), (1, 3, 9, 3, 48), (1, 3, 8, 3, 48), (1, 4, 6, 3, 48), (1, 3, 7, 4, 48),
    ]
    workload = sys.argv[4] if len(sys.argv) > 4 else "clifford_t"
    min_tail = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    search = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    extra = json.loads(sys.argv[7]) if len(sys.argv) > 7 else {}
    text = W.random_rotation_qasm(n, g, 20261019) if workload == "rot" else W.random_clifford_t_qasm(n, g, 20261019)
    qc = host.Circuit(qasm_text=text)
    gs = qc.gate_array()
    print("n", n, "gates", g, workload, flush=True)
    for cfg in configs:
        fuse, R, tile, w, cap = cfg[:5]
        pol = cfg[5] if len(cfg) > 5 else 0
        dense = cfg[6] if len(cfg) > 6 else 0
        opt = Q.PlanOptions(fuse=fuse, reg_qubits=R, tile_qubits=tile, log2_tile_cols=w, max_ops_per_sweep=cap, no_perm_folding=pol,
                            dense_block=dense, min_tail_ops=min_tail, search=search, **extra)
        try:
            r = time_plan(n, gs, opt)
        except Exception as e:  # noqa
            r = {"error": str(e)}
        r.update(fuse=fuse, R=R, tile=tile, log2w=w, cap=cap, pol=pol, dense=dense)
        print(json.dumps(r), flush=True)
    qft_c = host.Circuit(qasm_text=W.qft_qasm(n))
    qft = qft_c.gate_array()
    print("qft", json.dumps(time_plan(n, qft, Q.PlanOptions(max_ops_per_sweep=128))), flush=True)
    a, b = Q.Unitary(n), Q.Unitary(n)
    a.equiv_partial(b)
    t0 = time.perf_counter()
    for _ in range(5):
        a.equiv_partial(b)
    dt = (time.perf_counter() - t0) / 5
    print("equiv ms", round(dt * 1e3, 3), "GB/s", round(2 * 16 * 4**n / dt / 1e9, 1))


if __name__ == "__main__":
    main()
