"""Tile / width / ops-per-sweep sweep for the specialised sweeps (needs a B200).  One process per QSX_JIT_MINCTAS
value (the occupancy target is baked into a kernel when it is compiled and is not part of the cache key).
    python scripts/jit_tile_sweep.py c4 "8,3,96 9,3,96 9,2,96 10,2,96 10,2,200" > profiles/...jsonl"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsyn_b200 as Q
from qsyn_b200 import abi, host, workloads as W

name = sys.argv[1]
configs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[2].split()]
n, desc, text = W.workload_qasm(name)
circ = host.Circuit(qasm_text=text)
gs = circ.gate_array()
u = Q.Unitary(n)
for cfg in configs:
    tile, log2w, cap = cfg[:3]
    reg = cfg[3] if len(cfg) > 3 else 4
    opt = Q.PlanOptions(tile_qubits=tile, log2_tile_cols=log2w, max_ops_per_sweep=cap, reg_qubits=reg, jit=1)
    plan = Q.Plan(n, 2**n, gs, opt)
    st = plan.stats()
    t0 = time.time()
    plan.compile(True)
    tc = time.time() - t0
    u.set_identity()
    plan.run(u)
    ms = min(plan.run(u).device_ms for _ in range(2))
    rs = plan.run(u)
    print(json.dumps({"workload": name, "tile": tile, "log2w": log2w, "cap": cap, "reg": reg, "minctas": os.environ.get("QSX_JIT_MINCTAS"),
                      "sweeps": st.n_sweeps, "stages": st.n_stages, "uops": st.n_micro_ops, "construct_ms": round(ms, 3),
                      "jit_launches": rs.jit_launches, "launches": rs.launches, "ms_per_sweep": round(ms / max(st.n_sweeps, 1), 4),
                      "GBps": round(st.alg_bytes_per_run / ms / 1e6, 1), "compile_s": round(tc, 2), "failed": abi.jit_stats().failed}), flush=True)
    plan.close()
