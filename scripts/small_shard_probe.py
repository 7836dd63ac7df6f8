"""C2's 1/8 column shard (32 MiB: what one of 8 GPUs holds) on one GPU: ms per operand for a few sweep variants.
    QSX_JIT_PERSIST=1 python scripts/small_shard_probe.py [shards]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsyn_b200 as Q
from qsyn_b200 import abi, host, workloads as W

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, desc, text = W.workload_qasm("c2")
circ = host.Circuit(qasm_text=text)
gs = circ.gate_array()
ncols = 2**n // shards
u = Q.Unitary(n, 0, 0, ncols)
for tile, log2w in ((9, 3), (8, 3), (9, 2), (8, 2)):
    plan = Q.Plan(n, ncols, gs, Q.PlanOptions(tile_qubits=tile, log2_tile_cols=log2w, jit=1))
    plan.compile(True)
    u.set_identity()
    plan.run(u)
    ms = min(plan.run(u).device_ms for _ in range(10))
    st = plan.stats()
    print(json.dumps({"shards": shards, "tile": tile, "log2w": log2w, "persist": os.environ.get("QSX_JIT_PERSIST"), "sweeps": st.n_sweeps,
                      "construct_ms": round(ms, 4), "us_per_sweep": round(ms * 1e3 / st.n_sweeps, 2), "failed": abi.jit_stats().failed}), flush=True)
    plan.close()
