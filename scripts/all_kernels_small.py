"""Small invocations of every kernel family in one process, each checked for unitarity (a quick all-kernels smoke; it was
written for compute-sanitizer, which is closed on this GPU pool):
interpreted sweeps (all three instantiations), specialised sweeps, dense blocks k = 3..6, identity, equivalence, adjoint."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qsyn_b200 as Q  # noqa: E402
from qsyn_b200 import host, workloads as W  # noqa: E402

n = 10
ca = host.Circuit(qasm_text=W.random_clifford_t_qasm(n, 300, 5))  # (the gate arrays point into the circuits: keep them alive)
cr = host.Circuit(qasm_text=W.random_rotation_qasm(n, 120, 6))
ga, gr = ca.gate_array(), cr.gate_array()
ref = None
for name, arr, opt in [("interpreted", ga, Q.PlanOptions(jit=-1)), ("interpreted rotations", gr, Q.PlanOptions(jit=-1)),
                       ("specialised", ga, Q.PlanOptions(jit=2)), ("specialised rotations", gr, Q.PlanOptions(jit=2)),
                       ("specialised m=7 R=3", ga, Q.PlanOptions(jit=2, tile_qubits=7, reg_qubits=3)),
                       ("dense k=3", gr, Q.PlanOptions(dense_block=3)), ("dense k=4", gr, Q.PlanOptions(dense_block=4)),
                       ("dense k=5", gr, Q.PlanOptions(dense_block=5)), ("dense k=6", gr, Q.PlanOptions(dense_block=6))]:
    u = Q.Unitary(n)
    plan = Q.Plan(n, 2**n, arr, opt)
    st = plan.run(u)
    m = u.download()
    err = np.abs(m.conj().T @ m - np.eye(2**n)).max()
    print(f"{name}: launches={st.launches} specialised={st.jit_launches} unitarity_err={err:.1e}", flush=True)
    assert err < 1e-11
    v = Q.Unitary(n)
    r = Q.equiv_finish(Q.equiv_reduce([u], [v]))
    u.adjoint_inplace() if hasattr(u, "adjoint_inplace") else None
    for x in (plan, u, v):
        x.close()
print("done")
