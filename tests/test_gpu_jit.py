"""Specialised (run-time compiled) sweeps against the interpreter and the oracle.

qsx_plan_options.jit = 2 compiles every sweep program of a plan with NVRTC into straight-line code (jit_codegen.cpp)
and waits for the compiler; jit = -1 interprets the same programs (sweep_kernel).  The generated code spells out the
interpreter's arithmetic operation by operation, so both executors must produce the SAME values (== on every element;
only the sign of a zero may differ, which == ignores).  Needs a B200 and libnvrtc."""
import numpy as np
import pytest

import qsyn_b200 as Q
from oracle import oracle as O
from qsyn_b200 import abi
from tests.test_abi_cpu import _rand_circuit

pytestmark = pytest.mark.gpu
TOL = 1e-12


def frob_rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _run(n, gs, col0=0, ncols=None, **opts):
    u = Q.Unitary(n, 0, col0, ncols)
    plan = Q.Plan(n, 2**n if ncols is None else ncols, gs, Q.PlanOptions(**opts))
    st = plan.run(u)
    out = u.download()
    u.close()
    plan.close()
    return out, st


def test_jit_is_available_on_the_gpu_box():
    s = abi.jit_stats()
    assert s.available == 1, "libnvrtc / driver entry points missing: specialised sweeps cannot be tested"


@pytest.mark.parametrize("n,g,seed,extra", [(4, 60, 1, True), (6, 150, 2, True), (8, 300, 3, True), (10, 500, 4, False), (10, 300, 5, True),
                                            (11, 400, 6, True)])
@pytest.mark.parametrize("opts", [{}, {"reg_qubits": 3, "tile_qubits": 7}, {"reg_qubits": 5, "tile_qubits": 9}, {"snap_tol": -1.0},
                                  {"no_perm_folding": 1}, {"reg_qubits": 2, "tile_qubits": 5, "log2_tile_cols": 2, "max_ops_per_sweep": 40},
                                  {"tile_qubits": 10, "log2_tile_cols": 2}, {"fuse": 0},
                                  {"tile_qubits": 10}, {"tile_qubits": 9, "reg_qubits": 3}, {"tile_qubits": 10, "no_perm_folding": 1}],
                         ids=["default", "R3m7", "R5m9", "nosnap", "noperm", "R2m5w4", "m10w4", "unfused", "m10pair", "R3m9pair", "m10pair-noperm"])
def test_specialised_sweeps_equal_the_interpreter(n, g, seed, extra, opts, request, monkeypatch):
    if opts.get("fuse") == 0 and g > 200:
        pytest.skip("one kernel per gate: covered at the small sizes")
    if "pair" in request.node.callspec.id:
        monkeypatch.setenv("QSX_JIT_PAIR", "1")  # 512-thread tiles as clusters of two CTAs (distributed shared memory)
    qc = _rand_circuit(n, g, seed, extra=extra)
    gs = O.gate_stream(qc)
    want = O.to_matrix_rowform(qc)
    opts = dict(opts)
    opts.setdefault("tile_qubits", 9)  # the default tile depends on the executor (9 specialised, 8 interpreted): same PLAN for both
    interp, st_i = _run(n, gs, jit=-1, **opts)
    spec, st_j = _run(n, gs, jit=2, **opts)
    assert st_i.jit_launches == 0
    assert st_j.jit_launches == st_j.launches > 0, (st_j.jit_launches, st_j.launches, abi.jit_stats().failed)
    assert np.array_equal(spec, interp)
    assert frob_rel(spec, want) < TOL


def test_specialised_sweeps_on_shards_blocks_and_existing_data():
    n = 10
    qc, qc2 = _rand_circuit(n, 400, 21), _rand_circuit(n, 150, 22)
    gs, gs2 = O.gate_stream(qc), O.gate_stream(qc2)
    want = O.to_matrix_rowform(qc2) @ O.to_matrix_rowform(qc)
    for col0, ncols in ((0, 2**n), (256, 256), (2**n - 8, 8)):
        for blk in (-1, 1):
            outs = []
            for jit in (-1, 2):
                u = Q.Unitary(n, 0, col0, ncols)
                u.apply(gs, Q.PlanOptions(jit=jit, l2_block_mib=blk, tile_qubits=9))   # from the lazy identity
                u.apply(gs2, Q.PlanOptions(jit=jit, l2_block_mib=blk, tile_qubits=9))  # on existing data
                outs.append(u.download())
                u.close()
            assert np.array_equal(outs[0], outs[1]), (col0, ncols, blk)
            assert frob_rel(outs[1], want[:, col0:col0 + ncols]) < TOL


def test_kernel_cache_and_async_mode():
    """jit = 1 never blocks: the first run may interpret, after qsx_plan_compile(wait) every sweep is specialised;
    a second plan of the same circuit hits the cache."""
    n = 9
    qc = _rand_circuit(n, 300, 31)
    gs = O.gate_stream(qc)
    want = O.to_matrix_rowform(qc)
    before = abi.jit_stats()
    plan = Q.Plan(n, 2**n, gs, Q.PlanOptions(jit=1))
    u = Q.Unitary(n)
    st = plan.run(u)
    assert 0 <= st.jit_launches <= st.launches
    assert frob_rel(u.download(), want) < TOL
    plan.compile(True)
    u.set_identity()
    st = plan.run(u)
    assert st.jit_launches == st.launches
    assert frob_rel(u.download(), want) < TOL
    mid = abi.jit_stats()
    made = lambda a, b: (a.compiled + a.disk_hits) - (b.compiled + b.disk_hits)  # compiled here, or found in the on-disk cache
    assert made(mid, before) >= 1 and mid.failed == before.failed
    plan2 = Q.Plan(n, 2**n, gs, Q.PlanOptions(jit=2))
    u.set_identity()
    st = plan2.run(u)
    after = abi.jit_stats()
    assert st.jit_launches == st.launches and made(after, mid) == 0 and after.cache_hits > mid.cache_hits
    # and the default (jit = 0) leaves small circuits to the interpreter
    st = Q.Unitary(n).apply(gs)
    assert st.jit_launches == 0


def test_c2_specialised_equals_interpreted_at_full_size():
    from qsyn_b200 import workloads
    n, desc, text = workloads.workload_qasm("c2")
    gs = O.gate_stream(O.qcir_from_qasm_text(text))
    outs, sums = [], []
    for jit in (-1, 2):
        u = Q.Unitary(n)
        st = u.apply(gs, Q.PlanOptions(jit=jit, tile_qubits=9))
        assert (st.jit_launches == st.launches) == (jit == 2)
        outs.append(u.download_block(0, 2**n, 1000, 64))
        sums.append(Q.equiv_reduce([u], [u]))
        u.close()
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(sums[0], sums[1])
