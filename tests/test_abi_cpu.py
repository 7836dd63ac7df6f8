"""CPU-only tests of the C-ABI library: it loads, exports every symbol include/qsx.h
declares, and its host logic (gate lowering, scheduling, equivalence finish, phase
conversion) matches the oracle.  No kernels are launched here."""
import os
import re

import numpy as np
import pytest

import qsyn_b200 as Q
from oracle import oracle as O
from tests import plan_sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qsx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(qsx_[a-z_0-9]+)\s*\(", hdr))
    names -= {"qsx_stop_fn"}
    assert len(names) >= 20
    L = Q.lib()
    for nm in sorted(names):
        assert hasattr(L, nm), f"libqsx.so does not export {nm}"
    assert L.qsx_abi_version() == 2


def test_no_cpu_fallback_when_no_device():
    if Q.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(Q.QsxError) as ei:
        Q.Unitary(3)
    assert ei.value.status == "QSX_ERR_CUDA" and "no CPU fallback" in str(ei.value)


def test_apply_gates_ex_checks_its_arguments_before_touching_a_device():
    """qsx_apply_gates_ex (the general form to_tensor(QCir) calls): unknown run flags and null handles are QSX_ERR_BAD_ARG
    on any host; the flag values are those of include/qsx.h."""
    import ctypes as C

    from qsyn_b200 import abi

    L = Q.lib()
    assert (abi.QSX_RUN_ASYNC, abi.QSX_RUN_ASYNC_IF_SHORT) == (1, 2)
    hdr = open(os.path.join(ROOT, "include", "qsx.h")).read()
    assert "#define QSX_RUN_ASYNC 1u" in hdr and "#define QSX_RUN_ASYNC_IF_SHORT 2u" in hdr
    null_stop = C.cast(None, abi.STOP_FN)
    assert L.qsx_apply_gates_ex(None, None, 0, None, null_stop, None, 64, None) != 0
    assert b"unknown run flag" in L.qsx_last_error()
    assert L.qsx_apply_gates_ex(None, None, 0, None, null_stop, None, abi.QSX_RUN_ASYNC_IF_SHORT, None) != 0
    assert b"null unitary" in L.qsx_last_error()


def _rand_circuit(n, g, seed, extra=True):
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(n, g, seed))
    if extra:  # sprinkle the rest of the reference gate set
        rng = np.random.default_rng(seed + 1)
        names1 = ["y", "sx", "sy", "sxdg", "tx", "ty", "id"]
        namesp = ["rz", "rx", "ry", "p", "px", "py"]
        for _ in range(g // 3):
            k = rng.integers(0, 7)
            if k == 0:
                qc.append(O.str_to_operation(names1[rng.integers(len(names1))]), [int(rng.integers(n))])
            elif k == 1:
                ph = O.Phase(int(rng.integers(-7, 8)), int(rng.integers(1, 9)))
                qc.append(O.str_to_operation(namesp[rng.integers(len(namesp))], [ph]), [int(rng.integers(n))])
            elif k == 2 and n >= 2:
                a, b = rng.choice(n, 2, replace=False)
                qc.append(O.str_to_operation(["swap", "ecr", "cy", "ch"][rng.integers(4)]), [int(a), int(b)])
            elif k == 3 and n >= 2:
                a, b = rng.choice(n, 2, replace=False)
                ph = O.Phase(int(rng.integers(-7, 8)), int(rng.integers(1, 9)))
                qc.append(O.str_to_operation(["cp", "crz", "crx", "cry"][rng.integers(4)], [ph]), [int(a), int(b)])
            elif k == 4 and n >= 3:
                a, b, c = rng.choice(n, 3, replace=False)
                qc.append(O.str_to_operation(["ccx", "ccz", "cswap", "cecr", "ccy"][rng.integers(5)]), [int(a), int(b), int(c)])
            elif k == 5 and n >= 4:
                q = rng.choice(n, 4, replace=False)
                qc.append(O.str_to_operation(["cccx", "cccz", "ccswap"][rng.integers(3)]), [int(x) for x in q])
            else:
                qc.append(O.str_to_operation("h"), [int(rng.integers(n))])
    return qc


@pytest.mark.parametrize("n,g,seed", [(1, 12, 1), (2, 30, 2), (3, 40, 3), (5, 120, 4), (6, 200, 5), (7, 150, 6), (9, 300, 16)])
@pytest.mark.parametrize("fuse", [0, 1])
def test_schedule_preserves_the_unitary(n, g, seed, fuse):
    qc = _rand_circuit(n, g, seed)
    want = O.to_matrix_rowform(qc)
    for opt in (Q.PlanOptions(fuse=fuse), Q.PlanOptions(fuse=fuse, reg_qubits=3, tile_qubits=5, max_ops_per_sweep=9),
                Q.PlanOptions(fuse=fuse, snap_tol=-1.0, reg_qubits=2), Q.PlanOptions(fuse=fuse, reg_qubits=1, tile_qubits=2),
                Q.PlanOptions(fuse=fuse, no_perm_folding=1), Q.PlanOptions(fuse=fuse, no_perm_folding=2, reg_qubits=3),
                Q.PlanOptions(fuse=fuse, reg_qubits=4, tile_qubits=8, max_ops_per_sweep=200),
                Q.PlanOptions(fuse=fuse, no_perm_folding=2, tile_qubits=7),
                # every search / fusion knob, and the R = 5 instantiation (a planner that stops making
                # progress shows up here, on CPU, as a timeout instead of a hung GPU test)
                Q.PlanOptions(fuse=fuse, reg_qubits=5, tile_qubits=9, stage_search=1),
                Q.PlanOptions(fuse=fuse, stage_search=1, search=3, tile_qubits=5, reg_qubits=3),
                Q.PlanOptions(fuse=fuse, stage_search=-1, search=-1, min_tail_ops=-1, no_1q_fusion=1, no_diag_conjugation=1)):
        plan = Q.Plan(n, 2**n, O.gate_stream(qc), opt)
        d = plan.dump()
        st = plan.stats()
        plan_sim.check_invariants(d, st.n_ops)
        assert st.n_sweeps == len(d["sweeps"]) and st.n_gates == len(qc.gates)
        if fuse == 0:
            assert st.n_sweeps == st.n_ops
        got = plan_sim.run_plan(d, np.eye(2**n, dtype=np.complex128))
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err < 1e-13, err
        # ... and so do the merged micro-ops the kernel actually executes
        got = plan_sim.run_plan_uops(d, np.eye(2**n, dtype=np.complex128))
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err < 1e-13, err


@pytest.mark.parametrize("k", [3, 4, 5, 6])
def test_dense_block_folding_preserves_the_unitary(k):
    """dense_block = k: the whole gate stream becomes dense 2^k x 2^k blocks (host products)."""
    for n, g, seed in [(5, 120, 4), (7, 150, 6), (9, 300, 16)]:
        if n < k:
            continue
        qc = _rand_circuit(n, g, seed)  # seed 16 contains cccx / ccswap: wider than a k = 3 / 4 block
        want = O.to_matrix_rowform(qc)
        plan = Q.Plan(n, 2**n, O.gate_stream(qc), Q.PlanOptions(dense_block=k))
        d, st = plan.dump(), plan.stats()
        assert 0 < st.n_dense_blocks <= st.n_sweeps == len(d["sweeps"])
        seen = plan_sim.check_invariants(d)
        assert len(seen) == st.n_ops
        got = plan_sim.run_plan_uops(d, np.eye(2**n, dtype=np.complex128))
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13
    with pytest.raises(Q.QsxError):
        Q.Plan(5, 32, [], Q.PlanOptions(dense_block=2))
    with pytest.raises(Q.QsxError):
        Q.Plan(4, 16, [], Q.PlanOptions(dense_block=5))
    with pytest.raises(Q.QsxError):
        Q.Plan(8, 256, [], Q.PlanOptions(dense_block=7))


def _rand_unitary(k, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((2**k, 2**k)) + 1j * rng.standard_normal((2**k, 2**k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def _circuit_with_dense_gates(n, seed):
    """random Clifford+T stream with user gates of 3..6 qubits (targets + controls) sprinkled in: what a nested circuit
    handed over as ONE matrix looks like (reference: to_tensor(QCir) used as a gate, qcir_to_tensor.cpp:147)."""
    rng = np.random.default_rng(seed)
    gs = O.gate_stream(O.qcir_from_qasm_text(O.random_clifford_t_qasm(n, 60, seed)))
    out = []
    for i, g in enumerate(gs):
        out.append(g)
        if i % 12 == 5:
            k = int(rng.integers(3, min(n, 6) + 1))
            nc = int(rng.integers(0, k - 2))  # at least 3 targets
            qs = [int(q) for q in rng.choice(n, k, replace=False)]
            out.append((nc, qs, _rand_unitary(k - nc, seed * 100 + i)))
    return out


def _oracle_matrix(n, gs):
    u = np.eye(2**n, dtype=np.complex128)
    for nc, qs, m in gs:
        u = O.apply_gate_rows(u, n, O.mat_control(m, nc), qs)
    return u


@pytest.mark.parametrize("n,seed", [(3, 1), (5, 2), (6, 3), (8, 4), (9, 5)])
def test_dense_user_gates_become_dense_sweeps(n, seed):
    gs = _circuit_with_dense_gates(n, seed)
    want = _oracle_matrix(n, gs)
    for opt in (None, Q.PlanOptions(fuse=0), Q.PlanOptions(dense_block=4 if n >= 4 else 3), Q.PlanOptions(reg_qubits=3, tile_qubits=6)):
        plan = Q.Plan(n, 2**n, gs, opt)
        d, st = plan.dump(), plan.stats()
        n_dense_gates = sum(1 for nc, qs, m in gs if len(qs) - nc >= 3)
        assert st.n_dense_blocks >= n_dense_gates > 0
        seen = plan_sim.check_invariants(d)
        assert len(seen) == st.n_ops
        got = plan_sim.run_plan_uops(d, np.eye(2**n, dtype=np.complex128))
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13
    # too wide for a dense block: refused with the reference's "not supported" status, not a wrong answer
    with pytest.raises(Q.QsxError) as ei:
        Q.Plan(9, 512, [(0, list(range(7)), _rand_unitary(7, 1))])
    assert ei.value.status == "QSX_ERR_UNSUPPORTED"


def test_snapping_is_optional_and_small():
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(4, 300, 9))
    want = O.to_matrix_rowform(qc)
    d_exact = Q.Plan(4, 16, O.gate_stream(qc), Q.PlanOptions(snap_tol=-1.0)).dump()
    d_snap = Q.Plan(4, 16, O.gate_stream(qc)).dump()
    kinds = lambda d: {op["kind"] for sw in d["sweeps"] for st in sw["stages"] for op in st["ops"]}
    ukinds = {u["kind"] for sw in d_snap["sweeps"] for st in sw["stages"] for u in st["uops"]}
    assert "H" in ukinds and ukinds & {"DTABH", "DTABF"}  # n = R = 4: every bit is a register bit
    d6 = Q.Plan(6, 64, O.gate_stream(O.qcir_from_qasm_text(O.random_clifford_t_qasm(6, 300, 9)))).dump()
    assert "SCALN" in {u["kind"] for sw in d6["sweeps"] for st in sw["stages"] for u in st["uops"]}
    assert "X" in kinds(d_snap) and "NEG" in kinds(d_snap) and "MULI" in kinds(d_snap)
    assert "X" not in kinds(d_exact) and "NEG" not in kinds(d_exact)
    for d in (d_exact, d_snap):
        got = plan_sim.run_plan(d, np.eye(16, dtype=np.complex128))
        assert np.linalg.norm(got - want) / 4.0 < 1e-13
        got = plan_sim.run_plan_uops(d, np.eye(16, dtype=np.complex128))
        assert np.linalg.norm(got - want) / 4.0 < 1e-13


def test_plan_rejects_bad_input():
    h = O.mat_h()
    with pytest.raises(Q.QsxError) as ei:
        Q.Plan(3, 8, [(0, [3], h)])
    assert ei.value.status == "QSX_ERR_BAD_ARG"
    with pytest.raises(Q.QsxError):
        Q.Plan(3, 8, [(1, [1, 1], h)])
    with pytest.raises(Q.QsxError):
        Q.Plan(3, 6, [(0, [0], h)])  # ncols not a power of two
    with pytest.raises(Q.QsxError) as ei:  # wider than the widest dense block (6 qubits incl. controls)
        Q.Plan(8, 256, [(1, list(range(7)), _rand_unitary(6, 3))])
    assert ei.value.status == "QSX_ERR_UNSUPPORTED"
    assert Q.Plan(3, 8, [(0, [0, 1, 2], np.eye(8))]).stats().n_sweeps == 0  # a 3-target identity vanishes like any identity
    assert Q.Plan(3, 8, [(0, [0, 1, 2], _rand_unitary(3, 4))]).stats().n_dense_blocks == 1
    # identity gates vanish, empty stream is legal (zero gates -> identity, Appendix A.1)
    p = Q.Plan(3, 8, [(0, [0], np.eye(2)), (1, [0, 1], np.eye(2))])
    assert p.stats().n_ops == 0 and p.stats().n_sweeps == 0
    assert Q.Plan(3, 8, []).stats().n_sweeps == 0


def test_qft_needs_few_sweeps():
    """Controlled phases need no locality: QFT-14 (105 gates) collapses to a couple of sweeps."""
    qc = O.qcir_from_qasm_text(O.qft_qasm(14))
    st = Q.Plan(14, 2**14, O.gate_stream(qc), Q.PlanOptions(max_ops_per_sweep=128)).stats()
    assert st.n_gates == 105 and st.n_sweeps <= 2


def test_l2_column_blocking_decision():
    """qsx_plan_options.l2_block_mib: off / forced sizes are powers of two between the tile width and the
    shard; the default only blocks plans the cost model calls memory-bound (few micro-ops per sweep)."""
    n = 12
    qc = _rand_circuit(n, 300, 5)
    gs = O.gate_stream(qc)
    assert Q.Plan(n, 2**n, gs).dump()["block_cols"] == 0                      # compute-bound: not blocked
    assert Q.Plan(n, 2**n, gs, Q.PlanOptions(l2_block_mib=-1)).dump()["block_cols"] == 0
    col_bytes = 16 * 2**n
    for mib in (1, 2, 32, 64):
        bc = Q.Plan(n, 2**n, gs, Q.PlanOptions(l2_block_mib=mib)).dump()["block_cols"]
        assert bc == min(mib * 2**20 // col_bytes, 2**n) % 2**n               # (== shard: reported as 0)
    assert Q.Plan(n, 2**n, gs, Q.PlanOptions(l2_block_mib=4096)).dump()["block_cols"] == 0
    assert Q.Plan(n, 8, gs, Q.PlanOptions(l2_block_mib=1)).dump()["block_cols"] == 0   # one tile width: nothing to split
    assert Q.Plan(n, 2**n, gs, Q.PlanOptions(l2_block_mib=1, dense_block=4)).dump()["block_cols"] == 0
    # one gate per sweep (fuse = 0) is the memory-bound extreme: blocked by default
    assert Q.Plan(n, 2**n, gs, Q.PlanOptions(fuse=0)).dump()["block_cols"] == 64 * 2**20 // col_bytes
    assert Q.Plan(n, 2**n, gs[:1], Q.PlanOptions(fuse=0)).dump()["block_cols"] == 0    # a single sweep reads nothing twice


@pytest.mark.parametrize("case", ["t_x_t_x", "rz_small", "qft3crz", "zero_sum"])
def test_equiv_finish_matches_oracle(case, goldens):
    if case == "t_x_t_x":
        a = O.qc2ts(O.QCir(1))
        qc = O.QCir(1)
        for g in "txtx":
            qc.append(O.str_to_operation(g), [0])
        b = O.qc2ts(qc)
    elif case == "rz_small":
        a = O.qc2ts(O.QCir(1))
        qc = O.QCir(1)
        qc.append(O.str_to_operation("rz", [O.str_to_phase("0.001")]), [0])
        b = O.qc2ts(qc)
    elif case == "qft3crz":
        a = O.qc2ts(O.qcir_from_qasm_text(goldens["inputs"]["benchmark/qft/qft_3.qasm"]))
        b = O.qc2ts(O.qcir_from_qasm_text(goldens["inputs"]["benchmark/qft/qft_3crz.qasm"]))
    else:
        a = np.array([[1, 0], [0, -1]], dtype=np.complex128)  # sums to exactly zero
        b = np.eye(2, dtype=np.complex128)
    sums = O.equiv_sums(a, b)
    for eps in (1e-2, 1e-6, 1e-8, 1e-10):
        for strict in (False, True):
            r = Q.equiv_finish(sums, eps, strict)
            if case == "zero_sum":
                # Appendix C trap 2: defined behaviour instead of the reference's NaN -> UB
                assert r.phase_defined == 0 and not np.isfinite(r.norm)
                if strict:
                    assert r.equivalent == 0
                continue
            w = O.finish_equiv(sums, eps, strict)
            assert bool(r.equivalent) == w.equivalent
            assert r.cosine == pytest.approx(w.cosine, abs=1e-15)
            assert r.norm == pytest.approx(w.norm, abs=1e-15)
            assert (r.phase_num, r.phase_den) == (w.phase.numerator, w.phase.denominator)
