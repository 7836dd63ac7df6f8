"""SURVEY 8 rows a13 / f3 / f4 on the device: `qcir equiv` (the tensor half of qcir::is_equivalent,
reference src/qcir/qcir_equiv.cpp:57-65: to_tensor(qc1^-1 . qc2) against QTensor::identity(n) on the interleaved
rank-2n shape), element access on a device-resident tensor (src/tensor/decomposer.hpp:64-69 reads matrix(i, j)),
and `tensor write` -> `tensor read` -> `tensor equiv` (src/tensor/tensor.hpp:326-384)."""
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from qsyn_b200 import host
from tests.test_cli_golden import SHIM, run_dof

pytestmark = pytest.mark.gpu


def _circuit(text):
    return host.Circuit(qasm_text=text)


def test_qcir_equiv_tensor_half(goldens):
    inp = goldens["inputs"]
    q5, q5dec, q5crz = (_circuit(inp[f"benchmark/qft/qft_5{s}.qasm"]) for s in ("", "dec", "crz"))
    assert q5.equiv(q5dec)            # exact decomposition
    assert q5.equiv(q5crz)            # equal up to a global phase
    assert q5.equiv(q5)
    other = _circuit(O.random_clifford_t_qasm(5, 40, 3))
    assert not q5.equiv(other)
    # one extra T gate breaks it; T then Tdg restores it
    a = _circuit(O.random_clifford_t_qasm(7, 120, 5))
    b = _circuit(O.random_clifford_t_qasm(7, 120, 5))
    b.append("t", [3])
    assert not a.equiv(b)
    b.append("tdg", [3])
    assert a.equiv(b)
    # different qubit counts: not equivalent, nothing is launched (qcir_equiv.cpp:20-23)
    assert not a.equiv(q5)
    # beyond the reference's 7-qubit cut-off the device still answers
    c = _circuit(O.random_clifford_t_qasm(11, 300, 9))
    d = _circuit(O.random_clifford_t_qasm(11, 300, 9))
    for g in "txtx":
        d.append(g, [0])  # global phase pi/4
    assert c.equiv(d)
    d.append("h", [10])
    assert not c.equiv(d)


def test_circuit_adjoint_is_the_inverse(goldens):
    text = goldens["inputs"]["benchmark/qft/qft_8crz.qasm"]
    fwd, inv = _circuit(text), _circuit(text).adjoint_inplace()
    u, v = fwd.qc2ts().to_numpy(), inv.qc2ts().to_numpy()
    assert np.allclose(v, u.conj().T, rtol=0, atol=1e-13)
    assert np.allclose(v @ u, np.eye(2**8), rtol=0, atol=1e-12)


def test_interleaved_device_tensor_against_host_identity():
    """a13 in isolation: is_equivalent(to_tensor(qc), QTensor::identity(n)) -- the left operand is device-resident with the
    interleaved logical shape, the right one is built on the host and lifted (qtensor.hpp equivalence_sums)."""
    n = 6
    ident = host.Circuit(n)
    for q in range(n):  # h h = identity on every qubit, s s s s = identity
        ident.append("h", [q]).append("h", [q])
        for _ in range(4):
            ident.append("s", [q])
    empty = host.Circuit(n)
    assert empty.equiv(ident)
    ident.append("z", [2])
    assert not empty.equiv(ident)


def test_element_access_reads_single_elements_from_the_device():
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(9, 250, 13))
    want = O.to_matrix_rowform(qc)
    t = _circuit(O.random_clifford_t_qasm(9, 250, 13)).qc2ts()
    rng = np.random.default_rng(1)
    for _ in range(40):
        i, j = (int(x) for x in rng.integers(0, 2**9, 2))
        assert abs(t.element(i, j) - want[i, j]) < 1e-13
    assert abs(t.element(0, 0) - want[0, 0]) < 1e-13 and abs(t.element(511, 511) - want[511, 511]) < 1e-13
    with pytest.raises(Exception):
        t.element(512, 0)


def test_tensor_write_read_roundtrip(tmp_path, goldens):
    """`tensor write` streams the device matrix row block by row block; `tensor read` + `tensor equiv --strict`
    gives Equivalent / norm 1 / phase 0, and the file is the reference's "re+imj" CSV."""
    text = goldens["inputs"]["benchmark/qft/qft_8.qasm"]
    t = _circuit(text).qc2ts()
    path = tmp_path / "qft8.csv"
    t.write(path)
    lines = path.read_text().split("\n")
    assert len(lines) == 257 and lines[-1] == ""
    first = lines[0].split(", ")
    m = t.to_numpy()

    def shortest(x):  # fmt's "{}" of a double: shortest round-trip digits, no trailing ".0"
        s = repr(float(x))
        return s[:-2] if s.endswith(".0") else s

    def cell(z):  # tensor.hpp:335-339: "{space if re >= 0}{re}+{|im|}j" / "{re}{im}j"
        return (" " if z.real >= 0 else "") + shortest(z.real) + ("+" + shortest(abs(z.imag)) if z.imag >= 0 else shortest(z.imag)) + "j"

    assert len(first) == 256 and first[0] == cell(m[0, 0]) == " 0.06250000000000001+0j"
    assert lines[1].split(", ")[1] == cell(m[1, 1]) and lines[200].split(", ")[77] == cell(m[200, 77])
    parsed = np.array([[complex(c.strip().replace("j", "j")) for c in ln.split(", ")] for ln in lines[:-1]])
    assert np.array_equal(parsed, m)  # shortest round-trip formatting loses nothing
    back = host.Tensor.read(path)
    assert back.shape == [256, 256]
    r, msg = t.equiv(back, 1e-6, True)
    assert r.equivalent == 1 and msg == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n"
    assert np.array_equal(back.to_numpy(), m)
    # a non-square file is refused like the reference does
    bad = tmp_path / "bad.csv"
    bad.write_text(" 1+0j,  0+0j\n 0+0j\n")
    with pytest.raises(Exception):
        host.Tensor.read(bad)


def test_cli_qcir_equiv_and_tensor_write_read(tmp_path, goldens):
    dof = "\n".join([
        "qcir read benchmark/qft/qft_5.qasm", "qcir read benchmark/qft/qft_5dec.qasm", "qcir read benchmark/qft/qft_5crz.qasm",
        "qcir equiv 0 1", "qcir equiv 0", "qcir checkout 0", "qcir gate add t 2", "qcir equiv 1",
        "qc2ts", "tensor write t.csv", "tensor read t.csv", "tensor equiv 0 1 --strict", "quit -f"])
    rc, out = run_dof(tmp_path, goldens, dof)
    assert rc == 0, out
    assert out.count("The two circuits are equivalent!!") == 2 and out.count("The two circuits are not equivalent!!") == 1
    assert "qsyn> tensor equiv 0 1 --strict\nEquivalent\n- Global Norm : 1\n- Global Phase: 0\n" in out


def test_nested_circuit_gate_through_qc2ts():
    from tests.test_host_layer import _nested_example
    outer, want = _nested_example()
    got = outer.qc2ts().to_numpy()
    assert np.abs(got - want).max() < 1e-14


def test_short_conversions_are_enqueued_long_ones_are_waited_for():
    """to_tensor(QCir) calls qsx_apply_gates_ex with QSX_RUN_ASYNC_IF_SHORT: a stream of less than 64 GB of HBM traffic is
    only enqueued (device_ms is not measured, the result is ordered behind it on the device's stream), a longer one is
    synchronous with the stop callback polled between sweeps.  Either way the values are those of the plain call."""
    import qsyn_b200 as Q
    from qsyn_b200 import abi

    n = 11
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(n, 500, 17))
    gs = O.gate_stream(qc)
    ref = Q.Unitary(n)
    st_ref = ref.apply(gs)
    assert st_ref.device_ms > 0
    want = ref.download()
    polls = []
    u = Q.Unitary(n)
    st = u.apply_ex(gs, abi.QSX_RUN_ASYNC_IF_SHORT, should_stop=lambda: polls.append(1) and False)
    assert st.device_ms == 0 and st.launches == st_ref.launches and len(polls) >= st.launches - 1  # enqueued, callback polled per sweep
    assert np.array_equal(u.download(), want)
    # one gate per sweep at n = 12: ~500 passes over (a fraction of) 256 MiB, well over 64 GB -> a long run: synchronous
    gs12 = O.gate_stream(O.qcir_from_qasm_text(O.random_clifford_t_qasm(12, 500, 18)))
    ref12, v = Q.Unitary(12), Q.Unitary(12)
    ref12.apply(gs12)
    st = v.apply_ex(gs12, abi.QSX_RUN_ASYNC_IF_SHORT, Q.PlanOptions(fuse=0, l2_block_mib=-1))
    assert st.device_ms > 0 and st.launches >= 400
    assert np.abs(v.download() - ref12.download()).max() < 1e-12
    ref12.close(), v.close()
    # the stop callback still interrupts an enqueue-only run
    w = Q.Unitary(n)
    with pytest.raises(Q.QsxError) as ei:
        w.apply_ex(gs, abi.QSX_RUN_ASYNC_IF_SHORT, should_stop=lambda: True)
    assert ei.value.status == "QSX_ERR_INTERRUPTED"
    with pytest.raises(Q.QsxError):
        w.apply_ex(gs, 64)
    # through the host layer: two conversions back to back and the check give the text of the synchronous path
    text = O.random_clifford_t_qasm(n, 500, 17)
    ca, cb = _circuit(text), _circuit(text + "t q[0];\nx q[0];\nt q[0];\nx q[0];\n")
    ta, tb = ca.qc2ts(), cb.qc2ts()
    r, verdict = ta.equiv(tb)
    assert verdict == "Equivalent\n- Global Norm : 1\n- Global Phase: π/4\n"
    assert np.array_equal(ta.to_numpy(), want)
