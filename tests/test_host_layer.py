"""The C++ host layer (qsyn_b200/host: reference-shaped QCir / QASM reader / to_tensor / QTensor /
`tensor equiv`) through its C ABI (include/qsx_host.h), against the oracle."""
import os
import re

import numpy as np
import pytest

import qsyn_b200 as Q
from oracle import oracle as O
from qsyn_b200 import host, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qsx_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(qsxh_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 14
    for nm in sorted(names):
        assert hasattr(host.lib(), nm), nm


def test_workload_generators_match_the_oracle():
    """bench.py's product arm (qsyn_b200.workloads) and the CPU arm (oracle) see the same circuits."""
    assert workloads.random_clifford_t_qasm(12, 2000, 20261019) == O.random_clifford_t_qasm(12, 2000, 20261019)
    assert workloads.random_clifford_t_qasm(1, 20, 3) == O.random_clifford_t_qasm(1, 20, 3)
    assert workloads.qft_qasm(14) == O.qft_qasm(14)
    assert workloads.random_rotation_qasm(7, 90, 5) == O.random_rotation_qasm(7, 90, 5)


def test_qasm_reader_gate_order_and_matrices_match_oracle(goldens):
    """Every golden input circuit: same gate order (non-monotone ids), same reprs (float angle ->
    rational), and gate matrices equal to the oracle's reference formulas to the last ulp or two
    (libm vs numpy complex exp / product rounding)."""
    for path, text in goldens["inputs"].items():
        if not path.endswith(".qasm"):
            continue
        c = host.Circuit(qasm_text=text)
        qc = O.qcir_from_qasm_text(text)
        assert (c.n_qubits, c.n_gates) == (qc.n_qubits, len(qc.gates)), path
        order = qc.topological_order()
        assert c.gate_reprs() == [(g.gid, g.op.repr()) for g in order], path
        for (nc, qs, m), (onc, oqs, om) in zip(c.gate_stream(), O.gate_stream(qc)):
            assert nc == onc and qs == oqs and np.allclose(m, om, rtol=0, atol=5e-16), path


def test_whole_gate_table_and_phase_strings():
    cases = [("h", None), ("x", None), ("y", None), ("z", None), ("s", None), ("sdg", None), ("t", None), ("tdg", None),
             ("sx", None), ("sxdg", None), ("sy", None), ("sydg", None), ("tx", None), ("ty", None), ("id", None), ("not", None),
             ("x_1_2", None), ("rz", "pi/3"), ("rx", "6.28*0.46181601443868003"), ("ry", "-3*pi/7"), ("p", "0.001"),
             ("px", "2.3145031"), ("py", "pi"), ("pz", "-pi/4")]
    for name, ph in cases:
        c = host.Circuit(3).append(name, [1], ph)
        op = O.str_to_operation(name, [O.str_to_phase(ph)] if ph else [])
        assert c.gate_reprs()[0][1] == op.repr(), (name, ph)
        assert np.allclose(c.gate_stream()[0][2], op.target_matrix(), rtol=0, atol=5e-16), (name, ph)
    for name, qs, ph in [("cx", [0, 2], None), ("ccz", [2, 0, 1], None), ("cswap", [1, 0, 2], None), ("swap", [0, 1], None),
                         ("ecr", [2, 1], None), ("crz", [0, 1], "pi/8"), ("cccx", [3, 0, 1, 2], None), ("cecr", [0, 1, 2], None)]:
        c = host.Circuit(4).append(name, qs, ph)
        op = O.str_to_operation(name, [O.str_to_phase(ph)] if ph else [])
        nc, q, m = c.gate_stream()[0]
        assert (nc, q) == (op.n_ctrls, qs) and np.allclose(m, op.target_matrix(), rtol=0, atol=5e-16)
        assert c.gate_reprs()[0][1] == op.repr()
    for bad in [("foo", [0], None), ("rz", [0], None), ("h", [0], "pi"), ("cx", [0, 0], None), ("h", [9], None), ("rz", [0], "abc")]:
        with pytest.raises(Q.QsxError):
            host.Circuit(2).append(bad[0], bad[1], bad[2])
    with pytest.raises(Q.QsxError):
        host.Circuit(qasm_text="OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[2];\nh q[5];\n")


def test_reader_quirks_match_oracle():
    text = ("OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[3];\ncreg c[3];\n// a comment\n  h q[0]; // trailing\n"
            "mcx q[0],q[1],q[2];\nCX q[0] , q[1];\nrz(pi/2) q[2];\nbarrier q[0];\nt q[2];\n")
    c, qc = host.Circuit(qasm_text=text), O.qcir_from_qasm_text(text)
    assert c.gate_reprs() == [(g.gid, g.op.repr()) for g in qc.topological_order()]
    assert c.n_gates == len(qc.gates) == 4  # mcx and barrier are skipped silently (Appendix C trap 7)
    # a phase expression with blanks is "invalid phase on line" for the reference reader: both reject it
    bad = text.replace("rz(pi/2)", "rz(pi / 2)")
    assert O.qcir_from_qasm_text(bad) is None
    with pytest.raises(Q.QsxError):
        host.Circuit(qasm_text=bad)


def test_reader_differential_fuzz_against_the_oracle():
    """Random QASM text with the dialect's quirks -- blanks and tabs anywhere the reference's tokeniser tolerates them,
    `//` comments, upper-case names, unknown gates, parametrised gates, c^n prefixes, creg/qreg lines, out-of-range and
    repeated qubit ids, malformed operands -- through the C++ reader (string-view tokeniser, shared operations) and the
    oracle's restatement of qcir_reader.cpp:47-120: same accept / reject decision, same gates in the same order."""
    rng = np.random.default_rng(20261020)
    names1 = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "sxdg", "sy", "ty", "id", "H", "Tdg", "not", "x_1_2"]
    namesp = ["rz", "rx", "ry", "p", "px", "py", "pz", "RZ"]
    names2 = ["cx", "cz", "cy", "swap", "ecr", "ch", "cs", "CX", "ct"]
    namespc = ["crz", "cp", "crx", "ccp"]
    names3 = ["ccx", "ccz", "cswap", "cch"]
    phases = ["pi", "pi/2", "-pi/4", "3*pi/8", "0.25", "-1.5707963", "pi*3/7", "2*pi", "0", "1e-3", "pi / 2", "abc", "", "pi/0.5"]
    n_accept = n_reject = 0
    for trial in range(400):
        n = int(rng.integers(1, 7))
        lines = []
        expect_gate_lines = 0
        for _ in range(int(rng.integers(0, 25))):
            kind = rng.random()
            sp = lambda: " " * int(rng.integers(0, 3)) + ("\t" if rng.random() < 0.1 else "")  # noqa: E731
            def q(k):
                ids = [int(x) for x in rng.choice(max(n, 1), size=min(k, n), replace=False)]
                while len(ids) < k:
                    ids.append(int(rng.integers(n)))  # forces a repeated id when the gate needs more qubits than exist
                if rng.random() < 0.03:
                    ids[0] = n + int(rng.integers(0, 3))  # out of range
                sep = "," + (" " if rng.random() < 0.3 else "")
                return sep.join(f"q[{i}]" for i in ids)
            if kind < 0.45:
                line = f"{names1[int(rng.integers(len(names1)))]} {q(1)};"
            elif kind < 0.60:
                line = f"{namesp[int(rng.integers(len(namesp)))]}({phases[int(rng.integers(len(phases)))]}) {q(1)};"
            elif kind < 0.78:
                line = f"{names2[int(rng.integers(len(names2)))]} {q(2)};"
            elif kind < 0.84:
                nm = namespc[int(rng.integers(len(namespc)))]
                line = f"{nm}({phases[int(rng.integers(len(phases)))]}) {q(3 if nm == 'ccp' else 2)};"
            elif kind < 0.90:
                line = f"{names3[int(rng.integers(len(names3)))]} {q(3)};"
            elif kind < 0.93:
                line = ["barrier q;", "measure q[0] -> c[0];", "creg c[2];", "foo q[0];", "mcx q[0],q[1];", "u3(0,0,0) q[0];"][int(rng.integers(6))]
            elif kind < 0.96:
                line = ["", "   ", "// only a comment", "\t"][int(rng.integers(4))]
            else:
                line = ["h q0;", "h q[a];", "h q[];", "h ;", "cx q[0] q[1];", "h q[0]", "h  q [0];", "c q[0];"][int(rng.integers(8))]
            if rng.random() < 0.2:
                line = sp() + line
            if rng.random() < 0.15:
                line = line + sp() + "// note" + ("//x" if rng.random() < 0.3 else "")
            lines.append(line)
        text = "OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[%d];\n" % n + "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")
        try:
            qc = O.qcir_from_qasm_text(text)
        except Exception:  # the oracle mirrors a throw of the reference (e.g. a name of only 'c's): not a case for the diff
            continue
        try:
            c = host.Circuit(qasm_text=text)
        except Q.QsxError:
            c = None
        assert (c is None) == (qc is None), (text, qc is None)
        if qc is None:
            n_reject += 1
            continue
        n_accept += 1
        assert (c.n_qubits, c.n_gates) == (qc.n_qubits, len(qc.gates)), text
        assert c.gate_reprs() == [(g.gid, g.op.repr()) for g in qc.topological_order()], text
        for (nc, qs, m), (onc, oqs, om) in zip(c.gate_stream(), O.gate_stream(qc)):
            assert nc == onc and qs == oqs and np.allclose(m, om, rtol=0, atol=5e-16), text
        c.close()
    assert n_accept >= 100 and n_reject >= 30, (n_accept, n_reject)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["benchmark/qft/qft_5.qasm", "benchmark/qasm/qc2ts/toffoli_qiskit.qasm",
                                  "benchmark/SABRE/small/3_17_13.qasm", "benchmark/qft/qft_8crz.qasm"])
def test_qc2ts_through_the_host_layer(goldens, name):
    """qsyn::to_tensor(QCir) + to_matrix() of the C++ host layer == the reference's literal algorithm."""
    text = goldens["inputs"][name]
    t = host.Circuit(qasm_text=text).qc2ts()
    want = O.qc2ts(O.qcir_from_qasm_text(text))
    assert t.shape == list(want.shape)
    got = t.to_numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-12


@pytest.mark.gpu
def test_tensor_equiv_text_through_the_host_layer(goldens):
    ins = goldens["inputs"]
    a = host.Circuit(qasm_text=ins["benchmark/qft/qft_8.qasm"]).qc2ts()
    b = host.Circuit(qasm_text=ins["benchmark/qft/qft_8dec.qasm"]).qc2ts()
    c = host.Circuit(qasm_text=ins["benchmark/qft/qft_8crz.qasm"]).qc2ts()
    r, text = a.equiv(b, 1e-6, True)
    assert text == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n" and r.equivalent == 1
    r, text = a.equiv(c)
    assert text == "Equivalent\n- Global Norm : 1\n- Global Phase: 126π/253\n"
    one = host.Circuit(1).qc2ts()
    txtx = host.Circuit(1)
    for g in "txtx":
        txtx.append(g, [0])
    r, text = one.equiv(txtx.qc2ts())
    assert text == "Equivalent\n- Global Norm : 1\n- Global Phase: π/4\n"
    r, text = one.equiv(txtx.qc2ts(), 1e-6, True)
    assert text == "Not Equivalent\n- Cosine Similarity: 1\n"
    r, text = one.equiv(host.Circuit(2).qc2ts())
    assert text == "Not Equivalent\n- Shape Mismatch: [2, 2] vs [4, 4]\n"
    with pytest.raises(Q.QsxError) as ei:
        host.Circuit(0).qc2ts()
    assert "QCir is empty!!" in str(ei.value)
    b.adjoint_inplace()
    r, text = a.equiv(b)
    assert text.startswith("Not Equivalent\n- Cosine Similarity: ")


def _nested_example():
    inner = host.Circuit(3)
    inner.append("h", [0]).append("cx", [0, 2]).append("t", [1]).append("ccx", [2, 0, 1]).append("rz", [2], "pi/3")
    outer = host.Circuit(6)
    outer.append("h", [5]).append_circuit(inner, [4, 1, 3]).append("cx", [1, 0]).append_circuit(inner, [5, 2, 0, 3, 1], n_ctrls=2)

    def mat_of(c):
        u = np.eye(2**c.n_qubits, dtype=complex)
        for nc, qs, m in c.gate_stream():
            u = O.apply_gate_rows(u, c.n_qubits, O.mat_control(m, nc), qs)
        return u

    u_in = mat_of(inner)
    want = np.eye(64, dtype=complex)
    want = O.apply_gate_rows(want, 6, O.mat_h(), [5])
    want = O.apply_gate_rows(want, 6, u_in, [4, 1, 3])  # the reference contracts to_tensor(inner) as ONE gate tensor
    want = O.apply_gate_rows(want, 6, O.mat_control(O.mat_px(O.Phase(1)), 1), [1, 0])
    want = O.apply_gate_rows(want, 6, O.mat_control(u_in, 2), [5, 2, 0, 3, 1])
    return outer, want


def test_nested_circuit_gate_is_flattened_into_the_stream():
    """A QCir is an Operation in the reference (qcir.hpp:181-183); its gate tensor is to_tensor(inner) (qcir_to_tensor.cpp:147).
    Here the (controlled) nested circuit is flattened: every inner gate carries the outer controls -- the same operator."""
    outer, want = _nested_example()
    assert outer.n_gates == 4
    gs = outer.gate_stream()
    assert len(gs) == 12 and max(len(qs) for _, qs, _ in gs) == 5  # ccx inside a doubly controlled circuit: 4 controls + 1 target
    got = np.eye(64, dtype=complex)
    for nc, qs, m in gs:
        got = O.apply_gate_rows(got, 6, O.mat_control(m, nc), qs)
    assert np.abs(got - want).max() < 1e-15


def test_tensor_read_write_print_on_the_host(tmp_path):
    """`tensor read` / `tensor write` / `tensor print` of a host tensor need no device (tensor.hpp:326-384, sk.log:19-20)."""
    src = tmp_path / "m.csv"
    src.write_text(" 0.126496-0.02764j,  0.169178-0.977043j\n-0.169178-0.977043j,  0.126496+0.02764j\n")
    t = host.Tensor.read(src)
    assert t.shape == [2, 2]
    m = t.to_numpy()
    assert m[0, 1] == 0.169178 - 0.977043j and m[1, 0] == -0.169178 - 0.977043j
    assert t.print_string() == "{{ 0.126496-0.02764i ,  0.169178-0.977043i},\n {-0.169178-0.977043i,  0.126496+0.02764i }}"
    out = tmp_path / "w.csv"
    t.write(out)
    assert out.read_text() == " 0.126496-0.02764j,  0.169178-0.977043j\n-0.169178-0.977043j,  0.126496+0.02764j\n"
    assert abs(t.element(1, 1) - (0.126496 + 0.02764j)) == 0
    # "j" alone after one number: purely imaginary (tensor.hpp:364-367)
    src.write_text("2j, 1+0j\n 0+0j, -1j\n")
    assert host.Tensor.read(src).to_numpy()[0, 0] == 2j


def test_host_library_links_the_shared_cxx_runtime():
    """A static libstdc++ inside libqsxhost.so (what the build image's compiler wrapper produces by default) never runs the
    stream / locale initialiser of GCC 13 and crashes in the first `istream >> double`: the Makefile must link the .so."""
    import subprocess
    for path in (os.path.join(os.path.dirname(Q.lib_path()), "libqsxhost.so"), Q.lib_path()):
        needed = subprocess.run(["readelf", "-d", path], capture_output=True, text=True).stdout
        assert "libstdc++.so.6" in needed, path
