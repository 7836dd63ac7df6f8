"""Pins the CPU oracle (oracle/oracle.py) against the reference's own golden
vectors for this path.  CPU only."""
import re

import numpy as np
import pytest

from oracle import oracle as O


def _qc_from_cmds(lines):
    qc = None
    for ln in lines:
        tok = ln.split()
        if tok[:3] == ["qcir", "qubit", "add"]:
            if qc is None:
                qc = O.QCir(0)
            qc.add_qubits(int(tok[3]) if len(tok) > 3 else 1)
        elif tok[:3] == ["qcir", "gate", "add"]:
            name, ph, qs, i = tok[3], None, [], 4
            while i < len(tok):
                if tok[i] == "-ph":
                    ph = O.str_to_phase(tok[i + 1])
                    i += 2
                else:
                    qs.append(int(tok[i]))
                    i += 1
            op = O.str_to_operation(name, [ph] if ph is not None else [])
            assert op is not None, ln
            qc.append(op, qs)
    return qc


def test_sk_known_answer_matrix(goldens):
    """tests/conversion/sk/ref/sk.log:12-20 : phases and the 2x2 matrix."""
    case = goldens["cases"]["sk"]
    qc = _qc_from_cmds(case["dof"].splitlines()[:4])
    assert [g.op.repr() for g in qc.gates][0] == "rx(601π/651)"
    assert "0 (t=1): rx(601π/651)" in case["log"]
    m = O.qc2ts(qc)
    want = np.array([[0.126496 - 0.02764j, 0.169178 - 0.977043j], [-0.169178 - 0.977043j, 0.126496 + 0.02764j]])
    # xtensor prints fixed-point with 6 decimals (trailing zeros stripped: 0.027640 -> 0.02764)
    for z, w in zip(m.ravel(), want.ravel()):
        assert round(z.real, 6) == pytest.approx(w.real, abs=1e-12)
        assert round(z.imag, 6) == pytest.approx(w.imag, abs=1e-12)
    # literal tensordot algorithm == in-place row-bit form
    assert np.array_equal(m, O.to_matrix_rowform(qc))


def test_global_norm_phase(goldens):
    """tests/tensor/equiv/ref/global_norm_phase.log:53-66."""
    log = goldens["cases"]["global_norm_phase"]["log"]
    a = O.qc2ts(_qc_from_cmds(["qcir qubit add"]))
    b = O.qc2ts(_qc_from_cmds(["qcir qubit add", "qcir gate add t 0", "qcir gate add x 0", "qcir gate add t 0", "qcir gate add x 0"]))
    r = O.tensor_equiv(a, b)
    assert r.text == "Equivalent\n- Global Norm : 1\n- Global Phase: π/4\n"
    assert "qsyn> tensor equiv 0 1\n" + r.text in log
    r = O.tensor_equiv(a, b, strict=True)
    assert r.text == "Not Equivalent\n- Cosine Similarity: 1\n"
    assert "qsyn> tensor equiv 0 1 --strict\n" + r.text in log
    # the fused-sum finish gives the same report
    assert O.finish_equiv(O.equiv_sums(a, b)).text == O.tensor_equiv(a, b).text


def test_eps_ladder(goldens):
    """tests/tensor/not_equiv/ref/error.log."""
    log = goldens["cases"]["error"]["log"]
    a = O.qc2ts(_qc_from_cmds(["qcir qubit add"]))
    b = O.qc2ts(_qc_from_cmds(["qcir qubit add", "qcir gate add rz -ph 0.001 0"]))
    for eps_s in ["1e-2", "1e-4", "1e-6", "1e-8", "1e-10"]:
        r = O.tensor_equiv(a, b, eps=float(eps_s))
        assert f"qsyn> tensor equiv 0 1 --eps {eps_s}\n{r.text}\n" in log, eps_s
        assert O.finish_equiv(O.equiv_sums(a, b), eps=float(eps_s)).text == r.text


def test_shape_mismatch(goldens):
    """tests/tensor/not_equiv/ref/size.log."""
    log = goldens["cases"]["size"]["log"]
    a = O.qc2ts(_qc_from_cmds(["qcir qubit add"]))
    b = O.qc2ts(_qc_from_cmds(["qcir qubit add", "qcir qubit add"]))
    r = O.tensor_equiv(a, b)
    assert r.text == "Not Equivalent\n- Shape Mismatch: [2, 2] vs [4, 4]\n"
    assert r.text in log


def test_qc2ts_trace_log(goldens):
    """tests/conversion/qc2ts/ref/qc2ts.log : gate order (non-monotone ids),
    reprs (float angle -> rational) and the pin-permutation bookkeeping, 5 circuits."""
    case = goldens["cases"]["qc2ts"]
    paths = re.findall(r"qcir read (\S+)", case["dof"])
    blocks = re.findall(r"qsyn> qc2ts\n(.*?)\nqsyn> ", case["log"], flags=re.S)
    assert len(paths) == len(blocks) == 5
    for path, block in zip(paths, blocks):
        qc = O.qcir_from_qasm_text(goldens["inputs"][path])
        trace = []
        t = O.to_tensor_literal(qc, trace)
        want = [ln for ln in block.split("\n") if ln.startswith(("[debug]", "[trace]"))]
        assert trace == want, path
        # and the value agrees with the row-bit form to rounding
        m = O.interleaved_to_matrix(t)
        assert np.allclose(m, O.to_matrix_rowform(qc), rtol=0, atol=1e-14)


def test_empty_circuit(goldens):
    """tests/conversion/ref/empty_mgr.log : zero qubits -> warn + nullopt."""
    trace = []
    assert O.to_tensor_literal(O.QCir(0), trace) is None
    assert trace == ["[warn]     QCir is empty!!"]
    assert "[warn]     QCir is empty!!" in goldens["cases"]["empty_mgr"]["log"]


def test_phase_parsing():
    """phase.hpp:167-233 + rational.hpp:97-128; values cross-checked with qc2ts.log:15 / sk.log:17."""
    assert O.str_to_phase("2.3145031").print_string() == "277π/376"
    assert O.str_to_phase("6.28*0.46181601443868003").print_string() == "601π/651"
    assert O.str_to_phase("pi/128").print_string() == "π/128"
    assert O.str_to_phase("-pi/4").print_string() == "-π/4"
    assert O.str_to_phase("3*pi/4").print_string() == "3π/4"
    assert O.str_to_phase("pi").print_string() == "π"
    assert O.str_to_phase("0").print_string() == "0"
    assert O.str_to_phase("2*pi").print_string() == "0"
    assert O.str_to_phase("-pi").print_string() == "π"  # normalised to (-pi, pi]
    assert O.str_to_phase("abc") is None
    assert O.str_to_phase("pi/") is None
    assert O.Phase.from_float(np.pi / 4).print_string() == "π/4"


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8])
def test_qft_families(goldens, n):
    """benchmark/qft: qft_N == qft_Ndec exactly (phase 0); qft_N ~ qft_Ncrz up to a global phase."""
    ins = goldens["inputs"]
    a = O.qc2ts(O.qcir_from_qasm_text(ins[f"benchmark/qft/qft_{n}.qasm"]))
    b = O.qc2ts(O.qcir_from_qasm_text(ins[f"benchmark/qft/qft_{n}dec.qasm"]))
    c = O.qc2ts(O.qcir_from_qasm_text(ins[f"benchmark/qft/qft_{n}crz.qasm"]))
    r = O.tensor_equiv(a, b)
    assert r.text == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n"
    r = O.tensor_equiv(a, c)
    assert r.equivalent and O.fmt_g6(r.norm) == "1"
    if n == 3:
        assert r.phase.print_string() == "-5π/16"
    if n == 8:
        assert r.phase.print_string() == "126π/253"  # Stern-Brocot artefact of 255π/512
    # generator restatement == the committed benchmark file
    assert O.qft_qasm(n) == ins[f"benchmark/qft/qft_{n}.qasm"]


def test_control_matrix_and_gate_table():
    cx = O.str_to_operation("cx")
    assert cx.n_ctrls == 1 and cx.repr() == "cx"
    m = cx.matrix()
    assert m.shape == (4, 4) and abs(m[2, 3] - 1) < 1e-15 and abs(m[0, 0] - 1) == 0
    ccz = O.str_to_operation("ccz").matrix()
    assert np.allclose(np.diag(ccz), [1, 1, 1, 1, 1, 1, 1, -1])
    assert O.str_to_operation("crz", [O.Phase(1, 2)]).repr() == "crz(π/2)"
    assert O.str_to_operation("cp", [O.Phase(1, 2)]).repr() == "cs"
    assert O.str_to_operation("mcx") is None  # silently skipped by the reader (Appendix C trap 7)
    # Px(pi) is not an exact permutation (Appendix C trap 3)
    x = O.mat_px(O.Phase(1))
    assert x[0, 1].real == 0.9999999999999998
    # every supported gate is unitary to rounding
    for name in ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "sy", "swap", "ecr", "id"]:
        u = O.str_to_operation(name).matrix()
        assert np.allclose(u @ u.conj().T, np.eye(u.shape[0]), atol=1e-15), name
    for name in ["rz", "rx", "ry", "px", "py", "p"]:
        u = O.str_to_operation(name, [O.Phase(3, 7)]).matrix()
        assert np.allclose(u @ u.conj().T, np.eye(2), atol=1e-15), name


def test_literal_equals_rowform_random():
    """Appendix A.1: literal tensordot+pins == in-place row-bit form on random circuits."""
    for seed in range(6):
        qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(5, 60, 100 + seed))
        assert len(qc.gates) == 60
        a = O.qc2ts(qc)
        b = O.to_matrix_rowform(qc)
        assert np.allclose(a, b, rtol=0, atol=1e-13)
        # column shard of the row form
        s = O.to_matrix_rowform(qc, col0=8, ncols=8)
        assert np.array_equal(s, b[:, 8:16])


def test_c_restatement_matches_numpy():
    """oracle/oracle_c.c (the plain-C port used as a CPU baseline) == oracle.py."""
    import cmath

    from oracle import oracle_c
    from tests.test_abi_cpu import _rand_circuit

    for n, g, seed in [(1, 10, 1), (3, 50, 2), (6, 200, 3)]:
        qc = _rand_circuit(n, g, seed)
        want = O.to_matrix_rowform(qc)
        got = oracle_c.to_matrix(O.gate_stream(qc), n)
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-14
        half = oracle_c.to_matrix(O.gate_stream(qc), n, 2**n // 2, 2**n // 2)
        assert np.array_equal(half, got[:, 2**n // 2:])
        other = O.to_matrix_rowform(_rand_circuit(n, g, seed + 100))
        cos, _, _ = oracle_c.equiv_reference_passes(want, other)
        assert cos == pytest.approx(O.cosine_similarity(want, other), rel=1e-12)
        # the sum ratio is only well conditioned between near-proportional tensors (Appendix C trap 2)
        other = want * cmath.exp(0.3j) * 0.75
        cos, norm, ph = oracle_c.equiv_reference_passes(want, other)
        if abs(np.sum(want)) > 1e-6:
            assert cos == pytest.approx(1.0, abs=1e-12)
            assert norm == pytest.approx(0.75, rel=1e-9) and ph == pytest.approx(0.3, abs=1e-9)


def test_column_block_form_of_the_literal_algorithm():
    """to_tensor_literal_columns (the bounded CPU-baseline sample of bench.py at n >= 14) is the same per-gate
    tensordot loop with the input axes collapsed: bit-identical to the row form on any column block."""
    from tests.test_abi_cpu import _rand_circuit
    for n, g, seed, col0, ncols in [(3, 40, 1, 0, 8), (6, 120, 2, 16, 16), (7, 150, 3, 100, 28)]:
        qc = _rand_circuit(n, g, seed)
        want = O.to_matrix_rowform(qc)[:, col0:col0 + ncols]
        got = O.to_tensor_literal_columns(qc, col0, ncols)
        assert np.array_equal(got, want)
        assert np.allclose(O.qc2ts(qc)[:, col0:col0 + ncols], got, rtol=0, atol=1e-14)
