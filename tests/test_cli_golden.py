"""Replays the reference's own dofiles for the tensor-verification path through the C++ host
layer (bin/qsyn-b200: reference-shaped QCir / to_tensor / QTensor / tensor equiv on top of the
C ABI) and diffs the output against the reference's golden logs, byte for byte.
Fixtures: tests/golden/reference_goldens.json (dofile, log and input circuits)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "bin", "qsyn-b200")


def run_dof(tmp_path, goldens, dof_text):
    for rel, text in goldens["inputs"].items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(text)
    dof = tmp_path / "case.dof"
    dof.write_text(dof_text)
    r = subprocess.run([SHIM, "--no-version", "--qsynrc-path", "/dev/null", "--verbose", str(dof)], cwd=tmp_path,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return r.returncode, r.stdout


def test_shim_is_built_and_fails_loudly_without_gpu(tmp_path, goldens):
    assert os.path.exists(SHIM), "run `make` first"
    import qsyn_b200 as Q
    if Q.device_count() > 0:
        pytest.skip("a CUDA device is present")
    rc, out = run_dof(tmp_path, goldens, "qcir qubit add\nqc2ts\nquit -f\n")
    assert rc != 0 and "no CPU fallback" in out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["qc2ts", "error", "size"])
def test_full_dofile_matches_golden_log(tmp_path, goldens, case):
    """603 lines of trace-level pin bookkeeping (qc2ts), the eps ladder (error) and the shape
    mismatch (size): identical to the reference's logs."""
    c = goldens["cases"][case]
    rc, out = run_dof(tmp_path, goldens, c["dof"])
    assert rc == 0
    assert out == c["log"]


@pytest.mark.gpu
def test_global_norm_phase_qcir_half(tmp_path, goldens):
    """tests/tensor/equiv/global_norm_phase: the second half of the dofile is ZX-free
    (t,x,t,x vs the empty circuit: Equivalent, norm 1, phase pi/4; --strict: Not Equivalent)."""
    c = goldens["cases"]["global_norm_phase"]
    lines = c["dof"].split("\n")
    start = lines.index("qcir qubit add")
    rc, out = run_dof(tmp_path, goldens, "\n".join(lines[start:]))
    assert rc == 0
    assert out == c["log"][c["log"].index("qsyn> qcir qubit add"):]


@pytest.mark.gpu
def test_sk_known_answer_matrix_print(tmp_path, goldens):
    """tests/conversion/sk: rx,ry,rz with float angles -> `tensor print` shows the reference's 2x2 matrix."""
    c = goldens["cases"]["sk"]
    lines = c["dof"].split("\n")
    keep = lines[0:4] + lines[6:8] + ["quit -f"]
    rc, out = run_dof(tmp_path, goldens, "\n".join(keep))
    assert rc == 0
    log = c["log"]
    a = log.index("qsyn> convert qcir tensor")
    b = log.index("qsyn> logger info")
    assert log[a:b] in out
    assert "{{ 0.126496-0.02764i ,  0.169178-0.977043i},\n {-0.169178-0.977043i,  0.126496+0.02764i }}" in out


@pytest.mark.gpu
def test_empty_manager_messages(tmp_path, goldens):
    """tests/conversion/empty_mgr: error text with no QCir, warning for a QCir without qubits."""
    rc, out = run_dof(tmp_path, goldens, "qc2ts\n\nqcir new\nqc2ts\n\nquit -f")
    log = goldens["cases"]["empty_mgr"]["log"]
    assert rc == 0
    for piece in ["qsyn> qc2ts\n[error]    QCir list is empty. Please create a QCir first!!\n\n",
                  "qsyn> \nqsyn> qcir new\n\n", "qsyn> qc2ts\n[warn]     QCir is empty!!\n\n", "qsyn> quit -f\n\n"]:
        assert piece in log and piece in out
    assert out.endswith("qsyn> quit -f\n\n")


@pytest.mark.gpu
def test_qft_verdicts_through_the_cli(tmp_path, goldens):
    """benchmark/qft families through `qcir read` + `qc2ts` + `tensor equiv`."""
    dof = "\n".join([
        "qcir read benchmark/qft/qft_8.qasm", "qc2ts", "qcir read benchmark/qft/qft_8dec.qasm", "qc2ts",
        "qcir read benchmark/qft/qft_8crz.qasm", "qc2ts", "tensor list", "tensor equiv 0 1 --strict", "tensor equiv 0 2",
        "tensor adjoint 1", "tensor equiv 0 1", "quit -f"])
    rc, out = run_dof(tmp_path, goldens, dof)
    assert rc == 0
    assert "qsyn> tensor equiv 0 1 --strict\nEquivalent\n- Global Norm : 1\n- Global Phase: 0\n" in out
    assert "qsyn> tensor equiv 0 2\nEquivalent\n- Global Norm : 1\n- Global Phase: 126π/253\n" in out
    assert "  0    qft_8               #Dim: 2   QC2TS\n" in out
    assert "★ 2    qft_8crz            #Dim: 2   QC2TS\n" in out
    assert "qsyn> tensor equiv 0 1\nNot Equivalent\n- Cosine Similarity:" in out  # QFT is not Hermitian
