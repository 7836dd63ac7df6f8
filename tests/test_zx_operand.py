"""The ZX operand of the verification path: `.zx` reader, ZXGraph construction
and ZX -> tensor.

* the CPU oracle (oracle/zx_oracle.py) is pinned on the reference's golden logs, whose trace
  output exposes the topological order, every frontier and every axis id;
* the C++ host layer (bin/qsyn-b200) replays the reference's dofiles byte for byte -- on CPU the
  ones that stop before `tensor equiv`, on the GPU (marker) the complete ones;
* `tensor equiv` of ZX-built tensors runs on the device (no CPU fallback)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import zx_oracle as Z

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "bin", "qsyn-b200")
_LEVEL = {"trace": 0, "debug": 1, "info": 2, "warn": 3, "error": 4}


def oracle_replay(dof_text, inputs):
    """The log lines (only) the reference prints for a zx dofile, from the oracle."""
    graphs, tensors, out = [], [], []
    cur, level = -1, 3
    for raw in dof_text.splitlines():
        t = raw.split("//")[0].split()
        if not t:
            continue
        log = []
        if t[:2] == ["zx", "new"]:
            graphs.append(Z.ZXGraph())
            cur = len(graphs) - 1
        elif t[:2] == ["zx", "read"]:
            graphs.append(Z.zxgraph_from_zx_text(inputs[t[2]]))
            cur = len(graphs) - 1
        elif t[:2] == ["zx", "checkout"]:
            cur = int(t[2])
        elif t[:2] == ["zx", "tensor-product"]:
            graphs[cur].tensor_product(graphs[int(t[2])])
        elif t[:3] == ["zx", "vertex", "add"]:
            g, k = graphs[cur], t[3][0]
            if k == "i":
                g.add_input(int(t[4]))
            elif k == "o":
                g.add_output(int(t[4]))
            else:
                ph = O.str_to_phase(t[4]) if len(t) > 4 else (O.Phase(1) if k == "h" else O.Phase(0))
                g.add_vertex({"z": Z.Z, "x": Z.X, "h": Z.HBOX}[k], ph)
        elif t[:3] == ["zx", "edge", "add"]:
            g = graphs[cur]
            g.add_edge(g.id_to_vertex[int(t[3])], g.id_to_vertex[int(t[4])], Z.SIMPLE if t[5][0] in "Ss" else Z.HADAMARD)
        elif t[:3] == ["zx", "edge", "remove"]:
            g = graphs[cur]
            g.remove_edge(g.id_to_vertex[int(t[3])], g.id_to_vertex[int(t[4])])
        elif t[0] == "zx2ts":
            log.append(f"[info]     Converting ZXGraph {cur} to Tensor {len(tensors)}...")
            m = Z.zx2ts(graphs[cur], log)
            if m is not None:
                log.append(f"[info]     Successfully created and checked out to Tensor {len(tensors)}")
                tensors.append(m)
        elif t[0] == "logger":
            level = _LEVEL[t[1]]
            log.append(f'[info]     Setting logger level to "{t[1]}"')
        out += [line for line in log if _LEVEL[line[1:line.index("]")]] >= level]
    return out, tensors


@pytest.mark.parametrize("case", ["zx2ts_h_edge", "zx2ts_disconnected", "zx2ts_invalid_graph"])
def test_oracle_reproduces_the_reference_trace_logs(goldens, case):
    """130 / 210 / 2 log lines: DFS order, frontier order after erasures, axis ids after every
    tensordot, Hadamard-edge handling, boundary-to-boundary wires, tensor-product id renumbering."""
    c = goldens["cases"][case]
    got, _ = oracle_replay(c["dof"], goldens["inputs"])
    want = [line for line in c["log"].splitlines() if line.startswith("[")]
    assert len(want) > 0 and got == want


def test_oracle_zx_tensors_match_the_golden_verdicts(goldens):
    zx = lambda name: Z.zx2ts(Z.zxgraph_from_zx_text(goldens["inputs"][f"benchmark/zx/{name}.zx"]))
    # tests/tensor/equiv/ref/tiny_circuits.log: 3 x CX == swap.zx (strict), tof3 == tof3_zh (strict)
    qc = O.QCir(2)
    for a, b in [(0, 1), (1, 0), (0, 1)]:
        qc.append(O.str_to_operation("cx"), [a, b])
    r = O.tensor_equiv(O.to_matrix_rowform(qc), zx("swap"), 1e-6, True)
    assert r.text == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n"
    for strict in (False, True):
        assert O.tensor_equiv(zx("tof3"), zx("tof3_zh"), 1e-6, strict).text == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n"
    # examples/tensor-equiv.dof (BASELINE.json configs[0]): tof3_op vs tof3_zh
    assert O.tensor_equiv(zx("tof3_op"), zx("tof3_zh")).text == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n"
    # the Toffoli itself: |110> <-> |111>
    t = zx("tof3")
    want = np.eye(8)
    want[[6, 7]] = want[[7, 6]]
    k = t[0, 0]
    assert abs(abs(k) - 1) < 1e-12 and np.allclose(t / k, want, atol=1e-12)
    # tests/tensor/manip/ref/adjoint.log: zx adjoint == tensor adjoint (a 4 x 2 isometry)
    g = Z.ZXGraph()
    g.add_input(0), g.add_output(0), g.add_output(1), g.add_vertex(Z.Z, O.str_to_phase("pi/2"))
    for a in (0, 1, 2):
        g.add_edge(g.id_to_vertex[a], g.id_to_vertex[3], Z.SIMPLE)
    m0 = Z.zx2ts(g)
    g.adjoint()
    m1 = Z.zx2ts(g)
    assert m0.shape == (4, 2) and m1.shape == (2, 4)
    assert O.tensor_equiv(m0.conj().T, m1, 1e-6, True).equivalent


def test_zx_reader_edge_cases():
    g = Z.zxgraph_from_zx_text("I0 (0,0) S2 // c\nO1 (0,2) S2\n\nZ2 (0,1) S0 S1 pi/4\n")
    assert [v.id for v in g.vertices] == [0, 1, 2] and g.id_to_vertex[2].phase == O.Phase(1, 4)
    # explicit qubit ids, keep_id, H box default phase pi, parallel Z-Z hadamard edges cancel
    g = Z.zxgraph_from_zx_text("I5 S7 3\nO6 S7 3\nZ7\nH9\n", keep_id=True)
    assert sorted(v.id for v in g.vertices) == [5, 6, 7, 9] and g.input_list[3].id == 5
    g = Z.ZXGraph()
    a, b = g.add_vertex(Z.Z), g.add_vertex(Z.Z)
    g.add_edge(a, b, Z.HADAMARD)
    g.add_edge(a, b, Z.HADAMARD)
    assert not g.is_neighbor(a, b)
    g.add_edge(a, b, Z.SIMPLE)
    g.add_edge(a, b, Z.HADAMARD)  # simple kept, hadamard becomes a pi phase
    assert g.get_edge_type(a, b) == Z.SIMPLE and a.phase == O.Phase(1)
    for bad in ["Q0", "Z", "Zx", "I0 (0 S1", "I0 0,0) S1", "Z0 (,1)", "Z0 T1", "Z0 S", "Z0\nZ0", "I0\nI1 (1,0) S0 0\nI2 (2,0) S0 0"]:
        assert Z.zxgraph_from_zx_text(bad) is None, bad


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


@pytest.mark.parametrize("case", ["zx2ts_disconnected", "zx2ts_invalid_graph", "zx_add_remove", "zx_tensor_product"])
def test_host_layer_replays_zx_dofiles_without_a_device(tmp_path, goldens, case):
    """ZX -> tensor is host work (a handful of axes): these dofiles never reach `tensor equiv`,
    so the C++ host layer must reproduce the golden logs byte for byte on any machine
    (zx_add_remove / zx_tensor_product: vertex and edge removal, id renumbering, `zx print -v/-i/-o`)."""
    c = goldens["cases"][case]
    rc, out = run_dof(tmp_path, goldens, c["dof"])
    assert rc == 0 and out == c["log"]


def test_zx_built_tensors_are_verified_on_the_device_only(tmp_path, goldens):
    import qsyn_b200 as Q
    if Q.device_count() > 0:
        pytest.skip("a CUDA device is present")
    c = goldens["cases"]["zx2ts_h_edge"]
    rc, out = run_dof(tmp_path, goldens, c["dof"])
    cut = c["log"].index("qsyn> qc2ts") + len("qsyn> qc2ts\n")  # the first command that needs the device
    assert out.startswith(c["log"][:cut]) and rc != 0 and "no CPU fallback" in out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["zx2ts_h_edge", "tiny_circuits", "tensor_adjoint"])
def test_zx_dofiles_match_golden_logs_on_device(tmp_path, goldens, case):
    c = goldens["cases"][case]
    rc, out = run_dof(tmp_path, goldens, c["dof"])
    assert rc == 0 and out == c["log"]


@pytest.mark.gpu
def test_baseline_config0_tensor_equiv_example(tmp_path, goldens):
    """examples/tensor-equiv.dof (BASELINE.json configs[0]) minus its `zx print -v` lines."""
    dof = "\n".join(["zx read benchmark/zx/tof3_op.zx", "zx read benchmark/zx/tof3_zh.zx", "convert zx tensor", "zx checkout 0",
                     "convert zx tensor", "tensor list", "tensor equiv 0 1", "quit -f", ""])
    rc, out = run_dof(tmp_path, goldens, dof)
    assert rc == 0
    assert "★ 1    tof3_op             #Dim: 2   ZX2TS\n" in out and "  0    tof3_zh             #Dim: 2   ZX2TS\n" in out
    assert "qsyn> tensor equiv 0 1\nEquivalent\n- Global Norm : 1\n- Global Phase: 0\n" in out


def _parse_printed_matrix(text):
    """xtensor's print of a complex matrix: {{a+bi, ...},\n {...}} with 6 decimals."""
    import re
    rows = []
    for line in text.strip().splitlines():
        vals = re.findall(r"(-?\d*\.\d*(?:e[-+]?\d+)?) *([-+]) *(\d*\.\d*(?:e[-+]?\d+)?)i", line)
        rows.append([complex(float(a), float(b) * (-1 if s == "-" else 1)) for a, s, b in vals])
    return np.array(rows)


@pytest.mark.parametrize("name", ["swap", "cnot", "tof3", "tof3_zh", "tof3_op"])
def test_host_layer_zx_tensor_values_match_oracle(tmp_path, goldens, name):
    """`zx read; zx2ts; tensor print` through the C++ host layer vs the oracle's matrix (6 printed decimals)."""
    rc, out = run_dof(tmp_path, goldens, f"zx read benchmark/zx/{name}.zx\nzx2ts\ntensor print\nquit -f\n")
    assert rc == 0
    body = out[out.index("qsyn> tensor print\n") + len("qsyn> tensor print\n"):out.index("\n\nqsyn> quit -f")]
    got = _parse_printed_matrix(body)
    want = Z.zx2ts(Z.zxgraph_from_zx_text(goldens["inputs"][f"benchmark/zx/{name}.zx"]))
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=0, atol=1.5e-6)


@pytest.mark.parametrize("name", ["swap", "cnot", "tof3", "tof3_zh", "tof3_op"])
def test_host_c_abi_zx2ts_matches_oracle_to_rounding(goldens, name):
    """qsxh_zxgraph_from_zx_text + qsxh_zx2ts (include/qsx_host.h): full double precision this time."""
    from qsyn_b200 import host
    text = goldens["inputs"][f"benchmark/zx/{name}.zx"]
    g = host.ZXGraph(zx_text=text)
    og = Z.zxgraph_from_zx_text(text)
    assert g.info == {"inputs": len(og.inputs), "outputs": len(og.outputs), "vertices": len(og.vertices),
                      "edges": sum(len(v.nbrs) for v in og.vertices) // 2}
    got = g.zx2ts().to_numpy()
    want = Z.zx2ts(og)
    assert got.shape == want.shape and np.allclose(got, want, rtol=0, atol=1e-14)
    with pytest.raises(Exception):
        host.ZXGraph(zx_text="Q0 S1\n")


@pytest.mark.gpu
def test_host_c_abi_zx_vs_qcir_equivalence_runs_on_the_device(goldens):
    from qsyn_b200 import host
    zx = host.ZXGraph(zx_text=goldens["inputs"]["benchmark/zx/swap.zx"]).zx2ts()
    qc = host.Circuit(n_qubits=2)
    for a, b in [(0, 1), (1, 0), (0, 1)]:
        qc.append("cx", [a, b])
    r, text = qc.qc2ts().equiv(zx, 1e-6, True)
    assert r.equivalent == 1 and text == "Equivalent\n- Global Norm : 1\n- Global Phase: 0\n"
