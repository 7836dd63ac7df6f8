"""One process, several GPUs: the path qsyn itself would take (a single process with one worker thread per command,
reference src/cli/cli_exec.cpp:192-217).  Column shards live on different devices, construction needs no
communication, `tensor equiv` is one qsx_equiv_reduce = concurrent per-device reductions + ONE ncclAllReduce of
8 doubles (communicators from ncclCommInitAll), and `tensor adjoint` exchanges blocks through peer memory.
Skipped on a one-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi_device`)."""
import math
import os
import subprocess

import numpy as np
import pytest

import qsyn_b200 as Q
from oracle import oracle as O
from qsyn_b200 import abi, workloads
from tests.test_abi_cpu import _rand_circuit
from tests.test_cli_golden import SHIM

pytestmark = pytest.mark.gpu


def _ndev():
    return Q.device_count()


needs2 = pytest.mark.skipif(_ndev() < 2, reason="needs >= 2 GPUs in one process")


def _sharded(n, gs, shards, devices):
    ncols = 2**n // shards
    us = []
    for k in range(shards):
        u = Q.Unitary(n, devices[k % len(devices)], k * ncols, ncols)
        u.apply_async(gs)
        us.append(u)
    return us


def test_equiv_reduce_on_one_device_is_the_sum_of_the_partials():
    """(runs on any box) several shard pairs on one device: one finish over all their partial blocks."""
    n, shards = 8, 4
    qa, qb = _rand_circuit(n, 150, 61), _rand_circuit(n, 150, 62)
    A, B = O.to_matrix_rowform(qa), O.to_matrix_rowform(qb)
    ua, ub = _sharded(n, O.gate_stream(qa), shards, [0]), _sharded(n, O.gate_stream(qb), shards, [0])
    total = Q.equiv_reduce(ua, ub)
    parts = sum(a.equiv_partial(b) for a, b in zip(ua, ub))
    want = O.equiv_sums(A, B)
    assert np.all(np.abs(total - want) < 1e-11) and np.all(np.abs(total - parts) < 1e-11)
    with pytest.raises(Q.QsxError) as ei:  # pairwise geometry is checked before anything is launched
        Q.equiv_reduce([ua[0], ua[1]], [ub[0], ub[2]])
    assert ei.value.status == "QSX_ERR_SHAPE"
    world, version = abi.comm_info()
    assert world == 1


@needs2
@pytest.mark.parametrize("n,shards", [(8, 2), (9, 4), (10, 8), (7, 16)])
def test_equiv_reduce_across_devices_uses_one_allreduce(n, shards):
    ndev = _ndev()
    devices = list(range(min(ndev, shards)))
    qa, qb = _rand_circuit(n, 200, 71), _rand_circuit(n, 200, 72)
    A, B = O.to_matrix_rowform(qa), O.to_matrix_rowform(qb)
    ua, ub = _sharded(n, O.gate_stream(qa), shards, devices), _sharded(n, O.gate_stream(qb), shards, devices)
    assert len({u.info.device for u in ua}) == len(devices) > 1
    total = Q.equiv_reduce(ua, ub)
    want = O.equiv_sums(A, B)
    assert np.all(np.abs(total - want) < 1e-10), total - want
    r, w = Q.equiv_finish(total), O.tensor_equiv(A, B)
    assert bool(r.equivalent) == w.equivalent and abs(r.cosine - w.cosine) < 1e-12
    # every shard's data is what a single-GPU build gives
    ncols = 2**n // shards
    for k in (0, shards - 1):
        assert np.allclose(ua[k].download(), A[:, k * ncols:(k + 1) * ncols], rtol=0, atol=1e-13)
    # NCCL was really used: a version code is reported once the library is loaded
    assert abi.comm_info()[1] >= 21000


@needs2
def test_qsx_init_builds_the_clique_eagerly():
    abi.init(0)
    n = 9
    qc = _rand_circuit(n, 120, 5)
    gs = O.gate_stream(qc)
    tail = [(0, [0], O.str_to_operation(g).target_matrix()) for g in "txtx"]
    ndev = _ndev()
    ua = _sharded(n, gs, ndev if ndev & (ndev - 1) == 0 else 2, list(range(ndev)))
    ub = _sharded(n, gs + tail, len(ua), list(range(ndev)))
    r = Q.equiv_finish(Q.equiv_reduce(ua, ub))
    assert r.equivalent == 1 and (r.phase_num, r.phase_den) == (1, 4) and abs(r.norm - 1) < 1e-12


@needs2
def test_adjoint_across_real_peers():
    n = 9
    ndev = _ndev()
    shards = 2 if ndev < 4 else 4
    rng = np.random.default_rng(9)
    dim = 2**n
    m = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    c = dim // shards
    us = []
    for k in range(shards):
        u = Q.Unitary(n, k % ndev, k * c, c, identity=False)
        u.upload(np.ascontiguousarray(m[:, k * c:(k + 1) * c]))
        us.append(u)
    assert len({u.info.device for u in us}) > 1
    Q.adjoint_sharded(us)
    want = m.conj().T
    for k, u in enumerate(us):
        assert np.array_equal(u.download(), want[:, k * c:(k + 1) * c]), k


def _run_shim(tmp_path, files, dof_text, n_gpus):
    for rel, text in files.items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(text)
    dof = tmp_path / "case.dof"
    dof.write_text(dof_text)
    env = dict(os.environ, QSX_NUM_GPUS=str(n_gpus))
    # stdout is the product's text interface; stderr is kept apart (NCCL prints its version banner there when the image
    # exports NCCL_DEBUG=VERSION -- comm.cpp points fd 1 at fd 2 while a communicator is created)
    r = subprocess.run([SHIM, "--no-version", "--qsynrc-path", "/dev/null", "--verbose", str(dof)], cwd=tmp_path, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    return r.returncode, r.stdout if r.returncode == 0 else r.stdout + r.stderr


def _qft_dec_qasm(n):
    """benchmark/qft/generateQFTqasm.py's decomposed variant: cp(t) c,t -> p(t/2) c; p(t/2) t; cx; p(-t/2) t; cx."""
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];"]
    for i in range(n):
        lines.append(f"h q[{i}];")
        for j in range(i + 1, n):
            d = 2 ** (j - i + 1)
            lines += [f"p(pi/{d}) q[{j}];", f"p(pi/{d}) q[{i}];", f"cx q[{j}],q[{i}];", f"p(-pi/{d}) q[{i}];", f"cx q[{j}],q[{i}];"]
    return "\n".join(lines) + "\n"


@needs2
@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_cli_with_several_gpus_in_one_process(tmp_path, goldens, n_gpus):
    """The reference's commands, QSX_NUM_GPUS devices behind every QTensor: same text as on one GPU."""
    if n_gpus > _ndev():
        pytest.skip(f"{n_gpus} GPUs not present")
    files = dict(goldens["inputs"])
    files["c3/qft_14.qasm"] = workloads.qft_qasm(14)
    files["c3/qft_14dec.qasm"] = _qft_dec_qasm(14)
    dof = "\n".join([
        "qcir read benchmark/qft/qft_8.qasm", "qc2ts", "qcir read benchmark/qft/qft_8dec.qasm", "qc2ts",
        "qcir read benchmark/qft/qft_8crz.qasm", "qc2ts", "tensor equiv 0 1 --strict", "tensor equiv 0 2",
        "tensor adjoint 1", "tensor equiv 0 1", "tensor adjoint 1", "tensor equiv 0 1 --strict",
        "qcir read c3/qft_14.qasm", "qc2ts", "qcir read c3/qft_14dec.qasm", "qc2ts", "tensor equiv 3 4 --strict",
        "quit -f"])
    rc, out = _run_shim(tmp_path, files, dof, n_gpus)
    assert rc == 0, out[-2000:]
    assert "qsyn> tensor equiv 0 1 --strict\nEquivalent\n- Global Norm : 1\n- Global Phase: 0\n" in out
    assert "qsyn> tensor equiv 0 2\nEquivalent\n- Global Norm : 1\n- Global Phase: 126π/253\n" in out
    assert "qsyn> tensor equiv 0 1\nNot Equivalent\n- Cosine Similarity:" in out
    assert out.count("qsyn> tensor equiv 0 1 --strict\nEquivalent\n- Global Norm : 1\n- Global Phase: 0\n") == 2
    assert "qsyn> tensor equiv 3 4 --strict\nEquivalent\n- Global Norm : 1\n- Global Phase: 0\n" in out
    one_rc, one_out = _run_shim(tmp_path, files, dof, 1)
    assert one_rc == 0 and one_out == out


@needs2
def test_golden_dofiles_with_qsx_num_gpus(tmp_path, goldens):
    """The golden logs are n <= 5 circuits: the shard rule (>= 64 columns per shard) keeps them on one device, and
    the text is byte-identical with QSX_NUM_GPUS set."""
    for case in ("qc2ts", "error", "size"):
        c = goldens["cases"][case]
        rc, out = _run_shim(tmp_path, goldens["inputs"], c["dof"], 2)
        assert rc == 0 and out == c["log"], case
