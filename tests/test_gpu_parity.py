"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200."""
import math

import numpy as np
import pytest

import qsyn_b200 as Q
from oracle import oracle as O
from tests.test_abi_cpu import _rand_circuit

pytestmark = pytest.mark.gpu

# north_star: Frobenius relative error <= 1e-10 vs the reference; we hold the kernels to a much
# tighter bar at test sizes so that the budget is left for 10k-gate circuits.
TOL = 1e-12


def frob_rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def run_device(qc, opt=None, col0=0, ncols=None):
    u = Q.Unitary(qc.n_qubits, 0, col0, ncols)
    u.apply(O.gate_stream(qc), opt)
    out = u.download()
    u.close()
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 9])
def test_identity_and_shards(n):
    dim = 2**n
    u = Q.Unitary(n)
    assert np.array_equal(u.download(), np.eye(dim))
    info = u.info
    assert (info.n_qubits, info.nrows, info.ncols, info.bytes) == (n, dim, dim, 16 * dim * dim)
    for ncols in {1, max(1, dim // 4), dim}:
        for col0 in {0, dim - ncols}:
            s = Q.Unitary(n, 0, col0, ncols)
            assert np.array_equal(s.download(), np.eye(dim)[:, col0:col0 + ncols])
            # partial row download
            assert np.array_equal(s.download(dim // 2, dim - dim // 2), np.eye(dim)[dim // 2:, col0:col0 + ncols])


GATES_1Q = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "sxdg", "sy", "sydg", "tx", "ty", "id"]
GATES_1P = ["rz", "rx", "ry", "p", "px", "py"]
GATES_2Q = ["cx", "cz", "cy", "ch", "swap", "ecr", "csx", "ct"]
GATES_2P = ["cp", "crz", "crx", "cry"]
GATES_3Q = ["ccx", "ccz", "cswap", "cecr", "ccy", "cch"]


def _all_single_gate_circuits(n):
    ph = O.Phase(3, 7)
    for name in GATES_1Q:
        for q in range(n):
            yield name, [q], O.str_to_operation(name)
    for name in GATES_1P:
        for q in range(n):
            yield name, [q], O.str_to_operation(name, [ph])
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    for name in GATES_2Q:
        for a, b in pairs:
            yield name, [a, b], O.str_to_operation(name)
    for name in GATES_2P:
        for a, b in pairs[::3]:
            yield name, [a, b], O.str_to_operation(name, [ph])
    trip = [(a, b, c) for a in range(n) for b in range(n) for c in range(n) if len({a, b, c}) == 3]
    for name in GATES_3Q:
        for t in trip[::5]:
            yield name, list(t), O.str_to_operation(name)
    yield "cccx", [0, 2, 3, 1], O.str_to_operation("cccx")
    yield "ccswap", [3, 0, 1, 2], O.str_to_operation("ccswap")


@pytest.mark.parametrize("reg_qubits", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("snap", [1e-15, -1.0])
def test_every_gate_kind_every_position(reg_qubits, snap):
    """One gate on a NON-trivial input (a random unitary prefix) so that misplaced rows show up."""
    n = max(4, reg_qubits)
    prefix = _rand_circuit(n, 30, 77, extra=True)
    base = O.to_matrix_rowform(prefix)
    opt = Q.PlanOptions(reg_qubits=reg_qubits, snap_tol=snap)
    u0 = Q.Unitary(n)
    u0.upload(base)
    worst = 0.0
    for name, qs, op in _all_single_gate_circuits(n):
        u = u0.clone()
        u.apply([(op.n_ctrls, qs, op.target_matrix())], opt)
        got = u.download()
        u.close()
        want = O.apply_gate_rows(base, n, op.matrix(), qs)
        err = frob_rel(got, want)
        worst = max(worst, err)
        assert err < 2e-15 if snap < 0 else err < 1e-15 * 4, (name, qs, err)
    assert worst < 4e-15


@pytest.mark.parametrize("n,g,seed", [(1, 10, 11), (2, 40, 12), (3, 60, 13), (5, 150, 14), (7, 300, 15), (9, 300, 16)])
@pytest.mark.parametrize("fuse", [0, 1])
def test_random_circuits_match_oracle(n, g, seed, fuse):
    qc = _rand_circuit(n, g, seed)
    want = O.to_matrix_rowform(qc)
    for opt in (None if fuse else Q.PlanOptions(fuse=0),
                Q.PlanOptions(fuse=fuse, reg_qubits=3, tile_qubits=6, max_ops_per_sweep=17),
                Q.PlanOptions(fuse=fuse, reg_qubits=2, tile_qubits=4, log2_tile_cols=1, snap_tol=-1.0),
                Q.PlanOptions(fuse=fuse, reg_qubits=4, tile_qubits=8, log2_tile_cols=4, max_ops_per_sweep=200),
                Q.PlanOptions(fuse=fuse, reg_qubits=1, tile_qubits=3, log2_tile_cols=-1),
                Q.PlanOptions(fuse=fuse, reg_qubits=5, tile_qubits=9, stage_search=1),
                Q.PlanOptions(fuse=fuse, no_perm_folding=1), Q.PlanOptions(fuse=fuse, no_perm_folding=2, tile_qubits=7),
                Q.PlanOptions(fuse=fuse, no_perm_folding=2, reg_qubits=3, max_ops_per_sweep=200)):
        got = run_device(qc, opt)
        assert frob_rel(got, want) < TOL


@pytest.mark.parametrize("n,shards", [(4, 2), (6, 8), (8, 4), (8, 64)])
def test_column_shards_are_independent(n, shards):
    """every shard runs the same gate stream with no communication."""
    qc = _rand_circuit(n, 200, 21)
    want = O.to_matrix_rowform(qc)
    ncols = 2**n // shards
    for k in {0, 1, shards - 1}:
        got = run_device(qc, None, k * ncols, ncols)
        assert frob_rel(got, want[:, k * ncols:(k + 1) * ncols]) < TOL
        assert np.allclose(got, O.to_matrix_rowform(qc, k * ncols, ncols), rtol=0, atol=1e-13)


def test_qft_and_literal_reference_algorithm(goldens):
    """Device result vs the LITERAL per-gate tensordot algorithm of the reference (qc2ts)."""
    for name in ["benchmark/qft/qft_5.qasm", "benchmark/qft/qft_8.qasm", "benchmark/qft/qft_8dec.qasm",
                 "benchmark/qft/qft_8crz.qasm", "benchmark/qasm/qc2ts/toffoli_qiskit.qasm",
                 "benchmark/SABRE/small/3_17_13.qasm"]:
        qc = O.qcir_from_qasm_text(goldens["inputs"][name])
        want = O.qc2ts(qc)
        assert frob_rel(run_device(qc), want) < TOL, name


def _device_equiv(a_host, b_host, eps=1e-6, strict=False):
    n = int(round(math.log2(a_host.shape[0])))
    a, b = Q.Unitary(n, identity=False), Q.Unitary(n, identity=False)
    a.upload(a_host)
    b.upload(b_host)
    sums = a.equiv_partial(b)
    return sums, Q.equiv_finish(sums, eps, strict)


def test_equiv_sums_match_compensated_oracle():
    rng = np.random.default_rng(5)
    for n in (1, 3, 6, 9):
        a = rng.standard_normal((2**n, 2**n)) + 1j * rng.standard_normal((2**n, 2**n))
        b = rng.standard_normal((2**n, 2**n)) + 1j * rng.standard_normal((2**n, 2**n))
        sums, _ = _device_equiv(a, b)
        want = O.equiv_sums(a, b)
        scale = np.array([4.0**n] * 4 + [2.0**n * 2] * 4)
        assert np.all(np.abs(sums - want) <= 1e-15 * scale * 8), (n, sums - want)


def test_golden_equiv_cases_on_device(goldens):
    """tests/tensor/equiv/ref/global_norm_phase.log, tests/tensor/not_equiv/ref/{error,size}.log"""
    ident = Q.Unitary(1)
    qc = O.QCir(1)
    for g in "txtx":
        qc.append(O.str_to_operation(g), [0])
    b = Q.Unitary(1)
    b.apply(O.gate_stream(qc))
    r = Q.equiv_finish(ident.equiv_partial(b))
    assert r.equivalent == 1 and O.fmt_g6(r.norm) == "1" and (r.phase_num, r.phase_den) == (1, 4)
    r = Q.equiv_finish(ident.equiv_partial(b), 1e-6, True)
    assert r.equivalent == 0 and O.fmt_g6(r.cosine) == "1"
    qc = O.QCir(1)
    qc.append(O.str_to_operation("rz", [O.str_to_phase("0.001")]), [0])
    c = Q.Unitary(1)
    c.apply(O.gate_stream(qc))
    for eps, want in [(1e-2, 1), (1e-4, 1), (1e-6, 1), (1e-8, 0), (1e-10, 0)]:
        r = Q.equiv_finish(ident.equiv_partial(c), eps)
        assert r.equivalent == want, eps
        if want:
            assert O.fmt_g6(r.norm) == "1" and (r.phase_num, r.phase_den) == (0, 1)
    with pytest.raises(Q.QsxError) as ei:
        ident.equiv_partial(Q.Unitary(2))
    assert ei.value.status == "QSX_ERR_SHAPE"


def test_qft_equivalence_verdicts_on_device(goldens):
    def dev(name):
        qc = O.qcir_from_qasm_text(goldens["inputs"][name])
        u = Q.Unitary(qc.n_qubits)
        u.apply(O.gate_stream(qc))
        return u
    a, b, c = dev("benchmark/qft/qft_8.qasm"), dev("benchmark/qft/qft_8dec.qasm"), dev("benchmark/qft/qft_8crz.qasm")
    r = Q.equiv_finish(a.equiv_partial(b), 1e-6, True)
    assert r.equivalent == 1 and (r.phase_num, r.phase_den) == (0, 1) and abs(r.norm - 1) < 1e-9
    r = Q.equiv_finish(a.equiv_partial(c))
    assert r.equivalent == 1 and (r.phase_num, r.phase_den) == (126, 253) and abs(r.norm - 1) < 1e-9
    assert abs(r.phase_rad - 255 * math.pi / 512) < 1e-9


def test_sharded_equiv_sums_add_up():
    n, shards = 7, 4
    qc1, qc2 = _rand_circuit(n, 100, 31), _rand_circuit(n, 100, 32)
    A, B = O.to_matrix_rowform(qc1), O.to_matrix_rowform(qc2)
    total = np.zeros(8)
    ncols = 2**n // shards
    for k in range(shards):
        a = Q.Unitary(n, 0, k * ncols, ncols)
        b = Q.Unitary(n, 0, k * ncols, ncols)
        a.apply(O.gate_stream(qc1))
        b.apply(O.gate_stream(qc2))
        total += a.equiv_partial(b)
    want = O.equiv_sums(A, B)
    assert np.all(np.abs(total - want) < 1e-10)
    r, w = Q.equiv_finish(total), O.tensor_equiv(A, B)
    assert bool(r.equivalent) == w.equivalent and abs(r.cosine - w.cosine) < 1e-12


def test_adjoint_inplace():
    for n in (1, 3, 5, 7):
        qc = _rand_circuit(n, 80, 40 + n)
        m = O.to_matrix_rowform(qc)
        u = Q.Unitary(n, identity=False)
        u.upload(m)
        u.adjoint_inplace()
        assert np.array_equal(u.download(), m.conj().T)
    s = Q.Unitary(3, 0, 0, 4)
    with pytest.raises(Q.QsxError) as ei:
        s.adjoint_inplace()
    assert ei.value.status == "QSX_ERR_UNSUPPORTED"


@pytest.mark.parametrize("n,shards", [(2, 2), (5, 4), (7, 2), (8, 8), (9, 4), (6, 64)])
def test_adjoint_of_a_column_sharded_unitary(n, shards):
    """block (j, k) <-> block (k, j), in place.  Shards are spread round-robin over the
    visible devices (all on device 0 on a one-GPU box; over NVLink peers with --gpus 2+)."""
    rng = np.random.default_rng(n * 100 + shards)
    dim = 2**n
    m = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    c = dim // shards
    ndev = Q.device_count()
    us = []
    for k in range(shards):
        u = Q.Unitary(n, k % ndev, k * c, c, identity=False)
        u.upload(np.ascontiguousarray(m[:, k * c:(k + 1) * c]))
        us.append(u)
    Q.adjoint_sharded(us)
    want = m.conj().T
    for k, u in enumerate(us):
        assert np.array_equal(u.download(), want[:, k * c:(k + 1) * c]), k
    Q.adjoint_sharded(us)  # an involution
    for k, u in enumerate(us):
        assert np.array_equal(u.download(), m[:, k * c:(k + 1) * c]), k
        u.close()
    with pytest.raises(Q.QsxError) as ei:  # shards must be in column order and complete
        a, b = Q.Unitary(3, 0, 4, 4), Q.Unitary(3, 0, 0, 4)
        Q.adjoint_sharded([a, b])
    assert ei.value.status == "QSX_ERR_SHAPE"


def test_interrupt_and_oom():
    qc = _rand_circuit(6, 200, 50)
    u = Q.Unitary(6)
    calls = {"n": 0}

    def stop():
        calls["n"] += 1
        return calls["n"] > 2

    with pytest.raises(Q.QsxError) as ei:
        u.apply(O.gate_stream(qc), Q.PlanOptions(fuse=0), should_stop=stop)
    assert ei.value.status == "QSX_ERR_INTERRUPTED"
    with pytest.raises(Q.QsxError) as ei:
        Q.Unitary(20)  # 17.6 TB
    assert ei.value.status == "QSX_ERR_OOM"
    # the device is still usable afterwards
    assert np.array_equal(Q.Unitary(2).download(), np.eye(4))


def test_inverse_circuit_roundtrip_at_full_size():
    """Size-independent property at BASELINE config C2 (n=12, 2000 gates): C^-1 C == I,
    judged by the device's own fused reduction against a fresh identity."""
    n, g = 12, 2000
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(n, g, 20261019))
    gs = O.gate_stream(qc)
    inv = [(c, q, m.conj().T) for c, q, m in reversed(gs)]
    u = Q.Unitary(n)
    u.apply(gs)
    u.apply(inv)
    ident = Q.Unitary(n)
    sums = ident.equiv_partial(u)
    r = Q.equiv_finish(sums, 1e-12, True)
    assert r.equivalent == 1 and abs(r.norm - 1) < 1e-10 and (r.phase_num, r.phase_den) == (0, 1)
    assert abs(sums[3] - 2.0**n) < 1e-9 * 2.0**n
    # and element-wise on a band of rows
    band = u.download(1000, 8)
    want = np.zeros_like(band)
    for i in range(8):
        want[i, 1000 + i] = 1
    assert np.max(np.abs(band - want)) < 1e-12


@pytest.mark.parametrize("n,seed", [(3, 1), (5, 2), (6, 3), (8, 4), (9, 5), (11, 6)])
def test_dense_user_gates_on_the_tensor_pipe(n, seed):
    """Gates with >= 3 targets (a nested circuit as ONE matrix, reference qcir_to_tensor.cpp:147 used as a gate; controls
    folded into the block): each is a dense sweep of its own (dense_kernel<3..6>) between interpreted / specialised
    sweeps -- the reference contracts them like any other gate tensor."""
    from tests.test_abi_cpu import _circuit_with_dense_gates, _oracle_matrix
    gs = _circuit_with_dense_gates(n, seed)
    want = _oracle_matrix(n, gs)
    for opt in (None, Q.PlanOptions(jit=2), Q.PlanOptions(fuse=0), Q.PlanOptions(dense_block=3)):
        u = Q.Unitary(n)
        u.apply(gs, opt)
        assert frob_rel(u.download(), want) < TOL, (n, seed)
        u.close()
    if n >= 6:  # a column shard
        u = Q.Unitary(n, 0, 2**n // 2, 2**n // 4)
        u.apply(gs)
        assert frob_rel(u.download(), want[:, 2**n // 2: 2**n // 2 + 2**n // 4]) < TOL
    rng = np.random.default_rng(1)
    wide = np.linalg.qr(rng.standard_normal((128, 128)) + 1j * rng.standard_normal((128, 128)))[0]
    with pytest.raises(Q.QsxError) as ei:  # wider than the largest dense block: refused, like an unsupported gate in the reference
        Q.Unitary(9).apply([(0, list(range(7)), wide)])
    assert ei.value.status == "QSX_ERR_UNSUPPORTED"


@pytest.mark.parametrize("k", [3, 4, 5, 6])
def test_dense_blocks_on_the_tensor_pipe(k):
    """dense_block = k: every sweep is one dense 2^k x 2^k block applied with DMMA (dense_kernel<k>)."""
    for n, g, seed in [(5, 120, 4), (7, 200, 6), (9, 300, 16), (10, 200, 3)]:
        if n < k:
            continue
        qc = _rand_circuit(n, g, seed)
        want = O.to_matrix_rowform(qc)
        got = run_device(qc, Q.PlanOptions(dense_block=k))
        assert frob_rel(got, want) < TOL, (k, n)
    # column shard
    qc = _rand_circuit(8, 150, 9)
    want = O.to_matrix_rowform(qc)
    got = run_device(qc, Q.PlanOptions(dense_block=k), 64, 64)
    assert frob_rel(got, want[:, 64:128]) < TOL


def test_n10_against_oracle_default_schedule():
    n = 10
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(n, 600, 123))
    want = O.to_matrix_rowform(qc)
    got = run_device(qc)
    assert frob_rel(got, want) < TOL
    got = run_device(qc, Q.PlanOptions(snap_tol=-1.0))
    assert frob_rel(got, want) < TOL


@pytest.mark.parametrize("opts", [{}, {"fuse": 0}, {"tile_qubits": 7, "log2_tile_cols": 4}, {"dense_block": 5},
                                  {"reg_qubits": 3, "tile_qubits": 9, "no_perm_folding": 2}])
def test_repeated_runs_are_bitwise_identical(opts):
    """A folded permutation makes threads of a CTA read rows that other threads write: the result
    must not depend on warp timing (regression: a one-stage sweep lacked the barrier between its
    global loads and stores and every run of a 600-gate circuit differed)."""
    n = 10
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(n, 600, 123))
    plan = Q.Plan(n, 2**n, O.gate_stream(qc), Q.PlanOptions(**opts))
    first = None
    for _ in range(6):
        u = Q.Unitary(n)
        plan.run(u)
        got = u.download()
        u.close()
        if first is None:
            first = got
            assert frob_rel(got, O.to_matrix_rowform(qc)) < TOL
        else:
            assert np.array_equal(got.view(np.float64), first.view(np.float64))


def test_async_apply_matches_sync_and_reuses_staging():
    """qsx_apply_gates_async: several one-shot programs in flight through the two pinned staging
    buffers and the single device program buffer (reuse ordered by events / the stream)."""
    n = 9
    circuits = [_rand_circuit(n, 150 + 40 * k, 30 + k) for k in range(5)]
    want = [run_device(qc) for qc in circuits]
    us = [Q.Unitary(n) for _ in circuits]
    for u, qc in zip(us, circuits):  # no synchronisation between the five submissions
        st = u.apply_async(O.gate_stream(qc))
        assert st.launches > 0 and st.device_ms == 0
    for u, w in zip(us, want):
        assert np.array_equal(u.download().view(np.float64), w.view(np.float64))
        u.close()


def test_lazy_identity_is_invisible():
    """qsx_unitary_set_identity only marks the shard; the first sweep of the next gate stream synthesises
    the identity (sweep_kernel<.., IDENT>), every other consumer materialises it first."""
    n = 9
    dim = 2**n
    eye = np.eye(dim)
    qc, qc2 = _rand_circuit(n, 180, 77), _rand_circuit(n, 90, 78)
    want, want2 = O.to_matrix_rowform(qc), O.to_matrix_rowform(qc2)
    u = Q.Unitary(n)
    # consumers of a never-written identity: clone, equiv, adjoint, partial upload
    c = u.clone()
    assert np.array_equal(c.download(), eye)
    s = Q.Unitary(n).equiv_partial(Q.Unitary(n))
    assert s[0] == dim and s[1] == 0 and s[2] == dim and s[3] == dim and s[4] == dim and s[6] == dim
    a = Q.Unitary(n)
    a.adjoint_inplace()
    assert np.array_equal(a.download(), eye)
    p = Q.Unitary(n)
    rows = (np.arange(3 * dim).reshape(3, dim) + 1j).astype(np.complex128)
    p.upload(rows, 5)
    w = eye.astype(np.complex128)
    w[5:8] = rows
    assert np.array_equal(p.download(), w)
    # a second gate stream continues from the data, not from the identity
    u.apply(O.gate_stream(qc))
    u.apply(O.gate_stream(qc2))
    assert frob_rel(u.download(), want2 @ want) < TOL
    # set_identity after gates, twice, then gates; and an empty stream in between
    for opt in (None, Q.PlanOptions(reg_qubits=3), Q.PlanOptions(dense_block=4), Q.PlanOptions(fuse=0, tile_qubits=6)):
        u.set_identity()
        u.set_identity()
        u.apply([], opt)
        st = u.apply(O.gate_stream(qc), opt)
        assert st.launches > 0
        assert frob_rel(u.download(), want) < TOL
    u.set_identity()
    assert np.array_equal(u.download(), eye)
    # column shard: the synthesised identity sits at global column col0 + c
    for col0, ncols in ((dim // 2, dim // 4), (dim - 8, 8), (3, 1)):
        sh = Q.Unitary(n, 0, col0, ncols)
        sh.apply(O.gate_stream(qc))
        assert frob_rel(sh.download(), want[:, col0:col0 + ncols]) < TOL


@pytest.mark.parametrize("mib_per_n", [{10: 1}, {10: 4}])
def test_l2_column_blocking_is_the_same_computation(mib_per_n):
    """l2_block_mib runs all sweeps on one block of columns before the next: bit-identical to the unblocked run
    (same tiles, same programs; only the order of (block, sweep) pairs changes), also from a lazy identity,
    on a column shard, and for a second gate stream on existing data."""
    qc10, qc10b = O.qcir_from_qasm_text(O.random_clifford_t_qasm(10, 500, 5)), _rand_circuit(10, 120, 8, extra=True)
    n = 10
    base = Q.Unitary(n)
    base.apply(O.gate_stream(qc10), Q.PlanOptions(l2_block_mib=-1))
    base.apply(O.gate_stream(qc10b), Q.PlanOptions(l2_block_mib=-1))
    want = base.download()
    assert frob_rel(want, O.to_matrix_rowform(qc10b) @ O.to_matrix_rowform(qc10)) < TOL
    for opt in (Q.PlanOptions(l2_block_mib=mib_per_n[10]), Q.PlanOptions(l2_block_mib=mib_per_n[10], fuse=0)):
        u = Q.Unitary(n)
        plan = Q.Plan(n, 2**n, O.gate_stream(qc10), opt)
        assert plan.dump()["block_cols"] == mib_per_n[10] * 2**20 // (16 * 2**n)
        st = plan.run(u)
        assert 0 <= st.launches - plan.stats().n_sweeps * (2**n // plan.dump()["block_cols"]) <= 1
        u.apply(O.gate_stream(qc10b), opt)
        got = u.download()
        if opt.fuse:
            assert np.array_equal(got.view(np.float64), want.view(np.float64))
        else:
            assert frob_rel(got, want) < TOL
        u.close()
    # column shard (col0 != 0) with blocks of one tile width, reg_qubits != 4 (identity materialised first)
    col0, ncols = 256, 256
    for opt in (Q.PlanOptions(l2_block_mib=1), Q.PlanOptions(l2_block_mib=1, reg_qubits=3, tile_qubits=7)):
        sh = Q.Unitary(n, 0, col0, ncols)
        sh.apply(O.gate_stream(qc10), opt)
        assert frob_rel(sh.download(), O.to_matrix_rowform(qc10)[:, col0:col0 + ncols]) < TOL
        sh.close()


def _device_fuzz(seed, seconds, max_cases=10**9):
    import time
    rng = np.random.default_rng(seed)
    t0, n_ok, worst = time.time(), 0, 0.0
    while time.time() - t0 < seconds and n_ok < max_cases:
        n, g, cseed = int(rng.integers(1, 11)), int(rng.integers(10, 300)), int(rng.integers(1 << 30))
        qc = _rand_circuit(n, g, cseed, extra=bool(rng.integers(2)))
        want = O.to_matrix_rowform(qc)
        opts = dict(fuse=int(rng.integers(4) > 0), reg_qubits=int(rng.integers(1, 6)), tile_qubits=int(rng.integers(2, 11)),
                    log2_tile_cols=int(rng.choice([-1, 1, 2, 3, 4])), max_ops_per_sweep=int(rng.choice([5, 17, 96, 300])),
                    min_tail_ops=int(rng.choice([-1, 8])), search=int(rng.choice([-1, 2])), stage_search=int(rng.choice([-1, 1])),
                    no_1q_fusion=int(rng.integers(2)), no_diag_conjugation=int(rng.integers(2)),
                    no_perm_folding=int(rng.integers(3)), snap_tol=float(rng.choice([1e-15, -1.0])),
                    l2_block_mib=int(rng.choice([-1, 0, 1, 1, 2])))
        shards = int(rng.choice([1, 1, 2, 4])) if n >= 3 else 1
        ncols = 2**n // shards
        k = int(rng.integers(shards))
        got = run_device(qc, Q.PlanOptions(**opts), k * ncols, ncols)
        err = float(frob_rel(got, want[:, k * ncols:(k + 1) * ncols]))
        assert err < TOL, (n, g, cseed, opts, shards, k, err)
        worst = max(worst, err)
        n_ok += 1
    return n_ok, worst


def test_random_circuits_times_random_options_on_device():
    """Differential fuzz of the whole device path (scheduler options, every kernel instantiation,
    column shards) against the oracle; `python -m tests.test_gpu_parity SEED SECONDS` runs it longer."""
    n_ok, worst = _device_fuzz(20261019, 15.0, 300)
    assert n_ok >= 30 and worst < TOL


if __name__ == "__main__":
    import sys
    print(_device_fuzz(int(sys.argv[1]) if len(sys.argv) > 1 else 1, float(sys.argv[2]) if len(sys.argv) > 2 else 60.0))
