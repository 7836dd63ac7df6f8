"""World-size-2 (and 4) test of the multi-rank path on CPU with the gloo backend: column
sharding, per-shard scheduling through the C ABI (qsx_plan_create is pure host work), per-shard
partial sums, ONE all_reduce, and the host finish -- compared with the unsharded oracle.
The device kernels are replaced by the plan simulator here; the GPU tests cover them."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import qsyn_b200 as Q
from oracle import oracle as O
from qsyn_b200 import sharding
from tests import plan_sim
from tests.test_abi_cpu import _rand_circuit


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        qa, qb = _rand_circuit(n, 60, seed), _rand_circuit(n, 60, seed)
        for g in "txtx":
            qb.append(O.str_to_operation(g), [0])
        col0, ncols = sharding.shard_columns(n, rank, world)
        shards = []
        for qc in (qa, qb):
            plan = Q.Plan(n, ncols, O.gate_stream(qc))  # same gate stream on every rank
            u = np.zeros((2**n, ncols), dtype=np.complex128)
            for c in range(ncols):
                u[col0 + c, c] = 1
            shards.append(plan_sim.run_plan_uops(plan.dump(), u))
        part = O.equiv_sums(shards[0], shards[1])
        total = sharding.allreduce_sums(part)
        r = Q.equiv_finish(total, 1e-6, False)
        worst = sharding.max_over_ranks(float(rank))
        out_q.put((rank, col0, ncols, total.tolist(), r.equivalent, r.phase_num, r.phase_den, r.norm, worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_equivalence_matches_unsharded_oracle(world):
    n, seed = 5, 77
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    qa, qb = _rand_circuit(n, 60, seed), _rand_circuit(n, 60, seed)
    for g in "txtx":
        qb.append(O.str_to_operation(g), [0])
    A, B = O.to_matrix_rowform(qa), O.to_matrix_rowform(qb)
    want = O.equiv_sums(A, B)
    w = O.tensor_equiv(A, B)
    cols = sorted((c0, nc) for _, c0, nc, *_ in results)
    assert cols == [(k * 2**n // world, 2**n // world) for k in range(world)]
    for rank, _, _, total, equivalent, pn, pd, norm, worst in results:
        assert np.allclose(total, want, rtol=0, atol=1e-11)  # every rank holds the all-reduced sums
        assert bool(equivalent) == w.equivalent and (pn, pd) == (w.phase.numerator, w.phase.denominator) == (1, 4)
        assert abs(norm - 1) < 1e-9
        assert worst == world - 1


def test_shard_columns_rejects_bad_worlds():
    assert sharding.shard_columns(4, 3, 4) == (12, 4)
    with pytest.raises(ValueError):
        sharding.shard_columns(4, 0, 3)
    with pytest.raises(ValueError):
        sharding.shard_columns(2, 0, 8)
    with pytest.raises(ValueError):
        sharding.shard_columns(4, 4, 4)
    assert np.array_equal(sharding.allreduce_sums([1.0] * 8), np.ones(8))  # no process group: identity
