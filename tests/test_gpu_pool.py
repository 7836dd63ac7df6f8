"""The per-device pool of shard buffers (csrc/runtime.cu pool_alloc / pool_free): freed unitaries are parked for reuse up
to a fraction of the device memory, and over that cap the buffers parked LONGEST ago are released -- never the one
coming in.  The case that went wrong before: a 15-qubit job leaves the pool full of 16 GiB buffers and every tensor of
the following 12-qubit jobs paid a cudaFree + cudaMalloc (C2 end to end 8.5 -> 20.9 ms in bench.py's `extra.c2`)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = """
    import ctypes, time
    import numpy as np
    import qsyn_b200 as Q
    from oracle import oracle as O

    rt = ctypes.CDLL("libcudart.so.12")
    def free_bytes():
        f, t = ctypes.c_size_t(), ctypes.c_size_t()
        assert rt.cudaMemGetInfo(ctypes.byref(f), ctypes.byref(t)) == 0
        return f.value, t.value

    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(10, 60, 3))
    gs, want = O.gate_stream(qc), O.to_matrix_rowform(qc)
    def small_job():
        u = Q.Unitary(10, device=0)            # 16 MiB
        u.apply(gs)
        got = u.download()
        u.close()
        return float(np.linalg.norm(got - want))

    small_job()
    f0, total = free_bytes()
    cap = int({frac} * total)
    # big buffers fill the pool up to its cap (13 qubits: 1 GiB each; the cap is about 2.3 of them)
    big = [Q.Unitary(13, device=0) for _ in range(3)]
    for b in big:
        b.close()
    f1, _ = free_bytes()
    parked_big = f0 - f1
    assert 2**30 <= parked_big <= cap + 2**21, (parked_big, cap)      # two of the three stay parked, the oldest went
    # small tensors now recycle through the pool although stale big buffers hold most of the room
    errs = [small_job() for _ in range(4)]
    t0 = time.perf_counter()
    errs += [small_job() for _ in range(20)]
    per_job_ms = (time.perf_counter() - t0) / 20 * 1e3
    f2, _ = free_bytes()
    assert max(errs) < 1e-12, max(errs)
    assert f0 - f2 <= cap + 2**21, (f0 - f2, cap)                       # the pool never exceeds its cap
    # a buffer larger than the whole cap is released at once, and everything parked is reusable afterwards
    huge = Q.Unitary(14, device=0)                                      # 4 GiB > cap
    huge.close()
    f3, _ = free_bytes()
    assert f0 - f3 <= cap + 2**21, (f0 - f3, cap)
    assert small_job() < 1e-12
    print("ok parked_big_MiB=%d per_small_job_ms=%.2f" % (parked_big >> 20, per_job_ms))
"""


def test_pool_keeps_recycling_small_buffers_when_stale_large_ones_fill_it():
    import torch

    total = torch.cuda.mem_get_info(0)[1]
    frac = 2.4 * 2**30 / total  # cap = 2.4 GiB
    env = dict(os.environ, PYTHONPATH=ROOT, QSX_POOL_FRACTION=repr(frac), QSX_JIT="0")
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(CHILD.format(frac=repr(frac)))], capture_output=True, text=True,
                       env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.stdout[-500:], r.stderr[-1500:])
    print(r.stdout.strip())
