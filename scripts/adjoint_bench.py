"""Times qsx_unitary_adjoint_sharded: n qubits over G column shards spread over the visible GPUs."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import qsyn_b200 as Q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
ndev = Q.device_count()
for shards in sorted({1, ndev, 2 * ndev}):
    c = 2**n // shards
    us = [Q.Unitary(n, k % ndev, k * c, c) for k in range(shards)]
    Q.adjoint_sharded(us)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        Q.adjoint_sharded(us)
    dt = (time.perf_counter() - t0) / reps
    gb = 2 * 16 * 4**n / 1e9
    chk = us[-1].download(2**n - 2, 2)[:, -2:]
    print(f"n={n} devices={ndev} shards={shards}: {dt * 1e3:.2f} ms  {gb / dt:.0f} GB/s (read+write of the whole tensor)  corner {chk.real.tolist()}")
    for u in us:
        u.close()
