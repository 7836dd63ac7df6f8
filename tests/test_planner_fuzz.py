"""Randomised differential test of the host scheduler (CPU only): random circuits over the whole
gate set x random settings of EVERY scheduler option; the micro-ops the kernel would execute are
replayed in numpy (tests/plan_sim.py) and compared with the oracle.  A fixed seed and a time
budget keep it deterministic-ish and short; `python tests/test_planner_fuzz.py SEED SECONDS` runs
it for longer (4800 plans passed when the scheduler features of round 1 were frozen)."""
import sys
import time

import numpy as np

import qsyn_b200 as Q
from oracle import oracle as O
from tests import plan_sim
from tests.test_abi_cpu import _rand_circuit


def fuzz(seed: int, seconds: float, max_plans: int = 10**9):
    rng = np.random.default_rng(seed)
    t0, n_ok, worst = time.time(), 0, 0.0
    while time.time() - t0 < seconds and n_ok < max_plans:
        n, g, cseed = int(rng.integers(2, 8)), int(rng.integers(20, 220)), int(rng.integers(1 << 30))
        qc = _rand_circuit(n, g, cseed, extra=bool(rng.integers(2)))
        want = O.to_matrix_rowform(qc)
        opts = dict(fuse=1, reg_qubits=int(rng.integers(1, 6)), tile_qubits=int(rng.integers(2, 10)),
                    log2_tile_cols=int(rng.choice([-1, 1, 2, 3])), max_ops_per_sweep=int(rng.choice([5, 17, 48, 96, 300])),
                    min_tail_ops=int(rng.choice([-1, 2, 8, 20])), search=int(rng.choice([-1, 1, 3])),
                    stage_search=int(rng.choice([-1, 1])), no_1q_fusion=int(rng.integers(2)),
                    no_diag_conjugation=int(rng.integers(2)), no_perm_folding=int(rng.integers(3)),
                    snap_tol=float(rng.choice([1e-15, -1.0])))
        plan = Q.Plan(n, 2**n, O.gate_stream(qc), Q.PlanOptions(**opts))
        d = plan.dump()
        plan_sim.check_invariants(d, plan.stats().n_ops)
        got = plan_sim.run_plan_uops(d, np.eye(2**n, dtype=np.complex128))
        err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        assert err < 1e-12, (n, g, cseed, opts, err)
        worst = max(worst, err)
        n_ok += 1
    return n_ok, worst


def test_random_circuits_times_random_scheduler_options():
    # bounded by the NUMBER of plans (a slow CI host must not turn a time budget into a failure); ~15 s here
    n_ok, worst = fuzz(seed=20261019, seconds=600.0, max_plans=120)
    assert n_ok == 120 and worst < 1e-12


if __name__ == "__main__":
    print(fuzz(int(sys.argv[1]) if len(sys.argv) > 1 else 1, float(sys.argv[2]) if len(sys.argv) > 2 else 120.0))
