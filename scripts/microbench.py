"""Per-micro-op cost model of the sweep kernel at n qubits (one GPU).  Builds synthetic gate
streams that the scheduler turns into ONE sweep with a known number of micro-ops / stages and
reports the marginal time per op kind.  Guides tuning; not the bench."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import qsyn_b200 as Q

H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2.0)
T = np.diag([1.0, np.exp(0.25j * np.pi)])
X = np.array([[0, 1], [1, 0]], dtype=complex)
_c, _s = np.cos(np.pi / 6), np.sin(np.pi / 6)
RX = np.array([[_c, -1j * _s], [-1j * _s, _c]], dtype=complex)


def run(n, gates, opt, reps=5):
    plan = Q.Plan(n, 2**n, gates, opt)
    st = plan.stats()
    u = Q.Unitary(n)
    plan.run(u)
    best = min(plan.run(u).device_ms for _ in range(reps))
    u.close()
    return best * 1e3, st


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    k = 32
    tile = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    opt = lambda **kw: Q.PlanOptions(max_ops_per_sweep=1000, tile_qubits=tile, min_tail_ops=-1, no_1q_fusion=1, search=-1,
                                     stage_search=-1, **kw)  # the cost model wants every gate as its own micro-op
    q = lambda b: n - 1 - b  # row bit -> qubit id
    cases = {}
    # register bits will be the first 4 distinct target bits: use row bits 0..3 as register bits
    cases["1 H (one sweep, one stage)"] = [(0, [q(0)], H)]
    cases["H x%d same bit" % k] = [(0, [q(0)], H)] * k
    for kk in (4, 8, 16, 64, 128):
        cases["H x%d same bit" % kk] = [(0, [q(0)], H)] * kk
    cases["H x%d over 4 reg bits" % k] = [(0, [q(i % 4)], H) for i in range(k)]
    cases["(H,T) x%d same reg bit (H + DTABH)" % k] = [(0, [q(0)], m) for _ in range(k) for m in (H, T)]
    cases["(H0,T8) x%d (H + DSCAL outside)" % k] = [g for _ in range(k) for g in ((0, [q(0)], H), (0, [q(8)], T))]
    cases["(H0,X1) x%d in-reg X (no fold)" % k] = [g for _ in range(k) for g in ((0, [q(0)], H), (0, [q(0)], X))]
    cases["(H0,CX 8->0) x%d in-reg X thread ctrl" % k] = [g for _ in range(k) for g in ((0, [q(0)], H), (1, [q(8), q(0)], X))]
    cases["(H0,CX 1->0) x%d in-reg XC" % k] = [g for _ in range(k) for g in ((0, [q(0)], H), (1, [q(1), q(0)], X))]
    cases["RX x%d same bit (U1)" % k] = [(0, [q(0)], RX)] * k
    cases["H on bits 0..7 cyclic x%d (stage every 4)" % k] = [(0, [q(i % 8)], H) for i in range(k)]
    cases["H alternating bit0/bit4.. 8 bits, forces stage per 4 ops"] = [(0, [q((i * 4 + i // 2) % 8)], H) for i in range(k)]
    cases["X x%d on non-reg tile bits (all folded)" % k] = [(0, [q(0)], H)] + [(0, [q(4 + i % 4)], X) for i in range(k)]
    cases["CX x%d folded" % k] = [(0, [q(0)], H)] + [(1, [q(5 + i % 3), q(4)], X) for i in range(k)]
    base = None
    for name, gates in cases.items():
        for pol in (0,):
            us, st = run(n, gates, opt(no_perm_folding=1 if "no fold" in name or "in-reg" in name else 0))
            if base is None:
                base = us
            print(json.dumps({"case": name, "us": round(us, 1), "sweeps": st.n_sweeps, "stages": st.n_stages, "uops": st.n_micro_ops,
                              "folded": st.n_folded_perms, "us_over_base_per_uop": round((us - base) / max(st.n_micro_ops - 1, 1), 2)}), flush=True)


if __name__ == "__main__":
    main()
