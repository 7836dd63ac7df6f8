"""Where the end-to-end time of a job goes (host layer, QASM text -> verdict): per-phase wall times.
usage: python scripts/e2e_phases.py [workload[,workload...]] [reps]   (several workloads run one after another in ONE process)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsyn_b200 import abi, host, workloads  # noqa: E402

TAIL = "t q[0];\nx q[0];\nt q[0];\nx q[0];\n"
wls = (sys.argv[1] if len(sys.argv) > 1 else "c2").split(",")
reps0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
acc = {}
wl = wls[0]


def tick(name, t0):
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


def report(tag, count, j0):
    j1 = abi.jit_stats()
    print(wl, tag, {k: round(v / count * 1e3, 3) for k, v in acc.items()}, "total ms", round(sum(acc.values()) / count * 1e3, 3),
          "| specialised launches per job", (j1.launches - j0.launches) // count, "compiled", j1.compiled, "from disk", j1.disk_hits)
    acc.clear()
    return j1


for wl in wls:
  reps = 3 if wl in ("c4", "c5") else reps0
  text = workloads.workload_qasm(wl)[2]
  j0 = abi.jit_stats()
  for rep in range(reps + 2):
      if rep == 2:
          j0 = report("first 2 jobs", 2, j0)
      t = time.perf_counter()
      ca = host.Circuit(qasm_text=text); t = tick("parse A", t)
      ta = ca.qc2ts(); t = tick("qc2ts A", t)
      cb = host.Circuit(qasm_text=text + TAIL); t = tick("parse B", t)
      tb = cb.qc2ts(); t = tick("qc2ts B", t)
      r, verdict = ta.equiv(tb); t = tick("equiv", t)
      for x in (ta, tb):
          x.close()
      t = tick("close tensors", t)
      for x in (ca, cb):
          x.close()
      t = tick("close circuits", t)
  report("next %d jobs" % reps, reps, j0)
