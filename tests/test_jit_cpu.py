"""Specialised sweeps without a GPU: code generation + NVRTC compilation (NVRTC needs no device), the kernel cache on
disk, and a clean process exit while compilations are in flight.  The execution of the compiled kernels is covered by
tests/test_gpu_jit.py."""
import os
import subprocess
import sys
import textwrap

import pytest

import qsyn_b200 as Q
from oracle import oracle as O
from qsyn_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, tmp_path, **env):
    e = dict(os.environ, PYTHONPATH=ROOT)
    e["QSX_JIT_CACHE"] = str(tmp_path / "jit-cache")
    e.update(env)
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=e, cwd=ROOT, timeout=300)


COMPILE = """
    import qsyn_b200 as Q
    from oracle import oracle as O
    from qsyn_b200 import abi
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(12, 160, 5))
    plan = Q.Plan(12, 4096, O.gate_stream(qc))          # n = 12: specialised sweeps are on by default
    d = plan.dump()
    assert d["jit"] == 1 and d["m_cap"] == 9, (d["jit"], d["m_cap"])
    plan.compile(True)
    s = abi.jit_stats()
    assert s.available == 1 and s.failed == 0 and s.requested == plan.stats().n_sweeps
    print("compiled", s.compiled, "disk", s.disk_hits, "sweeps", plan.stats().n_sweeps)
"""


def test_sweeps_compile_and_land_in_the_disk_cache(tmp_path):
    if abi.jit_stats().available != 1:
        pytest.skip("libnvrtc is not loadable on this host")
    r = _run(COMPILE, tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    first = r.stdout.split()
    assert int(first[1]) >= 1 and int(first[3]) == 0
    files = os.listdir(tmp_path / "jit-cache")
    assert len(files) == int(first[1]) and all(f.endswith(".cubin") for f in files)
    r = _run(COMPILE, tmp_path)  # a second process finds every kernel on disk
    assert r.returncode == 0, r.stderr[-2000:]
    second = r.stdout.split()
    assert int(second[1]) == 0 and int(second[3]) == int(first[1])
    r = _run(COMPILE, tmp_path, QSX_JIT_CACHE="off")  # and compiles again when the cache is off
    assert r.returncode == 0 and int(r.stdout.split()[1]) == int(first[1])


EXIT_EARLY = """
    import atexit, time
    import qsyn_b200 as Q
    from oracle import oracle as O
    qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(13, 1500, 6))
    plan = Q.Plan(13, 8192, O.gate_stream(qc))
    if {explicit} == 0:
        atexit.unregister(Q.lib().qsx_shutdown)   # an embedder that never calls qsx_shutdown: the library's own exit hooks
    plan.compile(False)                            # ~30 kernels queued on the compiler threads
    time.sleep(0.4)
    print("bye")
"""


@pytest.mark.parametrize("explicit", [1, 0])
def test_exit_while_compilations_are_in_flight(tmp_path, explicit):
    """`quit` right after a conversion used to crash inside NVRTC: its static state was destroyed under the compiler
    threads.  The queue is dropped and the running compilations are awaited before the process tears down."""
    if abi.jit_stats().available != 1:
        pytest.skip("libnvrtc is not loadable on this host")
    r = _run(EXIT_EARLY.format(explicit=explicit), tmp_path)
    assert r.returncode == 0 and "bye" in r.stdout, (r.returncode, r.stderr[-1500:])


def test_generated_source_is_straight_line_code(tmp_path):
    """The generator's output for one sweep: one kernel, no loops, immediates for every constant, and the renaming of
    permutations / the sign tracking of +-1, +-i phases (nothing but FP64 arithmetic on named values is left)."""
    if abi.jit_stats().available != 1:
        pytest.skip("libnvrtc is not loadable on this host")
    code = """
        import os, glob
        import qsyn_b200 as Q
        from oracle import oracle as O
        qc = O.qcir_from_qasm_text(O.random_clifford_t_qasm(12, 160, 5))
        plan = Q.Plan(12, 4096, O.gate_stream(qc))
        plan.compile(True)
    """
    dump = tmp_path / "dump"
    dump.mkdir()
    r = _run(code, tmp_path, QSX_JIT_DUMP=str(dump), QSX_JIT_CACHE="off")
    assert r.returncode == 0, r.stderr[-1500:]
    srcs = sorted(dump.glob("*.cu"))
    assert srcs and len(list(dump.glob("*.cubin"))) == len(srcs)
    text = srcs[0].read_text()
    assert text.count("__global__") == 1 and "for (" not in text and "while" not in text
    assert "fma(" in text and "__ldcg(" in text and "__stcg(" in text and "0x1." in text  # hex-float immediates
