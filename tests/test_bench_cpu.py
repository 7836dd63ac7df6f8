"""bench.py contract checks that need no GPU: the reference arm (the oracle port timed on host cores) prints
exactly ONE JSON line on stdout with the keys the driver reads; the product arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=dict(os.environ, **(env or {})), timeout=300)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    for key in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1", "--gpus", "2",
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout, r.stderr[-300:])


def test_reference_arm_runs_a_bounded_column_sample_beyond_host_memory():
    """n = 15: the full tensordot scheme needs 48 GiB; the arm times the same per-gate loop on a column block."""
    r = _run("--impl", "reference", "--workload", "c4", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip())
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["n_qubits"] == 15
    assert "first 2048 of 32768 columns" in d["config"]["step"] and "not the reference binary" in d["cpu_baseline"]["sample"]


def test_product_arm_needs_a_device():
    import qsyn_b200 as Q
    if Q.device_count() > 0:
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in r.stderr
