"""debug: the multi-GPU CLI case of tests/test_gpu_multi_device.py, output printed, with env variants"""
import json, os, subprocess, sys, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = json.load(open(os.path.join(ROOT, "tests/golden/reference_goldens.json")))
SHIM = os.path.join(ROOT, "bin", "qsyn-b200")
dof = "\n".join([
    "qcir read benchmark/qft/qft_8.qasm", "qc2ts", "qcir read benchmark/qft/qft_8dec.qasm", "qc2ts",
    "tensor equiv 0 1 --strict", "tensor adjoint 1", "tensor equiv 0 1", "tensor adjoint 1", "tensor equiv 0 1 --strict", "quit -f"])
for extra in ({}, {"QSX_POOL_FRACTION": "0"}, {"QSX_JIT": "0"}):
    for ng in (1, 2):
        tmp = pathlib.Path(tempfile.mkdtemp())
        for rel, text in g["inputs"].items():
            p = tmp / rel; p.parent.mkdir(parents=True, exist_ok=True); p.write_text(text)
        (tmp / "case.dof").write_text(dof)
        env = dict(os.environ, QSX_NUM_GPUS=str(ng), **extra)
        r = subprocess.run([SHIM, "--no-version", "--qsynrc-path", "/dev/null", "--verbose", str(tmp / "case.dof")], cwd=tmp, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        lines = [l for l in r.stdout.split("\n") if l and not l.startswith("qsyn> qcir") and not l.startswith("qsyn> qc2ts")]
        print(extra, "gpus", ng, "rc", r.returncode, " | ".join(lines))
