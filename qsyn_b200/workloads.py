"""Synthetic input circuits of the benchmark configs, as QASM text in the dialect
qsyn's reader accepts.  Product-side generator used by bench.py; tests assert it is byte-identical
to the oracle's generator so that both arms of the bench see the same circuit."""
from __future__ import annotations

import numpy as np

_HEADER = ["OPENQASM 2.0;", 'include "qelib1.inc";']


def random_clifford_t_qasm(n: int, n_gates: int, seed: int) -> str:
    """Each gate 2-qubit with p = 0.35 (cx 80 % / cz 20 %, random ordered distinct pair), else
    1-qubit uniform over {h, s, sdg, t, tdg, x, z} on a uniform qubit; numpy default_rng(PCG64)."""
    rng = np.random.default_rng(seed)
    one = ["h", "s", "sdg", "t", "tdg", "x", "z"]
    lines = _HEADER + [f"qreg q[{n}];"]
    for _ in range(n_gates):
        if n >= 2 and rng.random() < 0.35:
            name = "cx" if rng.random() < 0.8 else "cz"
            a, b = rng.choice(n, size=2, replace=False)
            lines.append(f"{name} q[{int(a)}],q[{int(b)}];")
        else:
            name = one[int(rng.integers(len(one)))]
            lines.append(f"{name} q[{int(rng.integers(n))}];")
    return "\n".join(lines) + "\n"


def qft_qasm(n: int) -> str:
    """h + cp(pi/2^j), the plain variant of benchmark/qft/generateQFTqasm.py."""
    lines = _HEADER + [f"qreg q[{n}];"]
    for i in range(n):
        lines.append(f"h q[{i}];")
        for j in range(i + 1, n):
            lines.append(f"cp(pi/{2 ** (j - i)}) q[{j}],q[{i}];")
    return "\n".join(lines) + "\n"


def random_rotation_qasm(n: int, n_gates: int, seed: int) -> str:
    """rx / ry / rz with random rational angles, 25 % cx: the dense-rotation workload."""
    rng = np.random.default_rng(seed)
    lines = _HEADER + [f"qreg q[{n}];"]
    for _ in range(n_gates):
        if n >= 2 and rng.random() < 0.25:
            a, b = rng.choice(n, size=2, replace=False)
            lines.append(f"cx q[{int(a)}],q[{int(b)}];")
        else:
            name = ["rx", "ry", "rz"][int(rng.integers(3))]
            num, den = int(rng.integers(1, 64)), 64
            lines.append(f"{name}({num}*pi/{den}) q[{int(rng.integers(n))}];")
    return "\n".join(lines) + "\n"


# BASELINE.json configs C2..C5 (+ a tiny smoke config): name -> (n_qubits, generator, n_gates, seed, description)
WORKLOADS = {
    "c2": (12, "clifford_t", 2000, 20261019, "C2: 12-qubit random Clifford+T, 2000 gates (256 MiB unitary)"),
    "c3": (14, "qft", 105, 0, "C3: 14-qubit QFT (benchmark/qft generator), 105 gates (4 GiB unitary)"),
    "c4": (15, "clifford_t", 10000, 20261020, "C4: 15-qubit random Clifford+T, 10000 gates (16 GiB unitary)"),
    "c5": (16, "clifford_t", 10000, 20261021, "C5: 16-qubit random Clifford+T, 10000 gates (64 GiB unitary)"),
    "tiny": (8, "clifford_t", 400, 7, "tiny: 8-qubit random Clifford+T, 400 gates (smoke)"),
    # not a BASELINE config: the dense-rotation circuit the dense k-qubit blocks are compared on
    "rot": (12, "rotation", 2000, 20261022, "rot: 12-qubit random rx/ry/rz (75 %) + cx circuit, 2000 gates (256 MiB unitary)"),
    # parity-only: generic (non-Clifford) 2x2 matrices at a size where rounding has room to accumulate
    "rot14": (14, "rotation", 6000, 20261023, "rot14: 14-qubit random rx/ry/rz (75 %) + cx circuit, 6000 gates (4 GiB unitary)"),
}


def workload_qasm(name: str) -> tuple[int, str, str]:
    """(n_qubits, description, QASM text of circuit A)."""
    n, gen, g, seed, desc = WORKLOADS[name]
    text = random_clifford_t_qasm(n, g, seed) if gen == "clifford_t" else random_rotation_qasm(n, g, seed) if gen == "rotation" else qft_qasm(n)
    return n, desc, text
