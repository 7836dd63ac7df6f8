#!/usr/bin/env python
"""Builds tests/golden/reference_goldens.json from the reference's OWN test
fixtures for the tensor-verification path and its ZX operand (8 f2).

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):  python tests/golden/extract_reference_goldens.py

For each case we keep the dofile text, the golden log text, and the text of
every input circuit the dofile reads, so that tests can (a) pin the oracle and
(b) replay the dofile through the CLI shim and diff the log byte for byte.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")

CASES = {
    # name: (dof, log)
    "qc2ts": ("tests/conversion/qc2ts/dof/qc2ts.dof", "tests/conversion/qc2ts/ref/qc2ts.log"),
    "sk": ("tests/conversion/sk/dof/sk.dof", "tests/conversion/sk/ref/sk.log"),
    "global_norm_phase": ("tests/tensor/equiv/dof/global_norm_phase.dof", "tests/tensor/equiv/ref/global_norm_phase.log"),
    "error": ("tests/tensor/not_equiv/dof/error.dof", "tests/tensor/not_equiv/ref/error.log"),
    "size": ("tests/tensor/not_equiv/dof/size.dof", "tests/tensor/not_equiv/ref/size.log"),
    "empty_mgr": ("tests/conversion/dof/empty_mgr.dof", "tests/conversion/ref/empty_mgr.log"),
    # the ZX operand
    "zx2ts_h_edge": ("tests/conversion/zx2ts/dof/h_edge.dof", "tests/conversion/zx2ts/ref/h_edge.log"),
    "zx2ts_disconnected": ("tests/conversion/zx2ts/dof/disconnected.dof", "tests/conversion/zx2ts/ref/disconnected.log"),
    "zx2ts_invalid_graph": ("tests/conversion/zx2ts/dof/invalid_graph.dof", "tests/conversion/zx2ts/ref/invalid_graph.log"),
    "tiny_circuits": ("tests/tensor/equiv/dof/tiny_circuits.dof", "tests/tensor/equiv/ref/tiny_circuits.log"),
    "tensor_adjoint": ("tests/tensor/manip/dof/adjoint.dof", "tests/tensor/manip/ref/adjoint.log"),
    "zx_add_remove": ("tests/zx/graph/dof/add_remove.dof", "tests/zx/graph/ref/add_remove.log"),
    "zx_tensor_product": ("tests/zx/graph/dof/tensor_product.dof", "tests/zx/graph/ref/tensor_product.log"),
}

# in-repo circuit families that are equivalent by construction (SURVEY 8c last bullet)
EXTRA_INPUTS = [f"benchmark/qft/qft_{n}{v}.qasm" for n in (2, 3, 4, 5, 8) for v in ("", "dec", "crz")] + [
    "benchmark/qasm/qc2ts/cnot.qasm",
    "benchmark/qasm/qc2ts/swap.qasm",
    "benchmark/qasm/qc2ts/toffoli.qasm",
    "benchmark/qasm/qc2ts/test.qasm",
    # examples/tensor-equiv.dof (BASELINE.json configs[0]) reads these two
    "benchmark/zx/tof3_op.zx",
    "benchmark/zx/tof3_zh.zx",
]


def main():
    out = {"_source": "DVLab-NTU/qsyn v0.8.1 tests/ and benchmark/ (fixtures only, no source code)", "cases": {}, "inputs": {}}
    for name, (dof, log) in CASES.items():
        dof_text = open(os.path.join(REF, dof)).read()
        log_text = open(os.path.join(REF, log)).read()
        out["cases"][name] = {"dof_path": dof, "dof": dof_text, "log": log_text}
        for m in re.finditer(r"(?:qcir|zx) read (\S+)", dof_text):
            p = m.group(1)
            out["inputs"][p] = open(os.path.join(REF, p)).read()
    for p in EXTRA_INPUTS:
        out["inputs"][p] = open(os.path.join(REF, p)).read()
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False, sort_keys=True)
    print("wrote", OUT, len(out["cases"]), "cases", len(out["inputs"]), "inputs")


if __name__ == "__main__":
    main()
