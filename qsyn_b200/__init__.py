"""qsyn_b200 -- B200-native engine for qsyn's tensor-verification hot path
(`convert qcir tensor` + `tensor equiv`).

The product is the C-ABI shared library ``qsyn_b200/lib/libqsx.so`` (include/qsx.h):
hand-written sm_100a CUDA kernels plus the host-side gate scheduler, and the C++
host layer under ``qsyn_b200/host`` that mirrors the reference's QTensor /
to_tensor / is_equivalent interface.  This Python package is only the ctypes binding
used by tests/ and bench.py; it contains no arithmetic and has no CPU fallback.
"""
from .abi import (  # noqa: F401
    QsxError,
    Plan,
    PlanOptions,
    Unitary,
    adjoint_sharded,
    device_count,
    equiv_finish,
    equiv_reduce,
    lib,
    lib_path,
)
