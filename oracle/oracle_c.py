"""ctypes access to oracle/_build/liboracle_c.so (plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_build", "liboracle_c.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.dirname(_HERE), "oracle"])
        _lib = C.CDLL(path)
        _lib.orc_apply_gate.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        _lib.orc_apply_gate.restype = C.c_int
        _lib.orc_equiv_reference_passes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]
        _lib.orc_equiv_reference_passes.restype = None
    return _lib


def apply_gate(u: np.ndarray, n: int, n_ctrls: int, qubits, target_matrix: np.ndarray) -> None:
    """In place on a C-contiguous (2^n, ncols) complex128 array."""
    assert u.flags.c_contiguous and u.dtype == np.complex128
    m = np.ascontiguousarray(target_matrix, dtype=np.complex128)
    q = (C.c_int * len(qubits))(*[int(x) for x in qubits])
    rc = lib().orc_apply_gate(u.ctypes.data, n, u.shape[1], n_ctrls, len(qubits) - n_ctrls, q, m.ctypes.data)
    assert rc == 0, rc


def to_matrix(gate_stream, n: int, col0: int = 0, ncols=None) -> np.ndarray:
    dim = 2**n
    ncols = dim - col0 if ncols is None else ncols
    u = np.zeros((dim, ncols), dtype=np.complex128)
    for c in range(ncols):
        u[col0 + c, c] = 1.0
    for n_ctrls, qubits, m in gate_stream:
        apply_gate(u, n, n_ctrls, qubits, m)
    return u


def equiv_reference_passes(a: np.ndarray, b: np.ndarray):
    """(cosine, norm, phase_rad) via 8 single-thread passes, as the reference evaluates them."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    out = (C.c_double * 3)()
    lib().orc_equiv_reference_passes(a.ctypes.data, b.ctypes.data, a.size, out)
    return out[0], out[1], out[2]
