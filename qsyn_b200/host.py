"""ctypes binding of include/qsx_host.h (libqsxhost.so): the C++ host layer -- qsyn-shaped QCir,
QASM reader, to_tensor and `tensor equiv` -- for tests and bench.py.  No arithmetic here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        abi.lib()  # libqsx.so first (dependency)
        path = os.environ.get("QSXHOST_LIB", os.path.join(os.path.dirname(abi.lib_path()), "libqsxhost.so"))
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `make`")
        L = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        L.qsxh_last_error.restype = C.c_char_p
        sig = {
            "qsxh_set_process_shard": [C.c_int, C.c_int, C.c_int],
            "qsxh_circuit_new": [sz, C.POINTER(vp)],
            "qsxh_circuit_from_qasm_text": [C.c_char_p, C.POINTER(vp)],
            "qsxh_circuit_from_qasm_file": [C.c_char_p, C.POINTER(vp)],
            "qsxh_circuit_append": [vp, C.c_char_p, C.c_char_p, C.POINTER(sz), sz],
            "qsxh_circuit_append_circuit": [vp, vp, sz, C.POINTER(sz), sz],
            "qsxh_circuit_info": [vp, C.POINTER(sz), C.POINTER(sz)],
            "qsxh_circuit_gate_repr": [vp, sz, C.c_char_p, sz, C.POINTER(sz)],
            "qsxh_circuit_gate_stream": [vp, C.POINTER(C.POINTER(abi.GateStruct)), C.POINTER(sz)],
            "qsxh_circuit_adjoint_inplace": [vp],
            "qsxh_circuit_equiv": [vp, vp, C.POINTER(C.c_int)],
            "qsxh_circuit_destroy": [vp],
            "qsxh_qc2ts": [vp, C.POINTER(vp)],
            "qsxh_zxgraph_from_zx_text": [C.c_char_p, C.c_int, C.POINTER(vp)],
            "qsxh_zxgraph_from_zx_file": [C.c_char_p, C.c_int, C.POINTER(vp)],
            "qsxh_zxgraph_info": [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)],
            "qsxh_zx2ts": [vp, C.POINTER(vp)],
            "qsxh_zxgraph_destroy": [vp],
            "qsxh_tensor_shape": [vp, C.POINTER(sz), sz, C.POINTER(sz)],
            "qsxh_tensor_download": [vp, C.POINTER(C.c_double), sz],
            "qsxh_tensor_element": [vp, sz, sz, C.POINTER(C.c_double)],
            "qsxh_tensor_write": [vp, C.c_char_p],
            "qsxh_tensor_read": [C.c_char_p, C.POINTER(vp)],
            "qsxh_tensor_print": [vp, C.c_char_p, sz, C.POINTER(sz)],
            "qsxh_tensor_adjoint_inplace": [vp],
            "qsxh_tensor_equiv": [vp, vp, C.c_double, C.c_int, C.POINTER(abi.EquivResult), C.c_char_p, sz],
            "qsxh_tensor_destroy": [vp],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise abi.QsxError(rc, lib().qsxh_last_error().decode("utf-8", "replace"))


class Circuit:
    """qsyn::qcir::QCir behind the host C ABI."""

    def __init__(self, n_qubits: int | None = None, qasm_text: str | None = None, qasm_file: str | None = None):
        self._h = C.c_void_p()
        if qasm_text is not None:
            _check(lib().qsxh_circuit_from_qasm_text(qasm_text.encode(), C.byref(self._h)))
        elif qasm_file is not None:
            _check(lib().qsxh_circuit_from_qasm_file(qasm_file.encode(), C.byref(self._h)))
        else:
            _check(lib().qsxh_circuit_new(int(n_qubits or 0), C.byref(self._h)))

    def append(self, gate: str, qubits, phase: str | None = None) -> "Circuit":
        q = (C.c_size_t * len(qubits))(*[int(x) for x in qubits])
        _check(lib().qsxh_circuit_append(self._h, gate.encode(), phase.encode() if phase is not None else None, q, len(qubits)))
        return self

    def append_circuit(self, sub: "Circuit", qubits, n_ctrls: int = 0) -> "Circuit":
        """`sub` as ONE gate on `qubits` (controls first): a nested QCir used as an Operation."""
        q = (C.c_size_t * len(qubits))(*[int(x) for x in qubits])
        _check(lib().qsxh_circuit_append_circuit(self._h, sub._h, n_ctrls, q, len(qubits)))
        return self

    @property
    def n_qubits(self) -> int:
        n, g = C.c_size_t(), C.c_size_t()
        _check(lib().qsxh_circuit_info(self._h, C.byref(n), C.byref(g)))
        return n.value

    @property
    def n_gates(self) -> int:
        n, g = C.c_size_t(), C.c_size_t()
        _check(lib().qsxh_circuit_info(self._h, C.byref(n), C.byref(g)))
        return g.value

    def gate_reprs(self):
        """[(gate id, repr)] in topological (execution) order."""
        out = []
        buf = C.create_string_buffer(128)
        gid = C.c_size_t()
        for i in range(self.n_gates):
            _check(lib().qsxh_circuit_gate_repr(self._h, i, buf, 128, C.byref(gid)))
            out.append((gid.value, buf.value.decode()))
        return out

    def gate_array(self):
        """(pointer to qsx_gate[], count): the gate stream in the layout the device ABI takes."""
        p = C.POINTER(abi.GateStruct)()
        n = C.c_size_t()
        _check(lib().qsxh_circuit_gate_stream(self._h, C.byref(p), C.byref(n)))
        return p, n.value

    def gate_stream(self):
        """The same stream as python tuples (n_ctrls, qubits, matrix) -- for tests."""
        p, n = self.gate_array()
        out = []
        for i in range(n):
            g = p[i]
            k = g.n_ctrls + g.n_targets
            d = 1 << g.n_targets
            m = np.ctypeslib.as_array(g.matrix, shape=(2 * d * d,)).copy().view(np.complex128).reshape(d, d)
            out.append((g.n_ctrls, [g.qubits[j] for j in range(k)], m))
        return out

    def qc2ts(self) -> "Tensor":
        h = C.c_void_p()
        _check(lib().qsxh_qc2ts(self._h, C.byref(h)))
        return Tensor(h)

    def adjoint_inplace(self) -> "Circuit":
        _check(lib().qsxh_circuit_adjoint_inplace(self._h))
        return self

    def equiv(self, other: "Circuit") -> bool:
        """`qcir equiv`: the tensor half of qcir::is_equivalent on the device."""
        out = C.c_int(0)
        _check(lib().qsxh_circuit_equiv(self._h, other._h, C.byref(out)))
        return bool(out.value)

    def close(self):
        if getattr(self, "_h", None):
            lib().qsxh_circuit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZXGraph:
    """qsyn::zx::ZXGraph behind the host C ABI (`.zx` reader + zx2ts)."""

    def __init__(self, zx_text: str | None = None, zx_file: str | None = None, keep_id: bool = False):
        self._h = C.c_void_p()
        if zx_text is not None:
            _check(lib().qsxh_zxgraph_from_zx_text(zx_text.encode(), int(keep_id), C.byref(self._h)))
        else:
            _check(lib().qsxh_zxgraph_from_zx_file(str(zx_file).encode(), int(keep_id), C.byref(self._h)))

    @property
    def info(self):
        v = [C.c_size_t() for _ in range(4)]
        _check(lib().qsxh_zxgraph_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("inputs", "outputs", "vertices", "edges"), (x.value for x in v)))

    def zx2ts(self) -> "Tensor":
        h = C.c_void_p()
        _check(lib().qsxh_zx2ts(self._h, C.byref(h)))
        return Tensor(h)

    def close(self):
        if getattr(self, "_h", None):
            lib().qsxh_zxgraph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tensor:
    """qsyn::tensor::QTensor<double> behind the host C ABI."""

    def __init__(self, handle):
        self._h = handle

    @property
    def shape(self):
        dims = (C.c_size_t * 64)()
        nd = C.c_size_t()
        _check(lib().qsxh_tensor_shape(self._h, dims, 64, C.byref(nd)))
        return [dims[i] for i in range(nd.value)]

    def to_numpy(self) -> np.ndarray:
        shp = self.shape
        out = np.empty(int(np.prod(shp)), dtype=np.complex128)
        _check(lib().qsxh_tensor_download(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), 2 * out.size))
        return out.reshape(shp)

    def adjoint_inplace(self):
        _check(lib().qsxh_tensor_adjoint_inplace(self._h))

    def element(self, i: int, j: int) -> complex:
        v = (C.c_double * 2)()
        _check(lib().qsxh_tensor_element(self._h, i, j, v))
        return complex(v[0], v[1])

    def write(self, path: str) -> None:
        _check(lib().qsxh_tensor_write(self._h, str(path).encode()))

    @staticmethod
    def read(path: str) -> "Tensor":
        h = C.c_void_p()
        _check(lib().qsxh_tensor_read(str(path).encode(), C.byref(h)))
        return Tensor(h)

    def print_string(self) -> str:
        need = C.c_size_t(0)
        _check(lib().qsxh_tensor_print(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().qsxh_tensor_print(self._h, buf, need.value, None))
        return buf.value.decode()

    def equiv(self, other: "Tensor", eps: float = 1e-6, strict: bool = False):
        """`tensor equiv`: (qsx_equiv_result, the text qsyn prints)."""
        r = abi.EquivResult()
        buf = C.create_string_buffer(256)
        _check(lib().qsxh_tensor_equiv(self._h, other._h, eps, int(strict), C.byref(r), buf, 256))
        return r, buf.value.decode()

    def close(self):
        if getattr(self, "_h", None):
            lib().qsxh_tensor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
