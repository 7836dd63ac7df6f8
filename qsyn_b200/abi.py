"""ctypes binding of include/qsx.h.  No arithmetic here: every call goes to libqsx.so."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
QSX_MAX_GATE_QUBITS = 16
QSX_RUN_ASYNC = 1
QSX_RUN_ASYNC_IF_SHORT = 2

STATUS = {
    0: "QSX_OK", 1: "QSX_ERR_BAD_ARG", 2: "QSX_ERR_OOM", 3: "QSX_ERR_INTERRUPTED", 4: "QSX_ERR_CUDA",
    5: "QSX_ERR_NCCL", 6: "QSX_ERR_SHAPE", 7: "QSX_ERR_UNSUPPORTED",
}


class QsxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.status = STATUS.get(code, str(code))


def lib_path() -> str:
    return os.environ.get("QSX_LIB", os.path.join(_HERE, "lib", "libqsx.so"))


_lib = None


def lib() -> C.CDLL:
    """Loads libqsx.so; fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `make` (or __graft_entry__.build()); "
                              "there is no Python or CPU fallback for this path")
        _lib = C.CDLL(path)
        _declare(_lib)
        import atexit
        atexit.register(_lib.qsx_shutdown)  # compiler threads and communicators go before the interpreter tears down
    return _lib


class GateStruct(C.Structure):
    _fields_ = [("n_ctrls", C.c_int32), ("n_targets", C.c_int32), ("qubits", C.c_int32 * QSX_MAX_GATE_QUBITS),
                ("matrix", C.POINTER(C.c_double))]


class PlanOptions(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("fuse", C.c_int32), ("snap_tol", C.c_double), ("reg_qubits", C.c_int32),
                ("tile_qubits", C.c_int32), ("log2_tile_cols", C.c_int32), ("max_ops_per_sweep", C.c_int32),
                ("lookahead", C.c_int32), ("dense_block", C.c_int32), ("no_perm_folding", C.c_int32), ("min_tail_ops", C.c_int32),
                ("search", C.c_int32), ("no_diag_conjugation", C.c_int32),
                ("no_1q_fusion", C.c_int32), ("stage_search", C.c_int32),
                ("l2_block_mib", C.c_int32), ("jit", C.c_int32)]

    def __init__(self, fuse: int = 1, **kw):
        super().__init__()
        self.struct_size = C.sizeof(PlanOptions)
        self.fuse = fuse
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(k)
            setattr(self, k, v)


class PlanStats(C.Structure):
    _fields_ = [("n_gates", C.c_uint64), ("n_ops", C.c_uint64), ("n_sweeps", C.c_uint64), ("n_stages", C.c_uint64),
                ("n_micro_ops", C.c_uint64), ("n_folded_perms", C.c_uint64), ("n_dense_blocks", C.c_uint64), ("alg_bytes_per_run", C.c_double), ("gate_bytes_per_run", C.c_double),
                ("plan_host_ms", C.c_double)]


class JitStats(C.Structure):
    _fields_ = [("requested", C.c_uint64), ("cache_hits", C.c_uint64), ("compiled", C.c_uint64), ("failed", C.c_uint64),
                ("launches", C.c_uint64), ("compile_seconds", C.c_double), ("workers", C.c_int32), ("available", C.c_int32),
                ("disk_hits", C.c_uint64)]


class RunStats(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("device_ms", C.c_double), ("jit_launches", C.c_uint64), ("plan_cache_hit", C.c_int32),
                ("reserved", C.c_int32)]


class UnitaryInfo(C.Structure):
    _fields_ = [("n_qubits", C.c_int32), ("device", C.c_int32), ("col0", C.c_uint64), ("ncols", C.c_uint64),
                ("nrows", C.c_uint64), ("bytes", C.c_uint64), ("data", C.c_void_p)]


class EquivResult(C.Structure):
    _fields_ = [("equivalent", C.c_int32), ("phase_defined", C.c_int32), ("cosine", C.c_double), ("norm", C.c_double),
                ("phase_rad", C.c_double), ("phase_num", C.c_int32), ("phase_den", C.c_int32)]


STOP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


def _declare(L: C.CDLL) -> None:
    vp, i32, u64, dp = C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_double)
    L.qsx_abi_version.restype = C.c_int
    L.qsx_last_error.restype = C.c_char_p
    sig = {
        "qsx_device_count": [C.POINTER(C.c_int)],
        "qsx_device_stream": [i32, C.POINTER(vp)],
        "qsx_device_synchronize": [i32],
        "qsx_device_trim": [i32],
        "qsx_unitary_create": [i32, i32, u64, u64, C.POINTER(vp)],
        "qsx_unitary_set_identity": [vp],
        "qsx_unitary_clone": [vp, C.POINTER(vp)],
        "qsx_unitary_destroy": [vp],
        "qsx_unitary_get_info": [vp, C.POINTER(UnitaryInfo)],
        "qsx_unitary_upload": [vp, u64, u64, dp],
        "qsx_unitary_download": [vp, u64, u64, dp],
        "qsx_plan_create": [i32, u64, C.POINTER(GateStruct), C.c_size_t, C.POINTER(PlanOptions), C.POINTER(vp)],
        "qsx_plan_destroy": [vp],
        "qsx_plan_get_stats": [vp, C.POINTER(PlanStats)],
        "qsx_plan_dump": [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)],
        "qsx_plan_run": [vp, vp, STOP_FN, vp, C.c_uint32, C.POINTER(RunStats)],
        "qsx_apply_gates": [vp, C.POINTER(GateStruct), C.c_size_t, C.POINTER(PlanOptions), STOP_FN, vp,
                            C.POINTER(RunStats)],
        "qsx_apply_gates_async": [vp, C.POINTER(GateStruct), C.c_size_t, C.POINTER(PlanOptions), C.POINTER(RunStats)],
        "qsx_apply_gates_ex": [vp, C.POINTER(GateStruct), C.c_size_t, C.POINTER(PlanOptions), STOP_FN, vp, C.c_uint32,
                               C.POINTER(RunStats)],
        "qsx_plan_compile": [vp, i32],
        "qsx_jit_get_stats": [C.POINTER(JitStats)],
        "qsx_init": [i32],
        "qsx_shutdown": [],
        "qsx_unitary_download_block": [vp, u64, u64, u64, u64, dp, u64],
        "qsx_unitary_upload_block": [vp, u64, u64, u64, u64, dp, u64],
        "qsx_equiv_partial": [vp, vp, dp],
        "qsx_equiv_reduce": [C.POINTER(vp), C.POINTER(vp), i32, dp],
        "qsx_comm_get_unique_id": [vp, C.c_size_t],
        "qsx_comm_init_rank": [i32, i32, i32, vp, C.c_size_t],
        "qsx_comm_info": [C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "qsx_equiv_finish": [dp, C.c_double, i32, C.POINTER(EquivResult)],
        "qsx_unitary_adjoint_inplace": [vp],
        "qsx_unitary_adjoint_sharded": [C.POINTER(vp), C.c_int],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int


def _check(rc: int) -> None:
    if rc != 0:
        raise QsxError(rc, lib().qsx_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().qsx_device_count(C.byref(n)))
    return n.value


def pack_gates(gates: Iterable[tuple]):
    """gates: iterable of (n_ctrls, qubits[controls first], target_matrix 2^t x 2^t complex128).
    Returns (ctypes array, keepalive list)."""
    gates = list(gates)
    arr = (GateStruct * max(len(gates), 1))()
    keep = []
    for i, (n_ctrls, qubits, mat) in enumerate(gates):
        m = np.ascontiguousarray(np.asarray(mat, dtype=np.complex128))
        nt = len(qubits) - n_ctrls
        assert m.shape == (2**nt, 2**nt), (m.shape, nt)
        keep.append(m)
        arr[i].n_ctrls = n_ctrls
        arr[i].n_targets = nt
        for k, q in enumerate(qubits):
            arr[i].qubits[k] = int(q)
        arr[i].matrix = m.ctypes.data_as(C.POINTER(C.c_double))
    return arr, keep, len(gates)


class Plan:
    def __init__(self, n_qubits: int, ncols: int, gates, options: PlanOptions | None = None):
        """gates: iterable of (n_ctrls, qubits, matrix) tuples, or a (qsx_gate pointer, count) pair
        as returned by qsyn_b200.host.Circuit.gate_array()."""
        if isinstance(gates, tuple) and len(gates) == 2 and isinstance(gates[1], int):
            arr, n = gates
        else:
            arr, keep, n = pack_gates(gates)
        self._h = C.c_void_p()
        _check(lib().qsx_plan_create(n_qubits, ncols, arr, n, C.byref(options) if options is not None else None,
                                     C.byref(self._h)))
        self.n_qubits, self.ncols = n_qubits, ncols

    def stats(self) -> PlanStats:
        s = PlanStats()
        _check(lib().qsx_plan_get_stats(self._h, C.byref(s)))
        return s

    def dump(self) -> dict:
        need = C.c_size_t(0)
        _check(lib().qsx_plan_dump(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().qsx_plan_dump(self._h, buf, need.value, None))
        return json.loads(buf.value.decode())

    def run(self, u: "Unitary", should_stop=None, flags: int = 0) -> RunStats:
        st = RunStats()
        cb = STOP_FN(lambda _arg: int(bool(should_stop()))) if should_stop is not None else C.cast(None, STOP_FN)
        _check(lib().qsx_plan_run(self._h, u._h, cb, None, flags, C.byref(st)))
        return st

    def compile(self, wait: bool = True) -> None:
        """qsx_plan_compile: request (and optionally wait for) the specialised kernels of every sweep."""
        _check(lib().qsx_plan_compile(self._h, int(wait)))

    def close(self):
        if self._h:
            lib().qsx_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Unitary:
    """One column shard of U[out,in] on one device."""

    def __init__(self, n_qubits: int, device: int = 0, col0: int = 0, ncols: int | None = None, identity: bool = True,
                 _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            ncols = 2**n_qubits if ncols is None else ncols
            self._h = C.c_void_p()
            _check(lib().qsx_unitary_create(device, n_qubits, col0, ncols, C.byref(self._h)))
            if identity:
                _check(lib().qsx_unitary_set_identity(self._h))

    @property
    def info(self) -> UnitaryInfo:
        i = UnitaryInfo()
        _check(lib().qsx_unitary_get_info(self._h, C.byref(i)))
        return i

    def set_identity(self):
        _check(lib().qsx_unitary_set_identity(self._h))

    def clone(self) -> "Unitary":
        h = C.c_void_p()
        _check(lib().qsx_unitary_clone(self._h, C.byref(h)))
        return Unitary(0, _handle=h)

    def download(self, row0: int = 0, nrows: int | None = None) -> np.ndarray:
        i = self.info
        nrows = i.nrows - row0 if nrows is None else nrows
        out = np.empty((nrows, i.ncols), dtype=np.complex128)
        _check(lib().qsx_unitary_download(self._h, row0, nrows, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def download_block(self, row0: int, nrows: int, c0: int, nc: int) -> np.ndarray:
        """rows [row0, row0+nrows) x shard-local columns [c0, c0+nc): qsx_unitary_download_block."""
        out = np.empty((nrows, nc), dtype=np.complex128)
        _check(lib().qsx_unitary_download_block(self._h, row0, nrows, c0, nc, out.ctypes.data_as(C.POINTER(C.c_double)), 0))
        return out

    def upload(self, host: np.ndarray, row0: int = 0):
        i = self.info
        h = np.ascontiguousarray(host, dtype=np.complex128)
        assert h.ndim == 2 and h.shape[1] == i.ncols
        _check(lib().qsx_unitary_upload(self._h, row0, h.shape[0], h.ctypes.data_as(C.POINTER(C.c_double))))

    def apply(self, gates: Iterable[tuple], options: PlanOptions | None = None, should_stop=None) -> RunStats:
        arr, keep, n = pack_gates(gates)
        st = RunStats()
        cb = STOP_FN(lambda _arg: int(bool(should_stop()))) if should_stop is not None else C.cast(None, STOP_FN)
        _check(lib().qsx_apply_gates(self._h, arr, n, C.byref(options) if options is not None else None, cb, None,
                                     C.byref(st)))
        return st

    def apply_ex(self, gates: Iterable[tuple], flags: int, options: PlanOptions | None = None, should_stop=None) -> RunStats:
        """qsx_apply_gates_ex: apply with run flags (QSX_RUN_ASYNC, QSX_RUN_ASYNC_IF_SHORT) and the stop callback."""
        arr, keep, n = pack_gates(gates)
        st = RunStats()
        cb = STOP_FN(lambda _arg: int(bool(should_stop()))) if should_stop is not None else C.cast(None, STOP_FN)
        _check(lib().qsx_apply_gates_ex(self._h, arr, n, C.byref(options) if options is not None else None, cb, None, flags,
                                        C.byref(st)))
        return st

    def apply_async(self, gates, options: PlanOptions | None = None) -> RunStats:
        """qsx_apply_gates_async: returns once the sweeps are enqueued (gates: tuples or (qsx_gate*, count))."""
        if isinstance(gates, tuple) and len(gates) == 2 and isinstance(gates[1], int):
            arr, n, keep = gates[0], gates[1], None
        else:
            arr, keep, n = pack_gates(gates)
        st = RunStats()
        _check(lib().qsx_apply_gates_async(self._h, arr, n, C.byref(options) if options is not None else None, C.byref(st)))
        return st  # (the gate matrices are only read while scheduling, i.e. before the call returns)

    def equiv_partial(self, other: "Unitary") -> np.ndarray:
        out = np.zeros(8, dtype=np.float64)
        _check(lib().qsx_equiv_partial(self._h, other._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def adjoint_inplace(self):
        _check(lib().qsx_unitary_adjoint_inplace(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().qsx_unitary_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def adjoint_sharded(shards: Sequence["Unitary"]) -> None:
    """qsx_unitary_adjoint_sharded: shards[k] holds columns [k c, (k+1) c), possibly on different devices."""
    arr = (C.c_void_p * len(shards))(*[u._h for u in shards])
    _check(lib().qsx_unitary_adjoint_sharded(arr, len(shards)))


def equiv_reduce(a: Sequence["Unitary"], b: Sequence["Unitary"]) -> np.ndarray:
    """qsx_equiv_reduce: the fused reduction on every shard pair + ONE NCCL all-reduce of the 8 sums (over the
    devices of this process, or over the ranks after comm_init_rank)."""
    assert len(a) == len(b) and len(a) >= 1
    aa = (C.c_void_p * len(a))(*[u._h for u in a])
    bb = (C.c_void_p * len(b))(*[u._h for u in b])
    out = np.zeros(8, dtype=np.float64)
    _check(lib().qsx_equiv_reduce(aa, bb, len(a), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


QSX_COMM_ID_BYTES = 128


def comm_get_unique_id() -> bytes:
    buf = C.create_string_buffer(QSX_COMM_ID_BYTES)
    _check(lib().qsx_comm_get_unique_id(buf, QSX_COMM_ID_BYTES))
    return buf.raw


def comm_init_rank(device: int, world: int, rank: int, uid: bytes) -> None:
    buf = C.create_string_buffer(uid, QSX_COMM_ID_BYTES)
    _check(lib().qsx_comm_init_rank(device, world, rank, buf, QSX_COMM_ID_BYTES))


def comm_info() -> tuple[int, int]:
    w, v = C.c_int(0), C.c_int(0)
    _check(lib().qsx_comm_info(C.byref(w), C.byref(v)))
    return w.value, v.value


def jit_stats() -> JitStats:
    s = JitStats()
    _check(lib().qsx_jit_get_stats(C.byref(s)))
    return s


def init(n_gpus: int = 0) -> None:
    _check(lib().qsx_init(n_gpus))


def shutdown() -> None:
    _check(lib().qsx_shutdown())


def equiv_finish(sums: Sequence[float], eps: float = 1e-6, strict: bool = False) -> EquivResult:
    s = np.ascontiguousarray(np.asarray(sums, dtype=np.float64))
    r = EquivResult()
    _check(lib().qsx_equiv_finish(s.ctypes.data_as(C.POINTER(C.c_double)), eps, int(strict), C.byref(r)))
    return r


def device_stream(device: int = 0) -> int:
    p = C.c_void_p()
    _check(lib().qsx_device_stream(device, C.byref(p)))
    return p.value


def device_trim(device: int = 0) -> None:
    _check(lib().qsx_device_trim(device))


def device_synchronize(device: int = 0) -> None:
    _check(lib().qsx_device_synchronize(device))
