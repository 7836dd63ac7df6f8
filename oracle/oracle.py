"""CPU oracle for qsyn's tensor-verification hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (DVLab-NTU/qsyn
v0.8.1) behind ``convert qcir tensor`` and ``tensor equiv``.  It is the
*checker* for the CUDA path in ``qsyn_b200/``: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``qsyn_b200/`` does.

Why a restatement: the arithmetic of the reference lives in un-vendored
third-party code (xtensor 0.26.0, xtensor-blas 0.22.0, xtl 0.8.0 on a system
OpenBLAS/LAPACK, CMakeLists.txt:87-110) that is absent from /root/reference,
so the reference cannot be compiled here.  numpy's ``tensordot`` is the same
transpose -> reshape -> zgemm scheme as ``xt::linalg::tensordot``.

Parity pinning: checked by tests/test_oracle_golden.py against every ZX-free
golden vector the reference's own tests hold for this path:
sk.log 2x2 known-answer matrix, global_norm_phase.log (pi/4), error.log (eps
ladder), size.log (shape mismatch), qc2ts.log (gate order + pin trace, 603
lines).  Large-n values are "parity unpinned" by the reference's own tests
(largest pinned numeric matrix is 2x2); they are judged against this oracle.

All ``file:line`` citations are relative to /root/reference/.
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np

# --------------------------------------------------------------------------
# dvlab::Rational / dvlab::Phase   (src/util/rational.hpp, src/util/phase.hpp)
# --------------------------------------------------------------------------


class Rational:
    """src/util/rational.hpp:27-148.  int numerator/denominator, always reduced."""

    __slots__ = ("num", "den")

    def __init__(self, num: int = 0, den: int = 1):
        if den == 0:
            raise ZeroDivisionError("Rational with zero denominator")
        if den < 0:  # reduce(), rational.hpp:130-138
            num, den = -num, -den
        g = math.gcd(int(num), int(den))
        if g == 0:
            g = 1
        self.num = int(num) // g
        self.den = int(den) // g

    # rational.hpp:97-128 : Stern-Brocot mediant search in [f-eps, f+eps]
    @staticmethod
    def to_rational(f: float, eps: float = 1e-4) -> "Rational":
        if not math.isfinite(f):
            # Appendix C trap 2: the reference has UB / never terminates here.
            raise ValueError("to_rational of a non-finite value")
        integral_part = int(math.floor(f))
        f -= integral_part
        lower, upper, med = (0, 1), (1, 1), (1, 2)

        def in_lower(q):
            return (f - eps) <= q[0] / q[1]

        def in_upper(q):
            return (f + eps) >= q[0] / q[1]

        if in_lower(lower) and in_upper(lower):
            return Rational(lower[0] + integral_part * lower[1], lower[1])
        if in_lower(upper) and in_upper(upper):
            return Rational(upper[0] + integral_part * upper[1], upper[1])
        while True:
            if not in_lower(med):
                lower = med
            elif not in_upper(med):
                upper = med
            else:
                return Rational(med[0] + integral_part * med[1], med[1])
            med = (lower[0] + upper[0], lower[1] + upper[1])  # rational.cpp:24-26

    def __mul__(self, o: "Rational") -> "Rational":
        return Rational(self.num * o.num, self.den * o.den)

    def __sub__(self, o: "Rational") -> "Rational":
        return Rational(self.num * o.den - self.den * o.num, self.den * o.den)

    def __add__(self, o: "Rational") -> "Rational":
        return Rational(self.num * o.den + self.den * o.num, self.den * o.den)

    def __eq__(self, o) -> bool:
        return isinstance(o, Rational) and self.num == o.num and self.den == o.den

    def __hash__(self):
        return hash((self.num, self.den))

    def to_float(self) -> float:  # rational.hpp:78
        return float(self.num) / float(self.den)

    def __repr__(self):
        return f"Rational({self.num},{self.den})"


class Phase:
    """src/util/phase.hpp:31-106.  A rational multiple of pi in (-pi, pi]."""

    __slots__ = ("r",)

    def __init__(self, num: int = 0, den: int = 1):
        self.r = Rational(num, den)
        self._normalize()

    @staticmethod
    def from_rational(r: Rational) -> "Phase":
        return Phase(r.num, r.den)

    @staticmethod
    def from_float(f: float, eps: float = 1e-4) -> "Phase":
        """phase.hpp:38-40 : Rational(f/pi, eps/pi) then normalize."""
        return Phase.from_rational(Rational.to_rational(f / math.pi, eps / math.pi))

    def _normalize(self):  # phase.hpp:239-243
        q = Rational(self.r.num, self.r.den * 2)  # _rational / 2
        # (num - (num >= 0 ? 0 : den)) / den with C++ truncating integer division
        n = q.num - (0 if q.num >= 0 else q.den)
        fl = n // q.den if n >= 0 else -((-n) // q.den)
        self.r = self.r - Rational(fl * 2, 1)
        if self.r.num > self.r.den:  # _rational > 1
            self.r = self.r - Rational(2, 1)

    @property
    def numerator(self) -> int:
        return self.r.num

    @property
    def denominator(self) -> int:
        return self.r.den

    def to_float(self) -> float:  # phase.hpp:65-67 : pi * (num/den)
        return math.pi * self.r.to_float()

    def __neg__(self) -> "Phase":
        return Phase(-self.r.num, self.r.den)

    def __eq__(self, o) -> bool:
        return isinstance(o, Phase) and self.r == o.r

    def __hash__(self):
        return hash(self.r)

    def times(self, q: Rational) -> "Phase":
        return Phase.from_rational(self.r * q)

    def print_string(self) -> str:  # phase.cpp:39-46
        n, d = self.r.num, self.r.den
        s = "" if n == 1 else ("-" if n == -1 else str(n))
        s += "π" if n != 0 else ""
        s += f"/{d}" if d != 1 else ""
        return s

    def __repr__(self):
        return f"Phase({self.print_string()})"


def _stoi_full(s: str):
    """std::stoi + 'whole string consumed' (dvlab_string.hpp:186-197)."""
    t = s.lstrip(" \t\n\v\f\r")
    if t != s and not t:
        return None
    body = t
    if body[:1] in "+-":
        body = body[1:]
    if not body or not body.isdigit() or not body.isascii():
        return None
    v = int(t)
    if v < -(2**31) or v > 2**31 - 1:
        return None
    return v


def _stod_full(s: str):
    """std::stod + 'whole string consumed'."""
    t = s.lstrip(" \t\n\v\f\r")
    if not t or t[-1:].isspace():
        return None
    try:
        return float(t)
    except ValueError:
        return None


def str_to_phase(s: str):
    """phase.hpp:167-233.  Returns Phase or None."""
    number_strings, operators = [], []
    curr = 0
    while True:
        idxs = [i for i in (s.find("*", curr), s.find("/", curr)) if i != -1]
        if idxs:
            op = min(idxs)
            operators.append(s[op])
            number_strings.append(s[curr:op])
            curr = op + 1
        else:
            number_strings.append(s[curr:])
            break
    if len(operators) >= len(number_strings):
        return None
    n_pis, numerator, denominator, temp_float = 0, 1, 1, 1.0
    for i, tok in enumerate(number_strings):
        do_div = i != 0 and operators[i - 1] == "/"
        low = tok.lower()
        if low == "pi":
            n_pis += -1 if do_div else 1
        elif low == "-pi":
            numerator *= -1
            n_pis += -1 if do_div else 1
        elif (iv := _stoi_full(tok)) is not None:
            if do_div:
                denominator *= iv
            else:
                numerator *= iv
        elif (fv := _stod_full(tok)) is not None:
            if do_div:
                temp_float /= fv
            else:
                temp_float *= fv
        else:
            return None
    tmp = Rational.to_rational(temp_float * math.pi ** (n_pis - 1), 1e-4 / math.pi)
    return Phase(numerator, denominator).times(tmp)


# --------------------------------------------------------------------------
# Gate matrices   (src/tensor/qtensor.hpp:121-277,346-372; qcir_to_tensor.cpp:28-94)
# All returned as 2^k x 2^k complex128, first gate qubit = MSB.
# --------------------------------------------------------------------------


def _nu_pow(n: int) -> float:  # qtensor.hpp:431-433
    return math.pow(2.0, -0.25 * n)


def _expi(phi: float) -> complex:
    # std::exp(1.0i * phi)
    return cmath.exp(1j * phi)


def mat_h() -> np.ndarray:  # hbox(2), qtensor.hpp:184-193
    t = np.ones((2, 2), dtype=np.complex128)
    t[1, 1] = -1.0
    return t * _nu_pow(2)


def mat_pz(p: Phase) -> np.ndarray:  # zspider(2), qtensor.hpp:142-153
    t = np.zeros((2, 2), dtype=np.complex128)
    t[0, 0] = 1.0
    t[1, 1] = _expi(p.to_float())
    return t * _nu_pow(0)


def mat_px(p: Phase) -> np.ndarray:  # xspider(2), qtensor.hpp:164-173
    t = np.ones((2, 2), dtype=np.complex128)
    ket_minus = np.array([1.0 + 0j, -1.0 + 0j])
    tmp = np.tensordot(ket_minus, ket_minus, axes=0)
    t = t + tmp * _expi(p.to_float())
    t = t / math.pow(math.sqrt(2), 2)  # 2.0000000000000004, Appendix C trap 3
    return t * _nu_pow(0)


def mat_py(p: Phase) -> np.ndarray:  # qtensor.hpp:259-265
    sdg = mat_pz(Phase(-1, 2))
    px = mat_px(p)
    s = mat_pz(Phase(1, 2))
    return s @ (px @ sdg)


def mat_rz(p: Phase) -> np.ndarray:  # qtensor.hpp:232-238
    return mat_pz(p) * cmath.exp(-0.5j * p.to_float())


def mat_rx(p: Phase) -> np.ndarray:  # qtensor.hpp:203-209
    return mat_px(p) * cmath.exp(-0.5j * p.to_float())


def mat_ry(p: Phase) -> np.ndarray:  # qtensor.hpp:217-223
    return mat_py(p) * cmath.exp(-0.5j * p.to_float())


def mat_swap() -> np.ndarray:  # qcir_to_tensor.cpp:38-45
    return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def mat_ecr() -> np.ndarray:  # qcir_to_tensor.cpp:47-55
    r = math.sqrt(2)
    a, b = 1.0 / r, 1j / r
    return np.array([[0, 0, a, b], [0, 0, b, a], [a, -b, 0, 0], [-b, a, 0, 0]], dtype=np.complex128)


def mat_control(m: np.ndarray, n_ctrls: int) -> np.ndarray:
    """qtensor.hpp:346-372 : direct_sum(I, M), controls are the MSBs."""
    if n_ctrls == 0:
        return m
    gs = m.shape[0]
    ident = gs * (2**n_ctrls - 1)
    out = np.zeros((ident + gs, ident + gs), dtype=np.complex128)
    out[:ident, :ident] = np.eye(ident)
    out[ident:, ident:] = m
    return out


_FIXED = {"z": (1, 1), "s": (1, 2), "sdg": (-1, 2), "t": (1, 4), "tdg": (-1, 4)}


@dataclass
class Op:
    """A type-erased Operation: basic kind + phase + number of controls
    (src/qcir/operation.hpp:61-188, basic_gate_type.hpp)."""

    kind: str  # id h swap ecr pz px py rz rx ry
    phase: Phase | None = None
    n_ctrls: int = 0

    @property
    def n_targets(self) -> int:
        return 2 if self.kind in ("swap", "ecr") else 1

    @property
    def n_qubits(self) -> int:
        return self.n_targets + self.n_ctrls

    def target_matrix(self) -> np.ndarray:
        k = self.kind
        if k == "id":
            return np.eye(2, dtype=np.complex128)
        if k == "h":
            return mat_h()
        if k == "swap":
            return mat_swap()
        if k == "ecr":
            return mat_ecr()
        return {"pz": mat_pz, "px": mat_px, "py": mat_py, "rz": mat_rz, "rx": mat_rx, "ry": mat_ry}[k](self.phase)

    def matrix(self) -> np.ndarray:
        return mat_control(self.target_matrix(), self.n_ctrls)

    def repr(self) -> str:  # get_repr(), basic_gate_type.hpp:52-66 etc.
        k = self.kind
        if k in ("id", "h", "swap", "ecr"):
            base = k
        elif k in ("pz", "px", "py"):
            axis = k[1]
            names = {
                (1, 1): {"z": "z", "x": "x", "y": "y"},
                (1, 2): {"z": "s", "x": "sx", "y": "sy"},
                (-1, 2): {"z": "sdg", "x": "sxdg", "y": "sydg"},
                (1, 4): {"z": "t", "x": "tx", "y": "ty"},
                (-1, 4): {"z": "tdg", "x": "txdg", "y": "tydg"},
            }
            key = (self.phase.numerator, self.phase.denominator)
            if key in names:
                base = names[key][axis]
            else:
                base = f"{'p' if k == 'pz' else k}({self.phase.print_string()})"
        else:
            base = f"{k}({self.phase.print_string()})"
        return "c" * self.n_ctrls + base


def str_to_operation(name: str, params: Sequence[Phase] = ()) -> Op | None:
    """src/qcir/gate_type.cpp:16-70."""
    s = name.lower()
    n_ctrls = len(s) - len(s.lstrip("c"))
    if n_ctrls == len(s):
        # find_first_not_of('c') == npos -> substr(npos) throws in the reference
        return None
    s = s[n_ctrls:]
    op = None
    if len(params) == 0:
        if s == "id":
            op = Op("id")
        elif s == "h":
            op = Op("h")
        elif s == "swap":
            op = Op("swap")
        elif s == "ecr":
            op = Op("ecr")
        elif s in _FIXED:
            op = Op("pz", Phase(*_FIXED[s]))
        elif s in ("x", "not"):
            op = Op("px", Phase(1, 1))
        elif s in ("sx", "x_1_2"):
            op = Op("px", Phase(1, 2))
        elif s == "sxdg":
            op = Op("px", Phase(-1, 2))
        elif s == "tx":
            op = Op("px", Phase(1, 4))
        elif s == "txdg":
            op = Op("px", Phase(-1, 4))
        elif s == "y":
            op = Op("py", Phase(1, 1))
        elif s in ("sy", "y_1_2"):
            op = Op("py", Phase(1, 2))
        elif s == "sydg":
            op = Op("py", Phase(-1, 2))
        elif s == "ty":
            op = Op("py", Phase(1, 4))
        elif s == "tydg":
            op = Op("py", Phase(-1, 4))
    if op is None and len(params) == 1:
        if s in ("p", "pz"):
            op = Op("pz", params[0])
        elif s in ("px", "py", "rz", "rx", "ry"):
            op = Op(s, params[0])
    if op is None:
        return None
    op.n_ctrls = n_ctrls
    return op


# --------------------------------------------------------------------------
# QCir: gate container + topological order
# (src/qcir/qcir.cpp:281-305, src/qcir/qcir_action.cpp:65-122)
# --------------------------------------------------------------------------


@dataclass
class Gate:
    gid: int
    op: Op
    qubits: list


@dataclass
class QCir:
    n_qubits: int = 0
    gates: list = field(default_factory=list)  # by id (append order)
    filename: str = ""
    procedures: list = field(default_factory=list)

    def add_qubits(self, n: int = 1):
        self.n_qubits += n

    def append(self, op: Op, qubits: Sequence[int]) -> int:
        assert op.n_qubits == len(qubits), (op, qubits)
        assert all(0 <= q < self.n_qubits for q in qubits)
        self.gates.append(Gate(len(self.gates), op, list(qubits)))
        return len(self.gates) - 1

    def topological_order(self) -> list:
        """Reverse DFS post-order, qcir_action.cpp:65-105."""
        n_g = len(self.gates)
        succ = [[None] * len(g.qubits) for g in self.gates]
        first = [None] * self.n_qubits
        last = [None] * self.n_qubits
        for g in self.gates:
            for pin, q in enumerate(g.qubits):
                if last[q] is not None:
                    pg = self.gates[last[q]]
                    succ[pg.gid][pg.qubits.index(q)] = g.gid
                else:
                    first[q] = g.gid
                last[q] = g.gid
        stack = [(False, first[q]) for q in range(self.n_qubits) if first[q] is not None]
        visited = [False] * n_g
        post = []
        while stack:
            done, node = stack.pop()
            if done:
                post.append(node)
                continue
            if visited[node]:
                continue
            visited[node] = True
            stack.append((True, node))
            for s in succ[node]:
                if s is not None and not visited[s]:
                    stack.append((False, s))
        post.reverse()
        return [self.gates[i] for i in post]


def _str_get_token(s: str, pos: int = 0, delim: str = " \t\n\v\f\r"):
    """dvlab_string.cpp:62-72.  Returns (token, end) ; end == -1 is npos."""
    if pos == -1 or pos > len(s):
        # find_first_not_of(delim, npos) -> npos
        return "", -1
    begin = -1
    for i in range(pos, len(s)):
        if s[i] not in delim:
            begin = i
            break
    if begin == -1:
        return "", -1
    end = -1
    for i in range(begin, len(s)):
        if s[i] in delim:
            end = i
            break
    tok = s[begin:] if end == -1 else s[begin:end]
    return tok, end


def qcir_from_qasm_text(text: str) -> QCir | None:
    """src/qcir/qcir_reader.cpp:47-120, quirks included."""
    # six whitespace-separated tokens: OPENQASM 2.0; include "qelib1.inc"; qreg q[N];
    pos, toks = 0, []
    n = len(text)
    for _ in range(6):
        while pos < n and text[pos].isspace():
            pos += 1
        start = pos
        while pos < n and not text[pos].isspace():
            pos += 1
        toks.append(text[start:pos])
    s = toks[-1]
    lb = s.find("[")
    nqubit = int(s[lb + 1: lb + 1 + (len(s) - lb - 3)])
    qc = QCir(nqubit)
    rest = text[pos:]
    lines = rest.split("\n")
    # getline() consumes the remainder of the qreg line
    for line in lines[1:]:
        cut = line.find("//")
        if cut != -1:
            line = line[:cut]
        line = line.strip(" \t\n\v\f\r")
        if not line:
            continue
        typ, type_end = _str_get_token(line)
        phase_str = "0"
        _, e = _str_get_token(line, 0, "(")
        if e != -1:
            typ, stop = _str_get_token(line, 0, "(")
            phase_str, _ = _str_get_token(line, stop + 1, ")")
        else:
            phase_str = "0"
        if typ in ("creg", "qreg", ""):
            continue
        qubit_ids = []
        token, npos = _str_get_token(line, type_end, ",")
        while token:
            _, lbr = _str_get_token(token, 0, "[")
            qid_str, _ = _str_get_token(token, lbr + 1, "]")
            if not (qid_str.isdigit() and qid_str.isascii()) or int(qid_str) >= nqubit:
                return None  # "invalid qubit id on line"
            qubit_ids.append(int(qid_str))
            token, npos = _str_get_token(line, npos, ",")
        if len(set(qubit_ids)) != len(qubit_ids):
            return None  # "duplicate qubit id on line"
        op = str_to_operation(typ)
        if op is not None:
            qc.append(op, qubit_ids)
            continue
        ph = str_to_phase(phase_str)
        if ph is None:
            return None  # "invalid phase on line"
        op = str_to_operation(typ, [ph])
        if op is not None:
            qc.append(op, qubit_ids)
            continue
        # silently skipped (Appendix C trap 7)
    return qc


def qcir_from_qasm(path: str) -> QCir | None:
    with open(path, "r") as f:
        qc = qcir_from_qasm_text(f.read())
    if qc is not None:
        qc.filename = path
    return qc


# --------------------------------------------------------------------------
# to_tensor(QCir)   (src/convert/qcir_to_tensor.cpp:147-214)
# --------------------------------------------------------------------------


def gate_tensor(op: Op) -> np.ndarray:
    """Rank-2k interleaved [o0,i0,o1,i1,...] gate tensor (to_qtensor(), qtensor.hpp:315-330)."""
    k = op.n_qubits
    m = op.matrix().reshape([2] * (2 * k))
    ax = []
    for i in range(k):
        ax += [i, i + k]
    return np.transpose(m, ax)


def to_tensor_literal(qc: QCir, trace: list | None = None) -> np.ndarray | None:
    """The reference's algorithm verbatim: one tensordot per gate with the
    pin bookkeeping of update_tensor_pin (qcir_to_tensor.cpp:116-136,147-214).
    Returns the rank-2n interleaved tensor.  If `trace` is a list, the
    [debug]/[trace] log lines are appended to it exactly as spdlog prints them.
    """
    n = qc.n_qubits
    if n == 0:
        if trace is not None:
            trace.append("[warn]     QCir is empty!!")
        return None
    if trace is not None:
        trace.append("[debug]    Add boundary")
    tensor = np.array(1.0 + 0j)
    ident = np.eye(2, dtype=np.complex128)
    for _ in range(n):
        tensor = np.tensordot(tensor, ident, axes=0)
    # unordered_map<qubit, (out,in)> with libstdc++: keys emplaced 0..n-1 iterate
    # in reverse insertion order while they share the initial single bucket chain
    # (Appendix C trap 5).  We keep insertion order and iterate reversed.
    pins = {}
    for q in range(n):
        pins[q] = (2 * q, 2 * q + 1)
        if trace is not None:
            trace.append(f"[trace]      - Add Qubit: {q} input: {2 * q + 1} output: {2 * q}")
    for g in qc.topological_order():
        if trace is not None:
            trace.append(f"[debug]    Gate {g.gid} ({g.op.repr()})")
        gt = gate_tensor(g.op)
        k = len(g.qubits)
        gate_in = [2 * p + 1 for p in range(k)]
        main_out = [pins[q][0] for q in g.qubits]
        nd_main = tensor.ndim
        tensor = np.tensordot(gt, tensor, axes=(gate_in, main_out))
        # _axis_history of the result (tensor.hpp:419-430)
        hist, nid = {}, 0
        for i in range(2 * k):
            if i in gate_in:
                continue
            hist[i] = nid
            nid += 1
        for i in range(nd_main):
            if i in main_out:
                continue
            hist[i + 2 * k] = nid
            nid += 1
        if trace is not None:
            trace.append("[trace]    Pin Permutation")
        for q in _unordered_map_iteration_order(n):
            old_out, old_in = pins[q]
            if q in g.qubits:
                new_out = hist[2 * g.qubits.index(q)]
            else:
                new_out = hist[2 * k + old_out]
            new_in = hist[2 * k + old_in]
            pins[q] = (new_out, new_in)
            if trace is not None:
                trace.append(f"[trace]      - Qubit: {q} input : {old_in} -> {new_in} output: {old_out} -> {new_out}")
    out_pins = [pins[q][0] for q in range(n)]
    in_pins = [pins[q][1] for q in range(n)]
    mat = np.transpose(tensor, out_pins + in_pins).reshape(2**n, 2**n)  # to_matrix, tensor.hpp:266-274
    return matrix_to_interleaved(mat)  # to_qtensor, qcir_to_tensor.cpp:208


def to_tensor_literal_columns(qc: QCir, col0: int, ncols: int, max_gates: int | None = None) -> np.ndarray:
    """The same per-gate tensordot loop (qcir_to_tensor.cpp:173-194) on a COLUMN BLOCK of the unitary: the n input
    axes, which no gate ever contracts, are collapsed into one trailing axis of `ncols` columns.  Every gate is still
    one np.tensordot(gate, main) over the gate's input pins and the main tensor's output pins (same transposed copy +
    zgemm scheme), and its cost is linear in the number of columns -- this is the bounded sample the CPU baseline of
    bench.py times at n >= 14, where the full tensor (3 x 16 x 4^n B) does not fit a host.  Returns U[:, col0:col0+ncols]."""
    n = qc.n_qubits
    dim = 2**n
    tensor = np.zeros((dim, ncols), dtype=np.complex128)
    for c in range(ncols):
        tensor[col0 + c, c] = 1.0
    tensor = tensor.reshape([2] * n + [ncols])
    out_pin = {q: q for q in range(n)}  # qubit -> axis of its output index
    gates = qc.topological_order()
    for g in gates[:max_gates]:
        gt = gate_tensor(g.op)
        k = len(g.qubits)
        gate_in = [2 * p + 1 for p in range(k)]
        main_out = [out_pin[q] for q in g.qubits]
        tensor = np.tensordot(gt, tensor, axes=(gate_in, main_out))
        # result axes: the gate's k output pins, then the surviving main axes in order (tensor.hpp:419-430)
        new_pin = {}
        for q in range(n):
            if q in g.qubits:
                new_pin[q] = g.qubits.index(q)
            else:
                new_pin[q] = k + out_pin[q] - sum(1 for a in main_out if a < out_pin[q])
        out_pin = new_pin
    return np.ascontiguousarray(np.transpose(tensor, [out_pin[q] for q in range(n)] + [n])).reshape(dim, ncols)


_LIBSTDCXX_PRIMES = [
    13, 29, 59, 127, 257, 541, 1109, 2357, 5087, 10273, 20753, 42043, 85229, 172933,
]


def _unordered_map_iteration_order(n: int) -> list:
    """Iteration order of a libstdc++ std::unordered_map<size_t, T> after
    emplacing keys 0..n-1 in order (qcir_to_tensor.cpp:118,166-171; Appendix C trap 5).

    libstdc++ keeps one singly linked list; a node hashed into an EMPTY bucket
    is linked at the list head.  Keys 0..n-1 (identity hash) never collide
    because the table is rehashed before the load factor exceeds 1 (first
    insertion: 13 buckets, then the next prime >= 2x: 29, 59, 127, ...), and a
    rehash relinks the nodes by walking the list head-first with the same rule,
    which reverses it.  Verified against g++ 13.3 for n <= 40 by
    oracle/check_unordered_map_order.sh: n <= 13 gives n-1..0 (the goldens).
    """
    order: list = []
    nb = 1
    for k in range(n):
        if k + 1 > nb:
            nb = next(p for p in _LIBSTDCXX_PRIMES if p >= max(k + 1, 2 * nb, 12))
            order.reverse()
        order.insert(0, k)
    return order


def matrix_to_interleaved(mat: np.ndarray) -> np.ndarray:
    """QTensor::to_qtensor, qtensor.hpp:315-330."""
    n = int(round(math.log2(mat.shape[0])))
    t = mat.reshape([2] * (2 * n))
    ax = []
    for i in range(n):
        ax += [i, i + n]
    return np.transpose(t, ax)


def interleaved_to_matrix(t: np.ndarray) -> np.ndarray:
    """QTensor::to_matrix(), qtensor.hpp:442-452 (conversion_cmd.cpp:89)."""
    n = t.ndim // 2
    ax = [2 * i for i in range(n)] + [2 * i + 1 for i in range(n)]
    return np.ascontiguousarray(np.transpose(t, ax)).reshape(2**n, 2**n)


def apply_gate_rows(u: np.ndarray, n: int, m: np.ndarray, qubits: Sequence[int]) -> np.ndarray:
    """U <- (G on row bits) . U ; Appendix A.1 'equivalent restatement'.
    u is (2^n, C).  qubit 0 is the MSB of the row index."""
    k = len(qubits)
    cols = u.shape[1]
    t = u.reshape([2] * n + [cols])
    t = np.moveaxis(t, list(qubits), list(range(k)))
    shp = t.shape
    t = (m @ t.reshape(2**k, -1)).reshape(shp)
    t = np.moveaxis(t, list(range(k)), list(qubits))
    return np.ascontiguousarray(t).reshape(2**n, cols)


def to_matrix_rowform(qc: QCir, col0: int = 0, ncols: int | None = None) -> np.ndarray | None:
    """qc2ts result as the stored 2^n x 2^n matrix U[out,in] (conversion_cmd.cpp:89),
    computed in the in-place row-bit form; optionally only columns
    [col0, col0+ncols) (a column shard)."""
    n = qc.n_qubits
    if n == 0:
        return None
    dim = 2**n
    ncols = dim - col0 if ncols is None else ncols
    u = np.zeros((dim, ncols), dtype=np.complex128)
    for c in range(ncols):
        u[col0 + c, c] = 1.0
    for g in qc.topological_order():
        u = apply_gate_rows(u, n, g.op.matrix(), g.qubits)
    return u


def qc2ts(qc: QCir) -> np.ndarray | None:
    """`convert qcir tensor`: to_tensor(QCir) then to_matrix() (conversion_cmd.cpp:84-98)."""
    t = to_tensor_literal(qc)
    return None if t is None else interleaved_to_matrix(t)


def gate_stream(qc: QCir) -> list:
    """The ordered gate stream the hot path consumes:
    [(n_ctrls, qubits[controls first], target 2^t x 2^t matrix)]."""
    out = []
    for g in qc.topological_order():
        out.append((g.op.n_ctrls, list(g.qubits), g.op.target_matrix()))
    return out


# --------------------------------------------------------------------------
# Equivalence   (src/tensor/tensor.hpp:393-403, qtensor.hpp:383-417, tensor_cmd.cpp:137-173)
# --------------------------------------------------------------------------


def inner_product(a: np.ndarray, b: np.ndarray) -> float:
    if a.shape != b.shape:
        raise ValueError("The two tensors should have the same shape")
    return abs(np.sum(np.conj(a) * b))


def cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    return inner_product(a, b) / math.sqrt(inner_product(a, a) * inner_product(b, b))


def global_scalar_factor(a: np.ndarray, b: np.ndarray) -> complex:
    return complex(np.sum(b)) / complex(np.sum(a))


def is_equivalent(a: np.ndarray, b: np.ndarray, eps: float = 1e-6) -> bool:
    if a.shape != b.shape:
        return False
    return cosine_similarity(a, b) >= (1 - eps)


def equiv_sums(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """The 8 doubles the fused device reduction produces (include/qsx.h):
    Re,Im sum conj(a) b ; sum |a|^2 ; sum |b|^2 ; Re,Im sum a ; Re,Im sum b.
    Compensated (math.fsum) so it can judge a Kahan device result (Appendix C trap 4)."""
    af, bf = a.ravel(), b.ravel()
    ar, ai, br, bi = af.real, af.imag, bf.real, bf.imag
    return np.array([
        math.fsum(ar * br) + math.fsum(ai * bi),
        math.fsum(ar * bi) - math.fsum(ai * br),
        math.fsum(ar * ar) + math.fsum(ai * ai),
        math.fsum(br * br) + math.fsum(bi * bi),
        math.fsum(ar), math.fsum(ai), math.fsum(br), math.fsum(bi),
    ])


@dataclass
class EquivReport:
    equivalent: bool
    shape_mismatch: bool
    cosine: float
    norm: float
    phase: Phase | None
    text: str


def fmt_g6(x: float) -> str:
    """fmt '{:.6}' for double: general format, 6 significant digits."""
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return format(x, ".6g")


def finish_equiv(sums: Sequence[float], eps: float = 1e-6, strict: bool = False) -> EquivReport:
    """Host finish from the 8 sums, with the formulas of tensor.hpp:393-403,
    qtensor.hpp:383-417 and the strict rule / text of tensor_cmd.cpp:152-173."""
    ipab = math.hypot(sums[0], sums[1])
    cos = ipab / math.sqrt(abs(sums[2]) * abs(sums[3]))
    equiv = cos >= (1 - eps)
    sa, sb = complex(sums[4], sums[5]), complex(sums[6], sums[7])
    if sa == 0:
        # Appendix C trap 2: reference divides by zero -> NaN phase -> UB.  Defined here:
        norm, phase = math.inf if sb != 0 else math.nan, None
    else:
        s = sb / sa
        norm = abs(s)
        phase = Phase.from_float(cmath.phase(s))
    if strict and (norm > 1 + eps or norm < 1 - eps or phase != Phase(0)):
        equiv = False
    if equiv:
        text = f"Equivalent\n- Global Norm : {fmt_g6(norm)}\n- Global Phase: {phase.print_string() if phase else 'nan'}\n"
    else:
        text = f"Not Equivalent\n- Cosine Similarity: {fmt_g6(cos)}\n"
    return EquivReport(equiv, False, cos, norm, phase, text)


def tensor_equiv(a: np.ndarray, b: np.ndarray, eps: float = 1e-6, strict: bool = False) -> EquivReport:
    """`tensor equiv` (tensor_cmd.cpp:137-173).  a = tensor1, b = tensor2."""
    if a.shape != b.shape:
        text = f"Not Equivalent\n- Shape Mismatch: [{', '.join(map(str, a.shape))}] vs [{', '.join(map(str, b.shape))}]\n"
        return EquivReport(False, True, math.nan, math.nan, None, text)
    # reference definitions, evaluated directly
    cos = cosine_similarity(a, b)
    equiv = cos >= (1 - eps)
    s = global_scalar_factor(a, b)
    norm = abs(s)
    phase = Phase.from_float(cmath.phase(s))
    if strict and (norm > 1 + eps or norm < 1 - eps or phase != Phase(0)):
        equiv = False
    if equiv:
        text = f"Equivalent\n- Global Norm : {fmt_g6(norm)}\n- Global Phase: {phase.print_string()}\n"
    else:
        text = f"Not Equivalent\n- Cosine Similarity: {fmt_g6(cos)}\n"
    return EquivReport(equiv, False, cos, norm, phase, text)


# --------------------------------------------------------------------------
# Synthetic workloads (C2/C4/C5 generator, QFT generator)
# --------------------------------------------------------------------------


def random_clifford_t_qasm(n: int, n_gates: int, seed: int) -> str:
    """each gate 2-qubit with p=0.35 (cx 80% / cz 20%, random
    ordered distinct pair) else 1-qubit uniform over {h,s,sdg,t,tdg,x,z}."""
    rng = np.random.default_rng(seed)
    one = ["h", "s", "sdg", "t", "tdg", "x", "z"]
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];"]
    for _ in range(n_gates):
        if n >= 2 and rng.random() < 0.35:
            name = "cx" if rng.random() < 0.8 else "cz"
            a, b = rng.choice(n, size=2, replace=False)
            lines.append(f"{name} q[{int(a)}],q[{int(b)}];")
        else:
            name = one[int(rng.integers(len(one)))]
            q = int(rng.integers(n))
            lines.append(f"{name} q[{q}];")
    return "\n".join(lines) + "\n"


def qft_qasm(n: int) -> str:
    """benchmark/qft/generateQFTqasm.py (plain variant): h + cp(pi/2^j)."""
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];"]
    for i in range(n):
        lines.append(f"h q[{i}];")
        for j in range(i + 1, n):
            lines.append(f"cp(pi/{2 ** (j - i)}) q[{j}],q[{i}];")
    return "\n".join(lines) + "\n"


def random_rotation_qasm(n: int, n_gates: int, seed: int) -> str:
    """Dense-rotation workload (where fused dense blocks pay off): rx/ry/rz with random rational
    angles on random qubits, 25 % cx on random pairs."""
    rng = np.random.default_rng(seed)
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', f"qreg q[{n}];"]
    for _ in range(n_gates):
        if n >= 2 and rng.random() < 0.25:
            a, b = rng.choice(n, size=2, replace=False)
            lines.append(f"cx q[{int(a)}],q[{int(b)}];")
        else:
            name = ["rx", "ry", "rz"][int(rng.integers(3))]
            num, den = int(rng.integers(1, 64)), 64
            lines.append(f"{name}({num}*pi/{den}) q[{int(rng.integers(n))}];")
    return "\n".join(lines) + "\n"
