"""CPU oracle for the ZX operand of the verification path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of what the reference does between `zx read` / `zx vertex add` / `zx edge add`
and the tensor `zx2ts` hands to `tensor equiv`:

  * ZXGraph construction        src/zx/zxgraph.cpp:339-500,625-640 (add_input/output/vertex/edge,
                                remove_edge), zxgraph_action.cpp:41-154 (lift_qubit, tensor_product,
                                adjoint), the copy constructor zxgraph.cpp:55-73
  * the `.zx` reader            src/zx/zx_io.cpp:83-436
  * topological order           src/zx/zx_traverse.cpp:22-74
  * ZX -> tensor                src/convert/zxgraph_to_tensor.cpp:95-385 (one tensordot per vertex with
                                frontier / axis-id bookkeeping), spider tensors qtensor.hpp:121-193

Containers follow the reference's ordered_hashmap / ordered_hashset (insertion order, erased
entries vanish, re-insertion goes to the end) = a python dict.  Log lines are produced exactly
as spdlog prints them so that the reference's golden logs pin the traversal and the axis
bookkeeping (tests/conversion/zx2ts/ref/*.log).  Only tests/, smoke() and bench.py's CPU legs
may import this module; the product never does.
"""
from __future__ import annotations

import math

import numpy as np

from .oracle import Phase, _expi, _nu_pow, _str_get_token, str_to_phase

SIMPLE, HADAMARD = "S", "H"
BOUNDARY, Z, X, HBOX = "B", "Z", "X", "H"
_VT_STR = {BOUNDARY: "●", Z: "Z", X: "X", HBOX: "H"}  # fmt of VertexType (zx_def.hpp)
_ET_STR = {SIMPLE: "-", HADAMARD: "H"}


class ZXVertex:
    __slots__ = ("id", "qubit", "type", "phase", "row", "col", "nbrs")

    def __init__(self, vid, qubit, vtype, phase, row, col):
        self.id, self.qubit, self.type, self.phase, self.row, self.col = vid, qubit, vtype, phase, row, col
        self.nbrs = {}  # ordered_hashset<(ZXVertex*, EdgeType)>

    def is_zx(self):
        return self.type in (Z, X)

    def is_boundary(self):
        return self.type == BOUNDARY


def _phase_add(a: Phase, b: Phase) -> Phase:
    return Phase.from_rational(a.r + b.r)


class ZXGraph:
    def __init__(self):
        self.vertices = {}  # ordered set
        self.inputs, self.outputs = {}, {}
        self.input_list, self.output_list = {}, {}  # qubit -> vertex
        self.id_to_vertex = {}
        self.next_v_id = 0
        self.filename = ""
        self.procedures = []

    # ---- construction (zxgraph.cpp) ----
    def _next_vertex_id(self):  # zxgraph.cpp:157-162 ; the caller post-increments
        while self.next_v_id in self.id_to_vertex:
            self.next_v_id += 1
        v = self.next_v_id
        self.next_v_id += 1
        return v

    def add_input(self, qubit, row=None, col=0.0, vid=None):
        vid = self._next_vertex_id() if vid is None else vid
        if vid in self.id_to_vertex or qubit in self.input_list:
            return None
        v = ZXVertex(vid, qubit, BOUNDARY, Phase(0), float(qubit) if row is None else row, col)
        self.inputs[v] = None
        self.input_list[qubit] = v
        self.vertices[v] = None
        self.id_to_vertex[vid] = v
        return v

    def add_output(self, qubit, row=None, col=0.0, vid=None):
        vid = self._next_vertex_id() if vid is None else vid
        if vid in self.id_to_vertex or qubit in self.output_list:
            return None
        v = ZXVertex(vid, qubit, BOUNDARY, Phase(0), float(qubit) if row is None else row, col)
        self.outputs[v] = None
        self.output_list[qubit] = v
        self.vertices[v] = None
        self.id_to_vertex[vid] = v
        return v

    def add_vertex(self, vtype, phase=None, row=0.0, col=0.0, vid=None):
        vid = self._next_vertex_id() if vid is None else vid
        if vid in self.id_to_vertex:
            return None
        v = ZXVertex(vid, 0, vtype, Phase(0) if phase is None else phase, row, col)
        self.vertices[v] = None
        self.id_to_vertex[vid] = v
        return v

    def is_neighbor(self, a, b, et=None):
        if et is None:
            return (b, SIMPLE) in a.nbrs or (b, HADAMARD) in a.nbrs
        return (b, et) in a.nbrs

    def get_edge_type(self, a, b):
        return SIMPLE if (b, SIMPLE) in a.nbrs else (HADAMARD if (b, HADAMARD) in a.nbrs else None)

    def add_edge(self, vs, vt, et):  # zxgraph.cpp:438-493
        if vs is vt:
            if not vs.is_zx():
                raise ValueError("Cannot add an edge between a boundary vertex and itself")
            vs.phase = _phase_add(vs.phase, Phase(1) if et == HADAMARD else Phase(0))
            return
        if vs.id > vt.id:
            vs, vt = vt, vs
        if not self.is_neighbor(vs, vt):
            vs.nbrs[(vt, et)] = None
            vt.nbrs[(vs, et)] = None
            return
        if not vs.is_zx() or not vt.is_zx():
            raise ValueError(f"Cannot add >1 between {_VT_STR[vs.type]}({vs.id}) and {_VT_STR[vt.type]}({vt.id})")
        existing = self.get_edge_type(vs, vt)
        same = vs.type == vt.type
        to_merge, to_cancel = (SIMPLE, HADAMARD) if same else (HADAMARD, SIMPLE)
        if existing == to_merge and et == to_merge:
            pass
        elif existing == to_cancel and et == to_cancel:
            self.remove_edge(vs, vt, to_cancel)
        else:
            if existing == to_cancel:
                self.remove_edge(vs, vt, to_cancel)
                vs.nbrs[(vt, to_merge)] = None
                vt.nbrs[(vs, to_merge)] = None
            vs.phase = _phase_add(vs.phase, Phase(1))

    def remove_edge(self, vs, vt, et=None):  # zxgraph.cpp:632-662
        if et is None:
            return self.remove_edge(vs, vt, SIMPLE) + self.remove_edge(vs, vt, HADAMARD)
        count = (1 if vs.nbrs.pop((vt, et), 0) is None else 0) + (1 if vt.nbrs.pop((vs, et), 0) is None else 0)
        if count == 1:
            raise IndexError("Graph connection error")
        return count // 2

    def copy(self) -> "ZXGraph":  # copy constructor, zxgraph.cpp:55-73
        g = ZXGraph()
        g.filename, g.procedures, g.next_v_id = self.filename, list(self.procedures), self.next_v_id
        old2new = {}
        for v in self.vertices:
            if v.is_boundary():
                if v in self.inputs:
                    old2new[v] = g.add_input(v.qubit, v.row, v.col, vid=v.id)
                else:
                    old2new[v] = g.add_output(v.qubit, v.row, v.col, vid=v.id)
            else:
                old2new[v] = g.add_vertex(v.type, v.phase, v.row, v.col, vid=v.id)
        for v in self.vertices:  # for_each_edge: each edge once, from its lower-id end (zxgraph.hpp:343-353)
            for (nb, et) in v.nbrs:
                if nb.id > v.id:
                    g.add_edge(old2new[v], old2new[nb], et)
        return g

    def lift_qubit(self, n):  # zxgraph_action.cpp:41-68
        for v in self.vertices:
            if v.row >= 0:
                v.row += float(n)
        for v in list(self.inputs) + list(self.outputs):
            v.qubit += n
        self.input_list = {q + n: v for q, v in self.input_list.items()}
        self.output_list = {q + n: v for q, v in self.output_list.items()}

    def tensor_product(self, target: "ZXGraph"):  # zxgraph_action.cpp:121-154
        c = target.copy()
        big, small = 2**31 - 1, -(2**31)
        ori_max, ori_min, cp_min = small, big, big
        for v in list(self.inputs) + list(self.outputs):
            ori_max, ori_min = max(ori_max, v.qubit), min(ori_min, v.qubit)
        for v in list(c.inputs) + list(c.outputs):
            cp_min = min(cp_min, v.qubit)
        c.lift_qubit((ori_max - ori_min + 1) - cp_min)
        for v in c.inputs:
            self.inputs[v] = None
        for q, v in c.input_list.items():
            self.input_list.setdefault(q, v)
        for v in c.outputs:
            self.outputs[v] = None
        for q, v in c.output_list.items():
            self.output_list.setdefault(q, v)
        for v in c.vertices:  # _move_vertices_from, zxgraph.cpp:511-521
            v.id = self.next_v_id
            self.id_to_vertex[v.id] = v
            self.next_v_id += 1
            self.vertices[v] = None
        return self

    def adjoint(self):  # zxgraph_action.cpp: swap inputs/outputs, negate phases, mirror columns
        self.inputs, self.outputs = self.outputs, self.inputs
        self.input_list, self.output_list = self.output_list, self.input_list
        for v in self.vertices:
            v.phase = -v.phase

    # ---- traversal (zx_traverse.cpp:22-74) ----
    def topological_order(self, log=None):
        order, visited = [], set()
        for root in list(self.inputs) + list(self.outputs):
            if root in visited:
                continue
            stack = [(False, root)]
            while stack:
                done, v = stack.pop()
                if done:
                    order.append(v)
                    continue
                if v in visited:
                    continue
                visited.add(v)
                stack.append((True, v))
                for (nb, _et) in v.nbrs:
                    if nb not in visited:
                        stack.append((False, nb))
        order.reverse()
        if log is not None:
            log.append("[trace]    Topological order from first input: " + " ".join(str(v.id) for v in order))
            log.append(f"[trace]    Size of topological order: {len(order)}")
        return order

    def is_io_connection_valid(self):  # zxgraph.cpp:209-216
        return all(len(v.nbrs) == 1 for v in self.inputs) and all(len(v.nbrs) == 1 for v in self.outputs)


# ---- the .zx reader (zx_io.cpp:83-436) ----
def _trim_comments(line: str) -> str:
    k = line.find("//")
    return line if k < 0 else line[:k]


def _to_float(tok: str):
    try:
        return float(tok)
    except ValueError:
        return None


def _to_int(tok: str):
    try:
        if tok.strip() != tok or tok == "":
            return None
        return int(tok, 10)
    except ValueError:
        return None


def zxgraph_from_zx_text(text: str, keep_id: bool = False):
    storage = {}
    max_in = max_out = 0
    taken_in = set()
    for line in text.splitlines():
        line = _trim_comments(line).strip(" \t\n\v\f\r")
        if not line:
            continue
        tok, pos = _str_get_token(line)
        tokens = [tok]
        lp = line.find("(", pos if pos is not None else len(line))
        rp = line.find(")", 0 if lp < 0 else lp)
        if lp < 0 and rp < 0:
            tokens += ["-", "-"]
        elif lp < 0 or rp < 0:
            return None
        else:
            comma = line.find(",", lp + 1)
            if comma < 0:
                return None
            a = line[lp + 1:comma].strip()
            b = line[comma + 1:rp].strip()
            if not a or not b:
                return None
            tokens += [a, b]
            pos = rp + 1
        while pos is not None:
            tok, pos = _str_get_token(line, pos)
            if not tok:
                break
            tokens.append(tok)
        t = tokens[0][0].upper()
        if t not in "IOZXH" or len(tokens[0]) < 2:
            return None
        vid = _to_int(tokens[0][1:])
        if vid is None or vid < 0 or vid in storage:
            return None
        info = {"type": t, "qubit": 0, "row": 0.0, "col": 0.0, "nbrs": [], "phase": Phase(1) if t == "H" else Phase(0)}
        if t in "IO":
            q = _to_int(tokens[-1])
            if q is not None and len(tokens) > 3:
                tokens.pop()
                info["qubit"] = q
                if t == "I":
                    max_in = max(max_in, q)
                else:
                    max_out = max(max_out, q)
            elif t == "I":
                info["qubit"] = max_in
                max_in += 1
            else:
                info["qubit"] = max_out
                max_out += 1
            if t == "I":
                if info["qubit"] in taken_in:
                    return None
                taken_in.add(info["qubit"])
            info["row"] = float(info["qubit"])
        else:
            ph = str_to_phase(tokens[-1]) if len(tokens) > 3 else None
            if ph is not None:
                tokens.pop()
                info["phase"] = ph
        if tokens[1] != "-":
            r = _to_float(tokens[1])
            if r is None:
                return None
            info["row"] = r
        if tokens[2] != "-":
            c = _to_float(tokens[2])
            if c is None:
                return None
            info["col"] = c
        for nt in tokens[3:]:
            et = nt[0].upper()
            nid = _to_int(nt[1:])
            if et not in "SH" or nid is None or nid < 0:
                return None
            info["nbrs"].append((et, nid))
        storage[vid] = info
    g = ZXGraph()
    id2v = {}
    for vid, info in storage.items():
        t = info["type"]
        if t == "I":
            v = g.add_input(info["qubit"], info["row"], info["col"])
        elif t == "O":
            v = g.add_output(info["qubit"], info["row"], info["col"])
        else:
            v = g.add_vertex({"Z": Z, "X": X, "H": HBOX}[t], info["phase"], info["row"], info["col"])
        if keep_id:
            v.id = vid
        id2v[vid] = v
    for vid, info in storage.items():
        for et, nid in info["nbrs"]:
            if nid not in id2v:
                return None
            if g.is_neighbor(id2v[vid], id2v[nid], et):
                continue
            g.add_edge(id2v[vid], id2v[nid], et)
    return g


# ---- spider tensors (qtensor.hpp:121-193) ----
def t_identity(n_qubits: int) -> np.ndarray:
    t = np.eye(2**n_qubits, dtype=np.complex128).reshape([2] * (2 * n_qubits))
    ax = []
    for i in range(n_qubits):
        ax += [i, i + n_qubits]
    return np.transpose(t, ax)


def t_zspider(arity: int, phase: Phase) -> np.ndarray:
    t = np.zeros([2] * arity, dtype=np.complex128)
    if arity == 0:
        t[()] = 1.0 + _expi(phase.to_float())
    else:
        t[(0,) * arity] = 1.0
        t[(1,) * arity] = _expi(phase.to_float())
    return t * _nu_pow(2 - arity)


def t_xspider(arity: int, phase: Phase) -> np.ndarray:
    t = np.ones([2] * arity, dtype=np.complex128)
    ket_minus = np.array([1.0 + 0j, -1.0 + 0j])
    tmp = np.array(1.0 + 0j)
    for _ in range(arity):
        tmp = np.tensordot(tmp, ket_minus, axes=0)
    t = t + tmp * _expi(phase.to_float())
    t = t / math.pow(math.sqrt(2), arity)
    return t * _nu_pow(2 - arity)


def t_hbox(arity: int, a: complex = -1.0) -> np.ndarray:
    t = np.ones([2] * arity, dtype=np.complex128)
    if arity == 0:
        t[()] = a
    else:
        t[(1,) * arity] = a
    return t * _nu_pow(arity)


def _tensor_form(g: ZXGraph, v: ZXVertex) -> np.ndarray:  # zxgraph_to_tensor.cpp:139-153
    k = len(v.nbrs)
    if v.type == Z:
        return t_zspider(k, v.phase)
    if v.type == X:
        return t_xspider(k, v.phase)
    if v.type == HBOX:
        return t_hbox(k)
    return t_identity(k)


def _tensordot(a, b, ax1, ax2):
    """tensordot + the axis history of the result (tensor.hpp:405-431): old id (b's shifted by
    a.ndim) -> new id; contracted axes are absent."""
    r = np.tensordot(a, b, axes=(list(ax1), list(ax2)))
    hist, nid = {}, 0
    for i in range(a.ndim):
        if i not in ax1:
            hist[i] = nid
            nid += 1
    for i in range(b.ndim):
        if i not in ax2:
            hist[i + a.ndim] = nid
            nid += 1
    return r, hist


_SIZE_MAX = 2**64 - 1


def _edge_key(v, nb, et):  # make_edge_pair: lower id first
    return ((v, nb), et) if v.id < nb.id else ((nb, v), et)


def zx2ts(g: ZXGraph, log: list | None = None):
    """ZX2TSMapper::map (zxgraph_to_tensor.cpp:95-135).  Returns the 2^n_out x 2^n_in matrix
    (or None); `log` receives the [error]/[debug]/[trace] lines."""
    def emit(s):
        if log is not None:
            log.append(s)

    if not g.vertices:
        emit("[error]    The ZXGraph is empty!!")
        return None
    if not g.is_io_connection_valid():
        emit("[error]    The ZXGraph is not valid!!")
        return None
    tensors = []  # [frontiers dict, tensor]
    boundary_edges = []
    pins = {}  # vertex -> tensor id

    for v in g.topological_order(log):
        # _get_tensor_id
        tid = len(pins)
        for (nb, _et) in v.nbrs:
            if nb in pins:
                tid = pins[nb]
                break
        if tid == len(pins):  # new subgraph (_initialize_subgraph)
            emit(f"[debug]    Mapping vertex {v.id:>4} ({_VT_STR[v.type]}): New Subgraph")
            (nb, et) = next(iter(v.nbrs))
            tid = len(tensors)
            t = np.tensordot(np.array(1.0 + 0j), t_identity(len(v.nbrs)), axes=0)
            key = _edge_key(v, nb, et)
            boundary_edges.append(key)
            tensors.append([{key: 1}, t])
        else:
            fr, ts = tensors[tid]
            simple_pins, had_pins, to_remove, to_add = [], [], [], []
            for (nb, et) in v.nbrs:
                key = _edge_key(v, nb, et)
                if nb not in pins:
                    to_add.append(key)
                else:
                    axid = fr[key]
                    (had_pins if key[1] == HADAMARD else simple_pins).append(axid)
                    to_remove.append(key)
            emit(f"[debug]    Mapping vertex {v.id:>4} ({_VT_STR[v.type]}): " + ("Boundary" if v.is_boundary() else "Tensordot"))
            # _dehadamardize
            hp = np.array(1.0 + 0j)
            for _ in had_pins:
                hp = np.tensordot(hp, t_hbox(2), axes=0)
            connect = [2 * i for i in range(len(had_pins))]
            tmp, hist = _tensordot(ts, hp, had_pins, connect)
            for key in fr:
                axid = fr[key]
                if axid in had_pins:
                    fr[key] = hist.get(ts.ndim + connect[had_pins.index(axid)] + 1, _SIZE_MAX)
                else:
                    fr[key] = hist.get(axid, _SIZE_MAX)
            had_new = [hist.get(ts.ndim + c + 1, _SIZE_MAX) for c in connect]
            simple_new = [hist.get(p, _SIZE_MAX) for p in simple_pins]
            all_pins = had_new + simple_new
            if v.is_boundary():
                tensors[tid][1] = tmp
            else:
                vt = _tensor_form(g, v)
                vpins = list(range(len(all_pins)))
                res, hist2 = _tensordot(tmp, vt, all_pins, vpins)
                for key in to_remove:
                    fr.pop(key, None)
                for key in fr:
                    fr[key] = hist2.get(fr[key], _SIZE_MAX)
                for t_i, key in enumerate(to_add):
                    if key not in fr:  # emplace
                        fr[key] = hist2.get(tmp.ndim + len(vpins) + t_i, _SIZE_MAX)
                tensors[tid][1] = res
        pins.setdefault(v, tid)
        fr, ts = tensors[tid]
        emit(f"[debug]    Done. Current tensor dimension: {ts.ndim}")
        emit("[trace]    Current frontiers:")
        for ((a, b), et), axid in fr.items():
            emit(f"[trace]      {a.id}--{b.id} ({_ET_STR[et]}) axis id: {axid}")

    result = np.array(1.0 + 0j)
    for _fr, ts in tensors:
        result = np.tensordot(result, ts, axes=0)
    for i, (fr, _ts) in enumerate(tensors):
        fr.setdefault(boundary_edges[i], 0)

    # _get_axis_orders (zxgraph_to_tensor.cpp:224-270)
    def table(vs):
        ordered = sorted(vs, key=lambda v: v.qubit)  # list::sort is stable
        return {v.qubit: i for i, v in enumerate(ordered)}

    in_tab, out_tab = table(list(g.inputs)), table(list(g.outputs))
    in_ids, out_ids = [0] * len(g.inputs), [0] * len(g.outputs)
    acc = 0
    for fr, _ts in tensors:
        b2b = False
        for ((v1, v2), _et), axid in fr.items():
            v1i, v2i, v1o, v2o = v1 in g.inputs, v2 in g.inputs, v1 in g.outputs, v2 in g.outputs
            if v1i:
                in_ids[in_tab[v1.qubit]] = axid + acc
            if v2i:
                in_ids[in_tab[v2.qubit]] = axid + acc
            if v1o:
                out_ids[out_tab[v1.qubit]] = axid + acc
            if v2o:
                out_ids[out_tab[v2.qubit]] = axid + acc
            if v1i and (v2i or v2o):
                in_ids[in_tab[v1.qubit]] -= 1
                b2b = True
            if v1o and (v2i or v2o):
                out_ids[out_tab[v1.qubit]] -= 1
                b2b = True
        acc += len(fr) + (1 if b2b else 0)
    emit("[trace]    Input  Axis IDs: " + " ".join(str(i) for i in in_ids))
    emit("[trace]    Output Axis IDs: " + " ".join(str(i) for i in out_ids))
    mat = np.transpose(result, out_ids + in_ids).reshape(2 ** len(out_ids), 2 ** len(in_ids))  # to_matrix
    return mat
