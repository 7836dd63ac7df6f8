// qcir.hpp -- the input contract of the tensor-verification path: a minimal QCir.
//
// the hot path consumes `QCir::get_gates()` (topological order),
// `QCirGate::{get_id,get_qubits,get_qubit,get_num_qubits,get_operation}` and a type-erased
// `Operation` with `get_type / get_repr / get_num_qubits` and a `to_tensor` hook.  This header
// provides exactly that surface with the reference's names and semantics
// (src/qcir/{qcir.hpp,qcir_gate.hpp,operation.hpp,basic_gate_type.hpp,gate_type.cpp,
// qcir_action.cpp,qcir_reader.cpp}); the optimiser / translation / oracle parts of qsyn's qcir
// package are out of scope.  Written from scratch: one header instead of the reference's many.
#pragma once

#include <algorithm>
#include <cstddef>
#include <filesystem>
#include <istream>
#include <format>
#include <memory>
#include <optional>
#include <set>
#include <stack>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../tensor/qtensor.hpp"
#include "../util/logger.hpp"
#include "../util/phase.hpp"

namespace qsyn {

using QubitIdType = std::size_t;
using QubitIdList = std::vector<QubitIdType>;

// per-type conversion hook, specialised in convert/qcir_to_tensor.cpp (operation.hpp:28-29)
template <typename T>
std::optional<tensor::QTensor<double>> to_tensor(T const& op);

namespace qcir {

class QCir;

// ---- type-erased gate (operation.hpp:61-188) ------------------------------------------------
class Operation {
public:
    Operation() = default;
    template <typename T>
    requires(!std::is_same_v<std::decay_t<T>, Operation>)
    Operation(T op) : _self(std::make_shared<Model<T>>(std::move(op))) {}

    std::string get_type() const { return _self->type(); }
    std::string get_repr() const { return _self->repr(); }
    size_t get_num_qubits() const { return _self->num_qubits(); }

    friend Operation adjoint(Operation const& op) { return op._self->adjoint_(); }
    friend bool is_clifford(Operation const& op) { return op._self->clifford(); }
    friend std::optional<tensor::QTensor<double>> to_tensor(Operation const& op) { return op._self->tensor_(); }

    bool operator==(Operation const& rhs) const { return get_repr() == rhs.get_repr() && get_num_qubits() == rhs.get_num_qubits(); }
    bool operator!=(Operation const& rhs) const { return !(*this == rhs); }

    template <typename T>
    bool is() const { return dynamic_cast<Model<T> const*>(_self.get()) != nullptr; }
    template <typename T>
    T get_underlying() const {
        if (auto const* m = dynamic_cast<Model<T> const*>(_self.get())) return m->value;
        spdlog::error("Operation type is {}, but expected {}", get_type(), typeid(T).name());
        throw std::bad_cast();
    }
    template <typename T>
    std::optional<T> get_underlying_if() const {
        if (auto const* m = dynamic_cast<Model<T> const*>(_self.get())) return m->value;
        return std::nullopt;
    }

private:
    struct Concept {
        virtual ~Concept()                                             = default;
        virtual std::string type() const                              = 0;
        virtual std::string repr() const                              = 0;
        virtual size_t num_qubits() const                             = 0;
        virtual Operation adjoint_() const                            = 0;
        virtual bool clifford() const                                 = 0;
        virtual std::optional<tensor::QTensor<double>> tensor_() const = 0;
    };
    template <typename T>
    struct Model final : Concept {
        explicit Model(T v) : value(std::move(v)) {}
        std::string type() const override { return value.get_type(); }
        std::string repr() const override { return value.get_repr(); }
        size_t num_qubits() const override { return value.get_num_qubits(); }
        Operation adjoint_() const override { return adjoint(value); }
        bool clifford() const override { return is_clifford(value); }
        std::optional<tensor::QTensor<double>> tensor_() const override { return qsyn::to_tensor(value); }
        T value;
    };
    std::shared_ptr<Concept const> _self;  // immutable payload, cheap copies
};

std::optional<Operation> str_to_operation(std::string str, std::vector<dvlab::Phase> const& params = {});

// ---- basic gate types (basic_gate_type.hpp:15-343) -------------------------------------------
struct IdGate {
    std::string get_type() const { return "id"; }
    std::string get_repr() const { return "id"; }
    size_t get_num_qubits() const { return 1; }
};
struct HGate {
    std::string get_type() const { return "h"; }
    std::string get_repr() const { return "h"; }
    size_t get_num_qubits() const { return 1; }
};
struct SwapGate {
    std::string get_type() const { return "swap"; }
    std::string get_repr() const { return "swap"; }
    size_t get_num_qubits() const { return 2; }
};
struct ECRGate {
    std::string get_type() const { return "ecr"; }
    std::string get_repr() const { return "ecr"; }
    size_t get_num_qubits() const { return 2; }
};

namespace detail {
// z/s/sdg/t/tdg style short names for the five fixed phases, else "<type>(<phase>)"
inline std::string phase_gate_repr(std::string const& type, char axis, dvlab::Phase const& ph) {
    static std::pair<dvlab::Phase, char const*> const fixed[] = {
        {dvlab::Phase(1), ""}, {dvlab::Phase(1, 2), "s"}, {dvlab::Phase(-1, 2), "sdg"}, {dvlab::Phase(1, 4), "t"}, {dvlab::Phase(-1, 4), "tdg"}};
    for (auto const& [p, name] : fixed) {
        if (ph != p) continue;
        std::string const n(name);
        if (n.empty()) return std::string(1, axis);                      // z / x / y
        if (axis == 'z') return n;                                       // s sdg t tdg
        return n.substr(0, 1) + std::string(1, axis) + n.substr(1);      // sx sxdg tx txdg ...
    }
    return std::format("{}({})", type, ph.get_print_string());
}
}  // namespace detail

#define QSX_PHASE_GATE(Name, type_str, axis, short_repr)                                                  \
    class Name {                                                                                          \
    public:                                                                                               \
        Name(dvlab::Phase phase) : _phase(phase) {}                                                       \
        std::string get_type() const { return type_str; }                                                 \
        std::string get_repr() const {                                                                    \
            if constexpr (short_repr) return detail::phase_gate_repr(type_str, axis, _phase);             \
            return std::format("{}({})", type_str, _phase.get_print_string());                            \
        }                                                                                                 \
        size_t get_num_qubits() const { return 1; }                                                       \
        auto get_phase() const { return _phase; }                                                         \
        void set_phase(dvlab::Phase phase) { _phase = phase; }                                            \
                                                                                                          \
    private:                                                                                              \
        dvlab::Phase _phase;                                                                              \
    };
QSX_PHASE_GATE(PZGate, "p", 'z', true)
QSX_PHASE_GATE(PXGate, "px", 'x', true)
QSX_PHASE_GATE(PYGate, "py", 'y', true)
QSX_PHASE_GATE(RZGate, "rz", 'z', false)
QSX_PHASE_GATE(RXGate, "rx", 'x', false)
QSX_PHASE_GATE(RYGate, "ry", 'y', false)
#undef QSX_PHASE_GATE

// NOLINTBEGIN(readability-identifier-naming)  pseudo classes, as in the reference
inline PZGate ZGate() { return PZGate(dvlab::Phase(1)); }
inline PZGate SGate() { return PZGate(dvlab::Phase(1, 2)); }
inline PZGate SdgGate() { return PZGate(dvlab::Phase(-1, 2)); }
inline PZGate TGate() { return PZGate(dvlab::Phase(1, 4)); }
inline PZGate TdgGate() { return PZGate(dvlab::Phase(-1, 4)); }
inline PXGate XGate() { return PXGate(dvlab::Phase(1)); }
inline PXGate SXGate() { return PXGate(dvlab::Phase(1, 2)); }
inline PXGate SXdgGate() { return PXGate(dvlab::Phase(-1, 2)); }
inline PXGate TXGate() { return PXGate(dvlab::Phase(1, 4)); }
inline PXGate TXdgGate() { return PXGate(dvlab::Phase(-1, 4)); }
inline PYGate YGate() { return PYGate(dvlab::Phase(1)); }
inline PYGate SYGate() { return PYGate(dvlab::Phase(1, 2)); }
inline PYGate SYdgGate() { return PYGate(dvlab::Phase(-1, 2)); }
inline PYGate TYGate() { return PYGate(dvlab::Phase(1, 4)); }
inline PYGate TYdgGate() { return PYGate(dvlab::Phase(-1, 4)); }
// NOLINTEND(readability-identifier-naming)

class ControlGate {
public:
    ControlGate(Operation op, size_t n_ctrls = 1) : _op(std::move(op)), _n_ctrls(n_ctrls) {
        if (n_ctrls == 0) throw std::invalid_argument("Cannot instantiate a control gate with zero controls");
    }
    std::string get_type() const { return std::string(_n_ctrls, 'c') + _op.get_type(); }
    std::string get_repr() const { return std::string(_n_ctrls, 'c') + _op.get_repr(); }
    size_t get_num_qubits() const { return _op.get_num_qubits() + _n_ctrls; }
    Operation get_target_operation() const { return _op; }
    void set_target_operation(Operation op) { _op = std::move(op); }
    size_t get_num_ctrls() const { return _n_ctrls; }

private:
    Operation _op;
    size_t _n_ctrls;
};
inline ControlGate CXGate() { return ControlGate(XGate()); }
inline ControlGate CYGate() { return ControlGate(YGate()); }
inline ControlGate CZGate() { return ControlGate(ZGate()); }
inline ControlGate CCXGate() { return ControlGate(XGate(), 2); }
inline ControlGate CCZGate() { return ControlGate(ZGate(), 2); }

inline Operation adjoint(IdGate const& op) { return op; }
inline Operation adjoint(HGate const& op) { return op; }
inline Operation adjoint(SwapGate const& op) { return op; }
inline Operation adjoint(ECRGate const& op) { return op; }
inline Operation adjoint(PZGate const& op) { return PZGate(-op.get_phase()); }
inline Operation adjoint(PXGate const& op) { return PXGate(-op.get_phase()); }
inline Operation adjoint(PYGate const& op) { return PYGate(-op.get_phase()); }
inline Operation adjoint(RZGate const& op) { return RZGate(-op.get_phase()); }
inline Operation adjoint(RXGate const& op) { return RXGate(-op.get_phase()); }
inline Operation adjoint(RYGate const& op) { return RYGate(-op.get_phase()); }
inline Operation adjoint(ControlGate const& op) { return ControlGate(adjoint(op.get_target_operation()), op.get_num_ctrls()); }

inline bool is_clifford(IdGate const&) { return true; }
inline bool is_clifford(HGate const&) { return true; }
inline bool is_clifford(SwapGate const&) { return true; }
inline bool is_clifford(ECRGate const&) { return true; }
template <typename G>
requires requires(G g) { g.get_phase(); }
inline bool is_clifford(G const& op) { return op.get_phase().denominator() <= 2; }
inline bool is_clifford(ControlGate const& op) {
    return op.get_num_ctrls() == 1 && (op.get_target_operation() == Operation(XGate()) ||
                                       op.get_target_operation() == Operation(YGate()) ||
                                       op.get_target_operation() == Operation(ZGate()));
}

// ---- gate instance + circuit container -------------------------------------------------------
class QCirGate {
public:
    QCirGate(size_t id, Operation op, QubitIdList qubits) : _id(id), _operation(std::move(op)), _qubits(std::move(qubits)) {}
    std::string get_type_str() const { return _operation.get_type(); }
    Operation const& get_operation() const { return _operation; }
    void set_operation(Operation const& op) { _operation = op; }
    size_t get_id() const { return _id; }
    QubitIdList get_qubits() const { return _qubits; }
    QubitIdType get_qubit(size_t pin_id) const { return _qubits[pin_id]; }
    size_t get_num_qubits() const { return _qubits.size(); }
    std::optional<size_t> get_pin_by_qubit(QubitIdType qubit) const {
        auto const it = std::find(_qubits.begin(), _qubits.end(), qubit);
        if (it == _qubits.end()) return std::nullopt;
        return static_cast<size_t>(it - _qubits.begin());
    }
    static bool qubit_id_is_unique(QubitIdList const& qubits) {
        if (qubits.size() > 8) return std::set<QubitIdType>(qubits.begin(), qubits.end()).size() == qubits.size();
        for (size_t i = 1; i < qubits.size(); ++i)
            for (size_t j = 0; j < i; ++j)
                if (qubits[i] == qubits[j]) return false;
        return true;
    }

private:
    size_t _id;
    Operation _operation;
    QubitIdList _qubits;
};

class QCirQubit {
public:
    QCirGate* get_first_gate() const { return _first; }
    QCirGate* get_last_gate() const { return _last; }
    void set_first_gate(QCirGate* g) { _first = g; }
    void set_last_gate(QCirGate* g) { _last = g; }

private:
    QCirGate *_first = nullptr, *_last = nullptr;
};

class QCir {
public:
    QCir() = default;
    explicit QCir(size_t n_qubits) { add_qubits(n_qubits); }
    QCir(QCir const& other) { *this = other; }
    QCir(QCir&&) noexcept            = default;
    QCir& operator=(QCir&&) noexcept = default;
    QCir& operator=(QCir const& other) {
        if (this == &other) return *this;
        reset();
        _filename   = other._filename;
        _procedures = other._procedures;
        add_qubits(other.get_num_qubits());
        for (auto const& g : other._gates_by_id) append(g->get_operation(), g->get_qubits());
        return *this;
    }

    size_t get_num_qubits() const { return _qubits.size(); }
    size_t get_num_gates() const { return _gates_by_id.size(); }
    bool is_empty() const { return _qubits.empty() || _gates_by_id.empty(); }
    std::vector<QCirQubit> const& get_qubits() const { return _qubits; }
    QCirGate* get_gate(std::optional<size_t> gid) const {
        return gid.has_value() && *gid < _gates_by_id.size() ? _gates_by_id[*gid].get() : nullptr;
    }
    std::vector<std::optional<size_t>> const& get_successors(size_t gid) const { return _successors.at(gid); }
    std::vector<std::optional<size_t>> const& get_predecessors(size_t gid) const { return _predecessors.at(gid); }

    // gates in topological order: reverse DFS post-order from the first gate of every qubit
    // (qcir.hpp:106-109, qcir_action.cpp:65-122)
    std::vector<QCirGate*> const& get_gates() const {
        _update_topological_order();
        return _gate_list;
    }

    void add_qubits(size_t num) { _qubits.resize(_qubits.size() + num); }

    // gate ids are assigned in append order; per-pin links to the previous / next gate on the same qubit
    size_t append(Operation const& op, QubitIdList const& bits) {
        if (op.get_num_qubits() != bits.size())
            throw std::invalid_argument(std::format("Operation {} requires {} qubits, but {} qubits are given.", op.get_repr(),
                                                    op.get_num_qubits(), bits.size()));
        size_t const id = _gates_by_id.size();
        _gates_by_id.push_back(std::make_unique<QCirGate>(id, op, bits));
        _predecessors.emplace_back(bits.size(), std::nullopt);
        _successors.emplace_back(bits.size(), std::nullopt);
        QCirGate* g = _gates_by_id.back().get();
        for (size_t pin = 0; pin < bits.size(); ++pin) {
            QubitIdType const qb = bits[pin];
            if (qb >= _qubits.size()) throw std::out_of_range(std::format("Qubit {} not found!!", qb));
            if (QCirGate* last = _qubits[qb].get_last_gate()) {
                _successors[last->get_id()][*last->get_pin_by_qubit(qb)] = id;
                _predecessors[id][pin]                                   = last->get_id();
            } else {
                _qubits[qb].set_first_gate(g);
            }
            _qubits[qb].set_last_gate(g);
        }
        _dirty = true;
        return id;
    }

    // Append `other` to this circuit, gate by gate in other's topological order (qcir_action.cpp:28-36)
    QCir& compose(QCir const& other) {
        if (get_num_qubits() < other.get_num_qubits()) add_qubits(other.get_num_qubits() - get_num_qubits());
        for (QCirGate* g : other.get_gates()) append(g->get_operation(), g->get_qubits());
        return *this;
    }

    // Every gate becomes its adjoint and the wiring is reversed: predecessors <-> successors, first <-> last gate of
    // every qubit; gate ids are kept (qcir_action.cpp:136-153)
    void adjoint_inplace() {
        for (auto& g : _gates_by_id) g->set_operation(adjoint(g->get_operation()));
        std::swap(_predecessors, _successors);
        for (auto& q : _qubits) {
            QCirGate* const first = q.get_first_gate();
            q.set_first_gate(q.get_last_gate());
            q.set_last_gate(first);
        }
        _dirty = true;
    }

    void reset() {
        _qubits.clear();
        _gates_by_id.clear();
        _predecessors.clear();
        _successors.clear();
        _gate_list.clear();
        _dirty = true;
    }

    // ASAP level of every gate: 1 + the latest level among its predecessors
    std::unordered_map<size_t, size_t> calculate_gate_times() const {
        std::unordered_map<size_t, size_t> level;
        for (QCirGate* g : get_gates()) {
            size_t latest = 0;
            for (auto const& pred : _predecessors[g->get_id()])
                if (pred.has_value()) latest = std::max(latest, level.at(*pred));
            level.emplace(g->get_id(), latest + 1);
        }
        return level;
    }

    // `qcir print --diagram`: one text lane per qubit, one fixed-width cell "-<name>(<id>)-" per
    // time step (same layout as the reference's qcir_print.cpp:116-165)
    std::vector<std::string> circuit_diagram_lines() const {
        auto const level = calculate_gate_times();
        size_t id_w = 0, name_w = 0, depth = 0;
        auto bare_name = [](QCirGate const& g) {
            auto r = g.get_operation().get_repr();
            return r.substr(0, r.find('('));
        };
        for (auto const& g : _gates_by_id) {
            id_w   = std::max(id_w, std::to_string(g->get_id()).size());
            name_w = std::max(name_w, bare_name(*g).size());
            depth  = std::max(depth, level.at(g->get_id()));
        }
        size_t const cell = 4 + name_w + id_w;
        std::vector<std::string> lines;
        for (size_t q = 0; q < _qubits.size(); ++q) {
            std::string line = std::format("Q{:>2}  ", q);
            size_t at        = 0;
            for (QCirGate* g = _qubits[q].get_first_gate(); g != nullptr;) {
                size_t const t = level.at(g->get_id());
                line += std::string(cell * (t - at - 1), '-');
                line += std::format("-{:>{}}({:>{}})-", bare_name(*g), name_w, g->get_id(), id_w);
                at            = t;
                auto const nx = _successors[g->get_id()][*g->get_pin_by_qubit(q)];
                g             = nx.has_value() ? _gates_by_id[*nx].get() : nullptr;
            }
            line += std::string(cell * (depth - at), '-');
            lines.push_back(std::move(line));
        }
        return lines;
    }

    void set_filename(std::string const& f) { _filename = f; }
    std::string get_filename() const { return _filename; }
    // additional APIs to make qcir::QCir an qcir::Operation (qcir.hpp:181-183): a circuit can be a gate of another one
    std::string get_type() const { return "qcir"; }
    std::string get_repr() const { return _filename; }
    void add_procedures(std::vector<std::string> const& ps) { _procedures.insert(_procedures.end(), ps.begin(), ps.end()); }
    void add_procedure(std::string const& p) { _procedures.emplace_back(p); }
    std::vector<std::string> const& get_procedures() const { return _procedures; }

private:
    void _update_topological_order() const {
        if (!_dirty) return;
        _gate_list.clear();
        std::vector<std::pair<bool, QCirGate*>> todo;  // (a stack; gate ids index _gates_by_id, so `seen` is a flat array)
        std::vector<char> seen(_gates_by_id.size(), 0);
        _gate_list.reserve(_gates_by_id.size());
        for (auto const& q : _qubits)
            if (q.get_first_gate() != nullptr) todo.emplace_back(false, q.get_first_gate());
        while (!todo.empty()) {
            auto const [expanded, node] = todo.back();
            todo.pop_back();
            if (expanded) {
                _gate_list.push_back(node);
                continue;
            }
            if (seen[node->get_id()]) continue;
            seen[node->get_id()] = 1;
            todo.emplace_back(true, node);
            for (auto const& succ : _successors[node->get_id()])
                if (succ.has_value() && !seen[*succ]) todo.emplace_back(false, _gates_by_id[*succ].get());
        }
        std::reverse(_gate_list.begin(), _gate_list.end());
        _dirty = false;
    }

    std::vector<QCirQubit> _qubits;
    std::vector<std::unique_ptr<QCirGate>> _gates_by_id;
    std::vector<std::vector<std::optional<size_t>>> _predecessors, _successors;
    mutable std::vector<QCirGate*> _gate_list;
    mutable bool _dirty = true;
    std::string _filename;
    std::vector<std::string> _procedures;
};

// a circuit used as a gate (qcir.cpp:563-567; is_clifford: every gate is)
inline Operation adjoint(QCir const& qcir) {
    auto copy = qcir;
    copy.adjoint_inplace();
    return copy;
}
inline bool is_clifford(QCir const& qcir) {
    for (QCirGate* g : qcir.get_gates())
        if (!is_clifford(g->get_operation())) return false;
    return true;
}

// QASM 2.0 reader with the reference's behaviour (qcir_reader.cpp:47-120)
std::optional<QCir> from_qasm(std::filesystem::path const& filepath);
std::optional<QCir> from_qasm_stream(std::istream& in, std::string const& name_for_messages);
std::optional<QCir> from_file(std::filesystem::path const& filepath);

}  // namespace qcir
}  // namespace qsyn
