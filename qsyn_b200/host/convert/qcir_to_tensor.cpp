// qcir_to_tensor.cpp -- QCir -> QTensor<double>, B200 path.
//
// Drop-in for the reference's src/convert/qcir_to_tensor.cpp.  Same entry points
// (to_tensor(QCirGate), to_tensor(QCir), the per-operation specialisations), same failure
// behaviour (std::nullopt + log line for an empty circuit, an unsupported gate, an interrupt
// or an allocation failure) and the same debug/trace log text -- but the main tensor is a
// device-resident unitary and the per-gate `tensordot(gate, main)` of the reference
// (qcir_to_tensor.cpp:184-193) becomes ONE gate stream handed to the C ABI
// (qsx_apply_gates), which schedules it into in-place multi-gate sweeps.
//
// The reference's axis bookkeeping (Qubit2TensorPinMap / update_tensor_pin,
// qcir_to_tensor.cpp:24,116-136) no longer drives any data movement -- gates act on row bits of
// U[out,in] in place -- but its [trace] lines are part of the golden logs
// (tests/conversion/qc2ts/ref/qc2ts.log), so the integer bookkeeping is replayed on the side,
// with the same container type so that the iteration order matches too.
#include "./qcir_to_tensor.hpp"

#include <complex>
#include <memory>
#include <unordered_map>

extern bool stop_requested();

namespace qsyn {

using Qubit2TensorPinMap = std::unordered_map<QubitIdType, std::pair<size_t, size_t>>;
using qsyn::tensor::QTensor;

// ---- per-operation tensors: the reference formulas (qtensor.hpp builders) ---------------------
template <>
std::optional<QTensor<double>> to_tensor(HGate const& /* op */) { return QTensor<double>::hbox(2); }
template <>
std::optional<QTensor<double>> to_tensor(IdGate const& /* op */) { return QTensor<double>::identity(1); }
template <>
std::optional<QTensor<double>> to_tensor(SwapGate const& /* op */) {
    QTensor<double> const m{{1.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 0.0}, {0.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 1.0}};
    return m.to_qtensor();
}
template <>
std::optional<QTensor<double>> to_tensor(ECRGate const& /* op */) {
    using C = std::complex<double>;
    C const a = 1.0 / std::sqrt(2), b = C(0, 1) / std::sqrt(2), z = 0.0;
    QTensor<double> const m{{z, z, a, b}, {z, z, b, a}, {a, -b, z, z}, {-b, a, z, z}};
    return m.to_qtensor();
}
template <>
std::optional<QTensor<double>> to_tensor(PZGate const& op) { return QTensor<double>::pzgate(op.get_phase()); }
template <>
std::optional<QTensor<double>> to_tensor(PXGate const& op) { return QTensor<double>::pxgate(op.get_phase()); }
template <>
std::optional<QTensor<double>> to_tensor(PYGate const& op) { return QTensor<double>::pygate(op.get_phase()); }
template <>
std::optional<QTensor<double>> to_tensor(RZGate const& op) { return QTensor<double>::rzgate(op.get_phase()); }
template <>
std::optional<QTensor<double>> to_tensor(RXGate const& op) { return QTensor<double>::rxgate(op.get_phase()); }
template <>
std::optional<QTensor<double>> to_tensor(RYGate const& op) { return QTensor<double>::rygate(op.get_phase()); }
template <>
std::optional<QTensor<double>> to_tensor(ControlGate const& op) {
    if (auto target = to_tensor(op.get_target_operation())) {
        return QTensor<double>::control(*target, op.get_num_qubits() - op.get_target_operation().get_num_qubits());
    }
    return std::nullopt;
}

std::optional<QTensor<double>> to_tensor(QCirGate const& gate) { return to_tensor(gate.get_operation()); }

namespace {

// One entry of the gate stream: the gate's controls are peeled off so that only the 2^t x 2^t
// TARGET matrix crosses the ABI -- a multi-controlled gate must not be materialised as a dense
// 2^k x 2^k block the way QTensor::control does (qtensor.hpp:363-371).
struct StreamGate {
    size_t n_ctrls = 0;
    std::shared_ptr<std::vector<double> const> matrix;  // row-major complex, first target = MSB; shared by every gate of the same operation
    size_t n_targets = 0;
};

std::optional<StreamGate> lower_to_stream(Operation const& op) {
    StreamGate g;
    Operation target = op;
    while (target.is<ControlGate>()) {
        auto const cg = target.get_underlying<ControlGate>();
        g.n_ctrls += cg.get_num_ctrls();
        target = cg.get_target_operation();
    }
    if (target.is<QCir>()) return std::nullopt;  // (flattened by append_flat, never lowered as one matrix)
    auto t = to_tensor(target);
    if (!t.has_value()) return std::nullopt;
    g.n_targets  = target.get_num_qubits();
    auto const m = t->to_matrix().host_elements();  // rank-2k [o0,i0,..] -> 2^k x 2^k
    std::vector<double> flat;
    flat.reserve(2 * m.size());
    for (auto const& z : m) {
        flat.push_back(z.real());
        flat.push_back(z.imag());
    }
    g.matrix = std::make_shared<std::vector<double> const>(std::move(flat));
    return g;
}

// The reference's pin bookkeeping after tensordot(gate, main): result axes are the gate's
// output pins 0..k-1 followed by the surviving main axes in order (tensor.hpp:419-430).
void update_tensor_pin(Qubit2TensorPinMap& qubit2pin, QCirGate const& gate, size_t n_main_axes) {
    spdlog::trace("Pin Permutation");
    auto const operands = gate.get_qubits();
    size_t const k      = operands.size();
    std::vector<size_t> contracted;
    for (auto q : operands) contracted.push_back(qubit2pin[q].first);
    auto shifted = [&](size_t x) {
        size_t below = 0;
        for (auto c : contracted) below += c < x ? 1 : 0;
        return k + x - below;
    };
    (void)n_main_axes;
    for (auto& [qubit, pin] : qubit2pin) {
        auto const [old_out, old_in] = pin;
        auto const it                = std::find(operands.begin(), operands.end(), qubit);
        size_t const new_out         = it != operands.end() ? static_cast<size_t>(it - operands.begin()) : shifted(old_out);
        size_t const new_in          = shifted(old_in);
        pin                          = {new_out, new_in};
        spdlog::trace("  - Qubit: {} input : {} -> {} output: {} -> {}", qubit, old_in, new_in, old_out, new_out);
    }
}

// A gate of the stream with its qubits (controls first).  A (controlled) nested circuit -- the reference contracts
// to_tensor(inner QCir) as ONE 2^k x 2^k gate tensor (qcir.hpp:181-183, qcir_to_tensor.cpp:147 through operation.hpp) --
// is FLATTENED into its own gates, each carrying the outer controls: C(AB) = C(A) C(B), same operator, and nothing wider
// than the widest basic gate ever crosses the ABI.
struct FlatGate {
    StreamGate g;  // (the matrix is shared with the per-operation cache below: no copy per gate)
    QubitIdList qubits;
};

bool append_flat(Operation const& op, QubitIdList const& qubits, std::vector<FlatGate>& out) {
    size_t n_ctrls   = 0;
    Operation target = op;
    while (target.is<ControlGate>()) {
        auto const cg = target.get_underlying<ControlGate>();
        n_ctrls += cg.get_num_ctrls();
        target = cg.get_target_operation();
    }
    if (target.is<QCir>()) {
        auto const inner = target.get_underlying<QCir>();
        if (inner.get_num_qubits() + n_ctrls != qubits.size()) return false;
        for (QCirGate* ig : inner.get_gates()) {
            QubitIdList mapped(qubits.begin(), qubits.begin() + static_cast<std::ptrdiff_t>(n_ctrls));
            for (auto q : ig->get_qubits()) mapped.push_back(qubits[n_ctrls + q]);
            Operation const wrapped = n_ctrls > 0 ? Operation(ControlGate(ig->get_operation(), n_ctrls)) : ig->get_operation();
            if (!append_flat(wrapped, mapped, out)) return false;
        }
        return true;
    }
    // A circuit has thousands of gates but a handful of distinct operations (h, t, cx, rz(pi/8), ...): the matrix of an
    // operation is built once through the QTensor builders and reused.  get_repr() identifies a basic operation completely
    // (type, phase as an exact rational, number of controls); nested circuits never get here.
    static thread_local std::unordered_map<std::string, StreamGate> cache;
    std::string const key = op.get_repr();
    auto it               = cache.find(key);
    if (it == cache.end()) {
        auto g = lower_to_stream(op);
        if (!g.has_value()) return false;
        if (cache.size() >= 4096) cache.clear();  // (matrices still referenced by a gate stream stay alive through it)
        it = cache.emplace(key, std::move(*g)).first;
    }
    if (qubits.size() > QSX_MAX_GATE_QUBITS || it->second.n_ctrls + it->second.n_targets != qubits.size()) return false;
    out.push_back(FlatGate{it->second, qubits});
    return true;
}

int stop_trampoline(void*) { return stop_requested() ? 1 : 0; }

}  // namespace

std::optional<GateStream> to_gate_stream(QCir const& qcir) {
    GateStream gs;
    std::vector<FlatGate> flat;
    flat.reserve(qcir.get_num_gates());
    for (auto const& gate : qcir.get_gates()) {
        if (!append_flat(gate->get_operation(), gate->get_qubits(), flat)) {
            spdlog::error("Conversion of Gate {} ({}) to Tensor is not supported yet!!", gate->get_id(), gate->get_operation().get_repr());
            return std::nullopt;
        }
    }
    gs.gates.reserve(flat.size());
    gs.matrices.reserve(flat.size());
    for (auto& f : flat) {
        qsx_gate d{};
        d.n_ctrls   = static_cast<int32_t>(f.g.n_ctrls);
        d.n_targets = static_cast<int32_t>(f.g.n_targets);
        for (size_t p = 0; p < f.qubits.size(); ++p) d.qubits[p] = static_cast<int32_t>(f.qubits[p]);
        d.matrix = f.g.matrix->data();  // shared storage the stream keeps alive
        gs.matrices.push_back(std::move(f.g.matrix));
        gs.gates.push_back(d);
    }
    return gs;
}

template <>
std::optional<QTensor<double>> to_tensor(QCir const& qcir) try {
    if (qcir.get_num_qubits() == 0) {
        spdlog::warn("QCir is empty!!");
        return std::nullopt;
    }
    spdlog::debug("Add boundary");
    size_t const n = qcir.get_num_qubits();
    if (stop_requested()) {
        spdlog::warn("Conversion interrupted.");
        return std::nullopt;
    }
    // identity on n qubits: one write pass on the device
    auto unitary = tensor::DeviceUnitary::identity(static_cast<int>(n));

    Qubit2TensorPinMap qubit_to_pins;  // qubit -> (output, input)
    for (size_t q = 0; q < n; ++q) {
        qubit_to_pins.emplace(q, std::make_pair(2 * q, 2 * q + 1));
        spdlog::trace("  - Add Qubit: {} input: {} output: {}", q, 2 * q + 1, 2 * q);
    }

    bool const tracing   = spdlog::get_level() <= spdlog::level::trace;
    bool const debugging = spdlog::get_level() <= spdlog::level::debug;
    std::vector<FlatGate> lowered;
    std::vector<qsx_gate> stream;
    lowered.reserve(qcir.get_num_gates());
    for (auto const& gate : qcir.get_gates()) {
        if (stop_requested()) {
            spdlog::warn("Conversion interrupted.");
            return std::nullopt;
        }
        if (debugging) spdlog::debug("Gate {} ({})", gate->get_id(), gate->get_operation().get_repr());
        if (!append_flat(gate->get_operation(), gate->get_qubits(), lowered)) {
            spdlog::error("Conversion of Gate {} ({}) to Tensor is not supported yet!!", gate->get_id(), gate->get_operation().get_repr());
            return std::nullopt;
        }
        if (tracing) update_tensor_pin(qubit_to_pins, *gate, 2 * n);
    }
    stream.reserve(lowered.size());
    for (auto const& f : lowered) {
        qsx_gate d{};
        d.n_ctrls   = static_cast<int32_t>(f.g.n_ctrls);
        d.n_targets = static_cast<int32_t>(f.g.n_targets);
        for (size_t p = 0; p < f.qubits.size(); ++p) d.qubits[p] = static_cast<int32_t>(f.qubits[p]);
        d.matrix = f.g.matrix->data();
        stream.push_back(d);
    }

    int const rc = unitary.apply(stream, &stop_trampoline, nullptr);
    if (rc == QSX_ERR_INTERRUPTED || stop_requested()) {
        spdlog::warn("Conversion interrupted.");
        return std::nullopt;
    }
    if (rc == QSX_ERR_UNSUPPORTED) {
        spdlog::error("Conversion of a gate to Tensor is not supported yet!! ({})", qsx_last_error());
        return std::nullopt;
    }
    if (rc != QSX_OK) {
        spdlog::error("Conversion failed: {}", qsx_last_error());
        return std::nullopt;
    }
    // rank-2n tensor with axes [o0,i0,o1,i1,...], like the reference after to_matrix + to_qtensor
    return QTensor<double>(std::move(unitary), /*as_matrix=*/false);
} catch (std::bad_alloc const& e) {
    spdlog::error("Memory allocation failed!!");
    return std::nullopt;
}

}  // namespace qsyn
