// qsx_host.cpp -- C ABI over the C++ host layer (include/qsx_host.h).
#include "../../../include/qsx_host.h"

#include <cstring>
#include <format>
#include <filesystem>
#include <sstream>
#include <string>
#include <utility>

#include "../convert/qcir_to_tensor.hpp"
#include "../convert/zxgraph_to_tensor.hpp"
#include "../qcir/qcir.hpp"
#include "../qcir/qcir_equiv.hpp"
#include "../tensor/qtensor.hpp"

// the embedding application owns cancellation (qsyn: SIGINT -> jthread stop token); a library user
// that has none links this weak default
__attribute__((weak)) bool stop_requested() { return false; }

struct qsxh_circuit {
    qsyn::qcir::QCir qcir;
    std::optional<qsyn::GateStream> stream;  // rebuilt lazily after a modification
};
struct qsxh_tensor {
    qsyn::tensor::QTensor<double> tensor;
};
struct qsxh_zxgraph {
    qsyn::zx::ZXGraph graph;
};

namespace {
thread_local std::string g_err;
int fail(int code, std::string msg) {
    g_err = std::move(msg);
    return code;
}
template <typename F>
int guarded(F f) {
    try {
        return f();
    } catch (std::bad_alloc const&) {
        return fail(QSX_ERR_OOM, "Memory allocation failed!!");
    } catch (std::invalid_argument const& e) {
        return fail(QSX_ERR_SHAPE, e.what());
    } catch (std::exception const& e) {
        return fail(QSX_ERR_CUDA, e.what());
    }
}
}  // namespace

extern "C" {

const char* qsxh_last_error(void) { return g_err.c_str(); }

int qsxh_set_process_shard(int rank, int world, int device) {
    if (world < 1 || rank < 0 || rank >= world || device < 0) return fail(QSX_ERR_BAD_ARG, "bad rank / world / device");
    qsyn::tensor::process_shard() = qsyn::tensor::ProcessShard{rank, world, device};
    return QSX_OK;
}

int qsxh_circuit_new(size_t n_qubits, qsxh_circuit** out) {
    if (out == nullptr) return fail(QSX_ERR_BAD_ARG, "null out-pointer");
    *out = new qsxh_circuit{qsyn::qcir::QCir(n_qubits), std::nullopt};
    return QSX_OK;
}

int qsxh_circuit_from_qasm_text(const char* text, qsxh_circuit** out) {
    if (text == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        std::istringstream in{std::string(text)};
        auto qc = qsyn::qcir::from_qasm_stream(in, "<text>");
        if (!qc.has_value()) return fail(QSX_ERR_BAD_ARG, "the QASM text has something wrong!!");
        *out = new qsxh_circuit{std::move(*qc), std::nullopt};
        return (int)QSX_OK;
    });
}

int qsxh_circuit_from_qasm_file(const char* path, qsxh_circuit** out) {
    if (path == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        auto qc = qsyn::qcir::from_qasm(path);
        if (!qc.has_value()) return fail(QSX_ERR_BAD_ARG, std::format("the format in \"{}\" has something wrong!!", path));
        *out = new qsxh_circuit{std::move(*qc), std::nullopt};
        return (int)QSX_OK;
    });
}

int qsxh_circuit_append(qsxh_circuit* c, const char* gate, const char* phase, const size_t* qubits, size_t n_qubits) {
    if (c == nullptr || gate == nullptr || (n_qubits > 0 && qubits == nullptr)) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        std::vector<dvlab::Phase> params;
        if (phase != nullptr) {
            auto const ph = dvlab::Phase::from_string(phase);
            if (!ph.has_value()) return fail(QSX_ERR_BAD_ARG, std::format("invalid phase \"{}\"!!", phase));
            params.push_back(*ph);
        }
        auto const op = qsyn::qcir::str_to_operation(gate, params);
        if (!op.has_value()) return fail(QSX_ERR_BAD_ARG, std::format("Invalid gate type {}!!", gate));
        qsyn::QubitIdList ids(qubits, qubits + n_qubits);
        if (op->get_num_qubits() != ids.size() || !qsyn::qcir::QCirGate::qubit_id_is_unique(ids))
            return fail(QSX_ERR_BAD_ARG, std::format("Gate {} requires {} distinct qubits", gate, op->get_num_qubits()));
        for (auto q : ids)
            if (q >= c->qcir.get_num_qubits()) return fail(QSX_ERR_BAD_ARG, std::format("Qubit ID {} does not exist!!", q));
        c->qcir.append(*op, ids);
        c->stream.reset();
        return (int)QSX_OK;
    });
}

int qsxh_circuit_append_circuit(qsxh_circuit* c, const qsxh_circuit* sub, size_t n_ctrls, const size_t* qubits, size_t n_qubits) {
    if (c == nullptr || sub == nullptr || (n_qubits > 0 && qubits == nullptr)) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        if (sub->qcir.get_num_qubits() + n_ctrls != n_qubits) return fail(QSX_ERR_BAD_ARG, "the sub-circuit's qubits plus the controls must match the qubit list");
        qsyn::QubitIdList ids(qubits, qubits + n_qubits);
        if (!qsyn::qcir::QCirGate::qubit_id_is_unique(ids)) return fail(QSX_ERR_BAD_ARG, "duplicate qubit ids");
        for (auto q : ids)
            if (q >= c->qcir.get_num_qubits()) return fail(QSX_ERR_BAD_ARG, std::format("Qubit ID {} does not exist!!", q));
        qsyn::qcir::Operation op = sub->qcir;
        if (n_ctrls > 0) op = qsyn::qcir::ControlGate(op, n_ctrls);
        c->qcir.append(op, ids);
        c->stream.reset();
        return (int)QSX_OK;
    });
}

int qsxh_circuit_info(const qsxh_circuit* c, size_t* n_qubits, size_t* n_gates) {
    if (c == nullptr) return fail(QSX_ERR_BAD_ARG, "null circuit");
    if (n_qubits != nullptr) *n_qubits = c->qcir.get_num_qubits();
    if (n_gates != nullptr) *n_gates = c->qcir.get_num_gates();
    return QSX_OK;
}

int qsxh_circuit_gate_repr(qsxh_circuit* c, size_t i, char* buf, size_t buf_size, size_t* gate_id) {
    if (c == nullptr) return fail(QSX_ERR_BAD_ARG, "null circuit");
    auto const& gates = c->qcir.get_gates();
    if (i >= gates.size()) return fail(QSX_ERR_BAD_ARG, "gate index out of range");
    std::string const r = gates[i]->get_operation().get_repr();
    if (buf != nullptr && buf_size > 0) {
        size_t const k = std::min(buf_size - 1, r.size());
        std::memcpy(buf, r.data(), k);
        buf[k] = '\0';
    }
    if (gate_id != nullptr) *gate_id = gates[i]->get_id();
    return QSX_OK;
}

int qsxh_circuit_gate_stream(qsxh_circuit* c, const qsx_gate** gates, size_t* n_gates) {
    if (c == nullptr || gates == nullptr || n_gates == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        if (!c->stream.has_value()) {
            c->stream = qsyn::to_gate_stream(c->qcir);
            if (!c->stream.has_value()) return fail(QSX_ERR_UNSUPPORTED, "Conversion of a gate to Tensor is not supported yet!!");
        }
        *gates   = c->stream->gates.data();
        *n_gates = c->stream->gates.size();
        return (int)QSX_OK;
    });
}

int qsxh_circuit_adjoint_inplace(qsxh_circuit* c) {
    if (c == nullptr) return fail(QSX_ERR_BAD_ARG, "null circuit");
    c->qcir.adjoint_inplace();
    c->stream.reset();
    return QSX_OK;
}

int qsxh_circuit_equiv(const qsxh_circuit* c1, const qsxh_circuit* c2, int* equivalent) {
    if (c1 == nullptr || c2 == nullptr || equivalent == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        *equivalent = qsyn::qcir::is_equivalent(c1->qcir, c2->qcir) ? 1 : 0;
        return (int)QSX_OK;
    });
}

int qsxh_circuit_destroy(qsxh_circuit* c) {
    delete c;
    return QSX_OK;
}

int qsxh_zxgraph_from_zx_text(const char* text, int keep_id, qsxh_zxgraph** out) {
    if (text == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        std::istringstream in{std::string(text)};
        auto g = qsyn::zx::from_zx(in, keep_id != 0);
        if (!g.has_value()) return fail(QSX_ERR_BAD_ARG, "the format of the ZX text has something wrong (see the log)");
        *out = new qsxh_zxgraph{std::move(*g)};
        return (int)QSX_OK;
    });
}

int qsxh_zxgraph_from_zx_file(const char* path, int keep_id, qsxh_zxgraph** out) {
    if (path == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        auto g = qsyn::zx::from_zx(std::filesystem::path(path), keep_id != 0);
        if (!g.has_value()) return fail(QSX_ERR_BAD_ARG, std::string("cannot read a ZXGraph from \"") + path + "\"");
        g->set_filename(std::filesystem::path(path).stem().string());
        *out = new qsxh_zxgraph{std::move(*g)};
        return (int)QSX_OK;
    });
}

int qsxh_zxgraph_info(const qsxh_zxgraph* g, size_t* n_inputs, size_t* n_outputs, size_t* n_vertices, size_t* n_edges) {
    if (g == nullptr) return fail(QSX_ERR_BAD_ARG, "null graph");
    if (n_inputs != nullptr) *n_inputs = g->graph.num_inputs();
    if (n_outputs != nullptr) *n_outputs = g->graph.num_outputs();
    if (n_vertices != nullptr) *n_vertices = g->graph.num_vertices();
    if (n_edges != nullptr) *n_edges = g->graph.num_edges();
    return QSX_OK;
}

int qsxh_zx2ts(const qsxh_zxgraph* g, qsxh_tensor** out) {
    if (g == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        auto t = qsyn::to_tensor(g->graph);  // conversion_cmd.cpp:226
        if (!t.has_value()) return fail(QSX_ERR_UNSUPPORTED, "conversion failed (see the log)");
        t->set_filename(g->graph.get_filename());
        t->add_procedures(g->graph.get_procedures());
        t->add_procedure("ZX2TS");
        *out = new qsxh_tensor{std::move(*t)};
        return (int)QSX_OK;
    });
}

int qsxh_zxgraph_destroy(qsxh_zxgraph* g) {
    delete g;
    return QSX_OK;
}

int qsxh_qc2ts(const qsxh_circuit* c, qsxh_tensor** out) {
    if (c == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        if (c->qcir.get_num_qubits() == 0) return fail(QSX_ERR_BAD_ARG, "QCir is empty!!");
        auto t = qsyn::to_tensor(c->qcir);  // conversion_cmd.cpp:86
        if (!t.has_value()) return fail(QSX_ERR_UNSUPPORTED, "conversion failed (see the log)");
        *t   = std::move(*t).to_matrix();   // conversion_cmd.cpp:89 (the rvalue form re-tags the device tensor, no clone)
        t->set_filename(c->qcir.get_filename());
        t->add_procedures(c->qcir.get_procedures());
        t->add_procedure("QC2TS");
        *out = new qsxh_tensor{std::move(*t)};
        return (int)QSX_OK;
    });
}

int qsxh_tensor_shape(const qsxh_tensor* t, size_t* dims, size_t max_dims, size_t* n_dims) {
    if (t == nullptr || n_dims == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    auto const s = t->tensor.shape();
    *n_dims      = s.size();
    for (size_t i = 0; i < s.size() && i < max_dims && dims != nullptr; ++i) dims[i] = s[i];
    return QSX_OK;
}

int qsxh_tensor_download(const qsxh_tensor* t, double* out, size_t n_doubles) {
    if (t == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        auto const v = t->tensor.host_elements();
        if (n_doubles != 2 * v.size()) return fail(QSX_ERR_BAD_ARG, "buffer size mismatch");
        std::memcpy(out, v.data(), sizeof(double) * n_doubles);
        return (int)QSX_OK;
    });
}

int qsxh_tensor_element(const qsxh_tensor* t, size_t i, size_t j, double re_im[2]) {
    if (t == nullptr || re_im == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        auto const sh = t->tensor.shape();
        if (sh.size() != 2 || i >= sh[0] || j >= sh[1]) return fail(QSX_ERR_BAD_ARG, "element index out of range (the tensor must be a matrix)");
        std::complex<double> const v = std::as_const(t->tensor)(i, j);
        re_im[0] = v.real(), re_im[1] = v.imag();
        return (int)QSX_OK;
    });
}

int qsxh_tensor_write(const qsxh_tensor* t, const char* path) {
    if (t == nullptr || path == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        if (!const_cast<qsxh_tensor*>(t)->tensor.tensor_write(path)) return fail(QSX_ERR_BAD_ARG, "Failed to open file");
        return (int)QSX_OK;
    });
}

int qsxh_tensor_read(const char* path, qsxh_tensor** out) {
    if (path == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return guarded([&] {
        auto t = std::make_unique<qsxh_tensor>();
        if (!t->tensor.tensor_read(path)) return fail(QSX_ERR_BAD_ARG, "the tensor file cannot be read (see the log)");
        *out = t.release();
        return (int)QSX_OK;
    });
}

int qsxh_tensor_print(const qsxh_tensor* t, char* buf, size_t buf_size, size_t* needed) {
    if (t == nullptr) return fail(QSX_ERR_BAD_ARG, "null tensor");
    return guarded([&] {
        std::ostringstream os;
        os << t->tensor;
        std::string const s = os.str();
        if (needed != nullptr) *needed = s.size() + 1;
        if (buf != nullptr && buf_size > 0) {
            size_t const k = std::min(buf_size - 1, s.size());
            std::memcpy(buf, s.data(), k);
            buf[k] = '\0';
        }
        return (int)QSX_OK;
    });
}

int qsxh_tensor_adjoint_inplace(qsxh_tensor* t) {
    if (t == nullptr) return fail(QSX_ERR_BAD_ARG, "null tensor");
    return guarded([&] {
        t->tensor.adjoint_inplace();
        return (int)QSX_OK;
    });
}

int qsxh_tensor_equiv(const qsxh_tensor* t1, const qsxh_tensor* t2, double eps, int strict, qsx_equiv_result* result, char* text,
                      size_t text_size) {
    if (t1 == nullptr || t2 == nullptr) return fail(QSX_ERR_BAD_ARG, "null tensor");
    return guarded([&] {
        using namespace qsyn::tensor;
        std::string msg;
        qsx_equiv_result r{};
        double s[8];
        if (!equivalence_sums(t1->tensor, t2->tensor, s)) {
            auto fmt_shape = [](std::vector<size_t> const& sh) {
                std::string o = "[";
                for (size_t i = 0; i < sh.size(); ++i) o += (i ? ", " : "") + std::to_string(sh[i]);
                return o + "]";
            };
            msg = std::format("Not Equivalent\n- Shape Mismatch: {} vs {}\n", fmt_shape(t1->tensor.shape()), fmt_shape(t2->tensor.shape()));
        } else {
            qsx_equiv_finish(s, eps, strict, &r);
            if (r.equivalent)
                msg = std::format("Equivalent\n- Global Norm : {:.6}\n- Global Phase: {}\n", r.norm,
                                  dvlab::Phase(r.phase_num, r.phase_den).get_print_string());
            else
                msg = std::format("Not Equivalent\n- Cosine Similarity: {:.6}\n", r.cosine);
        }
        if (result != nullptr) *result = r;
        if (text != nullptr && text_size > 0) {
            size_t const k = std::min(text_size - 1, msg.size());
            std::memcpy(text, msg.data(), k);
            text[k] = '\0';
        }
        return (int)QSX_OK;
    });
}

int qsxh_tensor_destroy(qsxh_tensor* t) {
    delete t;
    return QSX_OK;
}

}  // extern "C"
