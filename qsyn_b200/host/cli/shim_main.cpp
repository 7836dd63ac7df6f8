// shim_main.cpp -- `qsyn-b200`: a thin dofile runner around the B200 tensor-verification path.
//
// enough of qsyn's command surface to replay the reference's own dofiles for
// this path verbatim and diff the output against its golden logs:
//     qcir {new|read|delete|list|print|qubit add|gate add}, convert qcir tensor (alias qc2ts),
//     zx {new|read|delete|checkout|list|vertex add|edge add|edge remove|tensor-product|adjoint},
//     convert zx tensor (alias zx2ts)                                     [SURVEY.md 8 f2],
//     tensor {list|print|equiv|adjoint|delete|checkout|write|read}, logger, quit -f.
// Echo / spacing rules and message texts follow cli_exec.cpp:40-105, conversion_cmd.cpp:84-98,
// tensor_cmd.cpp:100-177, data_structure_manager.hpp and the golden logs.  The CLI framework
// of qsyn itself (argparse, history, aliases, tab completion ...) is out of scope.
#include <cstdio>
#include <cstdlib>
#include <format>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../convert/qcir_to_tensor.hpp"
#include "../convert/zxgraph_to_tensor.hpp"
#include "../qcir/qcir.hpp"
#include "../qcir/qcir_equiv.hpp"
#include "../tensor/qtensor.hpp"
#include "../util/logger.hpp"

bool stop_requested() { return false; }  // qsyn defines this in qsyn_main.cpp:53 (SIGINT -> jthread stop token)

namespace {

using qsyn::qcir::QCir;
using qsyn::tensor::QTensor;
using qsyn::zx::ZXGraph;

// minimal DataStructureManager with the reference's messages (data_structure_manager.hpp:35-198)
template <typename T>
class Manager {
public:
    explicit Manager(std::string name) : _type_name(std::move(name)) {}
    bool empty() const { return _list.empty(); }
    size_t size() const { return _list.size(); }
    size_t get_next_id() const { return _next_id; }
    size_t focused_id() const { return _focused_id; }
    bool is_id(size_t id) const { return _list.contains(id); }
    T* get() const { return size() ? _list.at(_focused_id).get() : nullptr; }
    T* find_by_id(size_t id) const {
        if (!is_id(id)) {
            std::fprintf(stderr, "Error: The ID provided does not exist!!\n");
            return nullptr;
        }
        return _list.at(id).get();
    }
    T* add(size_t id, std::unique_ptr<T> t = std::make_unique<T>()) {
        _list.insert_or_assign(id, std::move(t));
        _focused_id = id;
        if (id >= _next_id) _next_id = id + 1;
        spdlog::info("Successfully created and checked out to {0} {1}", _type_name, id);
        return get();
    }
    void set(std::unique_ptr<T> t) {
        if (_list.contains(_focused_id)) spdlog::info("Note: Replacing {} {}...", _type_name, _focused_id);
        _list.insert_or_assign(_focused_id, std::move(t));
    }
    void remove(size_t id) {
        if (!_list.contains(id)) {
            std::fprintf(stderr, "Error: The ID provided does not exist!!\n");
            return;
        }
        _list.erase(id);
        spdlog::info("Successfully removed {0} {1}", _type_name, id);
        if (size() && _focused_id == id) {
            std::cout << std::format("Note: Focused graph is deleted. Checked out to {} 0\n", _type_name);
            checkout(0);
        }
        if (empty()) std::cout << std::format("Note: The {} list is empty now\n", _type_name);
    }
    void clear() {
        _list.clear();
        _next_id = _focused_id = 0;
    }
    void checkout(size_t id) {
        if (!_list.contains(id)) {
            std::fprintf(stderr, "Error: The ID provided does not exist!!\n");
            return;
        }
        _focused_id = id;
        spdlog::info("Checked out to {} {}", _type_name, _focused_id);
    }
    template <typename F>
    void print_list(F info) const {
        if (empty()) {
            std::cout << std::format("The {} list is empty\n", _type_name);
            return;
        }
        for (auto const& [id, data] : _list) std::cout << std::format("{} {}    {}\n", id == _focused_id ? "★" : " ", id, info(*data));
    }
    std::string const& type_name() const { return _type_name; }

private:
    size_t _next_id = 0, _focused_id = 0;
    std::map<size_t, std::unique_ptr<T>> _list;
    std::string _type_name;
};

std::string join(std::vector<std::string> const& v, std::string const& sep) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) s += (i ? sep : "") + v[i];
    return s;
}

// left-align to `width` display columns (ASCII file names)
std::string pad_right(std::string s, size_t width) {
    if (s.size() < width) s.append(width - s.size(), ' ');
    return s;
}

bool is_prefix(std::string const& p, std::string const& full) { return !p.empty() && full.compare(0, p.size(), p) == 0; }

struct Shell {
    Manager<QCir> qcir_mgr{"QCir"};
    Manager<QTensor<double>> tensor_mgr{"Tensor"};
    Manager<ZXGraph> zx_mgr{"ZXGraph"};
    bool quit = false;

    bool need_zx() {
        if (!zx_mgr.empty()) return true;
        spdlog::error("ZXGraph list is empty. Please create a ZXGraph first!!");
        return false;
    }

    bool need_qcir() {
        if (!qcir_mgr.empty()) return true;
        spdlog::error("QCir list is empty. Please create a QCir first!!");
        return false;
    }
    bool need_tensor() {
        if (!tensor_mgr.empty()) return true;
        spdlog::error("Tensor list is empty. Please create a Tensor first!!");
        return false;
    }

    void cmd_logger(std::vector<std::string> const& a) {
        static std::pair<char const*, spdlog::level::level_enum> const levels[] = {
            {"trace", spdlog::level::trace}, {"debug", spdlog::level::debug},       {"info", spdlog::level::info},
            {"warning", spdlog::level::warn}, {"error", spdlog::level::err},         {"critical", spdlog::level::critical},
            {"off", spdlog::level::off}};
        if (a.empty()) {
            std::cout << "Logger Level: " << levels[spdlog::get_level()].first << "\n";
            return;
        }
        for (auto const& [name, lvl] : levels)
            if (is_prefix(a[0], name)) {
                spdlog::set_level(lvl);
                spdlog::info("Setting logger level to \"{}\"", name);
                return;
            }
        std::fprintf(stderr, "Error: invalid logger level \"%s\"!!\n", a[0].c_str());
    }

    void cmd_qcir(std::vector<std::string> a) {
        if (a.empty()) return;
        std::string const sub = a[0];
        a.erase(a.begin());
        if (is_prefix(sub, "new")) {
            qcir_mgr.add(a.empty() ? qcir_mgr.get_next_id() : std::stoul(a[0]));
        } else if (is_prefix(sub, "read")) {
            bool replace = false;
            std::string path;
            for (auto const& t : a) {
                if (t == "-r" || is_prefix(t, "--replace") )
                    replace = true;
                else
                    path = t;
            }
            auto qc = qsyn::qcir::from_file(path);
            if (!qc.has_value()) {
                std::cout << std::format("Error: the format in \"{}\" has something wrong!!\n", path);
                return;
            }
            auto up = std::make_unique<QCir>(std::move(*qc));
            if (qcir_mgr.empty() || !replace) {
                qcir_mgr.add(qcir_mgr.get_next_id(), std::move(up));
            } else {
                qcir_mgr.set(std::move(up));
            }
            qcir_mgr.get()->set_filename(std::filesystem::path(path).stem().string());
        } else if (is_prefix(sub, "delete")) {
            if (!a.empty() && (a[0] == "--all" || a[0] == "-a"))
                qcir_mgr.clear();
            else if (!a.empty())
                qcir_mgr.remove(std::stoul(a[0]));
        } else if (is_prefix(sub, "list")) {
            qcir_mgr.print_list([](QCir const& q) { return pad_right(q.get_filename().substr(0, 19), 19) + " " + join(q.get_procedures(), " ➔ "); });
        } else if (is_prefix(sub, "print")) {
            if (!need_qcir()) return;
            if (!a.empty() && (a[0] == "-d" || (a[0].size() > 2 && is_prefix(a[0], "--diagram")))) {
                for (auto const& l : qcir_mgr.get()->circuit_diagram_lines()) std::cout << l << "\n";
            } else {
                std::cout << std::format("QCir ({} qubits, {} gates)\n", qcir_mgr.get()->get_num_qubits(), qcir_mgr.get()->get_num_gates());
            }
        } else if (is_prefix(sub, "equiv")) {
            // qcir_cmd.cpp:603-650: one id = compare the focused circuit with it
            if (!need_qcir() || a.empty()) return;
            std::vector<size_t> ids;
            for (auto const& t : a) ids.push_back(std::stoul(t));
            for (size_t id : ids)
                if (qcir_mgr.find_by_id(id) == nullptr) {
                    std::fprintf(stderr, "Error: QCir %zu does not exist!!\n", id);
                    return;
                }
            if ((ids.size() == 2 && ids[0] == ids[1]) || (ids.size() == 1 && qcir_mgr.focused_id() == ids[0]))
                spdlog::info("Note: comparing the same circuit...");
            bool const eq = ids.size() == 1 ? qsyn::qcir::is_equivalent(*qcir_mgr.get(), *qcir_mgr.find_by_id(ids[0]))
                                            : qsyn::qcir::is_equivalent(*qcir_mgr.find_by_id(ids[0]), *qcir_mgr.find_by_id(ids[1]));
            std::cout << (eq ? "The two circuits are equivalent!!" : "The two circuits are not equivalent!!") << "\n";
        } else if (is_prefix(sub, "adjoint")) {
            if (!need_qcir()) return;
            qcir_mgr.get()->adjoint_inplace();
        } else if (is_prefix(sub, "checkout")) {
            if (!a.empty()) qcir_mgr.checkout(std::stoul(a[0]));
        } else if (is_prefix(sub, "qubit")) {
            if (a.empty() || !is_prefix(a[0], "add")) return;
            size_t const n = a.size() > 1 ? std::stoul(a[1]) : 1;
            if (qcir_mgr.empty()) qcir_mgr.add(qcir_mgr.get_next_id());  // qcir_cmd.cpp:507-527
            qcir_mgr.get()->add_qubits(n);
        } else if (is_prefix(sub, "gate")) {
            if (a.empty() || !is_prefix(a[0], "add") || !need_qcir()) return;
            std::string type;
            std::vector<dvlab::Phase> params;
            qsyn::QubitIdList qubits;
            for (size_t i = 1; i < a.size(); ++i) {
                if (a[i] == "-ph" || is_prefix(a[i], "--phase")) {
                    if (i + 1 >= a.size()) {
                        std::fprintf(stderr, "Error: missing argument for -ph!!\n");
                        return;
                    }
                    auto const ph = dvlab::Phase::from_string(a[++i]);
                    if (!ph.has_value()) {
                        std::fprintf(stderr, "Error: invalid phase \"%s\"!!\n", a[i].c_str());
                        return;
                    }
                    params.push_back(*ph);
                } else if (type.empty()) {
                    type = a[i];
                } else {
                    qubits.push_back(std::stoul(a[i]));
                }
            }
            auto const op = qsyn::qcir::str_to_operation(type, params);
            if (!op.has_value()) {
                spdlog::error("Invalid gate type {}!!", type);
                return;
            }
            if (op->get_num_qubits() != qubits.size()) {
                spdlog::error("Gate {} requires {} qubits, but {} qubits are given!!", type, op->get_num_qubits(), qubits.size());
                return;
            }
            for (auto q : qubits)
                if (q >= qcir_mgr.get()->get_num_qubits()) {
                    spdlog::error("Qubit ID {} does not exist!!", q);
                    return;
                }
            qcir_mgr.get()->append(*op, qubits);
        }
    }

    // zx_cmd.cpp:219-262 (read), 364-470 (vertex add), 512-560 (edge add/remove), zxgraph_mgr
    void cmd_zx(std::vector<std::string> a) {
        if (a.empty()) return;
        std::string const sub = a[0];
        a.erase(a.begin());
        using qsyn::zx::EdgeType;
        using qsyn::zx::VertexType;
        if (is_prefix(sub, "new")) {
            zx_mgr.add(a.empty() ? zx_mgr.get_next_id() : std::stoul(a[0]));
        } else if (is_prefix(sub, "read")) {
            bool keep_id = false, replace = false;
            std::string path;
            for (auto const& t : a) {
                if (t == "--keep-id")
                    keep_id = true;
                else if (t == "-r" || is_prefix(t, "--replace"))
                    replace = true;
                else
                    path = t;
            }
            auto graph = qsyn::zx::from_zx(std::filesystem::path(path), keep_id);
            if (!graph.has_value()) return;
            auto up = std::make_unique<ZXGraph>(std::move(*graph));
            if (zx_mgr.empty() || !replace) {
                zx_mgr.add(zx_mgr.get_next_id(), std::move(up));
            } else {
                zx_mgr.set(std::move(up));
            }
            zx_mgr.get()->set_filename(std::filesystem::path(path).stem().string());
        } else if (is_prefix(sub, "delete")) {
            if (!a.empty() && (a[0] == "--all" || a[0] == "-a"))
                zx_mgr.clear();
            else if (!a.empty())
                zx_mgr.remove(std::stoul(a[0]));
        } else if (is_prefix(sub, "checkout")) {
            if (!a.empty()) zx_mgr.checkout(std::stoul(a[0]));
        } else if (is_prefix(sub, "list")) {
            zx_mgr.print_list([](ZXGraph const& g) { return pad_right(g.get_filename().substr(0, 19), 19) + " " + join(g.get_procedures(), " ➔ "); });
        } else if (is_prefix(sub, "print")) {
            if (!need_zx()) return;
            ZXGraph const* g = zx_mgr.get();
            std::cout << std::flush;
            if (a.empty()) {
                g->print_graph();
            } else if (a[0] == "-v" || (a[0].size() > 2 && is_prefix(a[0], "--vertices"))) {
                std::vector<size_t> ids;
                for (size_t i = 1; i < a.size(); ++i) ids.push_back(std::stoul(a[i]));
                ids.empty() ? g->print_vertices() : g->print_vertices(ids);
            } else if (a[0] == "-i" || (a[0].size() > 2 && is_prefix(a[0], "--inputs"))) {
                g->print_boundary(true);
            } else if (a[0] == "-o" || (a[0].size() > 2 && is_prefix(a[0], "--outputs"))) {
                g->print_boundary(false);
            } else if (a[0] == "-e" || (a[0].size() > 2 && is_prefix(a[0], "--edges"))) {
                g->print_edges();
            } else {
                std::fprintf(stderr, "Error: `zx print %s` is outside the tensor-verification path of this build!!\n", a[0].c_str());
            }
            std::fflush(stdout);
        } else if (is_prefix(sub, "vertex") && a.size() >= 2 && is_prefix(a[0], "remove")) {
            if (!need_zx()) return;
            ZXGraph* g = zx_mgr.get();
            if (a[1] == "-i" || is_prefix(a[1], "--isolated")) {
                spdlog::info("Removing isolated vertices...");
                g->remove_isolated_vertices();
                return;
            }
            for (size_t i = 1; i < a.size(); ++i) {
                size_t const id = std::stoul(a[i]);
                if (!g->is_v_id(id)) {
                    spdlog::error("Cannot find vertex with ID {} in the ZXGraph!!", id);
                    return;
                }
            }
            for (size_t i = 1; i < a.size(); ++i) {
                spdlog::info("Removing vertex {}...", a[i]);
                g->remove_vertex(std::stoul(a[i]));
            }
        } else if (is_prefix(sub, "vertex")) {
            if (a.size() < 2 || !is_prefix(a[0], "add") || !need_zx()) return;
            ZXGraph* g           = zx_mgr.get();
            std::string const vt = a[1];
            if (is_prefix(vt, "input") || is_prefix(vt, "output")) {
                bool const in = is_prefix(vt, "input");
                int const qid = a.size() > 2 ? std::stoi(a[2]) : (in ? g->max_input_qubit() : g->max_output_qubit()) + 1;
                if (in ? g->is_input_qubit(qid) : g->is_output_qubit(qid)) {
                    if (in)
                        spdlog::error("Input vertex for qubit {} already exists!!", qid);
                    else
                        spdlog::error("Error: output vertex for qubit {} already exists!!", qid);
                    return;
                }
                in ? g->add_input(qid) : g->add_output(qid);
                spdlog::info("Adding {} vertex for qubit {}...", in ? "input" : "output", qid);
                return;
            }
            VertexType type;
            if (is_prefix(vt, "zspider"))
                type = VertexType::z;
            else if (is_prefix(vt, "xspider"))
                type = VertexType::x;
            else if (is_prefix(vt, "hbox"))
                type = VertexType::h_box;
            else {
                std::fprintf(stderr, "Error: invalid vertex type \"%s\"!!\n", vt.c_str());
                return;
            }
            dvlab::Phase phase = type == VertexType::h_box ? dvlab::Phase(1) : dvlab::Phase(0);
            if (a.size() > 2) {
                auto const ph = dvlab::Phase::from_string(a[2]);
                if (!ph.has_value()) {
                    std::fprintf(stderr, "Error: invalid phase \"%s\"!!\n", a[2].c_str());
                    return;
                }
                phase = *ph;
            }
            auto* v = g->add_vertex(type, phase);
            spdlog::info("Adding vertex {}...", v->get_id());
        } else if (is_prefix(sub, "edge")) {
            if (a.size() < 3 || !need_zx()) return;
            ZXGraph* g = zx_mgr.get();
            size_t const id0 = std::stoul(a[1]), id1 = std::stoul(a[2]);
            for (size_t id : {id0, id1})
                if (!g->is_v_id(id)) {
                    spdlog::error("Cannot find vertex with ID {} in the ZXGraph!!", id);
                    return;
                }
            auto parse_etype = [&](std::string t) -> std::optional<EdgeType> {
                for (auto& ch : t) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
                if (is_prefix(t, "simple")) return EdgeType::simple;
                if (is_prefix(t, "hadamard")) return EdgeType::hadamard;
                return std::nullopt;
            };
            if (is_prefix(a[0], "add")) {
                auto const et = a.size() > 3 ? parse_etype(a[3]) : std::nullopt;
                if (!et.has_value()) {
                    std::fprintf(stderr, "Error: missing or invalid edge type!!\n");
                    return;
                }
                try {
                    g->add_edge(g->vertex(id0), g->vertex(id1), *et);
                } catch (std::logic_error const& e) {
                    spdlog::error("{}", e.what());
                }
            } else if (is_prefix(a[0], "remove")) {
                auto const et = a.size() > 3 ? parse_etype(a[3]) : std::nullopt;
                if (et.has_value())
                    g->remove_edge(g->vertex(id0), g->vertex(id1), *et);
                else
                    g->remove_edge(g->vertex(id0), g->vertex(id1));
            }
        } else if (sub == "tensor-product") {
            if (a.empty() || !need_zx()) return;
            auto* other = zx_mgr.find_by_id(std::stoul(a[0]));
            if (other != nullptr) zx_mgr.get()->tensor_product(*other);
        } else if (is_prefix(sub, "adjoint")) {
            if (!need_zx()) return;
            zx_mgr.get()->adjoint();
        } else {
            std::fprintf(stderr, "Error: zx subcommand \"%s\" is outside the tensor-verification path of this build!!\n", sub.c_str());
        }
    }

    // conversion_cmd.cpp:221-240
    void cmd_zx2ts() {
        if (!need_zx()) return;
        spdlog::info("Converting ZXGraph {} to Tensor {}...", zx_mgr.focused_id(), tensor_mgr.get_next_id());
        auto tensor = qsyn::to_tensor(*zx_mgr.get());
        if (tensor.has_value()) {
            tensor_mgr.add(tensor_mgr.get_next_id(), std::make_unique<QTensor<double>>(std::move(tensor.value())));
            tensor_mgr.get()->set_filename(zx_mgr.get()->get_filename());
            tensor_mgr.get()->add_procedures(zx_mgr.get()->get_procedures());
            tensor_mgr.get()->add_procedure("ZX2TS");
        }
    }

    // conversion_cmd.cpp:84-98
    void cmd_qc2ts() {
        if (!need_qcir()) return;
        spdlog::info("Converting to QCir {} to Tensor {}...", qcir_mgr.focused_id(), tensor_mgr.get_next_id());
        auto tensor = qsyn::to_tensor(*qcir_mgr.get());
        if (tensor.has_value()) {
            *tensor = std::move(*tensor).to_matrix();  // (rvalue form: the device tensor is re-tagged, not cloned)
            tensor_mgr.add(tensor_mgr.get_next_id());
            tensor_mgr.set(std::make_unique<QTensor<double>>(std::move(tensor.value())));
            tensor_mgr.get()->set_filename(qcir_mgr.get()->get_filename());
            tensor_mgr.get()->add_procedures(qcir_mgr.get()->get_procedures());
            tensor_mgr.get()->add_procedure("QC2TS");
        }
    }

    void cmd_tensor(std::vector<std::string> a) {
        if (a.empty()) return;
        std::string const sub = a[0];
        a.erase(a.begin());
        if (is_prefix(sub, "list")) {
            tensor_mgr.print_list([](QTensor<double> const& t) {
                return std::format("{} #Dim: {}   {}", pad_right(t.get_filename().substr(0, 19), 19), t.dimension(), join(t.get_procedures(), " ➔ "));
            });
        } else if (is_prefix(sub, "print")) {
            if (!need_tensor()) return;
            auto* t = a.empty() ? tensor_mgr.get() : tensor_mgr.find_by_id(std::stoul(a[0]));
            if (t != nullptr) std::cout << *t << std::endl;
        } else if (is_prefix(sub, "adjoint")) {
            if (!need_tensor()) return;
            auto* t = a.empty() ? tensor_mgr.get() : tensor_mgr.find_by_id(std::stoul(a[0]));
            if (t != nullptr) t->adjoint_inplace();
        } else if (is_prefix(sub, "checkout")) {
            if (!a.empty()) tensor_mgr.checkout(std::stoul(a[0]));
        } else if (is_prefix(sub, "delete")) {
            if (!a.empty() && (a[0] == "--all" || a[0] == "-a"))
                tensor_mgr.clear();
            else if (!a.empty())
                tensor_mgr.remove(std::stoul(a[0]));
        } else if (is_prefix(sub, "write")) {
            if (!need_tensor() || a.empty()) return;
            tensor_mgr.get()->tensor_write(a[0]);
        } else if (is_prefix(sub, "read")) {
            if (a.empty()) return;
            auto t = std::make_unique<QTensor<double>>();
            if (!t->tensor_read(a[0])) return;
            tensor_mgr.add(tensor_mgr.get_next_id(), std::move(t));
        } else if (is_prefix(sub, "equiv")) {
            cmd_tensor_equiv(a);
        }
    }

    // tensor_cmd.cpp:120-177
    void cmd_tensor_equiv(std::vector<std::string> const& a) {
        double eps  = 1e-6;
        bool strict = false;
        std::vector<size_t> ids;
        for (size_t i = 0; i < a.size(); ++i) {
            if (a[i] == "-e" || (a[i].size() > 2 && is_prefix(a[i], "--epsilon"))) {
                if (i + 1 >= a.size()) return;
                eps = std::stod(a[++i]);
            } else if (a[i] == "-s" || (a[i].size() > 2 && is_prefix(a[i], "--strict"))) {
                strict = true;
            } else {
                ids.push_back(std::stoul(a[i]));
            }
        }
        if (ids.empty() || ids.size() > 2 || !need_tensor()) return;
        QTensor<double>*tensor1 = nullptr, *tensor2 = nullptr;
        if (ids.size() == 2) {
            tensor1 = tensor_mgr.find_by_id(ids[0]);
            tensor2 = tensor_mgr.find_by_id(ids[1]);
        } else {
            tensor1 = tensor_mgr.get();
            tensor2 = tensor_mgr.find_by_id(ids[0]);
        }
        if (tensor1 == nullptr || tensor2 == nullptr) return;

        using namespace qsyn::tensor;
        // ONE fused reduction provides every quantity the reference computes in 8 (12) passes
        double s[8];
        bool const same_shape = equivalence_sums(*tensor1, *tensor2, s);
        if (!same_shape) {
            std::cout << "Not Equivalent\n";
            auto fmt_shape = [](std::vector<size_t> const& sh) {
                std::string r = "[";
                for (size_t i = 0; i < sh.size(); ++i) r += (i ? ", " : "") + std::to_string(sh[i]);
                return r + "]";
            };
            std::cout << std::format("- Shape Mismatch: {} vs {}\n", fmt_shape(tensor1->shape()), fmt_shape(tensor2->shape()));
            return;
        }
        qsx_equiv_result r{};
        qsx_equiv_finish(s, eps, strict ? 1 : 0, &r);
        if (r.equivalent) {
            std::cout << "Equivalent\n";
            std::cout << std::format("- Global Norm : {:.6}\n", r.norm);
            std::cout << std::format("- Global Phase: {}\n", dvlab::Phase(r.phase_num, r.phase_den).get_print_string());
        } else {
            std::cout << "Not Equivalent\n";
            std::cout << std::format("- Cosine Similarity: {:.6}\n", r.cosine);
        }
    }

    void execute(std::string const& text) {
        std::istringstream ss(text);
        std::vector<std::string> tok;
        for (std::string t; ss >> t;) tok.push_back(t);
        if (tok.empty()) return;
        std::string const cmd = tok[0];
        tok.erase(tok.begin());
        if (cmd == "qc2ts") {
            cmd_qc2ts();
        } else if (cmd == "zx2ts") {
            cmd_zx2ts();
        } else if (is_prefix(cmd, "convert") && cmd.size() >= 3) {
            if (tok.size() >= 2 && is_prefix(tok[0], "qcir") && is_prefix(tok[1], "tensor"))
                cmd_qc2ts();
            else if (tok.size() >= 2 && tok[0] == "zx" && is_prefix(tok[1], "tensor"))
                cmd_zx2ts();
            else
                std::fprintf(stderr, "Error: this build only converts qcir / zx -> tensor!!\n");
        } else if (cmd == "zx") {
            cmd_zx(tok);
        } else if (cmd == "qcir") {
            cmd_qcir(tok);
        } else if (is_prefix(cmd, "tensor") && cmd.size() >= 3) {
            cmd_tensor(tok);
        } else if (is_prefix(cmd, "logger") && cmd.size() >= 3) {
            cmd_logger(tok);
        } else if (is_prefix(cmd, "quit") || cmd == "exit") {
            quit = true;
        } else {
            std::fprintf(stderr, "Error: command \"%s\" is outside the tensor-verification path of this build!!\n", cmd.c_str());
        }
    }

    // one dofile line: echo, then every `;`-separated command (cli_exec.cpp:40-105)
    void run_line(std::string const& raw) {
        std::cout << "qsyn> " << raw << "\n";
        std::string text    = raw.substr(0, raw.find("//"));
        size_t const b      = text.find_first_not_of(" \t\r");
        if (b == std::string::npos) return;  // empty / comment-only line: echo only
        text = text.substr(b, text.find_last_not_of(" \t\r") + 1 - b);
        std::cout << std::flush;
        std::istringstream parts(text);
        for (std::string one; std::getline(parts, one, ';');) {
            execute(one);
            if (quit) break;
        }
        std::cout << "\n" << std::flush;  // one blank line after the command's output
    }
};

}  // namespace

int main(int argc, char** argv) {
    std::string dofile, one_liner;
    for (int i = 1; i < argc; ++i) {
        std::string const a = argv[i];
        if (a == "--no-version" || a == "--verbose" || a == "-v") continue;
        if (a == "--qsynrc-path") {
            ++i;
            continue;
        }
        if (a == "-c" && i + 1 < argc) {
            one_liner = argv[++i];
            continue;
        }
        dofile = a;
    }
    Shell sh;
    if (!one_liner.empty()) {
        sh.execute(one_liner);
        return 0;
    }
    if (dofile.empty()) {
        std::fprintf(stderr, "usage: qsyn-b200 [--no-version] [--qsynrc-path P] [--verbose] <dofile> | -c \"command\"\n");
        return 2;
    }
    std::ifstream in(dofile);
    if (!in.is_open()) {
        std::fprintf(stderr, "Error: cannot open dofile \"%s\"!!\n", dofile.c_str());
        return 1;
    }
    try {
        for (std::string line; !sh.quit && std::getline(in, line);) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            sh.run_line(line);
        }
    } catch (std::exception const& e) {
        // e.g. no CUDA device: this build has no CPU fallback for the tensor path
        std::cout << std::flush;
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    // tensors (and their device shards) go first, then the communicators, while the CUDA runtime is still alive
    sh.tensor_mgr.clear();
    qsx_shutdown();
    return 0;
}
