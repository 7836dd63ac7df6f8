// zxgraph.hpp -- the ZX operand of the tensor-verification path.
//
// Just enough of qsyn's ZXGraph for `zx read`, `zx vertex/edge add|remove`, `zx tensor-product`,
// `zx adjoint` and ZX -> tensor: vertices with ordered neighbour sets, the edge merge/cancel rules
// of add_edge, the `.zx` reader and the topological order the converter walks.  The behaviour
// (including iteration orders, which the converter's trace log exposes) follows
// src/zx/zxgraph.{hpp,cpp}, zxgraph_action.cpp:41-154, zx_traverse.cpp:22-74 and
// zx_io.cpp:83-436; simplification, extraction, gadgets ... are out of scope.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <format>
#include <limits>
#include <filesystem>
#include <fstream>
#include <functional>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../util/logger.hpp"
#include "../util/phase.hpp"

namespace qsyn::zx {

using QubitIdType = int;

enum class VertexType { boundary, z, x, h_box };
enum class EdgeType { simple, hadamard };

inline char const* to_symbol(VertexType t) {
    switch (t) {
        case VertexType::z: return "Z";
        case VertexType::x: return "X";
        case VertexType::h_box: return "H";
        default: return "●";
    }
}
inline char const* to_symbol(EdgeType t) { return t == EdgeType::hadamard ? "H" : "-"; }

// dvlab::utils::ordered_hashmap / ordered_hashset: insertion order, erased entries vanish,
// emplace of an existing key is a no-op (ordered_hashtable.hpp:280-330).
template <typename K, typename V, typename Hash = std::hash<K>>
class OrderedMap {
public:
    using Entry = std::pair<K, V>;
    bool contains(K const& k) const { return _index.contains(k); }
    size_t size() const { return _index.size(); }
    bool empty() const { return _index.empty(); }
    bool emplace(K const& k, V const& v) {
        if (_index.contains(k)) return false;
        _index.emplace(k, _data.size());
        _data.emplace_back(Entry{k, v});
        return true;
    }
    size_t erase(K const& k) {
        auto it = _index.find(k);
        if (it == _index.end()) return 0;
        _data[it->second].reset();
        _index.erase(it);
        return 1;
    }
    V& at(K const& k) { return _data[_index.at(k)]->second; }
    V const& at(K const& k) const { return _data[_index.at(k)]->second; }
    template <typename F>
    void for_each(F&& f) {
        for (size_t i = 0; i < _data.size(); ++i)  // (index loop: f may emplace)
            if (_data[i].has_value()) f(_data[i]->first, _data[i]->second);
    }
    template <typename F>
    void for_each(F&& f) const {
        for (auto const& e : _data)
            if (e.has_value()) f(e->first, e->second);
    }
    std::vector<K> keys() const {
        std::vector<K> r;
        for (auto const& e : _data)
            if (e.has_value()) r.push_back(e->first);
        return r;
    }
    Entry const& front() const {
        for (auto const& e : _data)
            if (e.has_value()) return *e;
        throw std::out_of_range("empty ordered map");
    }

private:
    std::vector<std::optional<Entry>> _data;
    std::unordered_map<K, size_t, Hash> _index;
};

class ZXVertex;
using NeighborPair = std::pair<ZXVertex*, EdgeType>;
struct NeighborPairHash {
    size_t operator()(NeighborPair const& k) const {
        return std::hash<ZXVertex*>{}(k.first) ^ (k.second == EdgeType::hadamard ? 0x9e3779b97f4a7c15ull : 0ull);
    }
};
using Neighbors = OrderedMap<NeighborPair, char, NeighborPairHash>;
using EdgePair  = std::pair<std::pair<ZXVertex*, ZXVertex*>, EdgeType>;
struct EdgePairHash {
    size_t operator()(EdgePair const& k) const {
        return std::hash<ZXVertex*>{}(k.first.first) * 31 + std::hash<ZXVertex*>{}(k.first.second) * 7 +
               (k.second == EdgeType::hadamard ? 1 : 0);
    }
};

class ZXVertex {
    friend class ZXGraph;

public:
    ZXVertex(size_t id, QubitIdType qubit, VertexType vt, dvlab::Phase phase, float row, float col)
        : _id(id), _qubit(qubit), _type(vt), _phase(phase), _row(row), _col(col) {}
    size_t get_id() const { return _id; }
    void set_id(size_t id) { _id = id; }
    QubitIdType get_qubit() const { return _qubit; }
    void set_qubit(QubitIdType q) { _qubit = q; }
    VertexType type() const { return _type; }
    dvlab::Phase const& phase() const { return _phase; }
    dvlab::Phase& phase() { return _phase; }
    float get_row() const { return _row; }
    float get_col() const { return _col; }
    void set_row(float r) { _row = r; }
    void set_col(float c) { _col = c; }
    bool is_boundary() const { return _type == VertexType::boundary; }
    bool is_zx() const { return _type == VertexType::z || _type == VertexType::x; }

private:
    size_t _id;
    QubitIdType _qubit;
    VertexType _type;
    dvlab::Phase _phase;
    float _row, _col;
    Neighbors _neighbors;
};

inline EdgePair make_edge_pair(ZXVertex* v1, ZXVertex* v2, EdgeType et) {
    return v1->get_id() < v2->get_id() ? EdgePair{{v1, v2}, et} : EdgePair{{v2, v1}, et};
}

class ZXGraph {
public:
    using VertexSet = OrderedMap<ZXVertex*, char>;

    ZXGraph() = default;
    ~ZXGraph() { _delete_all(); }
    ZXGraph(ZXGraph const& other) : _filename(other._filename), _procedures(other._procedures), _next_v_id(other._next_v_id) {
        // zxgraph.cpp:55-73: vertices in order with their ids, then every edge once from its lower-id end
        std::unordered_map<ZXVertex*, ZXVertex*> o2n;
        other._vertices.for_each([&](ZXVertex* v, char) {
            if (v->is_boundary()) {
                o2n[v] = other._inputs.contains(v) ? add_input(v->get_id(), v->get_qubit(), v->get_row(), v->get_col())
                                                   : add_output(v->get_id(), v->get_qubit(), v->get_row(), v->get_col());
            } else {
                o2n[v] = add_vertex(v->get_id(), v->type(), v->phase(), v->get_row(), v->get_col());
            }
        });
        other._vertices.for_each([&](ZXVertex* v, char) {
            v->_neighbors.for_each([&](NeighborPair const& nb, char) {
                if (nb.first->get_id() > v->get_id()) add_edge(o2n[v], o2n[nb.first], nb.second);
            });
        });
    }
    ZXGraph(ZXGraph&& o) noexcept { _swap(o); }
    ZXGraph& operator=(ZXGraph o) noexcept {
        _swap(o);
        return *this;
    }

    std::string const& get_filename() const { return _filename; }
    void set_filename(std::string f) { _filename = std::move(f); }
    std::vector<std::string> const& get_procedures() const { return _procedures; }
    void add_procedure(std::string p) { _procedures.push_back(std::move(p)); }

    bool is_empty() const { return _vertices.empty(); }
    size_t num_vertices() const { return _vertices.size(); }
    size_t num_inputs() const { return _inputs.size(); }
    size_t num_outputs() const { return _outputs.size(); }
    VertexSet const& get_vertices() const { return _vertices; }
    VertexSet const& get_inputs() const { return _inputs; }
    VertexSet const& get_outputs() const { return _outputs; }
    bool is_v_id(size_t id) const { return _id_to_vertices.contains(id); }
    ZXVertex* vertex(size_t id) const { return is_v_id(id) ? _id_to_vertices.at(id) : nullptr; }
    bool is_input_qubit(QubitIdType q) const { return _input_list.contains(q); }
    bool is_output_qubit(QubitIdType q) const { return _output_list.contains(q); }
    QubitIdType max_input_qubit() const { return _max_key(_input_list); }
    QubitIdType max_output_qubit() const { return _max_key(_output_list); }

    Neighbors const& get_neighbors(ZXVertex* v) const { return v->_neighbors; }
    size_t num_neighbors(ZXVertex* v) const { return v->_neighbors.size(); }
    NeighborPair get_first_neighbor(ZXVertex* v) const { return v->_neighbors.front().first; }
    bool is_neighbor(ZXVertex* a, ZXVertex* b) const { return is_neighbor(a, b, EdgeType::simple) || is_neighbor(a, b, EdgeType::hadamard); }
    bool is_neighbor(ZXVertex* a, ZXVertex* b, EdgeType et) const { return a->_neighbors.contains({b, et}); }
    std::optional<EdgeType> get_edge_type(ZXVertex* a, ZXVertex* b) const {
        if (is_neighbor(a, b, EdgeType::simple)) return EdgeType::simple;
        if (is_neighbor(a, b, EdgeType::hadamard)) return EdgeType::hadamard;
        return std::nullopt;
    }

    // zxgraph.cpp:339-430
    ZXVertex* add_input(QubitIdType qubit, float col = 0.f) { return add_input(qubit, static_cast<float>(qubit), col); }
    ZXVertex* add_input(QubitIdType qubit, float row, float col) { return add_input(_next_vertex_id(), qubit, row, col); }
    ZXVertex* add_input(size_t id, QubitIdType qubit, float row, float col) {
        if (_id_to_vertices.contains(id)) {
            spdlog::warn("Vertex with id {} already exists", id);
            return nullptr;
        }
        if (is_input_qubit(qubit)) {
            spdlog::warn("Input qubit {} already exists", qubit);
            return nullptr;
        }
        auto* v = new ZXVertex(id, qubit, VertexType::boundary, dvlab::Phase(), row, col);
        _inputs.emplace(v, 0);
        _input_list.emplace(qubit, v);
        _vertices.emplace(v, 0);
        _id_to_vertices.emplace(id, v);
        return v;
    }
    ZXVertex* add_output(QubitIdType qubit, float col = 0.f) { return add_output(qubit, static_cast<float>(qubit), col); }
    ZXVertex* add_output(QubitIdType qubit, float row, float col) { return add_output(_next_vertex_id(), qubit, row, col); }
    ZXVertex* add_output(size_t id, QubitIdType qubit, float row, float col) {
        if (_id_to_vertices.contains(id)) {
            spdlog::warn("Vertex with id {} already exists", id);
            return nullptr;
        }
        if (is_output_qubit(qubit)) {
            spdlog::warn("Output qubit {} already exists", qubit);
            return nullptr;
        }
        auto* v = new ZXVertex(id, qubit, VertexType::boundary, dvlab::Phase(), row, col);
        _outputs.emplace(v, 0);
        _output_list.emplace(qubit, v);
        _vertices.emplace(v, 0);
        _id_to_vertices.emplace(id, v);
        return v;
    }
    ZXVertex* add_vertex(VertexType vt, dvlab::Phase phase = dvlab::Phase(), float row = 0.f, float col = 0.f) {
        return add_vertex(_next_vertex_id(), vt, phase, row, col);
    }
    ZXVertex* add_vertex(size_t id, VertexType vt, dvlab::Phase phase, float row, float col) {
        if (_id_to_vertices.contains(id)) {
            spdlog::warn("Vertex with id {} already exists", id);
            return nullptr;
        }
        auto* v = new ZXVertex(id, 0, vt, phase, row, col);
        _vertices.emplace(v, 0);
        _id_to_vertices.emplace(id, v);
        return v;
    }

    // zxgraph.cpp:438-493: parallel edges between Z/X spiders merge or cancel (Hopf / pi phase)
    void add_edge(ZXVertex* vs, ZXVertex* vt, EdgeType et) {
        if (vs == vt) {
            if (!vs->is_zx()) throw std::logic_error("Cannot add an edge between a boundary vertex and itself");
            vs->phase() += (et == EdgeType::hadamard ? dvlab::Phase(1) : dvlab::Phase(0));
            return;
        }
        if (vs->get_id() > vt->get_id()) std::swap(vs, vt);
        if (!is_neighbor(vs, vt)) {
            vs->_neighbors.emplace({vt, et}, 0);
            vt->_neighbors.emplace({vs, et}, 0);
            return;
        }
        if (!vs->is_zx() || !vt->is_zx())
            throw std::logic_error(std::format("Cannot add >1 between {}({}) and {}({})", to_symbol(vs->type()), vs->get_id(),
                                               to_symbol(vt->type()), vt->get_id()));
        auto const existing  = get_edge_type(vs, vt).value();
        bool const same_type = vs->type() == vt->type();
        auto const to_merge  = same_type ? EdgeType::simple : EdgeType::hadamard;
        auto const to_cancel = same_type ? EdgeType::hadamard : EdgeType::simple;
        if (existing == to_merge && et == to_merge) {
        } else if (existing == to_cancel && et == to_cancel) {
            remove_edge(vs, vt, to_cancel);
        } else {
            if (existing == to_cancel) {
                remove_edge(vs, vt, to_cancel);
                vs->_neighbors.emplace({vt, to_merge}, 0);
                vt->_neighbors.emplace({vs, to_merge}, 0);
            }
            vs->phase() += dvlab::Phase(1);
        }
    }
    size_t remove_edge(ZXVertex* vs, ZXVertex* vt, EdgeType et) {
        size_t const count = vs->_neighbors.erase({vt, et}) + vt->_neighbors.erase({vs, et});
        if (count == 1)
            throw std::out_of_range("Graph connection error in " + std::to_string(vs->get_id()) + " and " + std::to_string(vt->get_id()));
        return count / 2;
    }
    size_t remove_edge(ZXVertex* vs, ZXVertex* vt) { return remove_edge(vs, vt, EdgeType::simple) + remove_edge(vs, vt, EdgeType::hadamard); }

    // zxgraph.cpp:535-600
    size_t remove_vertex(ZXVertex* v) {
        if (!_vertices.contains(v)) return 0;
        for (auto const& nb : v->_neighbors.keys()) {
            v->_neighbors.erase(nb);
            nb.first->_neighbors.erase({v, nb.second});
        }
        _vertices.erase(v);
        _id_to_vertices.erase(v->get_id());
        if (_inputs.contains(v)) {
            _input_list.erase(v->get_qubit());
            _inputs.erase(v);
        }
        if (_outputs.contains(v)) {
            _output_list.erase(v->get_qubit());
            _outputs.erase(v);
        }
        delete v;
        return 1;
    }
    size_t remove_vertex(size_t id) {
        if (!is_v_id(id)) {
            spdlog::warn("Vertex with id {} does not exist", id);
            return 0;
        }
        return remove_vertex(vertex(id));
    }
    size_t remove_isolated_vertices() {
        size_t n = 0;
        for (ZXVertex* v : _vertices.keys())
            if (num_neighbors(v) == 0) n += remove_vertex(v);
        return n;
    }
    size_t num_edges() const {
        size_t n = 0;
        _vertices.for_each([&](ZXVertex* v, char) { n += v->_neighbors.size(); });
        return n / 2;
    }

    // zxgraph_print.cpp:27-143, zxvertex.cpp:62-78 (printed without a level prefix)
    static std::string vertex_line(ZXVertex const* v) {
        auto nbrs = v->_neighbors.keys();
        std::sort(nbrs.begin(), nbrs.end(), [](NeighborPair const& a, NeighborPair const& b) {
            return a.first->get_id() != b.first->get_id() ? a.first->get_id() < b.first->get_id() : a.second < b.second;
        });
        std::string list;
        for (size_t i = 0; i < nbrs.size(); ++i) list += std::format("{}({}, {})", i ? " " : "", nbrs[i].first->get_id(), to_symbol(nbrs[i].second));
        std::string const head = std::format("({}, {})", to_symbol(v->type()), v->phase().get_print_string());
        // the reference pads to 12 display columns; std::format would count the bytes of the bullet
        size_t cols = 0;
        for (unsigned char ch : head) cols += (ch & 0xC0) != 0x80;
        std::string const where = v->is_boundary() ? std::format("({}, {})", v->get_qubit(), v->get_col()) : std::format("({}, {})", v->get_row(), v->get_col());
        return std::format("ID: {:>4} {}{} (Qubit, Col): {:<14} #Neighbors: {:>3}    {}", v->get_id(), head, std::string(cols < 12 ? 12 - cols : 0, ' '), where,
                           nbrs.size(), list);
    }
    void print_vertices() const {
        std::puts("");
        _vertices.for_each([](ZXVertex* v, char) { std::puts(vertex_line(v).c_str()); });
        std::puts(std::format("Total #Vertices: {}", num_vertices()).c_str());
        std::puts("");
    }
    void print_vertices(std::vector<size_t> const& ids) const {
        std::puts("");
        for (size_t id : ids)
            if (is_v_id(id)) std::puts(vertex_line(vertex(id)).c_str());
        std::puts("");
    }
    void print_boundary(bool inputs) const {
        std::string ids;
        (inputs ? _inputs : _outputs).for_each([&](ZXVertex* v, char) { ids += (ids.empty() ? "" : ", ") + std::to_string(v->get_id()); });
        std::puts(std::format("{} ({})", inputs ? "Input: " : "Output:", ids).c_str());
        std::puts(std::format("Total #{}: {}", inputs ? "Inputs" : "Outputs", inputs ? num_inputs() : num_outputs()).c_str());
    }
    void print_edges() const {
        _vertices.for_each([&](ZXVertex* v, char) {
            v->_neighbors.for_each([&](NeighborPair const& nb, char) {
                if (nb.first->get_id() > v->get_id())
                    std::puts(std::format("{:<12} Type: {}", std::format("({}, {})", v->get_id(), nb.first->get_id()), to_symbol(nb.second)).c_str());
            });
        });
        std::puts(std::format("Total #Edges: {}", num_edges()).c_str());
    }
    void print_graph() const {
        std::puts(std::format("Graph ({} inputs, {} outputs, {} vertices, {} edges)", num_inputs(), num_outputs(), num_vertices(), num_edges()).c_str());
    }

    // zxgraph.cpp:672-681
    void adjoint() {
        std::swap(_inputs, _outputs);
        std::swap(_input_list, _output_list);
        float max_col = 0.f;
        bool first    = true;
        _vertices.for_each([&](ZXVertex* v, char) {
            max_col = first ? v->get_col() : std::max(max_col, v->get_col());
            first   = false;
        });
        _vertices.for_each([&](ZXVertex* v, char) {
            v->phase() = -v->phase();
            v->set_col(max_col - v->get_col());
        });
    }

    // zxgraph_action.cpp:41-68
    void lift_qubit(long n) {
        _vertices.for_each([&](ZXVertex* v, char) {
            if (v->get_row() >= 0) v->set_row(v->get_row() + static_cast<float>(n));
        });
        _inputs.for_each([&](ZXVertex* v, char) { v->set_qubit(v->get_qubit() + (int)n); });
        _outputs.for_each([&](ZXVertex* v, char) { v->set_qubit(v->get_qubit() + (int)n); });
        std::unordered_map<QubitIdType, ZXVertex*> ni, no;
        for (auto const& [q, v] : _input_list) ni[q + (int)n] = v;
        for (auto const& [q, v] : _output_list) no[q + (int)n] = v;
        _input_list  = ni;
        _output_list = no;
    }
    // zxgraph_action.cpp:121-154
    ZXGraph& tensor_product(ZXGraph const& target) {
        ZXGraph copied{target};
        QubitIdType ori_max = std::numeric_limits<QubitIdType>::min(), ori_min = std::numeric_limits<QubitIdType>::max();
        QubitIdType cp_min  = std::numeric_limits<QubitIdType>::max();
        auto span = [&](ZXVertex* v, char) {
            ori_max = std::max(ori_max, v->get_qubit());
            ori_min = std::min(ori_min, v->get_qubit());
        };
        _inputs.for_each(span);
        _outputs.for_each(span);
        auto low = [&](ZXVertex* v, char) { cp_min = std::min(cp_min, v->get_qubit()); };
        copied._inputs.for_each(low);
        copied._outputs.for_each(low);
        copied.lift_qubit(static_cast<long>(ori_max - ori_min + 1) - cp_min);
        copied._inputs.for_each([&](ZXVertex* v, char) { _inputs.emplace(v, 0); });
        for (auto const& kv : copied._input_list) _input_list.insert(kv);
        copied._outputs.for_each([&](ZXVertex* v, char) { _outputs.emplace(v, 0); });
        for (auto const& kv : copied._output_list) _output_list.insert(kv);
        // _move_vertices_from (zxgraph.cpp:511-521): fresh consecutive ids, ownership moves over
        copied._vertices.for_each([&](ZXVertex* v, char) {
            v->set_id(_next_v_id);
            _id_to_vertices.emplace(_next_v_id, v);
            ++_next_v_id;
            _vertices.emplace(v, 0);
        });
        copied._release();
        return *this;
    }

    // zxgraph.cpp:209-216
    bool is_io_connection_valid() const {
        bool ok = true;
        _inputs.for_each([&](ZXVertex* v, char) { ok = ok && num_neighbors(v) == 1; });
        _outputs.for_each([&](ZXVertex* v, char) { ok = ok && num_neighbors(v) == 1; });
        return ok;
    }

    // zx_traverse.cpp:22-74: reversed DFS post-order from the inputs, then the outputs
    std::vector<ZXVertex*> create_topological_order() const {
        std::vector<ZXVertex*> order;
        std::unordered_set<ZXVertex*> visited;
        auto dfs = [&](ZXVertex* root) {
            std::vector<std::pair<bool, ZXVertex*>> st;
            if (!visited.contains(root)) st.emplace_back(false, root);
            while (!st.empty()) {
                auto [done, v] = st.back();
                st.pop_back();
                if (done) {
                    order.push_back(v);
                    continue;
                }
                if (visited.contains(v)) continue;
                visited.insert(v);
                st.emplace_back(true, v);
                v->_neighbors.for_each([&](NeighborPair const& nb, char) {
                    if (!visited.contains(nb.first)) st.emplace_back(false, nb.first);
                });
            }
        };
        _inputs.for_each([&](ZXVertex* v, char) { dfs(v); });
        _outputs.for_each([&](ZXVertex* v, char) { dfs(v); });
        std::reverse(order.begin(), order.end());
        std::string ids;
        for (size_t i = 0; i < order.size(); ++i) ids += (i ? " " : "") + std::to_string(order[i]->get_id());
        spdlog::trace("Topological order from first input: {}", ids);
        spdlog::trace("Size of topological order: {}", order.size());
        return order;
    }

private:
    std::string _filename;
    std::vector<std::string> _procedures;
    size_t _next_v_id = 0;
    VertexSet _inputs, _outputs, _vertices;
    std::unordered_map<QubitIdType, ZXVertex*> _input_list, _output_list;
    std::unordered_map<size_t, ZXVertex*> _id_to_vertices;

    size_t _next_vertex_id() {  // zxgraph.cpp:157-162, post-incremented by the callers
        while (is_v_id(_next_v_id)) ++_next_v_id;
        return _next_v_id++;
    }
    static QubitIdType _max_key(std::unordered_map<QubitIdType, ZXVertex*> const& m) {
        QubitIdType r = -1;
        for (auto const& kv : m) r = std::max(r, kv.first);
        return r;
    }
    void _delete_all() {
        _vertices.for_each([](ZXVertex* v, char) { delete v; });
        _release();
    }
    void _release() {
        _inputs = _outputs = _vertices = VertexSet{};
        _input_list.clear();
        _output_list.clear();
        _id_to_vertices.clear();
    }
    void _swap(ZXGraph& o) {
        std::swap(_filename, o._filename);
        std::swap(_procedures, o._procedures);
        std::swap(_next_v_id, o._next_v_id);
        std::swap(_inputs, o._inputs);
        std::swap(_outputs, o._outputs);
        std::swap(_vertices, o._vertices);
        std::swap(_input_list, o._input_list);
        std::swap(_output_list, o._output_list);
        std::swap(_id_to_vertices, o._id_to_vertices);
    }
};

// zx_io.cpp:530-560
std::optional<ZXGraph> from_zx(std::istream& in, bool keep_id = false);
std::optional<ZXGraph> from_zx(std::filesystem::path const& path, bool keep_id = false);

}  // namespace qsyn::zx
