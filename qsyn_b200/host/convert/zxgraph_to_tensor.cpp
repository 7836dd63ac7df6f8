// zxgraph_to_tensor.cpp -- ZX diagram -> tensor, one tensordot per vertex in topological order
// with frontier / axis-id bookkeeping.  Algorithm, log texts and the order of every container
// walk follow src/convert/zxgraph_to_tensor.cpp:95-385 of the reference, so the debug / trace
// output is byte-identical (tests/conversion/zx2ts/ref/*.log).
#include "./zxgraph_to_tensor.hpp"

#include <map>

extern bool stop_requested();

namespace qsyn {

namespace {

using tensor::QTensor;
using tensor::TensorAxisList;
using zx::EdgePair;
using zx::EdgeType;
using zx::ZXGraph;
using zx::ZXVertex;

using Frontiers = zx::OrderedMap<EdgePair, size_t, zx::EdgePairHash>;

struct Piece {
    Frontiers frontiers;
    QTensor<double> tensor;
};

QTensor<double> tensor_form(ZXGraph const& g, ZXVertex* v) {
    switch (v->type()) {
        case zx::VertexType::z: return QTensor<double>::zspider(g.num_neighbors(v), v->phase());
        case zx::VertexType::x: return QTensor<double>::xspider(g.num_neighbors(v), v->phase());
        case zx::VertexType::h_box: return QTensor<double>::hbox(g.num_neighbors(v));
        default: return QTensor<double>::identity(g.num_neighbors(v));
    }
}

class Mapper {
public:
    std::optional<QTensor<double>> map(ZXGraph const& graph) {
        if (graph.is_empty()) {
            spdlog::error("The ZXGraph is empty!!");
            return std::nullopt;
        }
        if (!graph.is_io_connection_valid()) {
            spdlog::error("The ZXGraph is not valid!!");
            return std::nullopt;
        }
        for (ZXVertex* v : graph.create_topological_order()) map_one_vertex(graph, v);
        if (stop_requested()) {
            spdlog::error("Conversion is interrupted!!");
            return std::nullopt;
        }
        QTensor<double> result(std::complex<double>(1., 0.));
        for (auto const& p : _pieces) result = tensor::tensordot<std::complex<double>>(result, p.tensor);
        for (size_t i = 0; i < _pieces.size(); ++i) _pieces[i].frontiers.emplace(_boundary_edges[i], 0);

        auto const [input_ids, output_ids] = axis_orders(graph);
        auto join = [](TensorAxisList const& l) {
            std::string s;
            for (size_t i = 0; i < l.size(); ++i) s += (i ? " " : "") + std::to_string(l[i]);
            return s;
        };
        spdlog::trace("Input  Axis IDs: {}", join(input_ids));
        spdlog::trace("Output Axis IDs: {}", join(output_ids));
        return result.to_matrix(output_ids, input_ids);
    }

private:
    std::vector<EdgePair> _boundary_edges;
    std::vector<Piece> _pieces;
    std::unordered_map<ZXVertex*, size_t> _pins;
    size_t _cur = 0;

    struct MappingInfo {
        TensorAxisList simple_edge_pins, hadamard_edge_pins;
        std::vector<EdgePair> frontiers_to_remove, frontiers_to_add;
    };

    void map_one_vertex(ZXGraph const& graph, ZXVertex* v) {
        if (stop_requested()) return;
        _cur = _pins.size();
        for (auto const& nb : graph.get_neighbors(v).keys())
            if (_pins.contains(nb.first)) {
                _cur = _pins.at(nb.first);
                break;
            }
        if (_cur == _pins.size()) {
            initialize_subgraph(graph, v);
        } else {
            tensordot_vertex(graph, v);
        }
        _pins.emplace(v, _cur);
        spdlog::debug("Done. Current tensor dimension: {}", _pieces[_cur].tensor.dimension());
        spdlog::trace("Current frontiers:");
        _pieces[_cur].frontiers.for_each([](EdgePair const& e, size_t axid) {
            spdlog::trace("  {}--{} ({}) axis id: {}", e.first.first->get_id(), e.first.second->get_id(), zx::to_symbol(e.second), axid);
        });
    }

    void initialize_subgraph(ZXGraph const& graph, ZXVertex* v) {
        spdlog::debug("Mapping vertex {:>4} ({}): New Subgraph", v->get_id(), zx::to_symbol(v->type()));
        auto const [nb, etype] = graph.get_first_neighbor(v);
        _cur                   = _pieces.size();
        _pieces.emplace_back();
        _pieces[_cur].tensor = tensor::tensordot<std::complex<double>>(QTensor<double>(std::complex<double>(1., 0.)),
                                                                        QTensor<double>::identity(graph.num_neighbors(v)));
        auto const key = zx::make_edge_pair(v, nb, etype);
        _boundary_edges.push_back(key);
        _pieces[_cur].frontiers.emplace(key, 1);
    }

    MappingInfo mapping_info(ZXGraph const& graph, ZXVertex* v) {
        MappingInfo info;
        graph.get_neighbors(v).for_each([&](zx::NeighborPair const& nbr, char) {
            auto const key = zx::make_edge_pair(v, nbr.first, nbr.second);
            if (!_pins.contains(nbr.first)) {
                info.frontiers_to_add.push_back(key);
            } else {
                size_t const axid = _pieces[_cur].frontiers.at(key);
                (key.second == EdgeType::hadamard ? info.hadamard_edge_pins : info.simple_edge_pins).push_back(axid);
                info.frontiers_to_remove.push_back(key);
            }
        });
        return info;
    }

    // Hadamard edges become plain ones by contracting an H box onto each of their axes
    QTensor<double> dehadamardize(QTensor<double> const& ts, MappingInfo& info) {
        QTensor<double> const h_pow = tensor::tensor_product_pow<std::complex<double>>(QTensor<double>::hbox(2), info.hadamard_edge_pins.size());
        TensorAxisList connect;
        for (size_t i = 0; i < info.hadamard_edge_pins.size(); ++i) connect.push_back(2 * i);
        QTensor<double> tmp = tensor::tensordot<std::complex<double>>(ts, h_pow, info.hadamard_edge_pins, connect);
        _pieces[_cur].frontiers.for_each([&](EdgePair const&, size_t& axid) {
            auto const it = std::find(info.hadamard_edge_pins.begin(), info.hadamard_edge_pins.end(), axid);
            if (it != info.hadamard_edge_pins.end())
                axid = tmp.get_new_axis_id(ts.dimension() + connect[(size_t)(it - info.hadamard_edge_pins.begin())] + 1);
            else
                axid = tmp.get_new_axis_id(axid);
        });
        for (size_t i = 0; i < info.hadamard_edge_pins.size(); ++i)
            info.hadamard_edge_pins[i] = tmp.get_new_axis_id(ts.dimension() + connect[i] + 1);
        for (auto& pin : info.simple_edge_pins) pin = tmp.get_new_axis_id(pin);
        info.simple_edge_pins = tensor::concat_axis_list(info.hadamard_edge_pins, info.simple_edge_pins);
        return tmp;
    }

    void tensordot_vertex(ZXGraph const& graph, ZXVertex* v) {
        auto info = mapping_info(graph, v);
        if (v->is_boundary()) {
            spdlog::debug("Mapping vertex {:>4} ({}): Boundary", v->get_id(), zx::to_symbol(v->type()));
            _pieces[_cur].tensor = dehadamardize(_pieces[_cur].tensor, info);
            return;
        }
        spdlog::debug("Mapping vertex {:>4} ({}): Tensordot", v->get_id(), zx::to_symbol(v->type()));
        QTensor<double> const dehadamarded = dehadamardize(_pieces[_cur].tensor, info);
        TensorAxisList vertex_pins;  // every vertex tensor is symmetric: any of its axes will do
        for (size_t i = 0; i < info.simple_edge_pins.size(); ++i) vertex_pins.push_back(i);
        _pieces[_cur].tensor = tensor::tensordot<std::complex<double>>(dehadamarded, tensor_form(graph, v), info.simple_edge_pins, vertex_pins);
        for (auto const& e : info.frontiers_to_remove) _pieces[_cur].frontiers.erase(e);
        _pieces[_cur].frontiers.for_each([&](EdgePair const&, size_t& axid) { axid = _pieces[_cur].tensor.get_new_axis_id(axid); });
        for (size_t t = 0; t < info.frontiers_to_add.size(); ++t)
            _pieces[_cur].frontiers.emplace(info.frontiers_to_add[t],
                                            _pieces[_cur].tensor.get_new_axis_id(dehadamarded.dimension() + vertex_pins.size() + t));
    }

    // which tensor axis belongs to which input / output qubit (zxgraph_to_tensor.cpp:224-270)
    std::pair<TensorAxisList, TensorAxisList> axis_orders(ZXGraph const& g) {
        TensorAxisList inputs(g.num_inputs()), outputs(g.num_outputs());
        auto table = [](ZXGraph::VertexSet const& vs) {
            auto list = vs.keys();
            std::stable_sort(list.begin(), list.end(), [](ZXVertex* a, ZXVertex* b) { return a->get_qubit() < b->get_qubit(); });
            std::map<int, size_t> t;
            for (size_t i = 0; i < list.size(); ++i) t[list[i]->get_qubit()] = i;
            return t;
        };
        auto const in_tab = table(g.get_inputs()), out_tab = table(g.get_outputs());
        size_t acc = 0;
        for (auto& piece : _pieces) {
            bool b2b = false;
            piece.frontiers.for_each([&](EdgePair const& e, size_t& axid) {
                auto* const v1 = e.first.first;
                auto* const v2 = e.first.second;
                bool const v1i = g.get_inputs().contains(v1), v2i = g.get_inputs().contains(v2);
                bool const v1o = g.get_outputs().contains(v1), v2o = g.get_outputs().contains(v2);
                if (v1i) inputs[in_tab.at(v1->get_qubit())] = axid + acc;
                if (v2i) inputs[in_tab.at(v2->get_qubit())] = axid + acc;
                if (v1o) outputs[out_tab.at(v1->get_qubit())] = axid + acc;
                if (v2o) outputs[out_tab.at(v2->get_qubit())] = axid + acc;
                // a boundary-to-boundary wire has both ends on one frontier: shift one of them
                if (v1i && (v2i || v2o)) {
                    inputs[in_tab.at(v1->get_qubit())]--;
                    b2b = true;
                }
                if (v1o && (v2i || v2o)) {
                    outputs[out_tab.at(v1->get_qubit())]--;
                    b2b = true;
                }
            });
            acc += piece.frontiers.size() + (b2b ? 1 : 0);
        }
        return {inputs, outputs};
    }
};

}  // namespace

std::optional<tensor::QTensor<double>> to_tensor(zx::ZXGraph const& zxgraph) try {
    Mapper mapper;
    return mapper.map(zxgraph);  // a host tensor; `tensor equiv` lifts square operands to the device
} catch (std::bad_alloc const&) {
    spdlog::error("Memory allocation failed!!");
    return std::nullopt;
}

}  // namespace qsyn
