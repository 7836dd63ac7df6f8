// qcir_to_tensor.hpp -- conversion from QCir to Tensor.
// Same declarations as the reference's src/convert/qcir_to_tensor.hpp:21-23.
#pragma once

#include <memory>

#include "../qcir/qcir.hpp"
#include "../tensor/qtensor.hpp"

using namespace qsyn::qcir;

namespace qsyn {

// The ordered gate stream the device consumes (include/qsx.h): one entry per gate of
// QCir::get_gates(), controls peeled off so that only the 2^t x 2^t TARGET matrix (built with the
// reference formulas) crosses the ABI.  `gates[i].matrix` points into `*matrices[i]`; gates of the same operation share
// one matrix.
struct GateStream {
    std::vector<qsx_gate> gates;
    std::vector<std::shared_ptr<std::vector<double> const>> matrices;
};
// nullopt + error log if a gate has no tensor form (qcir_to_tensor.cpp:180-183 of the reference)
std::optional<GateStream> to_gate_stream(qsyn::qcir::QCir const& qcir);

std::optional<qsyn::tensor::QTensor<double>> to_tensor(QCirGate const& gate);
template <>
std::optional<qsyn::tensor::QTensor<double>> to_tensor(QCir const& qcir);

}  // namespace qsyn
