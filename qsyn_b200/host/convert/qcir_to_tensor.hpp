// qcir_to_tensor.hpp -- conversion from QCir to Tensor.
// Same declarations as the reference's src/convert/qcir_to_tensor.hpp:21-23.
#pragma once

#include "../qcir/qcir.hpp"
#include "../tensor/qtensor.hpp"

using namespace qsyn::qcir;

namespace qsyn {

std::optional<qsyn::tensor::QTensor<double>> to_tensor(QCirGate const& gate);
template <>
std::optional<qsyn::tensor::QTensor<double>> to_tensor(QCir const& qcir);

}  // namespace qsyn
