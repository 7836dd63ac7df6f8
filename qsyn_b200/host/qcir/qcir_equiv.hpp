// qcir_equiv.hpp -- `qcir equiv`: are two circuits the same operator up to a global phase?
// Reference: src/qcir/qcir_equiv.cpp:19-66.  The reference first tries to prove the equivalence by optimising the tableau
// of qcir1^-1 . qcir2 to the empty tableau (src/tableau, not part of this path) and falls back to tensors for <= 7
// qubits: to_tensor of the (optimised) composed circuit against QTensor::identity(n), on the interleaved rank-2n shape.
// Here the tableau step does not exist -- it can only ever turn a "not proven" into "equivalent" earlier -- so the
// tensor half runs directly on qcir1^-1 . qcir2, and, because the tensor lives on the GPU, it is not limited to 7 qubits.
#pragma once

#include "../convert/qcir_to_tensor.hpp"
#include "../tensor/qtensor.hpp"
#include "../util/logger.hpp"
#include "./qcir.hpp"

namespace qsyn::qcir {

// largest circuit checked by tensor contraction: 4^n x 16 B x 2 operands must fit the device (n = 15: 32 GiB)
inline constexpr size_t kMaxEquivQubits = 15;

inline bool is_equivalent(QCir const& qcir1, QCir const& qcir2) {
    if (qcir1.get_num_qubits() != qcir2.get_num_qubits()) {
        spdlog::info("The two circuits have different numbers of qubits.");
        return false;
    }
    spdlog::info("Tableau optimization is not part of this build; verifying equivalence via tensor contraction...");

    auto adjoint_composed = qcir1;
    adjoint_composed.adjoint_inplace();
    adjoint_composed.compose(qcir2);

    if (adjoint_composed.get_num_qubits() > kMaxEquivQubits) {
        spdlog::warn("The number of qubits is too large to check equivalence via tensor contraction.");
        spdlog::warn("Please note that this may be a false negative.");
        return false;
    }
    if (adjoint_composed.get_num_qubits() == 0) return true;

    auto const tensor1 = qsyn::to_tensor(adjoint_composed);
    if (!tensor1) {
        spdlog::error("Failed to convert the composed QCir to tensor.");
        return false;
    }
    // is qcir1^-1 . qcir2 close enough to the identity circuit?  (both operands have the interleaved rank-2n shape;
    // the host-built identity is lifted to the device by equivalence_sums, qtensor.hpp)
    return tensor::is_equivalent(*tensor1, tensor::QTensor<double>::identity(adjoint_composed.get_num_qubits()));
}

}  // namespace qsyn::qcir
