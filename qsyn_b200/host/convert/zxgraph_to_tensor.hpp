// zxgraph_to_tensor.hpp -- `convert zx tensor` / `zx2ts` (src/convert/zxgraph_to_tensor.{hpp,cpp}).
#pragma once

#include <optional>

#include "../tensor/qtensor.hpp"
#include "../zx/zxgraph.hpp"

namespace qsyn {

// ZXGraph -> (2^#outputs x 2^#inputs) matrix.  The diagram is contracted on the host (its
// intermediate tensors have a handful of axes); equivalence_sums() lifts a square 2^n x 2^n
// result to the device, so `tensor equiv` runs as the same fused GPU reduction as for QCirs.
std::optional<tensor::QTensor<double>> to_tensor(zx::ZXGraph const& zxgraph);

}  // namespace qsyn
