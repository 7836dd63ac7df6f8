// tensor_util.hpp -- shape / index / axis-list aliases and helpers of the tensor package.
// Same names and meaning as the reference's src/tensor/tensor_util.{hpp,cpp}:15-49, without
// the xtensor svector.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <vector>

namespace qsyn::tensor {

using TensorShape    = std::vector<size_t>;
using TensorIndex    = std::vector<size_t>;
using TensorAxisList = std::vector<size_t>;

inline TensorAxisList concat_axis_list(TensorAxisList const& ax1, TensorAxisList const& ax2) {
    TensorAxisList ax(ax1);
    ax.insert(ax.end(), ax2.begin(), ax2.end());
    return ax;
}

inline void print_axis_list(TensorAxisList const& ax) {
    for (size_t i = 0; i < ax.size(); ++i) std::printf("%s%zu", i ? ", " : "", ax[i]);
    std::printf("\n");
}

inline bool is_disjoint(TensorAxisList const& ax1, TensorAxisList const& ax2) {
    return std::none_of(ax1.begin(), ax1.end(), [&](size_t a) { return std::find(ax2.begin(), ax2.end(), a) != ax2.end(); });
}

}  // namespace qsyn::tensor
