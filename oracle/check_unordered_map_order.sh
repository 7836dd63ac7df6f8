#!/bin/sh
# Prints the iteration order of a libstdc++ unordered_map<size_t,...> after emplacing 0..n-1,
# for n = 1..40.  Used once to pin oracle._unordered_map_iteration_order (test infrastructure).
set -e
d=$(mktemp -d)
cat > "$d/um.cpp" <<'CPP'
#include <cstdio>
#include <unordered_map>
#include <utility>
int main() {
    for (size_t n = 1; n <= 40; n++) {
        std::unordered_map<size_t, std::pair<size_t, size_t>> m;
        for (size_t q = 0; q < n; q++) m.emplace(q, std::make_pair(2 * q, 2 * q + 1));
        std::printf("%zu:", n);
        for (auto& [k, v] : m) std::printf(" %zu", k);
        std::printf("\n");
    }
}
CPP
g++ -std=c++20 -O1 "$d/um.cpp" -o "$d/um" && "$d/um"
rm -rf "$d"
