// planner_internal.hpp -- pieces of the planner shared between its translation units
// (planner_lower.cpp, planner_schedule.cpp, planner_emit.cpp, planner.cpp).  Not installed.
#pragma once

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "planner.hpp"

namespace qsx::detail {

inline int popc(uint32_t x) { return __builtin_popcount(x); }

// entries within tol of 0 / +-1 are snapped (tol < 0: never)
inline double snap1(double x, double tol) {
    if (tol < 0) return x;
    if (std::fabs(x) <= tol) return 0.0;
    if (std::fabs(x - 1.0) <= tol) return 1.0;
    if (std::fabs(x + 1.0) <= tol) return -1.0;
    return x;
}

// log2 of a power of two, -1 otherwise
inline int ilog2_exact(uint64_t x) {
    if (x == 0 || (x & (x - 1)) != 0) return -1;
    int k = 0;
    while ((uint64_t{1} << k) < x) ++k;
    return k;
}

// planner_lower.cpp: one reference gate -> register-level ops (appended to out)
int lower_gate(int n, const qsx_gate& g, int gate_index, double tol, std::vector<LoweredOp>& out, std::string& msg);
// planner_schedule.cpp: ops -> sweeps -> stages (calls emit_uops per stage)
void schedule(Plan& plan);
void schedule_dense(Plan& plan);
// planner_emit.cpp: the micro-ops the kernel runs for one stage
void emit_uops(Plan const& plan, PlannedStage& st);

}  // namespace qsx::detail
