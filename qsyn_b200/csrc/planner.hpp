// planner.hpp -- host-side lowering + scheduling of a gate stream into sweeps.
// Pure C++ (no CUDA); see planner.cpp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qsx.h"
#include "program.hpp"

namespace qsx {

struct LoweredOp {
    int32_t  kind  = OP_U1;
    int      t0    = -1;  // non-diagonal target row bits (global positions)
    int      t1    = -1;
    int      dbit  = -1;  // OP_DIAG: row bit selecting d0/d1 (needs no locality)
    uint32_t req   = 0;   // required-one mask over global row bits (controls; phase-type: incl. the phase bit)
    double   c[8]  = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> mat;  // OP_U2: 4x4 complex, row-major; OP_DENSE: 2^k x 2^k complex, row-major
    std::vector<int> dbits;   // OP_DENSE: global row bit of matrix-index bit i (i = 0 is the least significant)
    int      src_gate = -1;
    double   frac     = 1.0;  // touched fraction f of SURVEY 8(d)

    uint32_t tmask() const {
        uint32_t m = (t0 >= 0 ? 1u << t0 : 0u) | (t1 >= 0 ? 1u << t1 : 0u);
        for (int b : dbits) m |= 1u << b;
        return m;
    }
    uint32_t dmask() const { return req | (dbit >= 0 ? 1u << dbit : 0u); }
};

// what the kernel executes (see MicroKind); masks are in GLOBAL row bits except creg
struct MicroOp {
    int32_t  kind   = K_U1;
    int      j0     = -1, j1 = -1;  // register-bit indices
    uint32_t creg   = 0;            // in-register required-one mask (non-diagonal kinds)
    uint32_t cother = 0;            // thread-level required-one mask (global row bits outside the registers)
    uint32_t czero  = 0;            // thread-level required-zero mask
    int      sel    = -1;           // K_DSCAL: global row bit choosing d1
    double   c[8]   = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> table;      // K_DTAB*: complex entries ; K_U2: 4x4 matrix
    int      src    = -1;           // non-diagonal kinds: position in PlannedStage::ops (of the first fused op)
    int      n_src  = 1;            // semantic ops multiplied into this micro-op (single-qubit fusion)
};

struct PlannedStage {
    std::vector<int> reg_bits;  // global row bit of register bit j (size R)
    std::vector<int> pre;       // X-type ops folded into this stage's load addressing (first stage only)
    std::vector<int> ops;       // indices into Plan::ops executed in registers (semantic order)
    std::vector<int> post;      // X-type ops folded into this stage's store addressing
    uint32_t post_taint = 0;    // row bits that depend on a register bit after the post list so far
    std::vector<MicroOp> uops;  // what the kernel runs for this stage
};

struct PlannedSweep {
    std::vector<int> tile_bits;  // global row bits in the tile, ascending (size m)
    uint32_t fixed_one = 0;
    std::vector<PlannedStage> stages;
    double alg_bytes = 0;
    // dense sweep (dense_block option): one 2^k x 2^k complex block applied on the tensor pipe
    std::vector<int> block_bits;        // global row bit of component-index bit i (size k), empty = interpreted sweep
    std::vector<double> block_matrix;   // complex, row-major, 2^k x 2^k
    std::vector<int> block_ops;         // the ops folded into it (indices into Plan::ops)
};

struct PlanConfig {
    int  n = 0;
    uint64_t ncols = 0;
    int  R = 4;          // register bits (kernel template parameter)
    int  log2w = 3;
    int  m_cap = 10;
    int  max_ops = 96;
    int  min_tail_ops = 8;   // a trailing stage with fewer ops is dropped from the sweep
    int  lookahead = 4096;
    int  search = 0;         // rounds of hill climbing over each sweep's tile bits
    bool fuse = true;
    bool fold_perms = true;  // fold X / CX / SWAP into the addressing of stage boundaries
    bool defer_perms = false;
    bool stage_search = false;        // pick each stage's register bits by trying every R-subset of the tile
    bool fuse_1q = true;              // multiply runs of uncontrolled single-qubit ops on one register bit
    bool conjugate_diagonals = true;  // carry pending diagonal tables through in-register permutations
    int  dense_k = 0;  // >= 3: fold the gate stream into dense k-qubit blocks (FP64 tensor pipe)  // never swap registers: an X-type op that cannot be hoisted ends the stage
    double snap_tol = 1e-15;
    int  l2_block_mib = 0;        // option as given: 0 auto, < 0 off, > 0 forced
    uint64_t block_cols = 0;      // columns per L2 block (0: the sweeps run over the whole shard)
    int  jit = 0;                 // specialised sweeps: 0 never, 1 when ready (never blocks), 2 wait for the compiler
};

struct Plan {
    PlanConfig cfg;
    std::vector<LoweredOp> ops;
    std::vector<PlannedSweep> sweeps;
    qsx_plan_stats stats{};
    // serialized device programs: one blob per sweep, concatenated; offsets are 256-B aligned
    std::vector<unsigned char> blob;
    std::vector<size_t> sweep_offset;
    std::vector<size_t> sweep_size;
    // specialised kernels of the sweeps (jit.hpp), requested on first use; nullptr = interpret
    std::vector<void*> jit_kernels;
};

// returns QSX_OK or an error code; msg receives the reason
int build_plan(int n_qubits, uint64_t ncols, const qsx_gate* gates, size_t n_gates,
               const qsx_plan_options* opt, Plan& plan, std::string& msg);
std::string dump_plan_json(const Plan& plan);

}  // namespace qsx
