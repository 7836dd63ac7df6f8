// program.hpp -- the device-visible "sweep program" produced by the scheduler
// (planner.cpp) and executed by the sweep interpreter kernel (kernels.cu).
//
// Vocabulary (row-bit view of SURVEY.md Appendix B):
//   row bit p     : bit p of the row (output) index of U; reference qubit q <-> p = n-1-q
//   sweep         : ONE kernel launch = one read+write pass over the shard in HBM
//   tile          : 2^m rows x W columns staged per CTA (m "tile bits" = row bits whose
//                   2^m combinations the CTA owns; the other n-m row bits select the tile)
//   stage         : within a sweep, each thread holds 2^r rows x 1 column in registers
//                   (r "register bits" out of the tile bits); stages are separated by a
//                   shared-memory transpose that changes which tile bits are in registers
//   op            : one lowered gate, applied in registers
#pragma once
#include <cstdint>

namespace qsx {

constexpr int kMaxRegBits  = 4;
constexpr int kMaxTileBits = 11;
constexpr int kMaxRowBits  = 26;

// CTA size limit of sweep_kernel<R>: 2^R complex128 live in registers per thread, so the
// R = 4 instantiation is compiled for <= 512 threads (128 registers), the others for 1024.
constexpr int max_threads_log2(int R) { return R >= 4 ? 9 : 10; }

enum OpKind : int32_t {
    OP_U1    = 0,  // general 2x2 on register bit j0:  c = m00,m01,m10,m11 (re,im)
    OP_H     = 1,  // s*(a+b), s*(a-b)                  c[0] = s
    OP_X     = 2,  // swap the pair (exact permutation)
    OP_AD    = 3,  // anti-diagonal: a' = m01*b, b' = m10*a   c = m01, m10
    OP_PHASE = 4,  // e *= w where all required bits are 1    c = w
    OP_DIAG  = 5,  // e *= (bit j0 ? d1 : d0)                 c = d0, d1
    OP_SWAP  = 6,  // swap register bits j0 and j1
    OP_U2    = 7,  // dense 4x4 on register bits (1,0) (first target = bit 1); matrix in the pool
    OP_NEG   = 8,  // e = -e where required bits are 1  (snapped Z / CZ)
    OP_MULI  = 9,  // e *= +i (c[0] = +1) or -i (c[0] = -1) where required bits are 1 (snapped S / Sdg)
};

// All masks are "required-one" masks: the op acts on an element only if every bit of the
// mask is 1 in the element's row index.
struct alignas(16) DevOp {
    int32_t  kind;
    int32_t  j0, j1;   // register-bit indices of the non-diagonal targets (or -1)
    uint32_t creg;     // required-one mask over the register index (bits 0..r-1)
    uint32_t cother;   // required-one mask over GLOBAL row bits that are not register bits this stage
    int32_t  pool;     // OP_U2: offset in doubles into the constant pool
    int32_t  pad0, pad1;
    double   c[8];
};
static_assert(sizeof(DevOp) == 96, "DevOp layout");

struct alignas(16) DevStage {
    int32_t op_begin, op_end;     // [begin,end) into the sweep's op array
    int32_t r;                    // register bits used (== kernel template R)
    int32_t n_t;                  // number of thread-row bits = m - r
    int32_t rl[kMaxRegBits];      // tile-local bit position of register bit j
    int32_t tl[kMaxTileBits];     // tile-local bit positions of the thread-row bits, ascending
    int32_t pad;
};
static_assert(sizeof(DevStage) % 16 == 0, "DevStage layout");

struct alignas(16) DevSweep {
    int32_t  n_stages;
    int32_t  m;                   // tile bits
    int32_t  log2w;               // log2(tile columns)
    int32_t  n_outer;             // free (enumerated) row bits outside the tile
    uint32_t fixed_one;           // outer row bits forced to 1 (tile skipping for all-controlled sweeps)
    int32_t  n_ops;
    int32_t  pool_doubles;
    int32_t  pad;
    int32_t  sp[kMaxTileBits];    // global row-bit position of tile-local bit l, ascending
    int32_t  outer[kMaxRowBits];  // global positions of the enumerated outer bits, ascending
    int32_t  pad2[3];
    // followed in memory by: DevStage stages[n_stages]; DevOp ops[n_ops]; double pool[pool_doubles]
};
static_assert(sizeof(DevSweep) % 16 == 0, "DevSweep layout");

}  // namespace qsx
