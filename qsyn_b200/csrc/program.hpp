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
//   op            : one lowered gate (semantic form, global row bits)
//   micro-op      : what the kernel executes in registers; diagonal ops of a stage are merged
//                   into per-thread scalars (K_DSCAL) and small phase tables (K_DTAB*)
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define QSX_HD __host__ __device__
#else
#define QSX_HD
#endif

namespace qsx {

constexpr int kMaxRegBits  = 5;
constexpr int kMaxTileBits = 11;
constexpr int kMaxRowBits  = 26;

// CTA size limit of sweep_kernel<R>: 2^R complex128 live in registers per thread, so the
// R = 3, 4 instantiations are compiled for <= 512 threads (128 registers), R = 5 (32 elements =
// 128 data registers) for 128 threads x 3 CTAs per SM (170 registers), the others for 1024.
QSX_HD constexpr int max_threads_log2(int R) { return R >= 5 ? 7 : R >= 3 ? 9 : 10; }
QSX_HD constexpr int min_ctas_per_sm(int R) { return R >= 5 ? 2 : 1; }

// semantic kinds of lowered gates (host side, planner + dumps)
enum OpKind : int32_t {
    OP_U1    = 0,  // general 2x2 on target bit t0:  c = m00,m01,m10,m11 (re,im)
    OP_H     = 1,  // s*(a+b), s*(a-b)                  c[0] = s
    OP_X     = 2,  // swap the pair (exact permutation)
    OP_AD    = 3,  // anti-diagonal: a' = m01*b, b' = m10*a   c = m01, m10
    OP_PHASE = 4,  // e *= w where all required bits are 1    c = w
    OP_DIAG  = 5,  // e *= (dbit ? d1 : d0) where required     c = d0, d1
    OP_SWAP  = 6,  // swap row bits t0 and t1
    OP_U2    = 7,  // dense 4x4 on (t0,t1) (t0 = MSB)
    OP_NEG   = 8,  // PHASE with w = -1  (snapped Z / CZ)
    OP_MULI  = 9,  // PHASE with w = +-i (snapped S / Sdg), c[0] = +-1
    OP_DENSE = 10, // dense 2^k x 2^k matrix on k = 3..kMaxDenseK row bits (user gates with >= 3 targets, controls folded
                   // in): never interpreted -- always a dense sweep of its own (dense_kernel<k>, FP64 tensor pipe)
};

// kinds the kernel interprets
enum MicroKind : int32_t {
    K_U1     = 0,   // 2x2 on register bit j0, creg = in-register controls
    K_H      = 1,
    K_X      = 2,
    K_AD     = 3,
    K_XC     = 4,   // X on register bit j0 controlled by register bit j1 (compile-time specialised CX)
    K_SWAP   = 5,   // swap register bits j0, j1
    K_U2     = 6,   // dense 4x4 on register bits (1,0); matrix at pool
    K_DSCAL  = 7,   // (scheduler-internal) one per-thread scalar: (sel >= 0 && row bit sel) ? d1 : d0, c = d0, d1;
                    // never sent to the device on its own: a stage's scalars travel as ONE K_SCALN
    K_DTABH  = 8,   // e[i] *= tab[i without bit j0] for the 8 (2^(R-1)) elements with bit j0 set; table at pool
    K_DTABF  = 9,   // e[i] *= tab[i] for all 2^R elements; table at pool
    K_PHJ    = 12,  // e[i] *= w for the elements with register bit j0 set (a K_DTABH whose entries are all equal)
    K_SCALN  = 11,  // all per-thread scalars of a stage in one micro-op: for k < j0 entries at pool
                    // {cother, czero, sel, pad, d0.re, d0.im, d1.re, d1.im} (48 B each):
                    //   s *= pass ? (sel >= 0 && row bit sel ? d1 : d0) : 1 ; then e[i] *= s
};

// Flat handler ids: the kernel dispatches every micro-op with ONE switch.  A handler is a (kind,
// register bit(s), "has in-register controls") combination, numbered for the maximum R = kJ;
// combinations that do not exist for a smaller R are never emitted.
constexpr int kJ = kMaxRegBits;
enum Handler : int32_t {
    // -- common handlers, numbered densely from 0 (the lean kernel contains only these) --
    H_H      = 0,                    // +J          uncontrolled inside the registers (no per-element tests)
    H_X      = H_H + kJ,             // +J
    H_XC     = H_X + kJ,             // +J*kJ+JC
    H_U1     = H_XC + kJ * kJ,       // +J
    H_DTABH  = H_U1 + kJ,            // +J
    H_DTABF  = H_DTABH + kJ,
    H_SCALN  = H_DTABF + 1,
    H_PHJ    = H_SCALN + 1,          // +J
    H_COMMON = H_PHJ + kJ,           // number of common handlers
    // -- rare handlers (full kernel only) --
    H_HG     = H_COMMON,             // +J          generic: runtime creg
    H_XG     = H_HG + kJ,            // +J
    H_U1G    = H_XG + kJ,            // +J
    H_AD     = H_U1G + kJ,           // +J
    H_ADG    = H_AD + kJ,            // +J
    H_SWAP   = H_ADG + kJ,           // +JA*kJ+JB  (JA > JB)
    H_SWAPG  = H_SWAP + kJ * kJ,     // +JA*kJ+JB
    H_U2     = H_SWAPG + kJ * kJ,
    H_COUNT  = H_U2 + 1,
};
static_assert(H_COMMON == 52 && H_COUNT == 128, "the kernel's switch enumerates these ranges");

// Thread-level predicate of every micro-op: (row & cother) == cother && (row & czero) == 0, where
// `row` is the thread's global row index with the register bits cleared.
struct alignas(16) DevOp {
    // first 16 bytes: everything the dispatch of a micro-op needs, one LDS.128
    int32_t  handler;  // Handler id (kind + register bits folded in)
    uint32_t cother;   // required-one mask over GLOBAL row bits that are not register bits this stage
    uint32_t czero;    // required-zero mask over GLOBAL row bits that are not register bits this stage
    int32_t  pool;     // offset in doubles into the constant pool (K_U2 matrix, K_DTAB* table)
    // second 16 bytes: read by the few handlers that need them
    int32_t  j0, j1;   // register-bit indices (or -1); K_SCALN: j0 = number of entries
    uint32_t creg;     // required-one mask over the register index (bits 0..r-1), generic non-diagonal kinds only
    int32_t  sel;      // K_DSCAL: global row bit selecting d1 (or -1)
    double   c[8];
};
static_assert(sizeof(DevOp) == 96, "DevOp layout");

// A permutation step folded into the addressing of a stage boundary: within the tile,
// row ^= tml  wherever (row & cml) == cml and the CTA's outer row bits contain cmo.
// X / CX / CCX.. are one step; SWAP is three.  Executed on ADDRESSES, never on data: the
// element a thread holds is simply stored to (loaded from) the permuted row.
struct alignas(16) DevPerm {
    uint32_t cml;  // required-one mask, tile-local bit positions
    uint32_t cmo;  // required-one mask, global positions of row bits outside the tile
    uint32_t tml;  // flipped bit, tile-local position (exactly one bit)
    uint32_t pad;
};

struct alignas(16) DevStage {
    int32_t op_begin, op_end;     // [begin,end) into the sweep's micro-op array
    int32_t r;                    // register bits used (== kernel template R)
    int32_t n_t;                  // number of thread-row bits = m - r
    int32_t rl[kMaxRegBits];      // tile-local bit position of register bit j
    int32_t tl[kMaxTileBits];     // tile-local bit positions of the thread-row bits, ascending
    int32_t pad[3];
    int32_t pre_begin, pre_end;   // permutation steps applied at this stage's LOAD (stage 0 only: global load)
    int32_t post_begin, post_end; // permutation steps applied at this stage's STORE
    int32_t rg[kMaxRegBits];      // GLOBAL row-bit position of register bit j   (= sp[rl[j]], precomputed)
    int32_t tg[kMaxTileBits + 1]; // GLOBAL row-bit position of thread-row bit k (= sp[tl[k]], precomputed)
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
    int32_t  n_perms;
    int32_t  sp[kMaxTileBits];    // global row-bit position of tile-local bit l, ascending
    int32_t  outer[kMaxRowBits];  // global positions of the enumerated outer bits, ascending
    int32_t  dense_k;             // 0: interpreted sweep.  k >= 3: ONE dense 2^k x 2^k block (dense_kernel<k>)
    int32_t  pad2[2];
    // interpreted sweep: followed by DevOp ops[n_ops]; DevStage stages[n_stages]; DevPerm perms[n_perms];
    //                    double pool[pool_doubles]   (ops first: their shared-memory address is o * 96 + a constant)
    // dense sweep:       followed by int32 block_local[8] (tile-local position of block bit i = bit i of the
    //                    component index); then the REAL form of the block matrix,
    //                    A[(2c'+p')][(2c+p)] with p = 0 re / 1 im, 2^(k+1) rows of dense_lda(k) doubles
};
constexpr int kMaxDenseK = 6;
QSX_HD constexpr int dense_lda(int k) { return (2 << k) + 2; }  // +2 doubles: conflict-free A fragments
static_assert(sizeof(DevSweep) % 16 == 0, "DevSweep layout");

}  // namespace qsx
