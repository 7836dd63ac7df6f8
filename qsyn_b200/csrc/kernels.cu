// kernels.cu -- hand-written sm_100a kernels of the tensor-verification hot path.
//
//   sweep_kernel<R>   K1: in-place strided gate-apply.  One launch = one read+write pass over
//                     the shard.  A CTA owns a tile of 2^m rows x W columns; each thread holds
//                     2^R rows x 1 column in registers and interprets the sweep's op list on
//                     them; between stages the tile is transposed through shared memory so a
//                     different set of R row bits becomes register-resident.  Every global and
//                     shared access is a 128-bit (one complex128) access, and adjacent lanes
//                     take adjacent columns, so accesses are coalesced for ANY row bit
//                     (replaces xt::linalg::tensordot, tensor.hpp:418).
//   set_identity      K2: one write pass (replaces qcir_to_tensor.cpp:154-164).
//   equiv_*           K4: fused 8-sum reduction with compensated accumulation
//                     (replaces the lazy xt::sum loops of tensor.hpp:393-403, qtensor.hpp:383-385).
//   adjoint_inplace   K5: tiled conjugate transpose (tensor.hpp:303-306).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "kernels.hpp"

namespace qsx {

namespace {


// ---------------------------------------------------------------------------------------
// complex helpers (explicit FMA forms so the arithmetic does not depend on contraction flags)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// a*b + c
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// In-place exchange of two register-resident elements, opaque to the compiler.  Written as
// PTX on purpose: a C++ swap lets the register allocator "rename" the two values inside the
// swapping branch of the interpreter switch and then repair the naming on every other path with
// full copies of the 2^R-element register tile (measured: 65 % of all executed instructions
// were such MOVs).  An opaque in-place swap keeps every element pinned to its registers.
__device__ __forceinline__ void swap_inplace(double2& a, double2& b, double nz) {
    // Moves go through the FP64 pipe as x + (-0.0), which is exact for every x (nz comes from the
    // op record so that it stays opaque): 6 DADD instead of 12 32-bit MOVs per exchanged pair.
    // Measured at n = 12: in-register X 5.3 us with MOVs, 4.2 us half/half, 3.9 us all DADD.
    asm volatile(
        "{\n\t.reg .f64 t;\n\t"
        "add.rn.f64 t, %0, %4; add.rn.f64 %0, %2, %4; add.rn.f64 %2, t, %4;\n\t"
        "add.rn.f64 t, %1, %4; add.rn.f64 %1, %3, %4; add.rn.f64 %3, t, %4;\n\t}"
        : "+d"(a.x), "+d"(a.y), "+d"(b.x), "+d"(b.y)
        : "d"(nz));
}

// In-place arithmetic on register-resident elements, also PTX with tied ("+d") operands.  With
// C++ value semantics every handler of the interpreter switch defines NEW values for e[i] while
// the old ones are still live; the phi elimination at the head of the op loop then copies half of
// the register tile on EVERY micro-op (measured: 32 IMAD.MOV per op, 30 % of all executed
// instructions).  Tied operands keep each element in one virtual register for the whole stage.
// The operation order (and therefore every rounding) is the one of cmul()/cfma() above.
__device__ __forceinline__ void cmul_inplace(double2& a, double wx, double wy) {
    asm("{\n\t.reg .f64 t, u;\n\t"
        "mul.f64 t, %1, %3;\n\t"
        "neg.f64 t, t;\n\t"
        "mul.f64 u, %1, %2;\n\t"
        "fma.rn.f64 %1, %0, %3, u;\n\t"
        "fma.rn.f64 %0, %0, %2, t;\n\t}"
        : "+d"(a.x), "+d"(a.y)
        : "d"(wx), "d"(wy));
}
__device__ __forceinline__ void cmul_inplace(double2& a, double2 w) { cmul_inplace(a, w.x, w.y); }

// (a, b) <- (s*(a+b), s*(a-b)) as DMUL + 2 DFMA per component
__device__ __forceinline__ void hadamard_inplace(double2& a, double2& b, double s) {
    asm("{\n\t.reg .f64 t, n;\n\t"
        "mul.f64 t, %2, %4;\n\t"
        "neg.f64 n, t;\n\t"
        "fma.rn.f64 %2, %4, %0, n;\n\t"
        "fma.rn.f64 %0, %4, %0, t;\n\t"
        "mul.f64 t, %3, %4;\n\t"
        "neg.f64 n, t;\n\t"
        "fma.rn.f64 %3, %4, %1, n;\n\t"
        "fma.rn.f64 %1, %4, %1, t;\n\t}"
        : "+d"(a.x), "+d"(a.y), "+d"(b.x), "+d"(b.y)
        : "d"(s));
}

// (a, b) <- (m00 a + m01 b, m10 a + m11 b), m = {m00, m01, m10, m11} as re/im pairs
__device__ __forceinline__ void u1_inplace(double2& a, double2& b, const double* __restrict__ m) {
    asm("{\n\t.reg .f64 t, p, q, r, s;\n\t"
        "mul.f64 t, %5, %1;\n\t"  // cmul(m00, a)
        "neg.f64 t, t;\n\t"
        "fma.rn.f64 p, %4, %0, t;\n\t"
        "mul.f64 t, %5, %0;\n\t"
        "fma.rn.f64 q, %4, %1, t;\n\t"
        "mul.f64 t, %9, %1;\n\t"  // cmul(m10, a)
        "neg.f64 t, t;\n\t"
        "fma.rn.f64 r, %8, %0, t;\n\t"
        "mul.f64 t, %9, %0;\n\t"
        "fma.rn.f64 s, %8, %1, t;\n\t"
        "neg.f64 t, %7;\n\t"  // cfma(m01, b, .)
        "fma.rn.f64 p, t, %3, p;\n\t"
        "fma.rn.f64 %0, %6, %2, p;\n\t"
        "fma.rn.f64 q, %7, %2, q;\n\t"
        "fma.rn.f64 %1, %6, %3, q;\n\t"
        "neg.f64 t, %11;\n\t"  // cfma(m11, b, .)
        "fma.rn.f64 r, t, %3, r;\n\t"
        "fma.rn.f64 r, %10, %2, r;\n\t"
        "fma.rn.f64 s, %11, %2, s;\n\t"
        "fma.rn.f64 %3, %10, %3, s;\n\t"
        "mov.f64 %2, r;\n\t}"
        : "+d"(a.x), "+d"(a.y), "+d"(b.x), "+d"(b.y)
        : "d"(m[0]), "d"(m[1]), "d"(m[2]), "d"(m[3]), "d"(m[4]), "d"(m[5]), "d"(m[6]), "d"(m[7]));
}

// ---------------------------------------------------------------------------------------
// register-level micro-ops.  E = 2^R elements per thread.  `creg` is a required-one mask over
// the register index; it is the same for every thread of the grid, so tests on it are uniform
// branches.  ALL = true is the common uncontrolled case: no tests, no selects.
// ---------------------------------------------------------------------------------------
template <int R, int J, bool ALL>
__device__ __forceinline__ void op_u1(double2 (&e)[1 << R], const double* __restrict__ c, uint32_t creg) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        if ((i >> J) & 1) continue;
        if constexpr (!ALL) {
            if ((i & creg) != creg) continue;
        }
        u1_inplace(e[i], e[i | (1 << J)], c);
    }
}

template <int R, int J, bool ALL>
__device__ __forceinline__ void op_h(double2 (&e)[1 << R], double s, uint32_t creg) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        if ((i >> J) & 1) continue;
        if constexpr (!ALL) {
            if ((i & creg) != creg) continue;
        }
        // 3 FP64 instructions per component (DMUL + 2 DFMA) instead of 4 (2 DADD + 2 DMUL)
        hadamard_inplace(e[i], e[i | (1 << J)], s);
    }
}

template <int R, int J, bool ALL>
__device__ __forceinline__ void op_x(double2 (&e)[1 << R], uint32_t creg, double nz) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        if ((i >> J) & 1) continue;
        if constexpr (!ALL) {
            if ((i & creg) != creg) continue;
        }
        swap_inplace(e[i], e[i | (1 << J)], nz);
    }
}

// X on register bit J controlled by register bit JC: the pairs are known at compile time
template <int R, int J, int JC>
__device__ __forceinline__ void op_xc(double2 (&e)[1 << R], double nz) {
    if constexpr (J != JC) {
#pragma unroll
        for (int i = 0; i < (1 << R); ++i) {
            if (((i >> J) & 1) || !((i >> JC) & 1)) continue;
            swap_inplace(e[i], e[i | (1 << J)], nz);
        }
    }
}

template <int R, int J, bool ALL>
__device__ __forceinline__ void op_ad(double2 (&e)[1 << R], const double* __restrict__ c, uint32_t creg) {
    double2 const m01 = make_double2(c[0], c[1]), m10 = make_double2(c[2], c[3]);
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        if ((i >> J) & 1) continue;
        if constexpr (!ALL) {
            if ((i & creg) != creg) continue;
        }
        double2 const a = e[i], b = e[i | (1 << J)];
        e[i]            = cmul(m01, b);
        e[i | (1 << J)] = cmul(m10, a);
    }
}

template <int R, int JA, int JB, bool ALL>
__device__ __forceinline__ void op_swap(double2 (&e)[1 << R], uint32_t creg, double nz) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        if (!((i >> JA) & 1) || ((i >> JB) & 1)) continue;
        if constexpr (!ALL) {
            if ((i & creg) != creg) continue;  // controls never include the swapped bits
        }
        constexpr int flip = (1 << JA) | (1 << JB);
        swap_inplace(e[i], e[i ^ flip], nz);
    }
}

template <int R>
__device__ __forceinline__ void op_u2(double2 (&e)[1 << R], const double* __restrict__ m, uint32_t creg) {
    if constexpr (R >= 2) {
#pragma unroll
        for (int g = 0; g < (1 << R); g += 4) {
            if ((g & creg) != creg) continue;
            double2 v[4], w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = e[g + k];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double2 const mk = make_double2(m[2 * (4 * r + k)], m[2 * (4 * r + k) + 1]);
                    acc              = cfma(mk, v[k], acc);
                }
                w[r] = acc;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) e[g + k] = w[k];
        }
    }
}

// merged diagonal: the 2^(R-1) elements with register bit J set
template <int R, int J>
__device__ __forceinline__ void op_dtab_half(double2 (&e)[1 << R], const double2* __restrict__ tab) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        if (!((i >> J) & 1)) continue;
        constexpr int lowmask = (1 << J) - 1;
        int const k           = ((i >> (J + 1)) << J) | (i & lowmask);
        cmul_inplace(e[i], tab[k]);
    }
}

// The handlers almost every sweep uses.  sweep_kernel<R, kLean> contains only these: the
// kernel is an interpreter whose hot loop jumps between handler bodies, and the full set is
// ~160 KB of SASS -- far beyond the instruction caches ("no_instruction" was the third largest
// stall reason).  A sweep that needs any other handler runs in the kFull instantiation.
__host__ __device__ constexpr bool is_common_handler(int h) { return h >= 0 && h < H_COMMON; }

// Three instantiations of the interpreter (template parameter MODE):
//   kFull  every handler;
//   kLean  the common handlers, 128 registers (R = 4: 4 CTAs of 128 threads per SM);
//   kLight the common handlers WITHOUT the general 2x2 (the only common handler that needs more
//          than ~96 registers: 16 live elements + 8 matrix entries + products), compiled for 96
//          registers = 5 CTAs (20 warps) per SM.  The interpreter is latency-bound (dispatch
//          chain, CTA ramp), so the extra warps are worth -5 % on Clifford+T sweeps, while the
//          general 2x2 spills at 96 registers (+33 % on rotation circuits): sweeps that contain
//          one stay in kLean.  Only CTAs of <= 128 threads (the default tile) can run kLight.
constexpr int kFull = 0, kLean = 1, kLight = 2;
__host__ __device__ constexpr bool is_light_handler(int h) { return is_common_handler(h) && !(h >= H_U1 && h < H_U1 + kJ); }
__host__ __device__ constexpr int mode_threads_log2(int R, int mode) { return mode == kLight ? 7 : max_threads_log2(R); }
__host__ __device__ constexpr int mode_min_ctas(int R, int mode) { return mode == kLight ? 5 : min_ctas_per_sm(R); }

// One handler = one straight-line body; everything about it is known at compile time.
template <int R, int H>
__device__ __forceinline__ void run_handler(double2 (&e)[1 << R], const DevOp* __restrict__ op, uint32_t creg,
                                            int pool_off, int sel, uint32_t grow, const double* __restrict__ pool) {
    const double* const c = op->c;
    if constexpr ((H >= H_H && H < H_H + kJ) || (H >= H_HG && H < H_HG + kJ)) {
        constexpr bool all = H < H_H + kJ;
        constexpr int J    = all ? H - H_H : H - H_HG;
        if constexpr (J < R) op_h<R, J, all>(e, c[0], creg);
    } else if constexpr ((H >= H_X && H < H_X + kJ) || (H >= H_XG && H < H_XG + kJ)) {
        constexpr bool all = H < H_X + kJ;
        constexpr int J    = all ? H - H_X : H - H_XG;
        if constexpr (J < R) op_x<R, J, all>(e, creg, c[7]);
    } else if constexpr (H >= H_XC && H < H_XC + kJ * kJ) {
        constexpr int J = (H - H_XC) / kJ, JC = (H - H_XC) % kJ;
        if constexpr (J < R && JC < R && J != JC) op_xc<R, J, JC>(e, c[7]);
    } else if constexpr ((H >= H_U1 && H < H_U1 + kJ) || (H >= H_U1G && H < H_U1G + kJ)) {
        constexpr bool all = H < H_U1 + kJ;
        constexpr int J    = all ? H - H_U1 : H - H_U1G;
        if constexpr (J < R) op_u1<R, J, all>(e, c, creg);
    } else if constexpr (H >= H_AD && H < H_ADG + kJ) {
        constexpr int J = (H - H_AD) % kJ;
        if constexpr (J < R) op_ad<R, J, (H < H_ADG)>(e, c, creg);
    } else if constexpr (H >= H_SWAP && H < H_SWAPG + kJ * kJ) {
        constexpr int JA = ((H - H_SWAP) % (kJ * kJ)) / kJ, JB = (H - H_SWAP) % kJ;
        if constexpr (JA < R && JB < JA) op_swap<R, JA, JB, (H < H_SWAPG)>(e, creg, c[7]);
    } else if constexpr (H == H_U2) {
        op_u2<R>(e, pool + pool_off, creg);
    } else if constexpr (H >= H_DTABH && H < H_DTABH + kJ) {
        constexpr int J = H - H_DTABH;
        if constexpr (J < R) op_dtab_half<R, J>(e, reinterpret_cast<const double2*>(pool + pool_off));
    } else if constexpr (H >= H_PHJ && H < H_PHJ + kJ) {
        constexpr int J = H - H_PHJ;
        if constexpr (J < R) {
            double2 const w = make_double2(c[0], c[1]);
#pragma unroll
            for (int i = 0; i < (1 << R); ++i)
                if ((i >> J) & 1) cmul_inplace(e[i], w);
        }
    } else if constexpr (H == H_DTABF) {
        const double2* const tab = reinterpret_cast<const double2*>(pool + pool_off);
#pragma unroll
        for (int i = 0; i < (1 << R); ++i) cmul_inplace(e[i], tab[i]);
    } else if constexpr (H == H_SCALN) {
        // every per-thread scalar of the stage, without a dispatch per gate
        const uint4* const ent = reinterpret_cast<const uint4*>(pool + pool_off);
        int const count        = op->j0;
        double2 sv             = make_double2(1.0, 0.0);
        for (int k = 0; k < count; ++k) {
            uint4 const w    = ent[3 * k];  // cother, czero, sel, pad
            bool const pass  = (grow & w.x) == w.x && (grow & w.y) == 0u;
            bool const hi    = (int)w.z >= 0 && ((grow >> (w.z & 31u)) & 1u);
            double2 const d  = *reinterpret_cast<const double2*>(ent + 3 * k + (hi ? 2 : 1));
            double2 const nv = cmul(sv, d);
            sv               = pass ? nv : sv;
        }
#pragma unroll
        for (int i = 0; i < (1 << R); ++i) cmul_inplace(e[i], sv);
    }
}

#define QSX_CASE(h)                                                                                   \
    case h:                                                                                           \
        if constexpr (MODE != kLight || is_light_handler(h)) run_handler<R, h>(e, op, creg, pool_off, sel, grow, pool); \
        break;
#define QSX_CASE8(b) QSX_CASE(b) QSX_CASE(b + 1) QSX_CASE(b + 2) QSX_CASE(b + 3) QSX_CASE(b + 4) QSX_CASE(b + 5) QSX_CASE(b + 6) QSX_CASE(b + 7)
#define QSX_CASE16(b) QSX_CASE8(b) QSX_CASE8(b + 8)

// hdr = {handler, cother, czero, pool}: the first 16 bytes of the op, loaded by the caller
template <int R, int MODE>
__device__ __forceinline__ void apply_op(double2 (&e)[1 << R], const DevOp* __restrict__ op, int4 hdr, uint32_t grow,
                                         const double* __restrict__ pool) {
    uint32_t const cother = (uint32_t)hdr.y, czero = (uint32_t)hdr.z;
    if ((grow & cother) != cother || (grow & czero) != 0u) return;
    int const pool_off  = hdr.w;
    uint32_t const creg = op->creg;  // (these two are loaded only inside the handlers that read them)
    int const sel       = op->sel;
    if constexpr (MODE != kFull) {
        switch (hdr.x) {  // dense 0..H_COMMON-1
            QSX_CASE16(0)
            QSX_CASE16(16)
            QSX_CASE16(32)
            QSX_CASE(48)
            QSX_CASE(49)
            QSX_CASE(50)
            QSX_CASE(51)
            default: break;
        }
    } else {
        switch (hdr.x) {
            QSX_CASE16(0)
            QSX_CASE16(16)
            QSX_CASE16(32)
            QSX_CASE16(48)
            QSX_CASE16(64)
            QSX_CASE16(80)
            QSX_CASE16(96)
            QSX_CASE16(112)
            default: break;
        }
    }
}
#undef QSX_CASE16
#undef QSX_CASE8
#undef QSX_CASE

// ---------------------------------------------------------------------------------------
// the sweep interpreter
// ---------------------------------------------------------------------------------------
// Address permutation of a stage boundary (DevPerm list), evaluated on tile-local row indices
// (boundaries to shared memory) or on global row indices (the load and store boundaries, whose
// steps the scheduler emits in global row bits with cmo = 0).
// rows[0] is the thread's base row (register bits clear), rows[1+j] = base ^ (register bit j).
// FORWARD applies the steps in list order (store side), !FORWARD in reverse (load side: every
// step is an involution, so the reversed list is the inverse permutation).
template <int R, bool FORWARD>
__device__ __forceinline__ void permute_rows(uint32_t (&rows)[R + 1], const DevPerm* __restrict__ perms, int begin, int end,
                                             uint32_t row_outer) {
    for (int k = 0; k < end - begin; ++k) {
        int const idx = FORWARD ? begin + k : end - 1 - k;
        uint4 const p = *reinterpret_cast<const uint4*>(perms + idx);  // cml, cmo, tml, pad
        if ((row_outer & p.y) != p.y) continue;                        // uniform per CTA
#pragma unroll
        for (int v = 0; v <= R; ++v) rows[v] ^= ((rows[v] & p.x) == p.x) ? p.z : 0u;
    }
}

// What the first stage's global load needs, passed BY VALUE (kernel parameters live in the constant bank: no
// memory latency), so that a CTA's tile loads are in flight while its program is still being staged.  The
// interpreter is latency-bound per CTA; this takes the program's L2 round trip off the path to the first load.
// (sizes overridable at build time for A/B runs: make variant NAME=wide DEFS="-DQSX_EARLY_PERMS=16 -DQSX_EARLY_ORUNS=8")
#ifndef QSX_EARLY_PERMS
#define QSX_EARLY_PERMS 8
#endif
#ifndef QSX_EARLY_ORUNS
#define QSX_EARLY_ORUNS 4
#endif
constexpr int kEarlyPerms = QSX_EARLY_PERMS, kEarlyRuns = 4, kEarlyORuns = QSX_EARLY_ORUNS;
struct FirstLoad {
    uint32_t fixed_one;
    int32_t  log2w;
    int32_t  n_pre;                     // folded permutation steps of the load boundary; < 0: take the staged path
    // bit deposits as runs of consecutive bits: row |= (x & mask[k]) << shift[k]; x = tile row index / thread row index
    uint32_t omask[kEarlyORuns], oshift[kEarlyORuns];
    uint32_t tmask[kEarlyRuns], tshift[kEarlyRuns];
    int32_t  rg[kMaxRegBits];
    // the steps in the order the LOAD side applies them (reversed list), global row bits, no outer condition
    uint32_t pre_cml[kEarlyPerms], pre_tml[kEarlyPerms];
};

// IDENT: the unitary is (lazily) the identity -- the first stage synthesises it instead of reading it.
template <int R, int MODE, bool IDENT>
__global__ void __launch_bounds__(1 << mode_threads_log2(R, MODE), mode_min_ctas(R, MODE))
    sweep_kernel(double2* __restrict__ u, const unsigned char* __restrict__ gprog, uint32_t prog_bytes, int log2_ncols,
                 int log2_ncb, uint32_t flags, uint32_t col0, const __grid_constant__ FirstLoad fl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 e[1 << R];
    bool const early = !IDENT && fl.n_pre >= 0;  // uniform
    if (early) {
        // the first stage's loads, addressed from the constant bank alone
        uint32_t const cw0   = threadIdx.x & ((1u << fl.log2w) - 1u);
        uint32_t const trow0 = threadIdx.x >> fl.log2w;
        uint64_t const t64   = (flags & 1u) ? (uint64_t)(gridDim.x - 1u - blockIdx.x) : (uint64_t)blockIdx.x;
        uint64_t const cb0   = t64 & ((1ull << log2_ncb) - 1ull);
        uint32_t const rc0   = (uint32_t)(t64 >> log2_ncb);
        uint32_t g0          = fl.fixed_one;
        if constexpr (kEarlyORuns == kEarlyRuns) {
#pragma unroll
            for (int k = 0; k < kEarlyRuns; ++k) g0 |= ((rc0 & fl.omask[k]) << fl.oshift[k]) | ((trow0 & fl.tmask[k]) << fl.tshift[k]);
        } else {
#pragma unroll
            for (int k = 0; k < kEarlyORuns; ++k) g0 |= (rc0 & fl.omask[k]) << fl.oshift[k];
#pragma unroll
            for (int k = 0; k < kEarlyRuns; ++k) g0 |= (trow0 & fl.tmask[k]) << fl.tshift[k];
        }
        uint32_t rows[R + 1];
        rows[0] = g0;
#pragma unroll
        for (int j = 0; j < R; ++j) rows[1 + j] = g0 | (1u << fl.rg[j]);
#pragma unroll
        for (int k = 0; k < kEarlyPerms; ++k) {
            if (k >= fl.n_pre) break;  // uniform
            uint32_t const cml = fl.pre_cml[k], tml = fl.pre_tml[k];
#pragma unroll
            for (int v = 0; v <= R; ++v) rows[v] ^= ((rows[v] & cml) == cml) ? tml : 0u;
        }
        const double2* const gp = u + ((cb0 << fl.log2w) + cw0);
#pragma unroll
        for (int i = 0; i < (1 << R); ++i) {
            uint32_t r = rows[0];
#pragma unroll
            for (int j = 0; j < R; ++j)
                if ((i >> j) & 1) r ^= rows[1 + j] ^ rows[0];
            e[i] = __ldcg(gp + ((size_t)r << log2_ncols));
        }
    }
    // the sweep program (a few KB, read by every thread for every op) is staged in shared
    // memory once per CTA; the tile follows it
    // (asynchronous copies: every 16-byte piece is in flight at once, ONE L2 round trip per CTA whatever the
    // program size; a register-staged copy loop was compiled to one dependent LDG -> STS round trip per iteration
    // for the usual 2..3 iterations, plus an integer division for its trip count)
    {
        uint32_t const dst = (uint32_t)__cvta_generic_to_shared(smem_raw);
#pragma unroll 1
        for (uint32_t off = threadIdx.x * 16u; off < prog_bytes; off += blockDim.x * 16u)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(gprog + off) : "memory");
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const unsigned char* const prog = smem_raw;
    double2* const tile             = reinterpret_cast<double2*>(smem_raw + prog_bytes);

    const DevSweep* const sw     = reinterpret_cast<const DevSweep*>(prog);
    // the ops come first: op o sits at a compile-time constant plus o * 96, no base register needed
    const DevOp* const ops       = reinterpret_cast<const DevOp*>(prog + sizeof(DevSweep));
    const DevStage* const stages = reinterpret_cast<const DevStage*>(ops + sw->n_ops);
    int const n_stages           = sw->n_stages;
    const DevPerm* const perms   = reinterpret_cast<const DevPerm*>(stages + n_stages);
    const double* const pool     = reinterpret_cast<const double*>(perms + sw->n_perms);

    int const log2w      = sw->log2w;
    uint32_t const tid   = threadIdx.x;
    uint32_t const cw    = tid & ((1u << log2w) - 1u);
    uint32_t const trow  = tid >> log2w;
    // odd sweeps walk the tiles backwards: what the previous sweep wrote last is still in L2
    uint64_t const tid64 = (flags & 1u) ? (uint64_t)(gridDim.x - 1u - blockIdx.x) : (uint64_t)blockIdx.x;
    uint64_t const cb    = tid64 & ((1ull << log2_ncb) - 1ull);
    uint32_t const rc    = (uint32_t)(tid64 >> log2_ncb);

    // row bits fixed by the tile id
    uint32_t row_outer = sw->fixed_one;
    {
        int const n_outer = sw->n_outer;
        for (int k = 0; k < n_outer; ++k) row_outer |= ((rc >> k) & 1u) << sw->outer[k];
    }
    uint64_t const col = (cb << log2w) + cw;

    for (int s = 0; s < n_stages; ++s) {
        const DevStage* const st = stages + s;
        // this thread's rows: tile-local index (register bits = 0) and the global row it maps to
        uint32_t lbase = 0, grow = row_outer;
        {
            // fixed trip count, independent loads (positions precomputed by the scheduler): one
            // shared-memory latency per stage instead of a chain of dependent ones
            int const n_t = st->n_t;
#pragma unroll
            for (int k = 0; k < kMaxTileBits; ++k) {
                if (k >= n_t) break;  // uniform; n_t = m - R (4 at the default tile)
                uint32_t const b = (trow >> k) & 1u;
                lbase |= b << st->tl[k];
                grow |= b << st->tg[k];
            }
        }
        int rl[R], rg[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            rl[j] = st->rl[j];
            rg[j] = st->rg[j];
        }

        if (s == 0 && early) {
            // (already in flight since before the program was staged)
        } else if (s == 0) {
            // global load; X-type ops scheduled before the first register op are folded into the addresses
            uint32_t gb = grow, gc[R];
            int const pb = st->pre_begin, pe = st->pre_end;
            if (pb == pe) {
#pragma unroll
                for (int j = 0; j < R; ++j) gc[j] = 1u << rg[j];
            } else {
                uint32_t rows[R + 1];  // global rows: this boundary's steps are in global row bits
                rows[0] = grow;
#pragma unroll
                for (int j = 0; j < R; ++j) rows[1 + j] = grow | (1u << rg[j]);
                permute_rows<R, false>(rows, perms, pb, pe, 0u);
                gb = rows[0];
#pragma unroll
                for (int j = 0; j < R; ++j) gc[j] = rows[1 + j] ^ rows[0];
            }
            const double2* const gp = u + col;
            if constexpr (IDENT) {
                uint32_t const gcol = col0 + (uint32_t)col;
#pragma unroll
                for (int i = 0; i < (1 << R); ++i) {
                    uint32_t r = gb;
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if ((i >> j) & 1) r ^= gc[j];
                    e[i] = make_double2(r == gcol ? 1.0 : 0.0, 0.0);
                }
            } else {
#pragma unroll
                for (int i = 0; i < (1 << R); ++i) {
                    uint32_t r = gb;
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if ((i >> j) & 1) r ^= gc[j];
                    e[i] = __ldcg(gp + ((size_t)r << log2_ncols));
                }
            }
        } else {
            uint32_t ss[R];
#pragma unroll
            for (int j = 0; j < R; ++j) ss[j] = (1u << rl[j]) << log2w;
            const double2* const sp_ = tile + ((lbase << log2w) + cw);
#pragma unroll
            for (int i = 0; i < (1 << R); ++i) {
                uint32_t off = 0;
#pragma unroll
                for (int j = 0; j < R; ++j)
                    if ((i >> j) & 1) off += ss[j];
                e[i] = sp_[off];
            }
            __syncthreads();  // all reads of the previous layout done before anyone overwrites it
        }

        {
            int const ob = st->op_begin, oe = st->op_end;
            // (prefetching the next header was measured: the 8 extra live registers spill, -12 %)
            for (int o = ob; o < oe; ++o) {
                int4 const hdr = *reinterpret_cast<const int4*>(ops + o);
                apply_op<R, MODE>(e, ops + o, hdr, grow, pool);
            }
        }

        // store; X-type ops scheduled after this stage's register ops are folded into the addresses
        // (global row bits at the last stage, tile-local row bits towards shared memory)
        bool const last = s == n_stages - 1;
        uint32_t rows[R + 1];
        rows[0] = last ? grow : lbase;
#pragma unroll
        for (int j = 0; j < R; ++j) rows[1 + j] = rows[0] | (1u << (last ? rg[j] : rl[j]));
        int const qb = st->post_begin, qe = st->post_end;
        if (qb != qe) permute_rows<R, true>(rows, perms, qb, qe, last ? 0u : row_outer);

        if (last) {
            // A folded permutation makes a thread load (s == 0) or store (here) rows that another
            // thread of the CTA stores / loads.  In a multi-stage sweep the transposes' barriers
            // already separate the global loads from the global stores; a one-stage sweep needs one.
            if (n_stages == 1 && (qb != qe || st->pre_begin != st->pre_end)) __syncthreads();
            uint32_t const gb = rows[0];
            uint32_t gc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) gc[j] = rows[1 + j] ^ rows[0];
            double2* const gp = u + col;
#pragma unroll
            for (int i = 0; i < (1 << R); ++i) {
                uint32_t r = gb;
#pragma unroll
                for (int j = 0; j < R; ++j)
                    if ((i >> j) & 1) r ^= gc[j];
                __stcg(gp + ((size_t)r << log2_ncols), e[i]);
            }
        } else {
            uint32_t const sb = rows[0];
            uint32_t sc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) sc[j] = rows[1 + j] ^ rows[0];
            double2* const sp_ = tile + cw;
#pragma unroll
            for (int i = 0; i < (1 << R); ++i) {
                uint32_t r = sb;
#pragma unroll
                for (int j = 0; j < R; ++j)
                    if ((i >> j) & 1) r ^= sc[j];
                sp_[r << log2w] = e[i];
            }
            __syncthreads();
        }
    }
}

template <int R, int MODE, bool IDENT = false>
cudaError_t launch_sweep_r(double2* u, int n, uint64_t ncols, uint64_t grid_ncols, const DevSweep& hs, const unsigned char* prog,
                           size_t prog_bytes, uint32_t flags, uint32_t col0, cudaStream_t s) {
    (void)n;
    int const m        = hs.m;
    int const threads  = 1 << (m - R + hs.log2w);
    size_t const smem  = prog_bytes + (hs.n_stages > 1 ? ((size_t)16 << (m + hs.log2w)) : 0);
    if (prog_bytes % 16 != 0 || smem > 200 * 1024) return cudaErrorInvalidValue;
    int log2_ncols     = 0;
    while ((1ull << log2_ncols) < ncols) ++log2_ncols;
    int log2_gcols     = 0;  // the grid covers grid_ncols columns starting at u (one L2 block, or the whole shard)
    while ((1ull << log2_gcols) < grid_ncols) ++log2_gcols;
    int const log2_ncb = log2_gcols - hs.log2w;
    uint64_t const nblocks = 1ull << (log2_ncb + hs.n_outer);
    if (threads < 1 || threads > (1 << mode_threads_log2(R, MODE)) || nblocks > 0x7fffffffull)
        return cudaErrorInvalidConfiguration;
    static bool attr_set[64] = {};  // per template instantiation, per device
    int dev                  = 0;
    if (cudaError_t const rc = cudaGetDevice(&dev); rc != cudaSuccess) return rc;
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t const rc =
            cudaFuncSetAttribute(sweep_kernel<R, MODE, IDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (rc != cudaSuccess) return rc;
        attr_set[dev] = true;
    }
    FirstLoad fl{};
    fl.n_pre = -1;
    if (!IDENT && hs.n_stages >= 1) {
        const unsigned char* const host = reinterpret_cast<const unsigned char*>(&hs);
        const DevOp* const ops          = reinterpret_cast<const DevOp*>(host + sizeof(DevSweep));
        const DevStage* const st0       = reinterpret_cast<const DevStage*>(ops + hs.n_ops);
        const DevPerm* const perms      = reinterpret_cast<const DevPerm*>(st0 + hs.n_stages);
        int const np                    = st0->pre_end - st0->pre_begin;
        bool ok                         = np >= 0 && np <= kEarlyPerms && st0->n_t >= 0 && st0->n_t <= kMaxTileBits;
        for (int k = 0; ok && k < np; ++k) {
            DevPerm const& pm = perms[st0->pre_end - 1 - k];  // load side: reversed
            ok                = pm.cmo == 0;
            fl.pre_cml[k] = pm.cml, fl.pre_tml[k] = pm.tml;
        }
        // source bit k -> row bit pos[k] (ascending) as runs of equal shift
        auto runs = [](const int32_t* pos, int count, uint32_t* mask, uint32_t* shift, int max_runs) {
            int n_runs = 0;
            for (int k = 0; k < count; ++k) {
                if (pos[k] < k || pos[k] >= 32) return false;
                uint32_t const sh = (uint32_t)(pos[k] - k);
                if (n_runs == 0 || shift[n_runs - 1] != sh) {
                    if (n_runs == max_runs) return false;
                    mask[n_runs] = 0u, shift[n_runs] = sh, ++n_runs;
                }
                mask[n_runs - 1] |= 1u << k;
            }
            return true;
        };
        ok = ok && hs.n_outer >= 0 && hs.n_outer <= kMaxRowBits && runs(hs.outer, hs.n_outer, fl.omask, fl.oshift, kEarlyORuns) &&
             runs(st0->tg, st0->n_t, fl.tmask, fl.tshift, kEarlyRuns);
        if (ok) {
            fl.fixed_one = hs.fixed_one, fl.log2w = hs.log2w, fl.n_pre = np;
            for (int k = 0; k < kMaxRegBits; ++k) fl.rg[k] = st0->rg[k];
        }
    }
#ifdef QSX_NO_EARLY_LOAD
    fl.n_pre = -1;
#endif
    sweep_kernel<R, MODE, IDENT><<<(unsigned)nblocks, threads, smem, s>>>(u, prog, (uint32_t)prog_bytes, log2_ncols, log2_ncb, flags, col0, fl);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// dense_kernel<K>: one dense 2^K x 2^K complex block on the FP64 tensor pipe (K1f), K = 3..6.
//
// A CTA works on tiles of 2^m rows x 8 columns (same tiles as sweep_kernel).  The block's K row bits are tile bits;
// for every combination j of the other m-K tile bits the 8 columns give 8 independent complex vectors v of length
// D = 2^K.  In real form [v'_re; v'_im] = A [v_re; v_im] with A = [[Gr, -Gi], [Gi, Gr]] interleaved per component,
// so one j is a (2D x 2D) x (2D x 8) real GEMM, computed with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4):
//   B fragment (k = lane%4, n = lane/4): component k of column n -> one 8-byte LDS from the input tile
//   A fragment (row = lane/4, k = lane%4): REGISTER resident for the whole launch
//   C fragment (row = lane/4, n = 2*(lane%4)+{0,1}): written to the output tile
// Round 1 read every A fragment from shared memory for every DMMA (one LDS per DMMA: ncu showed short_scoreboard on top
// and 55 % of the DMMA peak).  Now a warp owns a SET of output row blocks for the whole launch and keeps their A
// fragments in registers (<= 64 doubles): per j it loads the B fragments once and issues RB_W x KB DMMAs on RB_W
// accumulator chains (2 DMMAs per LDS).  Because several
// warps read the same input vector, the update is out of place (input tile -> output tile in shared memory), and the
// input tile is double buffered: the cp.async loads of tile i+1 are in flight while tile i is multiplied and stored.
// Persistent CTAs, two per SM (one at K = 6).
// Flops: 8 * 2^K per element; bytes: 32 per element.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int K, int RB_W, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
    dense_kernel(double2* __restrict__ u, const unsigned char* __restrict__ gprog, int log2_ncols, int log2_ncb,
                 uint32_t n_tiles) {
    constexpr int D    = 1 << K;        // complex components per vector
    constexpr int D2   = 2 * D;         // real GEMM size
    constexpr int LDA  = dense_lda(K);  // leading dimension of A in the program blob (doubles)
    constexpr int KB   = D2 / 4;        // k-blocks
    constexpr int RB   = D2 / 8;        // row-blocks
    constexpr int SETS = RB / RB_W;     // warps that share one input vector
    constexpr int NW   = 8;             // warps per CTA
    static_assert(RB % RB_W == 0 && NW % SETS == 0, "row blocks must split evenly over the warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevSweep sw;
    __shared__ int32_t block_local[8];
    __shared__ uint32_t comp_row[D];   // tile-local row offset of component c
    __shared__ uint32_t jrow[64];      // tile-local base row of vector group j (block bits clear)

    if (threadIdx.x < sizeof(DevSweep) / 16)
        reinterpret_cast<uint4*>(&sw)[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(gprog) + threadIdx.x);
    if (threadIdx.x < 8) block_local[threadIdx.x] = __ldg(reinterpret_cast<const int32_t*>(gprog + sizeof(DevSweep)) + threadIdx.x);
    __syncthreads();
    int const m    = sw.m;
    int const rows = 1 << m;
    size_t const tile_elems = (size_t)8 << m;
    double2* const tin0  = reinterpret_cast<double2*>(smem_raw);
    double2* const tin1  = tin0 + tile_elems;
    double2* const tout  = tin1 + tile_elems;
    uint32_t* const growl = reinterpret_cast<uint32_t*>(tout + tile_elems);  // tile-local row -> its global tile bits
    for (int l = threadIdx.x; l < rows; l += blockDim.x) {
        uint32_t g = 0;
        for (int b = 0; b < m; ++b) g |= ((l >> b) & 1u) << sw.sp[b];
        growl[l] = g;
    }
    uint32_t block_mask = 0;
    for (int i = 0; i < K; ++i) block_mask |= 1u << block_local[i];
    if (threadIdx.x < D) {
        uint32_t off = 0;
        for (int i = 0; i < K; ++i) off |= ((threadIdx.x >> i) & 1u) << block_local[i];
        comp_row[threadIdx.x] = off;
    }
    int const n_j = rows >> K;
    if ((int)threadIdx.x < n_j) {
        uint32_t base = 0;
        int jj = threadIdx.x;
        for (int b = 0; b < m; ++b)
            if (!((block_mask >> b) & 1u)) {
                base |= (uint32_t)(jj & 1) << b;
                jj >>= 1;
            }
        jrow[threadIdx.x] = base;
    }
    __syncthreads();

    uint32_t const cw = threadIdx.x & 7u;
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int const g = lane >> 2, t = lane & 3;
    int const set = warp % SETS, jlane = warp / SETS;   // this warp: row blocks [set * RB_W, ...), vectors jlane, jlane + NW/SETS, ...

    // A fragments of this warp's row blocks: registers for the whole launch
    double afrag[RB_W][KB];
    {
        const double* const Ag = reinterpret_cast<const double*>(gprog + sizeof(DevSweep) + 32);
#pragma unroll
        for (int r = 0; r < RB_W; ++r)
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) afrag[r][kb] = __ldg(Ag + (size_t)((set * RB_W + r) * 8 + g) * LDA + 4 * kb + t);
    }
    // per-lane fragment offsets in doubles relative to the vector's base row.  Output columns are swizzled with block bit
    // 0 of the row (= bit 0 of the component): the C fragments of one store instruction cover components 4rb .. 4rb+3 x
    // columns {0,2,4,6}; unswizzled that is a 4-way bank conflict (a tile row is exactly 32 banks wide).
    uint32_t boff[KB], ooff[RB_W], ooff2[RB_W];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        int const kidx = 4 * kb + t;  // real component index = 2*c + part
        boff[kb]       = (((comp_row[kidx >> 1] << 3) + (uint32_t)g) << 1) + (kidx & 1);
    }
#pragma unroll
    for (int r = 0; r < RB_W; ++r) {
        int const ridx = (set * RB_W + r) * 8 + g;  // output real component = 2*c' + part'
        int const c    = ridx >> 1;
        ooff[r]        = (((comp_row[c] << 3) + ((uint32_t)(2 * t) ^ (uint32_t)(c & 1))) << 1) + (ridx & 1);
        ooff2[r]       = (((comp_row[c] << 3) + ((uint32_t)(2 * t + 1) ^ (uint32_t)(c & 1))) << 1) + (ridx & 1);
    }
    int const swz_bit = block_local[0];

    auto tile_origin = [&](uint32_t tile_id, uint32_t& row_outer, uint64_t& col) {
        uint64_t const cb = tile_id & ((1u << log2_ncb) - 1u);
        uint32_t const rc = tile_id >> log2_ncb;
        row_outer         = 0;
        for (int k = 0; k < sw.n_outer; ++k) row_outer |= ((rc >> k) & 1u) << sw.outer[k];
        col = (cb << 3) + cw;
    };
    auto load_tile = [&](uint32_t tile_id, double2* dst_tile) {
        uint32_t row_outer;
        uint64_t col;
        tile_origin(tile_id, row_outer, col);
        for (int l = threadIdx.x >> 3; l < rows; l += blockDim.x >> 3) {
            const double2* const src = u + ((size_t)(row_outer | growl[l]) << log2_ncols) + col;
            uint32_t const dst       = (uint32_t)__cvta_generic_to_shared(dst_tile + (l << 3) + cw);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    uint32_t tile_id = blockIdx.x;
    if (tile_id < n_tiles) load_tile(tile_id, tin0);
    int buf = 0;
    for (; tile_id < n_tiles; tile_id += gridDim.x, buf ^= 1) {
        double2* const tin      = buf ? tin1 : tin0;
        uint32_t const next_id  = tile_id + gridDim.x;
        if (next_id < n_tiles) {
            load_tile(next_id, buf ? tin0 : tin1);  // (its previous contents were consumed two barriers ago)
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();  // tile `tile_id` is complete in tin; tout is free (the previous store phase is behind a barrier)

        const double* const tin_d = reinterpret_cast<const double*>(tin);
        double* const tout_d      = reinterpret_cast<double*>(tout);
        for (int j = jlane; j < n_j; j += NW / SETS) {
            uint32_t const base = jrow[j] << 4;  // (row * 8 columns) * 2 doubles
            double bfrag[KB];
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) bfrag[kb] = tin_d[base + boff[kb]];
            // One accumulator chain per row block.  Splitting a chain into NACC partial sums (more independent DMMAs in
            // flight) was measured and is SLOWER: k = 4 / 5 / 6 at n = 12: 104 / 174 / 322 us with 1 chain, 135 / 201 / 338 us
            // with 2, 136 / 180 / 347 us with 4 -- the final DADDs share the FP64 pipe with the DMMAs, and ncu shows the pipe
            // itself as the limiter (math_pipe_throttle on top, tensor pipe 67 % active).  -DQSX_DENSE_NACC=2|4 rebuilds it.
#ifndef QSX_DENSE_NACC
#define QSX_DENSE_NACC 1
#endif
            constexpr int NACC = QSX_DENSE_NACC;
            double acc[RB_W][NACC][2];
#pragma unroll
            for (int r = 0; r < RB_W; ++r)
#pragma unroll
                for (int a = 0; a < NACC; ++a) acc[r][a][0] = acc[r][a][1] = 0.0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int r = 0; r < RB_W; ++r) dmma_m8n8k4(acc[r][kb % NACC], afrag[r][kb], bfrag[kb]);
#pragma unroll
            for (int r = 0; r < RB_W; ++r) {
                double c0 = acc[r][0][0], c1 = acc[r][0][1];
                if constexpr (NACC == 4) {
                    c0 = (acc[r][0][0] + acc[r][1][0]) + (acc[r][2][0] + acc[r][3][0]);
                    c1 = (acc[r][0][1] + acc[r][1][1]) + (acc[r][2][1] + acc[r][3][1]);
                } else if constexpr (NACC == 2) {
                    c0 = acc[r][0][0] + acc[r][1][0];
                    c1 = acc[r][0][1] + acc[r][1][1];
                }
                tout_d[base + ooff[r]]  = c0;  // column 2t
                tout_d[base + ooff2[r]] = c1;  // column 2t + 1
            }
        }
        __syncthreads();

        uint32_t row_outer;
        uint64_t col;
        tile_origin(tile_id, row_outer, col);
#pragma unroll 8
        for (int l = threadIdx.x >> 3; l < rows; l += blockDim.x >> 3)
            __stcg(u + ((size_t)(row_outer | growl[l]) << log2_ncols) + col, tout[(l << 3) + (cw ^ (((uint32_t)l >> swz_bit) & 1u))]);
        // (the next iteration's first barrier separates these reads of tout from its next writes)
    }
}

template <int K, int RB_W, int MIN_CTAS>
cudaError_t launch_dense_k(double2* u, uint64_t ncols, const DevSweep& hs, const unsigned char* prog, cudaStream_t s) {
    int log2_ncols = 0;
    while ((1ull << log2_ncols) < ncols) ++log2_ncols;
    int const log2_ncb     = log2_ncols - 3;
    uint64_t const nblocks = 1ull << (log2_ncb + hs.n_outer);
    if (log2_ncb < 0 || hs.m > 8 || hs.m < K || nblocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    // two input tiles (double buffered), one output tile, the row table
    size_t const smem = 3 * ((size_t)16 * 8 << hs.m) + sizeof(uint32_t) * ((size_t)1 << hs.m);
    static bool attr_set[64] = {};
    int dev                  = 0;
    if (cudaError_t const rc = cudaGetDevice(&dev); rc != cudaSuccess) return rc;
    if (dev < 64 && !attr_set[dev]) {
        if (cudaError_t const rc = cudaFuncSetAttribute(dense_kernel<K, RB_W, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            rc != cudaSuccess)
            return rc;
        attr_set[dev] = true;
    }
    // persistent CTAs: the A fragments are loaded into registers once per CTA, not once per tile
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned const grid = (unsigned)std::min<uint64_t>(nblocks, (uint64_t)sms * MIN_CTAS);
    dense_kernel<K, RB_W, MIN_CTAS><<<grid, 256, smem, s>>>(u, prog, log2_ncols, log2_ncb, (uint32_t)nblocks);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// identity
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) set_identity_kernel(double2* __restrict__ u, uint64_t n_elems, uint64_t col0,
                                                           int log2_ncols) {
    uint64_t const stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t const cmask  = (1ull << log2_ncols) - 1ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        uint64_t const r = i >> log2_ncols, c = i & cmask;
        u[i]             = make_double2(r == col0 + c ? 1.0 : 0.0, 0.0);
    }
}

// ---------------------------------------------------------------------------------------
// equivalence reduction
// ---------------------------------------------------------------------------------------
constexpr int kEqThreads = 256;
constexpr int kEqBlocks  = 148 * 4;
constexpr int kEqUnroll  = 4;

struct Kahan {
    double s = 0.0, c = 0.0;
    __device__ __forceinline__ void add(double x) {
        // Neumaier variant: exact for |x| > |s| as well
        double const t = s + x;
        c += (fabs(s) >= fabs(x)) ? ((s - t) + x) : ((x - t) + s);
        s = t;
    }
    __device__ __forceinline__ double value() const { return s + c; }
};

__device__ __forceinline__ void accumulate8(Kahan (&acc)[8], const double2 (&a)[kEqUnroll], const double2 (&b)[kEqUnroll]) {
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < kEqUnroll; ++k) {
        p[0] += fma(a[k].x, b[k].x, a[k].y * b[k].y);     // Re conj(a) b
        p[1] += fma(a[k].x, b[k].y, -(a[k].y * b[k].x));  // Im conj(a) b
        p[2] += fma(a[k].x, a[k].x, a[k].y * a[k].y);
        p[3] += fma(b[k].x, b[k].x, b[k].y * b[k].y);
        p[4] += a[k].x;
        p[5] += a[k].y;
        p[6] += b[k].x;
        p[7] += b[k].y;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q].add(p[q]);
}

// two-sum merge of (s,c) pairs through the block: warp shuffle then shared memory
__device__ __forceinline__ void kahan_merge(Kahan& x, double s2, double c2) {
    x.add(s2);
    x.c += c2;
}

__global__ void __launch_bounds__(kEqThreads) equiv_partial_kernel(const double2* __restrict__ a,
                                                                   const double2* __restrict__ b, uint64_t n_elems,
                                                                   double* __restrict__ partials) {
    Kahan acc[8];
    uint64_t const stride = (uint64_t)gridDim.x * kEqThreads;
    uint64_t i            = (uint64_t)blockIdx.x * kEqThreads + threadIdx.x;
    // main loop: kEqUnroll independent 128-bit loads per operand in flight
    for (; i + (kEqUnroll - 1) * stride < n_elems; i += kEqUnroll * stride) {
        double2 va[kEqUnroll], vb[kEqUnroll];
#pragma unroll
        for (int k = 0; k < kEqUnroll; ++k) {
            va[k] = __ldcs(a + i + k * stride);
            vb[k] = __ldcs(b + i + k * stride);
        }
        accumulate8(acc, va, vb);
    }
    for (; i < n_elems; i += stride) {
        double2 va[kEqUnroll], vb[kEqUnroll];
#pragma unroll
        for (int k = 0; k < kEqUnroll; ++k) va[k] = vb[k] = make_double2(0.0, 0.0);
        va[0] = a[i];
        vb[0] = b[i];
        accumulate8(acc, va, vb);
    }
    // warp level
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double const s2 = __shfl_down_sync(0xffffffffu, acc[q].s, off);
            double const c2 = __shfl_down_sync(0xffffffffu, acc[q].c, off);
            kahan_merge(acc[q], s2, c2);
        }
    }
    __shared__ double sh[kEqThreads / 32][16];
    int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            sh[warp][2 * q]     = acc[q].s;
            sh[warp][2 * q + 1] = acc[q].c;
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        int const q = threadIdx.x;
        Kahan t;
        for (int w = 0; w < kEqThreads / 32; ++w) kahan_merge(t, sh[w][2 * q], sh[w][2 * q + 1]);
        partials[(size_t)blockIdx.x * 16 + 2 * q]     = t.s;
        partials[(size_t)blockIdx.x * 16 + 2 * q + 1] = t.c;
    }
}

// deterministic finish: one block, fixed order.  Warp q reduces quantity q: lane l merges the per-block partials
// l, l + 32, ... in order, then the 32 lanes are merged by a shuffle tree (5 dependent merges instead of a chain
// of n_blocks: 2 us instead of 9 us at 592 blocks, which is 10 % of the whole check at n = 12).
__global__ void __launch_bounds__(256) equiv_finish_kernel(const double* __restrict__ partials, int n_blocks,
                                                           double* __restrict__ out8) {
    int const q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Kahan t;
    for (int blk = lane; blk < n_blocks; blk += 32)
        kahan_merge(t, partials[(size_t)blk * 16 + 2 * q], partials[(size_t)blk * 16 + 2 * q + 1]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double const s2 = __shfl_down_sync(0xffffffffu, t.s, off);
        double const c2 = __shfl_down_sync(0xffffffffu, t.c, off);
        kahan_merge(t, s2, c2);
    }
    if (lane == 0) out8[q] = t.value();
}

// ---------------------------------------------------------------------------------------
// in-place conjugate transpose (full square matrix), 32x32 tiles, pairs of tiles swapped
// ---------------------------------------------------------------------------------------
constexpr int kTr = 32;
__global__ void __launch_bounds__(kTr * 8) adjoint_inplace_kernel(double2* __restrict__ u, uint64_t dim) {
    __shared__ double2 ta[kTr][kTr + 1];
    __shared__ double2 tb[kTr][kTr + 1];
    uint32_t const bx = blockIdx.x, by = blockIdx.y;
    if (bx > by) return;  // upper triangle of tile pairs (incl. diagonal)
    uint64_t const r0 = (uint64_t)by * kTr, c0 = (uint64_t)bx * kTr;
    int const tx = threadIdx.x & (kTr - 1), ty = threadIdx.x / kTr;  // 8 row groups
    for (int r = ty; r < kTr; r += 8) {
        if (r0 + r < dim && c0 + tx < dim) ta[r][tx] = u[(r0 + r) * dim + c0 + tx];
        if (c0 + r < dim && r0 + tx < dim) tb[r][tx] = u[(c0 + r) * dim + r0 + tx];
    }
    __syncthreads();
    for (int r = ty; r < kTr; r += 8) {
        // block (by,bx) <- conj(transpose(block (bx,by)))
        if (r0 + r < dim && c0 + tx < dim) {
            double2 const v           = tb[tx][r];
            u[(r0 + r) * dim + c0 + tx] = make_double2(v.x, -v.y);
        }
        if (bx != by && c0 + r < dim && r0 + tx < dim) {
            double2 const v           = ta[tx][r];
            u[(c0 + r) * dim + r0 + tx] = make_double2(v.x, -v.y);
        }
    }
}

// Off-diagonal block pair of a column-sharded matrix: A is a c x c row-major block of one shard,
// B the mirrored block of another shard (possibly in a PEER device's memory, read and written
// over NVLink): A <- conj(B^T), B <- conj(A^T), in place, every element moved exactly once.
__global__ void __launch_bounds__(kTr * 8)
    adjoint_block_pair_kernel(double2* __restrict__ a, double2* __restrict__ b, uint64_t c, uint32_t bx0, uint32_t by0) {
    __shared__ double2 ta[kTr][kTr + 1];
    __shared__ double2 tb[kTr][kTr + 1];
    // tile (r0, c0) of A, (c0, r0) of B; (bx0, by0) lets two devices split one block pair
    uint64_t const r0 = (uint64_t)(blockIdx.y + by0) * kTr, c0 = (uint64_t)(blockIdx.x + bx0) * kTr;
    int const tx = threadIdx.x & (kTr - 1), ty = threadIdx.x / kTr;
    for (int r = ty; r < kTr; r += 8) {
        if (r0 + r < c && c0 + tx < c) ta[r][tx] = a[(r0 + r) * c + c0 + tx];
        if (c0 + r < c && r0 + tx < c) tb[r][tx] = b[(c0 + r) * c + r0 + tx];
    }
    __syncthreads();
    for (int r = ty; r < kTr; r += 8) {
        if (r0 + r < c && c0 + tx < c) {
            double2 const v            = tb[tx][r];
            a[(r0 + r) * c + c0 + tx] = make_double2(v.x, -v.y);
        }
        if (c0 + r < c && r0 + tx < c) {
            double2 const v            = ta[tx][r];
            b[(c0 + r) * c + r0 + tx] = make_double2(v.x, -v.y);
        }
    }
}

}  // namespace

cudaError_t launch_set_identity(double2* u, int n, uint64_t col0, uint64_t ncols, cudaStream_t s) {
    int log2_ncols = 0;
    while ((1ull << log2_ncols) < ncols) ++log2_ncols;
    uint64_t const n_elems = (1ull << n) * ncols;
    uint64_t blocks        = (n_elems + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    set_identity_kernel<<<(unsigned)blocks, 256, 0, s>>>(u, n_elems, col0, log2_ncols);
    return cudaGetLastError();
}

cudaError_t launch_sweep(double2* u, int n, uint64_t ncols, uint64_t grid_ncols, int R, const DevSweep& hs,
                         const unsigned char* prog, size_t prog_bytes, bool reverse, bool from_identity, uint64_t col0,
                         cudaStream_t s) {
    if (grid_ncols == 0 || grid_ncols > ncols || (grid_ncols & (grid_ncols - 1)) != 0 || grid_ncols < (1ull << hs.log2w))
        return cudaErrorInvalidValue;
    uint32_t const flags = (reverse ? 1u : 0u) | (from_identity ? 2u : 0u);
    uint32_t const c0    = (uint32_t)col0;
    if (hs.dense_k != 0) {
        if (from_identity || grid_ncols != ncols) return cudaErrorInvalidValue;  // (the runtime materialises the identity first)
        switch (hs.dense_k) {
            // 2 output row blocks per warp, 2 CTAs per SM (tiles of 2^8 rows): measured against 4 row blocks per warp on one
            // CTA per SM with 2^9-row tiles: k = 3 / 4 / 5: 97 / 106 / 173 us against 106 / 133 / 194 us at n = 12
            case 3: return launch_dense_k<3, 2, 2>(u, ncols, hs, prog, s);
            case 4: return launch_dense_k<4, 2, 2>(u, ncols, hs, prog, s);
            case 5: return launch_dense_k<5, 2, 2>(u, ncols, hs, prog, s);
            case 6: return launch_dense_k<6, 2, 1>(u, ncols, hs, prog, s);
            default: return cudaErrorInvalidValue;
        }
    }
    // the smallest instantiation that has a handler for every micro-op of this sweep
    bool lean = true, light = R == 4 && (hs.m - R + hs.log2w) <= mode_threads_log2(R, kLight);
    {
        const unsigned char* const host = reinterpret_cast<const unsigned char*>(&hs);
        const DevOp* const ops = reinterpret_cast<const DevOp*>(host + sizeof(DevSweep));
        for (int i = 0; i < hs.n_ops && lean; ++i) {
            lean  = is_common_handler(ops[i].handler);
            light = light && is_light_handler(ops[i].handler);
        }
    }
    if (from_identity) {  // instantiated for the default register count only (sweep_can_start_from_identity)
        if (R != 4 || !lean) return cudaErrorInvalidValue;
        return light ? launch_sweep_r<4, kLight, true>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s)
                     : launch_sweep_r<4, kLean, true>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
    }
    if (lean && light) return launch_sweep_r<4, kLight>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
    switch (R * 2 + (lean ? 1 : 0)) {
        case 2: return launch_sweep_r<1, kFull>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 3: return launch_sweep_r<1, kLean>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 4: return launch_sweep_r<2, kFull>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 5: return launch_sweep_r<2, kLean>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 6: return launch_sweep_r<3, kFull>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 7: return launch_sweep_r<3, kLean>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 8: return launch_sweep_r<4, kFull>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 9: return launch_sweep_r<4, kLean>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 10: return launch_sweep_r<5, kFull>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        case 11: return launch_sweep_r<5, kLean>(u, n, ncols, grid_ncols, hs, prog, prog_bytes, flags, c0, s);
        default: return cudaErrorInvalidValue;
    }
}

bool sweep_geometry(uint64_t ncols, uint64_t grid_ncols, int R, const DevSweep& hs, SweepGeometry& g) {
    if (grid_ncols == 0 || grid_ncols > ncols || (grid_ncols & (grid_ncols - 1)) != 0 || grid_ncols < (1ull << hs.log2w)) return false;
    int const m = hs.m;
    if (m - R + hs.log2w < 0 || m - R + hs.log2w > 10) return false;
    g.threads    = 1u << (m - R + hs.log2w);
    g.tile_bytes = hs.n_stages > 1 ? ((size_t)16 << (m + hs.log2w)) : 0;
    g.log2_ncols = 0;
    while ((1ull << g.log2_ncols) < ncols) ++g.log2_ncols;
    int log2_gcols = 0;
    while ((1ull << log2_gcols) < grid_ncols) ++log2_gcols;
    g.log2_ncb             = log2_gcols - hs.log2w;
    uint64_t const nblocks = 1ull << (g.log2_ncb + hs.n_outer);
    if (nblocks > 0x7fffffffull || g.tile_bytes > 200 * 1024) return false;
    g.grid = (unsigned)nblocks;
    return true;
}

bool sweep_can_start_from_identity(int R, const DevSweep& hs) {
    if (R != 4 || hs.dense_k != 0 || hs.fixed_one != 0) return false;
    const unsigned char* const host = reinterpret_cast<const unsigned char*>(&hs);
    const DevOp* const ops          = reinterpret_cast<const DevOp*>(host + sizeof(DevSweep));
    for (int i = 0; i < hs.n_ops; ++i)
        if (!is_common_handler(ops[i].handler)) return false;
    return true;
}

size_t equiv_partials_doubles() { return (size_t)kEqBlocks * 16; }

cudaError_t launch_equiv_partials(const double2* a, const double2* b, uint64_t n_elems, double* partials, int* n_blocks,
                                  cudaStream_t s) {
    uint64_t blocks = (n_elems + kEqThreads - 1) / kEqThreads;
    if (blocks > (uint64_t)kEqBlocks) blocks = kEqBlocks;
    if (blocks == 0) blocks = 1;
    equiv_partial_kernel<<<(unsigned)blocks, kEqThreads, 0, s>>>(a, b, n_elems, partials);
    *n_blocks = (int)blocks;
    return cudaGetLastError();
}

cudaError_t launch_equiv_finish(const double* partials, int n_blocks, double* out8, cudaStream_t s) {
    equiv_finish_kernel<<<1, 256, 0, s>>>(partials, n_blocks, out8);
    return cudaGetLastError();
}

cudaError_t launch_adjoint_inplace(double2* u, uint64_t dim, cudaStream_t s) {
    unsigned const nb = (unsigned)((dim + kTr - 1) / kTr);
    dim3 grid(nb, nb);
    adjoint_inplace_kernel<<<grid, kTr * 8, 0, s>>>(u, dim);
    return cudaGetLastError();
}

cudaError_t launch_adjoint_block_pair(double2* a, double2* b, uint64_t c, int part, cudaStream_t s) {
    unsigned const nb = (unsigned)((c + kTr - 1) / kTr), half = nb / 2;
    // part 0: everything; part 1: tile rows [0, nb/2) of a; part 2 (called with the blocks swapped by
    // the other device): the tile COLUMNS [nb/2, nb) of its a = the rest of the pair
    if (part == 0 || half == 0) {
        if (part == 2) return cudaSuccess;
        adjoint_block_pair_kernel<<<dim3(nb, nb), kTr * 8, 0, s>>>(a, b, c, 0u, 0u);
    } else if (part == 1) {
        adjoint_block_pair_kernel<<<dim3(nb, half), kTr * 8, 0, s>>>(a, b, c, 0u, 0u);
    } else {
        adjoint_block_pair_kernel<<<dim3(nb - half, nb), kTr * 8, 0, s>>>(a, b, c, half, 0u);
    }
    return cudaGetLastError();
}

}  // namespace qsx
