// planner_lower.cpp -- lowers one reference-style gate (controls + a 2^t x 2^t target matrix) to
// register-level ops: classification into X / H / diagonal / anti-diagonal / general 2x2 / SWAP /
// dense 4x4, with optional snapping of entries that are 0 / +-1 / +-i up to rounding.
#include <algorithm>
#include <cassert>
#include <cstring>

#include "planner_internal.hpp"

namespace qsx::detail {

namespace {


struct C2 {
    double re, im;
};


bool is_zero(C2 z) { return z.re == 0.0 && z.im == 0.0; }
bool is_one(C2 z) { return z.re == 1.0 && z.im == 0.0; }


}  // namespace

// ---- lowering ---------------------------------------------------------------------------

int lower_gate(int n, const qsx_gate& g, int gate_index, double tol, std::vector<LoweredOp>& out, std::string& msg) {
    int const nc = g.n_ctrls, nt = g.n_targets;
    if (nc < 0 || nt < 1 || nc + nt > QSX_MAX_GATE_QUBITS || g.matrix == nullptr) {
        msg = "gate " + std::to_string(gate_index) + ": bad arity or null matrix";
        return QSX_ERR_BAD_ARG;
    }
    uint32_t seen = 0, req = 0;
    for (int i = 0; i < nc + nt; ++i) {
        int const q = g.qubits[i];
        if (q < 0 || q >= n) {
            msg = "gate " + std::to_string(gate_index) + ": qubit id out of range";
            return QSX_ERR_BAD_ARG;
        }
        uint32_t const b = 1u << (n - 1 - q);
        if (seen & b) {
            msg = "gate " + std::to_string(gate_index) + ": duplicate qubit id";
            return QSX_ERR_BAD_ARG;
        }
        seen |= b;
        if (i < nc) req |= b;
    }
    double const cfrac = std::ldexp(1.0, -nc);
    if (nt == 1) {
        int const tb = n - 1 - g.qubits[nc];
        C2 m[4];
        for (int i = 0; i < 4; ++i) m[i] = {snap1(g.matrix[2 * i], tol), snap1(g.matrix[2 * i + 1], tol)};
        LoweredOp op;
        op.src_gate = gate_index;
        op.req      = req;
        if (is_zero(m[1]) && is_zero(m[2])) {  // diagonal
            if (is_one(m[0]) && is_one(m[3])) return QSX_OK;  // identity: nothing to do
            if (is_one(m[0])) {
                op.req |= 1u << tb;
                op.frac = cfrac * 0.5;
                if (m[3].re == -1.0 && m[3].im == 0.0) {
                    op.kind = OP_NEG;
                } else if (m[3].re == 0.0 && (m[3].im == 1.0 || m[3].im == -1.0)) {
                    op.kind = OP_MULI;
                    op.c[0] = m[3].im;
                } else {
                    op.kind = OP_PHASE;
                    op.c[0] = m[3].re;
                    op.c[1] = m[3].im;
                }
            } else {
                op.kind = OP_DIAG;
                op.dbit = tb;
                op.frac = cfrac;
                op.c[0] = m[0].re, op.c[1] = m[0].im, op.c[2] = m[3].re, op.c[3] = m[3].im;
            }
        } else {
            op.t0   = tb;
            op.frac = cfrac;
            if (is_zero(m[0]) && is_zero(m[3])) {
                if (is_one(m[1]) && is_one(m[2])) {
                    op.kind = OP_X;
                } else {
                    op.kind = OP_AD;
                    op.c[0] = m[1].re, op.c[1] = m[1].im, op.c[2] = m[2].re, op.c[3] = m[2].im;
                }
            } else if (m[0].im == 0 && m[1].im == 0 && m[2].im == 0 && m[3].im == 0 && m[0].re == m[1].re &&
                       m[0].re == m[2].re && m[0].re == -m[3].re) {
                op.kind = OP_H;
                op.c[0] = m[0].re;
            } else {
                op.kind = OP_U1;
                for (int i = 0; i < 4; ++i) op.c[2 * i] = m[i].re, op.c[2 * i + 1] = m[i].im;
            }
        }
        out.push_back(std::move(op));
        return QSX_OK;
    }
    if (nt == 2) {
        int const b1 = n - 1 - g.qubits[nc], b0 = n - 1 - g.qubits[nc + 1];
        std::vector<double> m(32);
        for (int i = 0; i < 32; ++i) m[i] = snap1(g.matrix[i], tol);
        static double const kSwap[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
        bool is_swap = true, is_id = true;
        for (int i = 0; i < 16; ++i) {
            if (m[2 * i] != kSwap[i] || m[2 * i + 1] != 0.0) is_swap = false;
            if (m[2 * i] != ((i % 5 == 0) ? 1.0 : 0.0) || m[2 * i + 1] != 0.0) is_id = false;
        }
        if (is_id) return QSX_OK;
        LoweredOp op;
        op.src_gate = gate_index;
        op.req      = req;
        op.t0       = b1;
        op.t1       = b0;
        if (is_swap) {
            op.kind = OP_SWAP;
            op.frac = cfrac * 0.5;
        } else {
            op.kind = OP_U2;
            op.frac = cfrac;
            op.mat  = std::move(m);
        }
        out.push_back(std::move(op));
        return QSX_OK;
    }
    // >= 3 targets (a nested circuit handed over as one matrix, a user-defined k-qubit unitary): ONE dense block on the
    // FP64 tensor pipe.  Controls become further block bits: the block matrix is the identity except where every control
    // bit is 1 (qtensor.hpp:346-372: direct_sum(I, target) with the controls as most significant bits).
    int const k = nt + nc;
    if (k > kMaxDenseK) {
        msg = "gate " + std::to_string(gate_index) + ": a dense gate on " + std::to_string(k) + " qubits (targets + controls) exceeds the " +
              std::to_string(kMaxDenseK) + "-qubit dense block; decompose it on the host";
        return QSX_ERR_UNSUPPORTED;
    }
    if (n < 3) {
        msg = "gate " + std::to_string(gate_index) + ": dense blocks need at least 3 qubits";
        return QSX_ERR_UNSUPPORTED;
    }
    LoweredOp op;
    op.kind     = OP_DENSE;
    op.src_gate = gate_index;
    op.frac     = 1.0;
    // matrix-index bit i: targets first (the LAST target is the least significant bit of the gate matrix), then controls
    for (int i = 0; i < nt; ++i) op.dbits.push_back(n - 1 - g.qubits[nc + nt - 1 - i]);
    for (int i = 0; i < nc; ++i) op.dbits.push_back(n - 1 - g.qubits[i]);
    size_t const D = size_t{1} << k, Dt = size_t{1} << nt;
    op.mat.assign(2 * D * D, 0.0);
    bool identity = true;
    for (size_t r = 0; r < D; ++r)
        for (size_t c = 0; c < D; ++c) {
            double re = r == c ? 1.0 : 0.0, im = 0.0;
            if ((r >> nt) == (D >> nt) - 1 && (c >> nt) == (D >> nt) - 1) {  // every control bit set in both indices
                size_t const e = ((r & (Dt - 1)) * Dt + (c & (Dt - 1)));
                re = snap1(g.matrix[2 * e], tol), im = snap1(g.matrix[2 * e + 1], tol);
            }
            op.mat[2 * (r * D + c)] = re, op.mat[2 * (r * D + c) + 1] = im;
            if (re != (r == c ? 1.0 : 0.0) || im != 0.0) identity = false;
        }
    if (identity) return QSX_OK;
    out.push_back(std::move(op));
    return QSX_OK;
}

}  // namespace qsx::detail
