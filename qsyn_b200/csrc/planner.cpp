// planner.cpp -- lowers a reference-style gate stream to register-level ops and schedules
// them into blocked multi-gate sweeps.
//
// Replaces, on the host side, the per-gate `tensordot(gate, main)` dispatch of
// qcir_to_tensor.cpp:173-194: instead of one full-tensor contraction per gate, runs of
// gates are grouped so that ONE pass over the shard in HBM applies all of them.
//
// Scheduling rule (correctness): gate A may be moved before an earlier gate B iff
// they commute.  We use the sufficient condition "on every shared qubit both act
// diagonally" -- controls and diagonal gates are diagonal on that qubit.  With
//   T(g) = row bits g acts on non-diagonally,  D(g) = row bits g only tests,
// A and B commute when T(A) & (T(B)|D(B)) == 0 and T(B) & (T(A)|D(A)) == 0.
//
// Locality rule (performance): only the bits in T(g) have to be inside the shared-memory
// tile (and, for the stage that applies g, in registers); tested bits can be any row bit.
#include "planner.hpp"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <sstream>

namespace qsx {

namespace {

inline int popc(uint32_t x) { return __builtin_popcount(x); }

struct C2 {
    double re, im;
};

double snap1(double x, double tol) {
    if (tol < 0) return x;
    if (std::fabs(x) <= tol) return 0.0;
    if (std::fabs(x - 1.0) <= tol) return 1.0;
    if (std::fabs(x + 1.0) <= tol) return -1.0;
    return x;
}

bool is_zero(C2 z) { return z.re == 0.0 && z.im == 0.0; }
bool is_one(C2 z) { return z.re == 1.0 && z.im == 0.0; }

int ilog2_exact(uint64_t x) {
    if (x == 0 || (x & (x - 1))) return -1;
    int l = 0;
    while ((x >> l) != 1) ++l;
    return l;
}

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
    msg = "gate " + std::to_string(gate_index) + ": " + std::to_string(nt) +
          "-target dense gates are not supported (flatten nested circuits on the host)";
    return QSX_ERR_UNSUPPORTED;
}

// ---- scheduling -------------------------------------------------------------------------

struct Picker {
    uint32_t blockedT = 0, blockedA = 0;
    bool passes(LoweredOp const& op) const {
        uint32_t const T = op.tmask(), A = T | op.dmask();
        return (T & blockedA) == 0 && (A & blockedT) == 0;
    }
    void block(LoweredOp const& op) {
        blockedT |= op.tmask();
        blockedA |= op.tmask() | op.dmask();
    }
};

inline bool is_perm_op(LoweredOp const& op) { return op.kind == OP_X || op.kind == OP_SWAP; }

// X / CX / SWAP are permutations of rows.  At a stage boundary every element is written to
// (read from) an address the thread computes anyway, so such an op can be executed by
// permuting ADDRESSES instead of data, at no cost per element.  The kernel evaluates the
// boundary's permutation only at the thread's base row and at base ^ (register bit j) and
// XOR-combines the differences; that is exact iff the composed permutation is affine in the
// register bits.  `taint` = row bits whose value, after the steps accepted so far, depends on
// a register bit.  A step  t ^= AND(controls)  keeps every bit affine iff at most one of its
// controls is tainted (the others are per-thread constants); t then becomes tainted too.
bool boundary_accept(LoweredOp const& op, uint32_t& taint) {
    uint32_t tt = taint;
    auto step = [&](uint32_t ctrl, int t) {
        int const k = popc(ctrl & tt);
        if (k > 1) return false;
        if (k == 1) tt |= 1u << t;
        return true;
    };
    bool ok;
    if (op.kind == OP_SWAP) {  // CX(a,b) CX(b,a) CX(a,b), each with the explicit controls
        ok = step(op.req | (1u << op.t0), op.t1) && step(op.req | (1u << op.t1), op.t0) &&
             step(op.req | (1u << op.t0), op.t1);
    } else {
        ok = step(op.req, op.t0);
    }
    if (ok) taint = tt;
    return ok;
}

void build_stages(Plan& plan, std::vector<int> chosen, PlannedSweep& sw) {
    int const R = plan.cfg.R;
    bool first_come_once = false;  // set when a searched register set made no progress
    while (!chosen.empty()) {
        PlannedStage st;
        bool const first = sw.stages.empty();
        // load-side list of the first stage: its register bits are not chosen yet (and the list is
        // executed in reverse), so treat every bit as tainted: only steps with <= 1 control fold
        uint32_t pre_taint = ~0u;
        uint32_t rm = 0, chosenT = 0, chosenA = 0;
        int u2_t0 = -1, u2_t1 = -1;
        Picker pk;
        std::vector<int> rest;
        // Register bits of the stage: first come, first served -- or (stage_search) the R-subset of
        // the tile bits under which the stage takes the most register ops.  The trial pass ignores
        // the affinity test of folded permutations; it only ranks subsets.
        uint32_t allowed = 0;
        if (plan.cfg.stage_search && !first_come_once && (int)sw.tile_bits.size() > R) {
            uint32_t tile_mask = 0;
            for (int b : sw.tile_bits) tile_mask |= 1u << b;
            auto trial = [&](uint32_t cand) {
                Picker tp;
                uint32_t cT = 0, cA = 0;
                int score = 0;
                for (int idx : chosen) {
                    LoweredOp const& op = plan.ops[idx];
                    if (!tp.passes(op)) {
                        tp.block(op);
                        continue;
                    }
                    uint32_t const T = op.tmask(), A = T | op.dmask();
                    if (plan.cfg.fold_perms && is_perm_op(op) && (T & cA) == 0 && (A & cT) == 0 && popc(op.req) <= 1) continue;
                    if ((T & ~cand) == 0 && op.kind != OP_U2) {
                        cT |= T, cA |= A;
                        // a permutation gains nothing from sitting in registers (it may fold at the next boundary)
                        score += is_perm_op(op) || T == 0 ? 1 : 4;
                    } else {
                        tp.block(op);
                    }
                }
                return score;
            };
            int best = -1;
            for (uint32_t cand = tile_mask;; cand = (cand - 1) & tile_mask) {  // all submasks of the tile
                if (popc(cand) == R) {
                    int const sc = trial(cand);
                    if (sc > best) best = sc, allowed = cand;
                }
                if (cand == 0) break;
            }
            // two-target dense ops need the first-come rule (their pair is pinned to register bits 1, 0)
            for (int idx : chosen)
                if (plan.ops[idx].kind == OP_U2) allowed = 0;
        }
        for (int idx : chosen) {
            LoweredOp const& op = plan.ops[idx];
            if (!pk.passes(op)) {
                pk.block(op);
                rest.push_back(idx);
                continue;
            }
            uint32_t const T = op.tmask(), A = T | op.dmask();
            if (plan.cfg.fold_perms && is_perm_op(op)) {
                // commutes with everything already placed in this stage -> hoist to the boundary before it
                bool const commutes = (T & chosenA) == 0 && (A & chosenT) == 0;
                if (commutes && boundary_accept(op, first ? pre_taint : sw.stages.back().post_taint)) {
                    (first ? st.pre : sw.stages.back().post).push_back(idx);
                    ++plan.stats.n_folded_perms;
                    continue;
                }
                // not hoistable: either swap registers inside this stage, or (defer_perms) wait for
                // the next stage boundary, where it is first in line and always folds
                uint32_t probe = ~0u;
                if (plan.cfg.defer_perms && !st.ops.empty() && boundary_accept(op, probe)) {
                    pk.block(op);
                    rest.push_back(idx);
                    continue;
                }
            }
            bool ok = allowed != 0 ? (T & ~allowed) == 0 : popc(rm | T) <= R;
            if (ok && op.kind == OP_U2 && u2_t0 >= 0 && (u2_t0 != op.t0 || u2_t1 != op.t1)) ok = false;
            if (ok) {
                if (op.kind == OP_U2) u2_t0 = op.t0, u2_t1 = op.t1;
                rm |= T;
                chosenT |= T;
                chosenA |= A;
                st.ops.push_back(idx);
            } else {
                pk.block(op);
                rest.push_back(idx);
            }
        }
        if (allowed != 0 && rest.size() == chosen.size()) {  // nothing placed or folded: the ranking misled us
            first_come_once = true;
            continue;
        }
        first_come_once = false;
        chosen.swap(rest);
        if (st.ops.empty() && st.pre.empty()) continue;  // everything went to the previous boundary
        // register-bit assignment: a U2 pair must sit on register bits (1,0)
        std::vector<int> rb;
        if (u2_t0 >= 0) {
            rb.push_back(u2_t1);
            rb.push_back(u2_t0);
        }
        for (int b = 0; b < kMaxRowBits; ++b)
            if ((rm >> b) & 1u)
                if (std::find(rb.begin(), rb.end(), b) == rb.end()) rb.push_back(b);
        for (int b : sw.tile_bits) {
            if ((int)rb.size() >= R) break;
            if (std::find(rb.begin(), rb.end(), b) == rb.end()) rb.push_back(b);
        }
        st.reg_bits   = rb;
        st.post_taint = 0;
        for (int b : rb) st.post_taint |= 1u << b;
        // sink: an X-type op that had to stay in registers (something before it in this stage does
        // not commute with it) can still leave through the END of the stage if it commutes with
        // everything AFTER it.  Scan backwards so that "after" already excludes sunk ops.
        if (plan.cfg.fold_perms) {
            std::vector<int> sunk;  // collected back to front
            uint32_t laterT = 0, laterA = 0;
            for (size_t k = st.ops.size(); k-- > 0;) {
                LoweredOp const& op = plan.ops[st.ops[k]];
                uint32_t const T = op.tmask(), A = T | op.dmask();
                if (is_perm_op(op) && (T & laterA) == 0 && (A & laterT) == 0) {
                    sunk.push_back(st.ops[k]);
                    st.ops.erase(st.ops.begin() + (long)k);
                    continue;
                }
                laterT |= T;
                laterA |= A;
            }
            // `sunk` is back to front; in circuit order it is s_1 < ... < s_k.  The boundary list must
            // stay affine in the register bits, so sink the longest suffix s_j..s_k that the taint
            // check accepts and put s_1..s_{j-1} back where they were (a kept op precedes every
            // sunk one, so the order among them is unchanged).
            std::vector<int> order(sunk.rbegin(), sunk.rend());
            size_t first_sunk = order.size();
            for (size_t j = 0; j < order.size(); ++j) {
                uint32_t taint = st.post_taint;
                bool ok        = true;
                for (size_t k = j; k < order.size() && ok; ++k) ok = boundary_accept(plan.ops[order[k]], taint);
                if (ok) {
                    first_sunk = j;
                    break;
                }
            }
            for (size_t k = first_sunk; k < order.size(); ++k) {
                boundary_accept(plan.ops[order[k]], st.post_taint);
                st.post.push_back(order[k]);
                ++plan.stats.n_folded_perms;
            }
            if (first_sunk > 0) {
                // re-insert the kept ones at their original relative positions: merge by op index
                std::vector<int> merged;
                size_t a = 0;
                for (int idx : st.ops) {
                    while (a < first_sunk && order[a] < idx) merged.push_back(order[a++]);
                    merged.push_back(idx);
                }
                while (a < first_sunk) merged.push_back(order[a++]);
                st.ops.swap(merged);
            }
        }
        sw.stages.push_back(std::move(st));
    }
}

// ---- micro-op emission ------------------------------------------------------------------
//
// Non-diagonal ops map 1:1 to register micro-ops.  Diagonal ops (PHASE/NEG/MULI/DIAG) commute
// with each other and with any op that does not act non-diagonally on their bits, so they are
// held back and merged:
//   * a diagonal op whose bits are all OUTSIDE the registers multiplies every element a thread
//     holds by the same number: it becomes one complex multiply of a per-thread scalar
//     (K_DSCAL), applied to the registers once at the end of the stage (K_SAPPLY);
//   * the others are multiplied together, per distinct thread-level condition, into a table
//     indexed by the register index (K_DTABH when only elements with one register bit set are
//     affected, K_DTABF otherwise).

struct Cx {
    double re = 1.0, im = 0.0;
};
inline Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

Cx phase_of(LoweredOp const& op) {
    switch (op.kind) {
        case OP_NEG: return {-1.0, 0.0};
        case OP_MULI: return {0.0, op.c[0]};
        default: return {op.c[0], op.c[1]};
    }
}

void emit_uops(Plan const& plan, PlannedStage& st) {
    int const R = (int)st.reg_bits.size();
    int const E = 1 << R;
    int regidx_of[kMaxRowBits];
    for (int b = 0; b < kMaxRowBits; ++b) regidx_of[b] = -1;
    uint32_t regmask = 0;
    for (int j = 0; j < R; ++j) regidx_of[st.reg_bits[j]] = j, regmask |= 1u << st.reg_bits[j];
    auto creg_of = [&](uint32_t req) {
        uint32_t c = 0;
        for (int j = 0; j < R; ++j)
            if ((req >> st.reg_bits[j]) & 1u) c |= 1u << j;
        return c;
    };

    // per-thread scalars commute with every register op of the stage: they are all collected
    // into ONE micro-op at the end of the stage
    std::vector<MicroOp> scalars;

    // The diagonal ops seen so far and not yet applied, multiplied together: one table over the
    // register index per distinct thread-level condition.  A table stays live across
    //   * ops that do not act on a register bit it depends on (they commute), and
    //   * in-register permutations (X / CX / SWAP whose controls are all register bits):
    //     D then P  ==  P then D',  D'[P(i)] = D[i]  -- the table is permuted instead of applied;
    // it is applied right before an op that mixes a bit it depends on, or at the end of the stage.
    // A group is kept factored: value(i) = tab[i] * prod_j f[j][bit j of i].  Single-bit diagonals
    // (z, s, t, rz, p ...) go to the factors, so that before an op on bit j only the factor of j has
    // to be applied (one phase, K_PHJ) or absorbed (general 2x2), while the rest stays live.
    struct Group {
        uint32_t cother, czero;
        std::vector<Cx> tab;
        Cx f[kMaxRegBits][2];
    };
    std::vector<Group> groups;
    auto group_for = [&](uint32_t cother, uint32_t czero) -> Group& {
        for (Group& g : groups)
            if (g.cother == cother && g.czero == czero) return g;
        Group g;
        g.cother = cother, g.czero = czero;
        g.tab.assign((size_t)E, Cx{});
        groups.push_back(std::move(g));
        return groups.back();
    };
    auto is_unit = [](Cx z) { return z.re == 1.0 && z.im == 0.0; };
    auto same    = [](Cx a, Cx b) { return a.re == b.re && a.im == b.im; };
    // folds the factor of bit j into the table (needed before the table is permuted by a controlled op on j)
    auto merge_factor = [&](Group& g, int j) {
        if (is_unit(g.f[j][0]) && is_unit(g.f[j][1])) return;
        for (int i = 0; i < E; ++i) g.tab[i] = cmul(g.tab[i], g.f[j][(i >> j) & 1]);
        g.f[j][0] = Cx{}, g.f[j][1] = Cx{};
    };
    auto add_diagonal = [&](LoweredOp const& op) {
        uint32_t const creg   = creg_of(op.req);
        uint32_t const cother = op.req & ~regmask;
        bool const is_diag2   = op.kind == OP_DIAG;
        int const dj          = is_diag2 ? regidx_of[op.dbit] : -1;
        if (creg == 0 && dj < 0) {  // nothing in registers: per-thread scalar
            MicroOp u;
            u.kind   = K_DSCAL;
            u.cother = cother;
            if (is_diag2) {
                u.sel  = op.dbit;
                u.c[0] = op.c[0], u.c[1] = op.c[1], u.c[2] = op.c[2], u.c[3] = op.c[3];
            } else {
                Cx const w = phase_of(op);
                u.c[0] = w.re, u.c[1] = w.im;
            }
            scalars.push_back(std::move(u));
            return;
        }
        if (!is_diag2) {
            Group& g   = group_for(cother, 0);
            Cx const w = phase_of(op);
            if (popc(creg) == 1) {
                int const j = __builtin_ctz(creg);
                g.f[j][1]   = cmul(g.f[j][1], w);
            } else {
                for (int i = 0; i < E; ++i)
                    if (((uint32_t)i & creg) == creg) g.tab[i] = cmul(g.tab[i], w);
            }
        } else if (dj >= 0) {
            Group& g = group_for(cother, 0);
            Cx const d0{op.c[0], op.c[1]}, d1{op.c[2], op.c[3]};
            if (creg == 0) {
                g.f[dj][0] = cmul(g.f[dj][0], d0), g.f[dj][1] = cmul(g.f[dj][1], d1);
            } else {
                for (int i = 0; i < E; ++i)
                    if (((uint32_t)i & creg) == creg) g.tab[i] = cmul(g.tab[i], ((i >> dj) & 1) ? d1 : d0);
            }
        } else {  // selecting bit outside the registers, a control inside
            Cx const d0{op.c[0], op.c[1]}, d1{op.c[2], op.c[3]};
            Group& g1 = group_for(cother | (1u << op.dbit), 0);
            for (int i = 0; i < E; ++i)
                if (((uint32_t)i & creg) == creg) g1.tab[i] = cmul(g1.tab[i], d1);
            Group& g0 = group_for(cother, 1u << op.dbit);
            for (int i = 0; i < E; ++i)
                if (((uint32_t)i & creg) == creg) g0.tab[i] = cmul(g0.tab[i], d0);
        }
    };
    // Single-qubit fusion: the last uncontrolled non-diagonal op on register bit j that nothing
    // emitted since involves (no target, control or table dependence on j).  A later uncontrolled
    // op on j commutes back to it, so the two are multiplied on the host when one micro-op for the
    // product is cheaper than the separate ones (costs: the measured model of DESIGN.md 3.1).
    struct Fusable {
        int idx = -1;  // position in st.uops
        Cx m[4];
        double cost = 0;
    };
    Fusable fus[kMaxRegBits];
    auto table_depends_on = [&](Group const& g, int j) {
        for (int i = 0; i < E; ++i)
            if (!same(g.tab[i], g.tab[i ^ (1 << j)])) return true;
        return false;
    };
    auto depends_on = [&](Group const& g, int j) { return !same(g.f[j][0], g.f[j][1]) || table_depends_on(g, j); };
    auto emit_group = [&](Group g) {  // by value: factors are folded into the copy
        for (int j = 0; j < R; ++j) {
            if (depends_on(g, j)) fus[j].idx = -1;
            merge_factor(g, j);
        }
        uint32_t active = 0;
        for (int i = 0; i < E; ++i)
            if (g.tab[i].re != 1.0 || g.tab[i].im != 0.0) active |= 1u << i;
        if (active == 0) return;
        MicroOp u;
        u.cother = g.cother;
        u.czero  = g.czero;
        int half = -1;
        for (int j = 0; j < R && half < 0; ++j) {
            bool ok = true;
            for (int i = 0; i < E; ++i)
                if (((active >> i) & 1u) && !((i >> j) & 1)) ok = false;
            if (ok) half = j;
        }
        bool uniform = half >= 0;  // every entry of the half-table equal -> one scalar instead of a table
        for (int i = 0; uniform && i < E; ++i)
            if ((i >> half) & 1)
                uniform = g.tab[i].re == g.tab[1 << half].re && g.tab[i].im == g.tab[1 << half].im;
        if (uniform) {
            u.kind = K_PHJ;
            u.j0   = half;
            u.c[0] = g.tab[1 << half].re;
            u.c[1] = g.tab[1 << half].im;
        } else if (half >= 0) {
            u.kind = K_DTABH;
            u.j0   = half;
            for (int i = 0; i < E; ++i)
                if ((i >> half) & 1) u.table.push_back(g.tab[i].re), u.table.push_back(g.tab[i].im);
        } else {
            u.kind = K_DTABF;
            for (int i = 0; i < E; ++i) u.table.push_back(g.tab[i].re), u.table.push_back(g.tab[i].im);
        }
        st.uops.push_back(std::move(u));
    };
    // applies whatever depends on one of the register bits in jmask: the whole group when its
    // table does, else only the factors of those bits (a phase on one register bit each)
    auto flush = [&](uint32_t jmask) {
        std::vector<Group> keep;
        for (Group& g : groups) {
            bool whole = false;
            for (int j = 0; j < R && !whole; ++j)
                if ((jmask >> j) & 1u) whole = table_depends_on(g, j) || (!same(g.f[j][0], g.f[j][1]) && !is_unit(g.f[j][0]));
            if (whole) {
                emit_group(g);
                continue;
            }
            for (int j = 0; j < R; ++j) {
                if (!((jmask >> j) & 1u) || same(g.f[j][0], g.f[j][1])) continue;
                MicroOp u;  // f[j][0] == 1: a phase on the bit-j-set half
                u.kind   = K_PHJ;
                u.cother = g.cother;
                u.czero  = g.czero;
                u.j0     = j;
                u.c[0] = g.f[j][1].re, u.c[1] = g.f[j][1].im;
                st.uops.push_back(std::move(u));
                g.f[j][1]  = Cx{};
                fus[j].idx = -1;
            }
            keep.push_back(std::move(g));
        }
        groups.swap(keep);
    };
    auto conjugate_tables = [&](bool is_swap, uint32_t creg, int j0, int j1) {
        for (Group& g : groups) {
            if (creg != 0) {  // the factor of a conditionally moved bit stops being a one-bit function
                merge_factor(g, j0);
                if (is_swap) merge_factor(g, j1);
            } else if (is_swap) {
                std::swap(g.f[j0][0], g.f[j1][0]);
                std::swap(g.f[j0][1], g.f[j1][1]);
            } else {
                std::swap(g.f[j0][0], g.f[j0][1]);
            }
            std::vector<Cx> t((size_t)E);
            for (int i = 0; i < E; ++i) {
                int pi = i;
                if (((uint32_t)i & creg) == creg) {
                    if (!is_swap)
                        pi = i ^ (1 << j0);
                    else if (((i >> j0) & 1) != ((i >> j1) & 1))
                        pi = i ^ (1 << j0) ^ (1 << j1);
                }
                t[(size_t)pi] = g.tab[i];
            }
            g.tab.swap(t);
        }
    };
    auto matrix_of = [&](LoweredOp const& op, Cx (&m)[4]) {
        if (op.kind == OP_H) {
            m[0] = {op.c[0], 0}, m[1] = {op.c[0], 0}, m[2] = {op.c[0], 0}, m[3] = {-op.c[0], 0};
        } else if (op.kind == OP_X) {
            m[0] = {0, 0}, m[1] = {1, 0}, m[2] = {1, 0}, m[3] = {0, 0};
        } else {
            for (int i = 0; i < 4; ++i) m[i] = {op.c[2 * i], op.c[2 * i + 1]};
        }
    };
    auto cost_of = [](int kind) { return kind == K_H ? 4.1 : kind == K_X ? 3.7 : 8.1; };
    // a one-bit diagonal between two ops usually rides along with a table that is emitted anyway,
    // so absorbing it is not counted as a saving; fusion must win by a clear margin
    double const kDiagCost = 0.0, kMargin = 0.5;
    // B = op (an uncontrolled U1 / H / X on register bit u.j0).  Returns true when B was
    // multiplied into the previous op on that bit (or cancelled against it) and needs no micro-op.
    auto try_fuse = [&](LoweredOp const& op, MicroOp const& u) -> bool {
        int const j = u.j0;
        Fusable& f  = fus[j];
        // live diagonals that depend on bit j sit between the previous op and B: an unconditional
        // one-bit factor can be multiplied into B, anything else has to be applied first
        int absorb = -1;
        for (size_t gi = 0; gi < groups.size() && op.kind != OP_X; ++gi) {  // (a permutation carries them through)
            Group const& g = groups[gi];
            if (table_depends_on(g, j)) return false;
            if (same(g.f[j][0], g.f[j][1])) continue;
            if (g.cother != 0 || g.czero != 0 || absorb >= 0) return false;
            absorb = (int)gi;
        }
        Cx d0{1, 0}, d1{1, 0};
        if (absorb >= 0) d0 = groups[(size_t)absorb].f[j][0], d1 = groups[(size_t)absorb].f[j][1];
        auto absorbed = [&]() {
            if (absorb >= 0) groups[(size_t)absorb].f[j][0] = Cx{}, groups[(size_t)absorb].f[j][1] = Cx{};
        };
        Cx mb[4];
        matrix_of(op, mb);
        Cx bd[4] = {cmul(mb[0], d0), cmul(mb[1], d1), cmul(mb[2], d0), cmul(mb[3], d1)};  // B * D
        auto cadd = [](Cx a, Cx b) { return Cx{a.re + b.re, a.im + b.im}; };
        double const tol = plan.cfg.snap_tol;
        auto snap = [&](Cx z) { return Cx{snap1(z.re, tol), snap1(z.im, tol)}; };
        int const bkind = op.kind == OP_H ? K_H : op.kind == OP_X ? K_X : K_U1;
        if (f.idx >= 0) {
            Cx m[4] = {snap(cadd(cmul(bd[0], f.m[0]), cmul(bd[1], f.m[2]))), snap(cadd(cmul(bd[0], f.m[1]), cmul(bd[1], f.m[3]))),
                       snap(cadd(cmul(bd[2], f.m[0]), cmul(bd[3], f.m[2]))), snap(cadd(cmul(bd[2], f.m[1]), cmul(bd[3], f.m[3])))};
            auto zero = [](Cx z) { return z.re == 0.0 && z.im == 0.0; };
            auto one  = [](Cx z) { return z.re == 1.0 && z.im == 0.0; };
            bool const diag = zero(m[1]) && zero(m[2]);
            bool const isx  = zero(m[0]) && zero(m[3]) && one(m[1]) && one(m[2]);
            bool const ish  = m[0].im == 0.0 && m[1].im == 0.0 && m[2].im == 0.0 && m[3].im == 0.0 && m[0].re == m[1].re &&
                             m[0].re == m[2].re && m[0].re == -m[3].re;
            double const fused_cost = diag ? (one(m[0]) && one(m[3]) ? 0.0 : 0.5) : isx ? 3.7 : ish ? 4.1 : 8.1;
            double const separate   = f.cost + (absorb >= 0 ? kDiagCost : 0.0) + cost_of(bkind);
            if (fused_cost + kMargin < separate) {
                if (op.kind == OP_X && plan.cfg.conjugate_diagonals) conjugate_tables(false, 0u, j, -1);
                else if (op.kind == OP_X) return false;
                absorbed();
                MicroOp& a = st.uops[(size_t)f.idx];
                a.n_src += 1;
                if (diag) {
                    a.kind = -1;  // dropped below; the product joins the live tables
                    f.idx  = -1;
                    if (!(one(m[0]) && one(m[3]))) {
                        Group& g = group_for(0, 0);
                        g.f[j][0] = cmul(g.f[j][0], m[0]), g.f[j][1] = cmul(g.f[j][1], m[3]);
                    }
                    return true;
                }
                std::memset(a.c, 0, sizeof(a.c));
                if (isx) {
                    a.kind = K_X, a.c[7] = -0.0;
                } else if (ish) {
                    a.kind = K_H, a.c[0] = m[0].re;
                } else {
                    a.kind = K_U1;
                    for (int i = 0; i < 4; ++i) a.c[2 * i] = m[i].re, a.c[2 * i + 1] = m[i].im;
                }
                for (int i = 0; i < 4; ++i) f.m[i] = m[i];
                f.cost = fused_cost;
                return true;
            }
        }
        if (absorb >= 0 && op.kind == OP_U1) {  // a general 2x2 takes the one-bit diagonal for free
            absorbed();
            MicroOp v = u;
            v.kind    = K_U1;
            v.n_src   = 1;
            for (int i = 0; i < 4; ++i) v.c[2 * i] = bd[i].re, v.c[2 * i + 1] = bd[i].im;
            flush(1u << j);
            f.idx = (int)st.uops.size();
            for (int i = 0; i < 4; ++i) f.m[i] = bd[i];
            f.cost = 8.1;
            st.uops.push_back(std::move(v));
            return true;
        }
        return false;
    };

    for (int pos = 0; pos < (int)st.ops.size(); ++pos) {
        LoweredOp const& op = plan.ops[st.ops[pos]];
        if (op.tmask() == 0) {
            add_diagonal(op);
            continue;
        }
        MicroOp u;
        u.src    = pos;
        u.creg   = creg_of(op.req);
        u.cother = op.req & ~regmask;
        u.j0     = regidx_of[op.t0];
        u.j1     = op.t1 >= 0 ? regidx_of[op.t1] : -1;
        std::memcpy(u.c, op.c, sizeof(u.c));
        uint32_t const jmask = (1u << u.j0) | (u.j1 >= 0 ? 1u << u.j1 : 0u);
        bool const plain_1q = (op.kind == OP_U1 || op.kind == OP_H || op.kind == OP_X) && u.creg == 0 && u.cother == 0;
        if (plan.cfg.fuse_1q && plain_1q && try_fuse(op, u)) continue;
        if (plan.cfg.conjugate_diagonals && (op.kind == OP_X || op.kind == OP_SWAP) && u.cother == 0) {
            conjugate_tables(op.kind == OP_SWAP, u.creg, u.j0, u.j1);
        } else {
            flush(jmask);
        }
        // whatever this op touches in the registers can no longer be fused across
        for (int j = 0; j < R; ++j)
            if (((jmask | u.creg) >> j) & 1u) fus[j].idx = -1;
        switch (op.kind) {
            case OP_U1: u.kind = K_U1; break;
            case OP_H: u.kind = K_H; break;
            case OP_AD: u.kind = K_AD; break;
            case OP_SWAP: u.kind = K_SWAP; break;
            case OP_U2:
                u.kind  = K_U2;
                u.table = op.mat;
                break;
            default:  // OP_X
                if (__builtin_popcount(u.creg) == 1) {
                    u.kind = K_XC;
                    u.j1   = __builtin_ctz(u.creg);
                    u.creg = 0;
                } else {
                    u.kind = K_X;
                }
                break;
        }
        // the exchange handlers move half of the data through the FP64 pipe as x + (-0.0)
        if (u.kind == K_X || u.kind == K_XC || u.kind == K_SWAP) u.c[7] = -0.0;
        if (plain_1q) {
            Fusable& f = fus[u.j0];
            f.idx      = (int)st.uops.size();
            matrix_of(op, f.m);
            f.cost = cost_of(u.kind);
        }
        st.uops.push_back(std::move(u));
    }
    for (Group const& g : groups) emit_group(g);
    st.uops.erase(std::remove_if(st.uops.begin(), st.uops.end(), [](MicroOp const& v) { return v.kind < 0; }), st.uops.end());
    if (!scalars.empty()) {
        MicroOp u;
        u.kind = K_SCALN;
        u.j0   = (int)scalars.size();
        for (MicroOp const& e : scalars) {
            // entry = 2 doubles of packed ints {cother, czero, sel, 0} + d0 + d1
            uint32_t w[4] = {e.cother, e.czero, (uint32_t)e.sel, 0u};
            double packed[2];
            std::memcpy(packed, w, sizeof(packed));
            u.table.push_back(packed[0]);
            u.table.push_back(packed[1]);
            u.table.push_back(e.c[0]);
            u.table.push_back(e.c[1]);
            if (e.sel >= 0) {
                u.table.push_back(e.c[2]);
                u.table.push_back(e.c[3]);
            } else {
                u.table.push_back(e.c[0]);
                u.table.push_back(e.c[1]);
            }
        }
        st.uops.push_back(std::move(u));
    }
}

// Builds the sweep of `chosen` (tile bits S).  With `trim`, a short trailing stage is not worth
// its shared-memory transpose (~10 micro-ops of time): its ops are handed back through `returned`
// and the sweep is rebuilt without them; they are first in line for the next sweep, whose tile
// and first register set are then chosen around them.
void finish_sweep(Plan& plan, std::vector<int> chosen, uint32_t S, std::vector<int>* returned = nullptr) {
    PlanConfig const& cfg = plan.cfg;
    for (int attempt = 0;; ++attempt) {
        PlannedSweep sw;
        // pad the tile: at least R bits; one-stage sweeps get enough rows for a full CTA
        uint32_t T = S;
        int const m_target = std::min(cfg.n, std::min(cfg.m_cap, cfg.R + 8 - cfg.log2w));
        int m              = std::max(popc(T), m_target);
        m                  = std::max(m, std::min(cfg.n, cfg.R));
        for (int b = 0; b < cfg.n && popc(T) < m; ++b) T |= 1u << b;
        for (int b = 0; b < cfg.n; ++b)
            if ((T >> b) & 1u) sw.tile_bits.push_back(b);
        // tiles whose outer bits fail a control shared by every op are never touched
        uint32_t common = ~0u;
        for (int idx : chosen) common &= plan.ops[idx].req;
        sw.fixed_one = chosen.empty() ? 0u : (common & ~T);
        uint64_t const folded0 = plan.stats.n_folded_perms;
        build_stages(plan, chosen, sw);
        if (returned != nullptr && cfg.min_tail_ops > 0 && attempt < 4 && sw.stages.size() > 1) {
            PlannedStage const& tail = sw.stages.back();
            if ((int)(tail.ops.size() + tail.post.size()) < cfg.min_tail_ops) {
                std::vector<char> drop(plan.ops.size(), 0);
                for (int idx : tail.ops) drop[idx] = 1;
                for (int idx : tail.post) drop[idx] = 1;
                std::vector<int> keep;
                S = 0;
                for (int idx : chosen) {
                    if (drop[idx]) {
                        returned->push_back(idx);
                    } else {
                        keep.push_back(idx);
                        S |= plan.ops[idx].tmask();
                    }
                }
                chosen.swap(keep);
                plan.stats.n_folded_perms = folded0;
                continue;
            }
        }
        for (PlannedStage& st : sw.stages) {
            emit_uops(plan, st);
            plan.stats.n_micro_ops += st.uops.size();
        }
        sw.alg_bytes = 32.0 * std::ldexp(1.0, cfg.n - popc(sw.fixed_one)) * (double)cfg.ncols;
        plan.stats.n_stages += sw.stages.size();
        plan.stats.alg_bytes_per_run += sw.alg_bytes;
        plan.sweeps.push_back(std::move(sw));
        return;
    }
}

void schedule(Plan& plan) {
    PlanConfig const& cfg = plan.cfg;
    size_t const N        = plan.ops.size();
    if (!cfg.fuse) {
        for (size_t i = 0; i < N; ++i) finish_sweep(plan, {(int)i}, plan.ops[i].tmask());
        return;
    }
    std::vector<char> done(N, 0);
    size_t first = 0;
    // One scan of the pending ops: an op joins the sweep if it commutes past everything skipped
    // so far and its non-diagonal targets fit the tile.  `allowed` restricts the tile bits
    // (search); without it the tile grows first come, first served up to m_cap bits.
    auto scan = [&](uint32_t allowed, bool restricted, std::vector<int>* out, uint32_t& S_out) {
        uint32_t S = 0;
        int count = 0, scanned = 0;
        Picker pk;
        for (size_t i = first; i < N && scanned < cfg.lookahead; ++i) {
            if (done[i]) continue;
            ++scanned;
            LoweredOp const& op = plan.ops[i];
            bool const fits     = restricted ? (op.tmask() & ~allowed) == 0 : popc(S | op.tmask()) <= cfg.m_cap;
            if (pk.passes(op) && fits) {
                S |= op.tmask();
                if (out != nullptr) out->push_back((int)i);
                if (++count >= cfg.max_ops) break;
            } else {
                pk.block(op);
                if (popc(pk.blockedT) >= cfg.n) break;
            }
        }
        S_out = S;
        return count;
    };
    while (true) {
        while (first < N && done[first]) ++first;
        if (first >= N) break;
        uint32_t S = 0;
        std::vector<int> chosen;
        if (cfg.search > 0 && cfg.n > cfg.m_cap) {
            // hill climbing over the tile's bit set, seeded with the first-come choice: exchange one
            // tile bit for one outer bit while that lets the sweep take more ops
            uint32_t S0  = 0;
            int best     = scan(0, false, nullptr, S0);
            uint32_t cur = S0;
            for (int b = 0; b < cfg.n && popc(cur) < cfg.m_cap; ++b) cur |= 1u << b;
            best = std::max(best, scan(cur, true, nullptr, S0));
            for (int round = 0; round < cfg.search; ++round) {
                uint32_t best_set = cur;
                for (int a = 0; a < cfg.n; ++a) {
                    if (!((cur >> a) & 1u)) continue;
                    for (int b = 0; b < cfg.n; ++b) {
                        if ((cur >> b) & 1u) continue;
                        uint32_t const cand = (cur & ~(1u << a)) | (1u << b);
                        int const c         = scan(cand, true, nullptr, S0);
                        if (c > best) best = c, best_set = cand;
                    }
                }
                if (best_set == cur) break;
                cur = best_set;
            }
            scan(cur, true, &chosen, S);
        } else {
            scan(0, false, &chosen, S);
        }
        for (int idx : chosen) done[idx] = 1;
        // trimming pays only if the ops handed back join a sweep that exists anyway
        bool more = false;
        for (size_t i = first; i < N && !more; ++i) more = !done[i];
        std::vector<int> returned;
        finish_sweep(plan, chosen, S, more ? &returned : nullptr);
        for (int idx : returned) done[idx] = 0;
        if (!returned.empty()) first = std::min(first, (size_t)*std::min_element(returned.begin(), returned.end()));
    }
}

// ---- dense k-qubit blocks (north_star: "fused k-qubit blocks ... on the DMMA tensor pipe") ----
//
// With `dense_block = k` the gate stream is folded into dense 2^k x 2^k complex blocks: a block
// absorbs every op (taken in commutation-respecting order, like a sweep) ALL of whose bits --
// targets, controls and diagonal bits -- fit in the block's k row bits; the ops are multiplied
// together on the host and the block is applied as a real (2^(k+1))^2 GEMM by dense_kernel<k>.
// A dense block costs 8 * 2^k real flops per element no matter how many gates it absorbed, so it
// pays off only for circuits of dense rotations; for Clifford+T the interpreter is faster
// (B200's FP64 tensor pipe has the same peak as its FMA pipe).

using CMat = std::vector<std::complex<double>>;

// M <- (op acting on the block's local bits) * M, M is D x D row-major; lb(global bit) -> local bit
void fold_op_into_block(CMat& M, int k, LoweredOp const& op, int const* local_of_bit) {
    size_t const D = size_t{1} << k;
    auto lbit = [&](int g) { return local_of_bit[g]; };
    uint32_t req = 0;
    for (int b = 0; b < kMaxRowBits; ++b)
        if ((op.req >> b) & 1u) req |= 1u << lbit(b);
    using C = std::complex<double>;
    auto for_rows = [&](auto fn) {
        for (size_t r = 0; r < D; ++r)
            if (((uint32_t)r & req) == req) fn(r);
    };
    auto scale_row = [&](size_t r, C w) {
        for (size_t c = 0; c < D; ++c) M[r * D + c] *= w;
    };
    switch (op.kind) {
        case OP_PHASE: for_rows([&](size_t r) { scale_row(r, C(op.c[0], op.c[1])); }); return;
        case OP_NEG: for_rows([&](size_t r) { scale_row(r, C(-1, 0)); }); return;
        case OP_MULI: for_rows([&](size_t r) { scale_row(r, C(0, op.c[0])); }); return;
        case OP_DIAG: {
            int const d = lbit(op.dbit);
            for_rows([&](size_t r) { scale_row(r, ((r >> d) & 1) ? C(op.c[2], op.c[3]) : C(op.c[0], op.c[1])); });
            return;
        }
        default: break;
    }
    int const t0 = lbit(op.t0);
    if (op.kind == OP_SWAP || op.kind == OP_U2) {
        int const t1 = lbit(op.t1);
        C g[16];
        if (op.kind == OP_SWAP) {
            static double const sw[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
            for (int i = 0; i < 16; ++i) g[i] = sw[i];
        } else {
            for (int i = 0; i < 16; ++i) g[i] = C(op.mat[2 * i], op.mat[2 * i + 1]);
        }
        for_rows([&](size_t r) {
            if (((r >> t0) & 1) || ((r >> t1) & 1)) return;
            size_t const idx[4] = {r, r | (size_t{1} << t1), r | (size_t{1} << t0), r | (size_t{1} << t0) | (size_t{1} << t1)};
            for (size_t c = 0; c < D; ++c) {
                C v[4], w[4];
                for (int x = 0; x < 4; ++x) v[x] = M[idx[x] * D + c];
                for (int y = 0; y < 4; ++y) {
                    w[y] = 0;
                    for (int x = 0; x < 4; ++x) w[y] += g[4 * y + x] * v[x];
                }
                for (int y = 0; y < 4; ++y) M[idx[y] * D + c] = w[y];
            }
        });
        return;
    }
    C m00, m01, m10, m11;
    switch (op.kind) {
        case OP_H: m00 = m01 = m10 = op.c[0], m11 = -op.c[0]; break;
        case OP_X: m00 = m11 = 0, m01 = m10 = 1; break;
        case OP_AD: m00 = m11 = 0, m01 = C(op.c[0], op.c[1]), m10 = C(op.c[2], op.c[3]); break;
        default: m00 = C(op.c[0], op.c[1]), m01 = C(op.c[2], op.c[3]), m10 = C(op.c[4], op.c[5]), m11 = C(op.c[6], op.c[7]); break;
    }
    for_rows([&](size_t r) {
        if ((r >> t0) & 1) return;
        size_t const hi = r | (size_t{1} << t0);
        for (size_t c = 0; c < D; ++c) {
            C const a = M[r * D + c], b = M[hi * D + c];
            M[r * D + c]  = m00 * a + m01 * b;
            M[hi * D + c] = m10 * a + m11 * b;
        }
    });
}

void schedule_dense(Plan& plan) {
    PlanConfig const& cfg = plan.cfg;
    int const k           = cfg.dense_k;
    size_t const N        = plan.ops.size();
    std::vector<char> done(N, 0);
    size_t first = 0;
    while (true) {
        while (first < N && done[first]) ++first;
        if (first >= N) break;
        if (popc(plan.ops[first].tmask() | plan.ops[first].dmask()) > k) {
            // touches more than k bits (e.g. cccx for k = 3): no dense block can hold it ->
            // an interpreted sweep of its own
            done[first] = 1;
            finish_sweep(plan, {(int)first}, plan.ops[first].tmask());
            continue;
        }
        uint32_t S = 0;
        std::vector<int> chosen;
        Picker pk;
        int scanned = 0;
        for (size_t i = first; i < N && scanned < cfg.lookahead; ++i) {
            if (done[i]) continue;
            ++scanned;
            LoweredOp const& op = plan.ops[i];
            uint32_t const bits = op.tmask() | op.dmask();  // a dense matrix needs every touched bit inside
            if (pk.passes(op) && popc(S | bits) <= k) {
                S |= bits;
                chosen.push_back((int)i);
                done[i] = 1;
            } else {
                pk.block(op);
                if (popc(pk.blockedT) >= cfg.n) break;
            }
        }
        PlannedSweep sw;
        for (int b = 0; b < cfg.n && popc(S) < k; ++b) S |= 1u << b;  // pad the block to exactly k bits
        int local_of_bit[kMaxRowBits];
        for (int b = 0; b < kMaxRowBits; ++b) local_of_bit[b] = -1;
        for (int b = 0; b < cfg.n; ++b)
            if ((S >> b) & 1u) {
                local_of_bit[b] = (int)sw.block_bits.size();
                sw.block_bits.push_back(b);
            }
        size_t const D = size_t{1} << k;
        CMat M(D * D, 0.0);
        for (size_t r = 0; r < D; ++r) M[r * D + r] = 1.0;
        for (int idx : chosen) fold_op_into_block(M, k, plan.ops[idx], local_of_bit);
        for (auto const& z : M) {
            sw.block_matrix.push_back(z.real());
            sw.block_matrix.push_back(z.imag());
        }
        sw.block_ops = chosen;
        // tile: the block bits plus enough other row bits for a full CTA
        uint32_t T       = S;
        int const m_want = std::min(cfg.n, std::max(k, std::min(cfg.m_cap, 9)));
        for (int b = 0; b < cfg.n && popc(T) < m_want; ++b) T |= 1u << b;
        for (int b = 0; b < cfg.n; ++b)
            if ((T >> b) & 1u) sw.tile_bits.push_back(b);
        sw.alg_bytes = 32.0 * std::ldexp(1.0, cfg.n) * (double)cfg.ncols;
        plan.stats.alg_bytes_per_run += sw.alg_bytes;
        ++plan.stats.n_dense_blocks;
        plan.sweeps.push_back(std::move(sw));
    }
}

// ---- serialization ----------------------------------------------------------------------

int handler_of(MicroOp const& u) {
    bool const g = u.creg != 0;
    switch (u.kind) {
        case K_H: return (g ? H_HG : H_H) + u.j0;
        case K_X: return (g ? H_XG : H_X) + u.j0;
        case K_XC: return H_XC + u.j0 * kJ + u.j1;
        case K_U1: return (g ? H_U1G : H_U1) + u.j0;
        case K_AD: return (g ? H_ADG : H_AD) + u.j0;
        case K_SWAP: {
            int const ja = std::max(u.j0, u.j1), jb = std::min(u.j0, u.j1);
            return (g ? H_SWAPG : H_SWAP) + ja * kJ + jb;
        }
        case K_U2: return H_U2;
        case K_DSCAL: return H_DSCAL;
        case K_DTABH: return H_DTABH + u.j0;
        case K_DTABF: return H_DTABF;
        case K_SCALN: return H_SCALN;
        case K_PHJ: return H_PHJ + u.j0;
        default: return H_SAPPLY;
    }
}

void serialize(Plan& plan) {
    PlanConfig const& cfg = plan.cfg;
    for (PlannedSweep const& sw : plan.sweeps) {
        DevSweep hs;
        std::memset(&hs, 0, sizeof(hs));
        int const m   = (int)sw.tile_bits.size();
        if (!sw.block_bits.empty()) {
            int const k = (int)sw.block_bits.size();
            hs.m        = m;
            hs.log2w    = cfg.log2w;
            hs.dense_k  = k;
            uint32_t S  = 0;
            int local_of[kMaxRowBits];
            for (int b = 0; b < kMaxRowBits; ++b) local_of[b] = -1;
            for (int l = 0; l < m; ++l) {
                hs.sp[l]                   = sw.tile_bits[l];
                local_of[sw.tile_bits[l]] = l;
                S |= 1u << sw.tile_bits[l];
            }
            int no = 0;
            for (int b = 0; b < cfg.n; ++b)
                if (!((S >> b) & 1u)) hs.outer[no++] = b;
            hs.n_outer = no;
            int32_t bl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < k; ++i) bl[i] = local_of[sw.block_bits[i]];
            size_t const D = size_t{1} << k, lda = (size_t)dense_lda(k);
            std::vector<double> A(2 * D * lda, 0.0);
            for (size_t r = 0; r < D; ++r)
                for (size_t c = 0; c < D; ++c) {
                    double const gr = sw.block_matrix[2 * (r * D + c)], gi = sw.block_matrix[2 * (r * D + c) + 1];
                    A[(2 * r) * lda + 2 * c]         = gr;
                    A[(2 * r) * lda + 2 * c + 1]     = -gi;
                    A[(2 * r + 1) * lda + 2 * c]     = gi;
                    A[(2 * r + 1) * lda + 2 * c + 1] = gr;
                }
            hs.pool_doubles  = (int)A.size();
            size_t const off = (plan.blob.size() + 255) / 256 * 256;
            size_t const sz  = sizeof(DevSweep) + sizeof(bl) + A.size() * sizeof(double);
            plan.blob.resize(off + sz, 0);
            unsigned char* p = plan.blob.data() + off;
            std::memcpy(p, &hs, sizeof(hs));
            std::memcpy(p + sizeof(hs), bl, sizeof(bl));
            std::memcpy(p + sizeof(hs) + sizeof(bl), A.data(), A.size() * sizeof(double));
            plan.sweep_offset.push_back(off);
            plan.sweep_size.push_back(sz);
            continue;
        }
        hs.n_stages   = (int)sw.stages.size();
        hs.m          = m;
        hs.log2w      = cfg.log2w;
        hs.fixed_one  = sw.fixed_one;
        uint32_t S    = 0;
        int local_of[kMaxRowBits];
        for (int b = 0; b < kMaxRowBits; ++b) local_of[b] = -1;
        for (int l = 0; l < m; ++l) {
            hs.sp[l]                   = sw.tile_bits[l];
            local_of[sw.tile_bits[l]] = l;
            S |= 1u << sw.tile_bits[l];
        }
        int no = 0;
        for (int b = 0; b < cfg.n; ++b)
            if (!((S >> b) & 1u) && !((sw.fixed_one >> b) & 1u)) hs.outer[no++] = b;
        hs.n_outer = no;
        std::vector<DevStage> stages;
        std::vector<DevOp> ops;
        std::vector<DevPerm> perms;
        std::vector<double> pool;
        // Boundaries to shared memory address the tile (local row bits; controls outside the tile
        // are a per-CTA test); the two boundaries to global memory are kept in global row bits
        // so that the kernel needs no local -> global bit deposit.
        auto add_perm_steps = [&](int idx, bool global_rows) {
            LoweredOp const& lo = plan.ops[idx];
            auto step = [&](uint32_t ctrl_global, int target_global) {
                DevPerm d{0, 0, 0, 0};
                if (global_rows) {
                    d.cml = ctrl_global;
                    d.tml = 1u << target_global;
                    perms.push_back(d);
                    return;
                }
                for (int b = 0; b < cfg.n; ++b)
                    if ((ctrl_global >> b) & 1u) {
                        if (local_of[b] >= 0)
                            d.cml |= 1u << local_of[b];
                        else
                            d.cmo |= 1u << b;
                    }
                d.tml = 1u << local_of[target_global];
                perms.push_back(d);
            };
            if (lo.kind == OP_SWAP) {
                step(lo.req | (1u << lo.t0), lo.t1);
                step(lo.req | (1u << lo.t1), lo.t0);
                step(lo.req | (1u << lo.t0), lo.t1);
            } else {
                step(lo.req, lo.t0);
            }
        };
        for (size_t si = 0; si < sw.stages.size(); ++si) {
            PlannedStage const& st = sw.stages[si];
            DevStage ds;
            std::memset(&ds, 0, sizeof(ds));
            ds.op_begin = (int)ops.size();
            ds.r        = (int)st.reg_bits.size();
            uint32_t regmask_local = 0;
            for (int j = 0; j < ds.r; ++j) {
                ds.rl[j] = local_of[st.reg_bits[j]];
                regmask_local |= 1u << ds.rl[j];
            }
            int nt = 0;
            for (int l = 0; l < m; ++l)
                if (!((regmask_local >> l) & 1u)) {
                    ds.tg[nt]   = sw.tile_bits[l];
                    ds.tl[nt++] = l;
                }
            ds.n_t = nt;
            for (int j = 0; j < ds.r; ++j) ds.rg[j] = st.reg_bits[j];
            for (MicroOp const& mo : st.uops) {
                DevOp d;
                std::memset(&d, 0, sizeof(d));
                d.handler = handler_of(mo);
                d.j0     = mo.j0;
                d.j1     = mo.j1;
                d.creg   = mo.creg;
                d.cother = mo.cother;
                d.czero  = mo.czero;
                d.sel    = mo.sel;
                d.pool   = -1;
                std::memcpy(d.c, mo.c, sizeof(d.c));
                if (!mo.table.empty()) {
                    d.pool = (int)pool.size();
                    pool.insert(pool.end(), mo.table.begin(), mo.table.end());
                }
                ops.push_back(d);
            }
            ds.op_end   = (int)ops.size();
            ds.pre_begin = (int)perms.size();
            for (int idx : st.pre) add_perm_steps(idx, si == 0);
            ds.pre_end    = (int)perms.size();
            ds.post_begin = (int)perms.size();
            for (int idx : st.post) add_perm_steps(idx, si + 1 == sw.stages.size());
            ds.post_end = (int)perms.size();
            stages.push_back(ds);
        }
        hs.n_ops        = (int)ops.size();
        hs.n_perms      = (int)perms.size();
        hs.pool_doubles = (int)pool.size();
        size_t const off = (plan.blob.size() + 255) / 256 * 256;
        size_t const sz  = sizeof(DevSweep) + stages.size() * sizeof(DevStage) + ops.size() * sizeof(DevOp) +
                          perms.size() * sizeof(DevPerm) + pool.size() * sizeof(double);
        plan.blob.resize(off + sz, 0);
        unsigned char* p = plan.blob.data() + off;
        std::memcpy(p, &hs, sizeof(hs));
        p += sizeof(hs);
        if (!stages.empty()) std::memcpy(p, stages.data(), stages.size() * sizeof(DevStage));
        p += stages.size() * sizeof(DevStage);
        if (!ops.empty()) std::memcpy(p, ops.data(), ops.size() * sizeof(DevOp));
        p += ops.size() * sizeof(DevOp);
        if (!perms.empty()) std::memcpy(p, perms.data(), perms.size() * sizeof(DevPerm));
        p += perms.size() * sizeof(DevPerm);
        if (!pool.empty()) std::memcpy(p, pool.data(), pool.size() * sizeof(double));
        plan.sweep_offset.push_back(off);
        plan.sweep_size.push_back(sz);
    }
}

}  // namespace

int build_plan(int n_qubits, uint64_t ncols, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
               Plan& plan, std::string& msg) {
    auto const t_begin = std::chrono::steady_clock::now();
    if (n_qubits < 1 || n_qubits > kMaxRowBits) {
        msg = "n_qubits must be in [1, " + std::to_string(kMaxRowBits) + "]";
        return QSX_ERR_BAD_ARG;
    }
    int const lc = ilog2_exact(ncols);
    if (lc < 0 || lc > n_qubits) {
        msg = "ncols must be a power of two <= 2^n_qubits";
        return QSX_ERR_BAD_ARG;
    }
    if (n_gates > 0 && gates == nullptr) {
        msg = "null gate array";
        return QSX_ERR_BAD_ARG;
    }
    qsx_plan_options o;
    std::memset(&o, 0, sizeof(o));
    o.fuse = 1;
    if (opt != nullptr) {
        if (opt->struct_size != (int32_t)sizeof(qsx_plan_options)) {
            msg = "qsx_plan_options.struct_size mismatch";
            return QSX_ERR_BAD_ARG;
        }
        o = *opt;
    } else {
        o.snap_tol = 1e-15;
    }
    PlanConfig& cfg = plan.cfg;
    cfg.n           = n_qubits;
    cfg.ncols       = ncols;
    cfg.fuse        = o.fuse != 0;
    cfg.fold_perms  = o.no_perm_folding != 1;
    cfg.defer_perms = o.no_perm_folding == 2;
    cfg.snap_tol    = (opt != nullptr && o.snap_tol == 0.0) ? 1e-15 : o.snap_tol;
    cfg.R           = std::min(n_qubits, o.reg_qubits > 0 ? std::min((int)o.reg_qubits, kMaxRegBits) : 4);
    cfg.log2w       = std::min(lc, o.log2_tile_cols > 0 ? std::min((int)o.log2_tile_cols, 5) : 3);
    if (o.log2_tile_cols < 0) cfg.log2w = 0;
    cfg.m_cap       = o.tile_qubits > 0 ? (int)o.tile_qubits : 8;
    cfg.m_cap       = std::min(cfg.m_cap, kMaxTileBits);
    cfg.m_cap       = std::min(cfg.m_cap, 13 - cfg.log2w);            // 2^m * W * 16 B <= 128 KiB
    cfg.m_cap       = std::min(cfg.m_cap, max_threads_log2(cfg.R) - cfg.log2w + cfg.R);  // CTA size limit
    cfg.m_cap       = std::max(std::min(cfg.m_cap, n_qubits), cfg.R);
    cfg.max_ops     = o.max_ops_per_sweep > 0 ? (int)o.max_ops_per_sweep : 96;
    cfg.min_tail_ops = o.min_tail_ops > 0 ? (int)o.min_tail_ops : (o.min_tail_ops < 0 ? 0 : 8);
    cfg.lookahead   = o.lookahead > 0 ? (int)o.lookahead : 4096;
    // the search costs a few ms of host time: worth it once a sweep takes milliseconds (n >= 13)
    cfg.search      = o.search > 0 ? (int)o.search : (o.search < 0 || n_qubits < 13 ? 0 : 3);
    cfg.conjugate_diagonals = o.no_diag_conjugation == 0;
    cfg.fuse_1q             = o.no_1q_fusion == 0;
    cfg.stage_search        = o.stage_search > 0 || (o.stage_search == 0 && n_qubits >= 13);

    plan.stats          = qsx_plan_stats{};
    plan.stats.n_gates  = n_gates;
    double const shard_elems = std::ldexp(1.0, n_qubits) * (double)ncols;
    for (size_t i = 0; i < n_gates; ++i) {
        size_t const before = plan.ops.size();
        int const rc        = lower_gate(n_qubits, gates[i], (int)i, cfg.snap_tol, plan.ops, msg);
        if (rc != QSX_OK) return rc;
        for (size_t k = before; k < plan.ops.size(); ++k) plan.stats.gate_bytes_per_run += 32.0 * shard_elems * plan.ops[k].frac;
    }
    plan.stats.n_ops = plan.ops.size();
    for (LoweredOp const& op : plan.ops)
        if (op.t1 >= 0 && cfg.R < 2) {  // two-target ops need both bits in registers
            cfg.R     = 2;
            cfg.m_cap = std::max(cfg.m_cap, 2);
        }
    cfg.dense_k = 0;
    if (o.dense_block != 0) {
        if (o.dense_block < 3 || o.dense_block > kMaxDenseK) {
            msg = "dense_block must be 0 or 3.." + std::to_string(kMaxDenseK);
            return QSX_ERR_BAD_ARG;
        }
        if (cfg.log2w != 3 || n_qubits < o.dense_block) {
            msg = "dense blocks need >= k qubits and >= 8 columns per shard";
            return QSX_ERR_BAD_ARG;
        }
        cfg.dense_k = o.dense_block;
    }
    if (cfg.dense_k > 0)
        schedule_dense(plan);
    else
        schedule(plan);
    plan.stats.n_sweeps = plan.sweeps.size();
    serialize(plan);
    plan.stats.plan_host_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return QSX_OK;
}

std::string dump_plan_json(const Plan& plan) {
    static char const* const kNames[] = {"U1", "H", "X", "AD", "PHASE", "DIAG", "SWAP", "U2", "NEG", "MULI"};
    std::ostringstream os;
    os.precision(17);
    os << "{\"n\":" << plan.cfg.n << ",\"ncols\":" << plan.cfg.ncols << ",\"R\":" << plan.cfg.R
       << ",\"log2w\":" << plan.cfg.log2w << ",\"m_cap\":" << plan.cfg.m_cap << ",\"sweeps\":[";
    for (size_t s = 0; s < plan.sweeps.size(); ++s) {
        PlannedSweep const& sw = plan.sweeps[s];
        os << (s ? "," : "") << "{\"tile_bits\":[";
        for (size_t i = 0; i < sw.tile_bits.size(); ++i) os << (i ? "," : "") << sw.tile_bits[i];
        os << "],\"dense\":{\"bits\":[";
        for (size_t i = 0; i < sw.block_bits.size(); ++i) os << (i ? "," : "") << sw.block_bits[i];
        os << "],\"gates\":[";
        for (size_t i = 0; i < sw.block_ops.size(); ++i) os << (i ? "," : "") << plan.ops[sw.block_ops[i]].src_gate;
        os << "],\"mat\":[";
        for (size_t i = 0; i < sw.block_matrix.size(); ++i) os << (i ? "," : "") << sw.block_matrix[i];
        os << "]},\"fixed_one\":" << sw.fixed_one << ",\"stages\":[";
        for (size_t t = 0; t < sw.stages.size(); ++t) {
            PlannedStage const& st = sw.stages[t];
            os << (t ? "," : "") << "{\"reg_bits\":[";
            for (size_t i = 0; i < st.reg_bits.size(); ++i) os << (i ? "," : "") << st.reg_bits[i];
            os << "]";
            for (int which = 0; which < 2; ++which) {
                std::vector<int> const& lst = which == 0 ? st.pre : st.post;
                os << ",\"" << (which == 0 ? "pre" : "post") << "\":[";
                for (size_t i = 0; i < lst.size(); ++i) {
                    LoweredOp const& op = plan.ops[lst[i]];
                    os << (i ? "," : "") << "{\"kind\":\"" << kNames[op.kind] << "\",\"t0\":" << op.t0 << ",\"t1\":" << op.t1
                       << ",\"dbit\":-1,\"req\":" << op.req << ",\"gate\":" << op.src_gate << ",\"c\":[0,0,0,0,0,0,0,0],\"mat\":[]}";
                }
                os << "]";
            }
            os << ",\"ops\":[";
            for (size_t i = 0; i < st.ops.size(); ++i) {
                LoweredOp const& op = plan.ops[st.ops[i]];
                os << (i ? "," : "") << "{\"kind\":\"" << kNames[op.kind] << "\",\"t0\":" << op.t0 << ",\"t1\":" << op.t1
                   << ",\"dbit\":" << op.dbit << ",\"req\":" << op.req << ",\"gate\":" << op.src_gate << ",\"c\":[";
                for (int k = 0; k < 8; ++k) os << (k ? "," : "") << op.c[k];
                os << "],\"mat\":[";
                for (size_t k = 0; k < op.mat.size(); ++k) os << (k ? "," : "") << op.mat[k];
                os << "]}";
            }
            static char const* const kMicro[] = {"U1", "H", "X", "AD", "XC", "SWAP", "U2", "DSCAL", "DTABH", "DTABF", "SAPPLY", "SCALN", "PHJ"};
            os << "],\"uops\":[";
            for (size_t i = 0; i < st.uops.size(); ++i) {
                MicroOp const& u = st.uops[i];
                os << (i ? "," : "") << "{\"kind\":\"" << kMicro[u.kind] << "\",\"j0\":" << u.j0 << ",\"j1\":" << u.j1
                   << ",\"creg\":" << u.creg << ",\"cother\":" << u.cother << ",\"czero\":" << u.czero
                   << ",\"sel\":" << u.sel << ",\"src\":" << u.src << ",\"n_src\":" << u.n_src << ",\"c\":[";
                for (int k = 0; k < 8; ++k) os << (k ? "," : "") << u.c[k];
                os << "],\"table\":[";
                if (u.kind != K_SCALN)
                    for (size_t k = 0; k < u.table.size(); ++k) os << (k ? "," : "") << u.table[k];
                os << "],\"scalars\":[";
                if (u.kind == K_SCALN)
                    for (int k = 0; k < u.j0; ++k) {
                        uint32_t w[4];
                        std::memcpy(w, &u.table[(size_t)6 * k], sizeof(w));
                        os << (k ? "," : "") << "{\"cother\":" << w[0] << ",\"czero\":" << w[1] << ",\"sel\":" << (int)w[2]
                           << ",\"d\":[" << u.table[6 * k + 2] << "," << u.table[6 * k + 3] << "," << u.table[6 * k + 4] << ","
                           << u.table[6 * k + 5] << "]}";
                    }
                os << "]}";
            }
            os << "]}";
        }
        os << "]}";
    }
    os << "]}";
    return os.str();
}

}  // namespace qsx
