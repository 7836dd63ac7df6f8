// planner_emit.cpp -- turns the ops of one stage into the micro-ops the kernel interprets.
#include <algorithm>
#include <cassert>
#include <cstring>

#include "planner_internal.hpp"

namespace qsx::detail {

// ---- micro-op emission ------------------------------------------------------------------
//
// Non-diagonal ops become register micro-ops (runs of uncontrolled single-qubit ops on one bit
// are multiplied together first).  Diagonal ops (PHASE/NEG/MULI/DIAG) commute with each other
// and with any op that does not act non-diagonally on their bits, so they are held back and merged:
//   * a diagonal op whose bits are all OUTSIDE the registers multiplies every element a thread
//     holds by the same number: a per-thread scalar; all scalars of a stage travel as ONE
//     micro-op (K_SCALN) applied at the end of the stage;
//   * the others are multiplied together, per distinct thread-level condition, into a factored
//     table over the register index, emitted as K_PHJ (one phase on one bit), K_DTABH (the
//     elements with one register bit set) or K_DTABF (all 2^R elements).

namespace {

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

}  // namespace

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
    scalars.reserve(16);
    st.uops.reserve(st.ops.size() + 8);

    // The diagonal ops seen so far and not yet applied, multiplied together: one table over the
    // register index per distinct thread-level condition.  A table stays live across
    //   * ops that do not act on a register bit it depends on (they commute), and
    //   * in-register permutations (X / CX / SWAP whose controls are all register bits):
    //     D then P  ==  P then D',  D'[P(i)] = D[i]  -- the table is permuted instead of applied;
    // it is applied right before an op that mixes a bit it depends on, or at the end of the stage.
    // A group is kept factored: value(i) = tab[i] * prod_j f[j][bit j of i].  Single-bit diagonals
    // (z, s, t, rz, p ...) go to the factors, so that before an op on bit j only the factor of j has
    // to be applied (one phase, K_PHJ) or absorbed (general 2x2), while the rest stays live.
    // (fixed-size storage: a Group is copied, permuted and flushed for almost every op of a stage, and
    // heap-backed tables made the allocator the largest item of the planner's profile)
    constexpr int kMaxE = 1 << kMaxRegBits;
    struct Group {
        uint32_t cother, czero;
        Cx tab[kMaxE];
        Cx f[kMaxRegBits][2];
    };
    std::vector<Group> groups;
    groups.reserve(8);
    auto group_for = [&](uint32_t cother, uint32_t czero) -> Group& {
        for (Group& g : groups)
            if (g.cother == cother && g.czero == czero) return g;
        groups.emplace_back();
        Group& g = groups.back();
        g.cother = cother, g.czero = czero;
        return g;
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
        size_t n_keep = 0;  // compacted in place
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            Group& g = groups[gi];
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
            if (n_keep != gi) groups[n_keep] = g;
            ++n_keep;
        }
        groups.resize(n_keep);
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
            Cx t[kMaxE];
            for (int i = 0; i < E; ++i) {
                int pi = i;
                if (((uint32_t)i & creg) == creg) {
                    if (!is_swap)
                        pi = i ^ (1 << j0);
                    else if (((i >> j0) & 1) != ((i >> j1) & 1))
                        pi = i ^ (1 << j0) ^ (1 << j1);
                }
                t[pi] = g.tab[i];
            }
            for (int i = 0; i < E; ++i) g.tab[i] = t[i];
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

}  // namespace qsx::detail
