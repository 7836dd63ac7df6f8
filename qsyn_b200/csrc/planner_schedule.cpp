// planner_schedule.cpp -- schedules lowered ops into blocked multi-gate sweeps and stages.
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
#include <algorithm>
#include <cstdlib>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstring>

#include "planner_internal.hpp"

namespace qsx::detail {

namespace {

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

}  // namespace

// One dense sweep for a single OP_DENSE op (a user gate with >= 3 targets): dense_kernel<k>
void dense_op_sweep(Plan& plan, int idx) {
    PlanConfig const& cfg = plan.cfg;
    LoweredOp const& op   = plan.ops[idx];
    int const k           = (int)op.dbits.size();
    PlannedSweep sw;
    sw.block_bits   = op.dbits;
    sw.block_matrix = op.mat;
    sw.block_ops    = {idx};
    uint32_t T      = op.tmask();
    // tile: the block bits plus other row bits up to 2^8 rows (three tile buffers of 32 KiB: two CTAs per SM)
    int const m_want = std::min(cfg.n, std::max(k, 8));
    for (int b = 0; b < cfg.n && popc(T) < m_want; ++b) T |= 1u << b;
    for (int b = 0; b < cfg.n; ++b)
        if ((T >> b) & 1u) sw.tile_bits.push_back(b);
    sw.alg_bytes = 32.0 * std::ldexp(1.0, cfg.n) * (double)cfg.ncols;
    plan.stats.alg_bytes_per_run += sw.alg_bytes;
    ++plan.stats.n_dense_blocks;
    plan.sweeps.push_back(std::move(sw));
}

void schedule_range(Plan& plan, size_t begin, size_t N);

void schedule(Plan& plan) {
    // a dense user gate is a sweep of its own: the ops before it and after it are scheduled separately
    size_t begin = 0;
    for (size_t i = 0; i < plan.ops.size(); ++i)
        if (plan.ops[i].kind == OP_DENSE) {
            schedule_range(plan, begin, i);
            dense_op_sweep(plan, (int)i);
            begin = i + 1;
        }
    schedule_range(plan, begin, plan.ops.size());
}

void schedule_range(Plan& plan, size_t begin, size_t N) {
    PlanConfig const& cfg = plan.cfg;
    if (begin >= N) return;
    if (!cfg.fuse) {
        for (size_t i = begin; i < N; ++i) finish_sweep(plan, {(int)i}, plan.ops[i].tmask());
        return;
    }
    std::vector<char> done(N, 0);
    size_t first = begin;
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

namespace {

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

}  // namespace

void schedule_dense(Plan& plan) {
    PlanConfig const& cfg = plan.cfg;
    int const k           = cfg.dense_k;
    size_t const N        = plan.ops.size();
    std::vector<char> done(N, 0);
    size_t first = 0;
    while (true) {
        while (first < N && done[first]) ++first;
        if (first >= N) break;
        if (plan.ops[first].kind == OP_DENSE) {  // a user gate that already is a dense block
            done[first] = 1;
            dense_op_sweep(plan, (int)first);
            continue;
        }
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
            if (op.kind != OP_DENSE && pk.passes(op) && popc(S | bits) <= k) {
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
        int const m_want = std::min(cfg.n, std::max(k, 8));  // 2^8-row tiles (launch_dense_k, kernels.cu)
        for (int b = 0; b < cfg.n && popc(T) < m_want; ++b) T |= 1u << b;
        for (int b = 0; b < cfg.n; ++b)
            if ((T >> b) & 1u) sw.tile_bits.push_back(b);
        sw.alg_bytes = 32.0 * std::ldexp(1.0, cfg.n) * (double)cfg.ncols;
        plan.stats.alg_bytes_per_run += sw.alg_bytes;
        ++plan.stats.n_dense_blocks;
        plan.sweeps.push_back(std::move(sw));
    }
}

}  // namespace qsx::detail
