// planner.cpp -- entry points of the host-side planner: build_plan (lower -> schedule ->
// serialize into the device programs of program.hpp) and the JSON dump used by the tests.
// Lowering, scheduling and micro-op emission live in planner_{lower,schedule,emit}.cpp.
#include "planner.hpp"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstring>
#include <sstream>

#include "jit.hpp"
#include "planner_internal.hpp"

namespace qsx {

using namespace detail;

namespace {

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
        case K_DTABH: return H_DTABH + u.j0;
        case K_DTABF: return H_DTABF;
        case K_SCALN: return H_SCALN;
        case K_PHJ: return H_PHJ + u.j0;
        default: assert(false && "this micro-op kind is never sent to the device"); return -1;
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
        if (!ops.empty()) std::memcpy(p, ops.data(), ops.size() * sizeof(DevOp));
        p += ops.size() * sizeof(DevOp);
        if (!stages.empty()) std::memcpy(p, stages.data(), stages.size() * sizeof(DevStage));
        p += stages.size() * sizeof(DevStage);
        if (!perms.empty()) std::memcpy(p, perms.data(), perms.size() * sizeof(DevPerm));
        p += perms.size() * sizeof(DevPerm);
        if (!pool.empty()) std::memcpy(p, pool.data(), pool.size() * sizeof(double));
        plan.sweep_offset.push_back(off);
        plan.sweep_size.push_back(sz);
    }
}


}  // namespace

// Column blocking through the L2 (qsx_plan_options.l2_block_mib).  Columns are independent, so the runtime may
// run ALL sweeps on one block of columns before it touches the next block: with a block that fits the 126 MB L2
// only the first sweep reads it from HBM and only the final state has to be written back.  That pays when the
// sweeps are memory-bound; a compute-bound plan only loses (more launches, shorter grids), so the default is
// decided with the measured model of DESIGN.md 3.1: per sweep, 32 + 23 (stages - 1) + ~3.6 per micro-op
// against a memory floor of 90 (us at n = 12; both sides scale with the shard).
static void choose_l2_blocking(Plan& plan, int32_t opt_mib) {
    PlanConfig& cfg    = plan.cfg;
    cfg.l2_block_mib   = opt_mib;
    cfg.block_cols     = 0;
    if (opt_mib < 0 || cfg.dense_k != 0 || plan.sweeps.empty()) return;
    if (opt_mib == 0) {
        if (plan.sweeps.size() < 2) return;  // nothing is read twice
        double compute = 0.0, memory = 0.0;
        for (PlannedSweep const& sw : plan.sweeps) {
            double c = 32.0 + 23.0 * ((double)sw.stages.size() - 1.0);
            for (PlannedStage const& st : sw.stages) c += 3.6 * (double)st.uops.size();
            double const share = sw.alg_bytes / (32.0 * std::ldexp(1.0, cfg.n) * (double)cfg.ncols);  // skipped tiles
            compute += c * share;
            memory += 90.0 * share;
        }
        // measured (profiles/r1_l2_block_probe.jsonl): 2.5 micro-ops per sweep -11..-18 %, 5 per sweep +-2 %, 20 per sweep +5..+17 %
        if (compute > 0.5 * memory) return;
    }
    // 64 MiB: half of the 126 MB L2; 32 MiB blocks are grids of 1.4 waves and measured slower than no blocking
    double const mib        = opt_mib > 0 ? (double)opt_mib : 64.0;
    double const col_bytes  = 16.0 * std::ldexp(1.0, cfg.n);
    uint64_t const w        = 1ull << cfg.log2w;
    uint64_t cols           = w;
    while (cols * 2 <= cfg.ncols && (double)(cols * 2) * col_bytes <= mib * 1048576.0) cols *= 2;
    if (cols < cfg.ncols) cfg.block_cols = cols;
}

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
    // specialised sweeps (qsx_plan_options.jit): on by default from n = 12 (a sweep runs for >= 0.1 ms and a job repeats
    // its programs; below that the compiler would only burn host time)
    cfg.jit         = o.jit < 0 ? 0 : (o.jit == 0 ? (n_qubits >= 12 ? 1 : 0) : std::min((int)o.jit, 2));
    if (cfg.jit != 0 && !jit::available(nullptr)) cfg.jit = 0;  // no libnvrtc on this host: the interpreter runs everything
    // tile: 2^8 rows x 8 columns (32 KiB, 4-5 CTAs per SM) is the interpreter's optimum; compiled sweeps are cheap enough
    // per op that one more tile bit pays: C4 174 -> 142 sweeps at 5.9 -> 6.5 ms each (profiles/r2_tile_sweep_c4.jsonl)
    cfg.m_cap       = o.tile_qubits > 0 ? (int)o.tile_qubits : (cfg.jit != 0 ? 9 : 8);
    cfg.m_cap       = std::min(cfg.m_cap, kMaxTileBits);
    cfg.m_cap       = std::min(cfg.m_cap, 13 - cfg.log2w);            // 2^m * W * 16 B <= 128 KiB
    cfg.m_cap       = std::min(cfg.m_cap, max_threads_log2(cfg.R) - cfg.log2w + cfg.R);  // CTA size limit
    cfg.m_cap       = std::max(std::min(cfg.m_cap, n_qubits), cfg.R);
    cfg.max_ops     = o.max_ops_per_sweep > 0 ? (int)o.max_ops_per_sweep : 96;
    // a sweep's program is staged in shared memory next to the tile (200 KiB together, launch_sweep_r): cap the ops so
    // that the serialised program always fits (<= 96 B header + a 16-entry table per micro-op, scalars 48 B each)
    {
        long const tile_bytes = 16L << (cfg.m_cap + cfg.log2w);
        long const room       = 200L * 1024 - tile_bytes - 8192;
        cfg.max_ops           = (int)std::max(4L, std::min((long)cfg.max_ops, room / 400));
    }
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
    plan.ops.reserve(n_gates + n_gates / 8);
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
    for (LoweredOp const& op : plan.ops)
        if (op.kind == OP_DENSE && cfg.log2w != 3) {
            msg = "gate " + std::to_string(op.src_gate) + ": dense gates with >= 3 targets need tiles of 8 columns (>= 8 columns per shard, default log2_tile_cols)";
            return QSX_ERR_UNSUPPORTED;
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
    choose_l2_blocking(plan, o.l2_block_mib);
    serialize(plan);
    for (size_t k = 0; k < plan.sweeps.size(); ++k) {
        PlannedSweep const& sw = plan.sweeps[k];
        if (!sw.block_bits.empty()) continue;
        size_t const tile = sw.stages.size() > 1 ? ((size_t)16 << (sw.tile_bits.size() + (size_t)cfg.log2w)) : 0;
        if (plan.sweep_size[k] + tile > 200u * 1024u) {
            msg = "sweep " + std::to_string(k) + " needs " + std::to_string(plan.sweep_size[k] + tile) +
                  " B of shared memory (program + tile), more than the 200 KiB a CTA can have: lower max_ops_per_sweep or tile_qubits";
            return QSX_ERR_BAD_ARG;
        }
    }
    plan.stats.plan_host_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return QSX_OK;
}

std::string dump_plan_json(const Plan& plan) {
    static char const* const kNames[] = {"U1", "H", "X", "AD", "PHASE", "DIAG", "SWAP", "U2", "NEG", "MULI"};
    std::ostringstream os;
    os.precision(17);
    os << "{\"n\":" << plan.cfg.n << ",\"ncols\":" << plan.cfg.ncols << ",\"R\":" << plan.cfg.R
       << ",\"log2w\":" << plan.cfg.log2w << ",\"m_cap\":" << plan.cfg.m_cap << ",\"jit\":" << plan.cfg.jit << ",\"block_cols\":" << plan.cfg.block_cols << ",\"sweeps\":[";
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
