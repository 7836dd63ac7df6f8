// runtime.cu -- the C ABI of include/qsx.h: device contexts, shard storage, plan execution,
// equivalence reduction.  CUDA lives only behind this file, kernels.cu, comm.cpp and jit.cpp.
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/qsx.h"
#include "../host/util/phase.hpp"
#include "jit.hpp"
#include "kernels.hpp"
#include "planner.hpp"
#include "runtime_internal.hpp"

struct qsx_plan {
    qsx::Plan plan;
    std::mutex mu;
    std::map<int, unsigned char*> device_blob;  // device id -> uploaded program
};

namespace qsx {

thread_local std::string g_err;

namespace {
std::mutex g_mu;
std::map<int, DeviceCtx> g_ctx;  // (node-based: the contexts never move)
}  // namespace

int device_count_nothrow() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int fail(int code, std::string msg) {
    g_err = std::move(msg);
    return code;
}

int cuda_fail(cudaError_t e, char const* what) {
    cudaGetLastError();  // clear sticky-less error state
    std::string msg = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    if (e == cudaErrorMemoryAllocation) return fail(QSX_ERR_OOM, msg);
    return fail(QSX_ERR_CUDA, msg);
}

int get_ctx(int device, DeviceCtx** out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) {
        *out = &it->second;
        QSX_CUDA(cudaSetDevice(device));
        return QSX_OK;
    }
    int const n = device_count_nothrow();
    if (n == 0) return fail(QSX_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(QSX_ERR_BAD_ARG, "device index out of range");
    QSX_CUDA(cudaSetDevice(device));
    // created piece by piece; whatever exists is released again when a later piece fails
    cudaStream_t stream = nullptr;
    double *scratch = nullptr, *host8 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaError_t e   = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&scratch, (equiv_partials_doubles() * kMaxLocalShards + 8) * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocHost(&host8, 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaEventCreate(&ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ev1);
    if (e != cudaSuccess) {
        if (ev0 != nullptr) cudaEventDestroy(ev0);
        if (host8 != nullptr) cudaFreeHost(host8);
        if (scratch != nullptr) cudaFree(scratch);
        if (stream != nullptr) cudaStreamDestroy(stream);
        return cuda_fail(e, "creating the device context");
    }
    DeviceCtx& c = g_ctx[device];
    c.device = device, c.stream = stream, c.scratch = scratch, c.host8 = host8, c.ev0 = ev0, c.ev1 = ev1;
    *out = &c;
    return QSX_OK;
}

// Shard buffers are recycled: `tensor` commands create and delete 2^n x 2^n operands all the time (every qc2ts, every
// TensorMgr copy), and a cudaMalloc / cudaFree pair of 16 GiB costs milliseconds and a device-wide synchronisation.
// A freed buffer is parked (exact size match on reuse; all work on a device is ordered by its one stream, so the next
// owner's kernels run after the previous owner's) up to QSX_POOL_FRACTION (default 0.45) of the device memory.  Over the
// cap the buffers parked LONGEST ago are released, never the one coming in: after a 15-qubit job the pool is full of
// 16 GiB buffers, and a following run of 12-qubit jobs must not pay a cudaFree + cudaMalloc per tensor (measured:
// 8.5 -> 20.9 ms per C2 job) just because stale large buffers hold the room.
cudaError_t pool_alloc(DeviceCtx* c, void** out, size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(c->pool_mu);
        auto it = c->pool.find(bytes);
        if (it != c->pool.end()) {
            *out = it->second.p;
            c->pool.erase(it);
            c->pool_bytes -= bytes;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        pool_trim(c);
        e = cudaMalloc(out, bytes);
    }
    return e;
}

void pool_free(DeviceCtx* c, void* p, size_t bytes) {
    std::vector<void*> victims;
    {
        std::lock_guard<std::mutex> lk(c->pool_mu);
        if (c->pool_cap == 0) {
            size_t free_b = 0, total_b = 0;
            double frac   = 0.45;
            if (char const* e = std::getenv("QSX_POOL_FRACTION")) frac = std::atof(e);
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = 0, cudaGetLastError();
            c->pool_cap = frac <= 0 ? 1 : (size_t)(frac * (double)total_b) + 1;
        }
        if (bytes > c->pool_cap) {
            victims.push_back(p);
        } else {
            c->pool.emplace(bytes, DeviceCtx::Parked{p, ++c->pool_seq});
            c->pool_bytes += bytes;
            while (c->pool_bytes > c->pool_cap) {  // (terminates: the newcomer alone fits)
                auto oldest = c->pool.begin();
                for (auto it = c->pool.begin(); it != c->pool.end(); ++it)
                    if (it->second.seq < oldest->second.seq) oldest = it;
                victims.push_back(oldest->second.p);
                c->pool_bytes -= oldest->first;
                c->pool.erase(oldest);
            }
        }
    }
    if (victims.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (void* v : victims) cudaFree(v);
}

void pool_trim(DeviceCtx* c) {
    std::multimap<size_t, DeviceCtx::Parked> victims;
    {
        std::lock_guard<std::mutex> lk(c->pool_mu);
        victims.swap(c->pool);
        c->pool_bytes = 0;
    }
    if (victims.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& kv : victims) cudaFree(kv.second.p);
}

int materialize(qsx_unitary* u, DeviceCtx* c) {
    if (!u->lazy_identity) return QSX_OK;
    QSX_CUDA(launch_set_identity(u->data, u->n, u->col0, u->ncols, c->stream));
    u->lazy_identity = false;
    return QSX_OK;
}

}  // namespace qsx

namespace {
using qsx::cuda_fail;
using qsx::DeviceCtx;
using qsx::fail;
using qsx::get_ctx;
int materialize_(qsx_unitary* u, DeviceCtx* c) { return qsx::materialize(u, c); }
int materialize_(const qsx_unitary* u, DeviceCtx* c) { return qsx::materialize(const_cast<qsx_unitary*>(u), c); }
std::mutex g_prog_mu;  // (kept for the staging state's own consistency; the device mutex is held around it)
// Looks up / queues the specialised kernel of every interpreted sweep of the plan (never blocks).
void request_jit_kernels(qsx::Plan& pl) {
    if (pl.cfg.jit == 0 || !pl.jit_kernels.empty() || pl.sweeps.empty()) return;
    pl.jit_kernels.assign(pl.sweeps.size(), nullptr);
    if (!qsx::jit::available(nullptr)) return;
    // Order of compilation.  The compiler (16 threads: ~20 ms per kernel) is slower than the interpreter it competes with
    // (~12 ms per sweep at n = 15), so on a cold run a kernel is only useful if it is finished before the executor gets to
    // its sweep: the queue is filled from the LAST sweep backwards -- the executor and the compiler move towards each other
    // and everything behind the meeting point runs specialised already in the first conversion (in plan order the
    // compiler would trail the executor and nothing of the first operand would run compiled).  The processes of a
    // multi-GPU job request the same kernels: each rotates the order by its share (RANK / WORLD_SIZE of the launcher, else a
    // hash of the pid), so that together with the on-disk cache (jit.cpp) they compile different kernels first and pick up
    // each other's results.
    size_t const n_sw = pl.sweeps.size();
    size_t start      = 0;
    {
        char const* const rk = std::getenv("RANK");
        char const* const ws = std::getenv("WORLD_SIZE");
        long const world     = ws != nullptr ? std::atol(ws) : 1;
        long const rank      = rk != nullptr ? std::atol(rk) : 0;
        if (world > 1 && rank >= 0 && rank < world)
            start = (size_t)rank * n_sw / (size_t)world;
        else if (qsx::comm_world_size() > 1)
            start = ((size_t)getpid() * 2654435761u >> 7) % n_sw;
    }
    for (size_t i = 0; i < n_sw; ++i) {
        size_t const s          = n_sw - 1 - (start + i) % n_sw;
        qsx::DevSweep const& hs = *reinterpret_cast<qsx::DevSweep const*>(pl.blob.data() + pl.sweep_offset[s]);
        if (hs.dense_k != 0) continue;
        int const log2_threads = hs.m - pl.cfg.R + hs.log2w;
        if (log2_threads < 0 || log2_threads > 10) continue;
        // QSX_JIT_PAIR=1: a 512-thread tile is specialised as a cluster of two 256-thread CTAs (jit_codegen.cpp).  Off by
        // default: measured on C4 the pair runs a 2^10-row sweep no faster than one 512-thread CTA (8.95 vs 8.73 ms)
        char const* const pair_env = std::getenv("QSX_JIT_PAIR");
        bool const pair            = pair_env != nullptr && std::atoi(pair_env) != 0;
        int const cta_threads      = pair && log2_threads == 9 ? 256 : (1 << log2_threads);
        pl.jit_kernels[s]     = qsx::jit::request(pl.blob.data() + pl.sweep_offset[s], pl.sweep_size[s], pl.cfg.R, cta_threads);
    }
}

}  // namespace

extern "C" {

int qsx_abi_version(void) { return QSX_ABI_VERSION; }

const char* qsx_last_error(void) { return qsx::g_err.c_str(); }

int qsx_device_count(int* count) {
    if (count == nullptr) return fail(QSX_ERR_BAD_ARG, "null count");
    *count = qsx::device_count_nothrow();
    return QSX_OK;
}

int qsx_device_stream(int device, void** stream) {
    if (stream == nullptr) return fail(QSX_ERR_BAD_ARG, "null stream out-pointer");
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    *stream = (void*)c->stream;
    return QSX_OK;
}

int qsx_device_synchronize(int device) {
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    QSX_CUDA(cudaStreamSynchronize(c->stream));
    return QSX_OK;
}

int qsx_device_trim(int device) {
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    qsx::pool_trim(c);
    return QSX_OK;
}

// ---- storage ----------------------------------------------------------------------------

int qsx_unitary_create(int device, int n_qubits, uint64_t col0, uint64_t ncols, qsx_unitary** out) {
    if (out == nullptr) return fail(QSX_ERR_BAD_ARG, "null out-pointer");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > qsx::kMaxRowBits) return fail(QSX_ERR_BAD_ARG, "n_qubits out of range");
    uint64_t const dim = 1ull << n_qubits;
    if (ncols == 0 || (ncols & (ncols - 1)) != 0 || ncols > dim || col0 % ncols != 0 || col0 + ncols > dim)
        return fail(QSX_ERR_BAD_ARG, "column shard must be a power-of-two block aligned to its size");
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    auto* u = new (std::nothrow) qsx_unitary;
    if (u == nullptr) return fail(QSX_ERR_OOM, "host allocation failed");
    u->device = device, u->n = n_qubits, u->col0 = col0, u->ncols = ncols;
    cudaError_t const e = qsx::pool_alloc(c, reinterpret_cast<void**>(&u->data), u->bytes());
    if (e != cudaSuccess) {
        delete u;
        return cuda_fail(e, "cudaMalloc(unitary shard)");
    }
    *out = u;
    return QSX_OK;
}

int qsx_unitary_set_identity(qsx_unitary* u) {
    if (u == nullptr) return fail(QSX_ERR_BAD_ARG, "null unitary");
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(u->device, &c)) return rc;
    (void)c;
    u->lazy_identity = true;  // written by the first sweep of the next gate stream, or on first use
    return QSX_OK;
}

int qsx_unitary_clone(const qsx_unitary* u, qsx_unitary** out) {
    if (u == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    qsx_unitary* copy = nullptr;
    if (int rc = qsx_unitary_create(u->device, u->n, u->col0, u->ncols, &copy)) return rc;
    DeviceCtx* c = nullptr;
    int rc       = get_ctx(u->device, &c);
    if (rc == QSX_OK) rc = materialize_(u, c);
    if (rc == QSX_OK) {
        cudaError_t e = cudaMemcpyAsync(copy->data, u->data, u->bytes(), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(clone)");
    }
    if (rc != QSX_OK) {
        std::string const keep = qsx_last_error();
        qsx_unitary_destroy(copy);
        return fail(rc, keep);
    }
    *out = copy;
    return QSX_OK;
}

int qsx_unitary_destroy(qsx_unitary* u) {
    if (u == nullptr) return QSX_OK;
    if (u->data != nullptr) {
        DeviceCtx* c = nullptr;
        if (get_ctx(u->device, &c) == QSX_OK) qsx::pool_free(c, u->data, u->bytes());
    }
    delete u;
    return QSX_OK;
}

int qsx_unitary_get_info(const qsx_unitary* u, qsx_unitary_info* info) {
    if (u == nullptr || info == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    if (u->lazy_identity) {  // the raw pointer is handed out: the data must be there
        DeviceCtx* c = nullptr;
        if (int rc = get_ctx(u->device, &c)) return rc;
        if (int rc = materialize_(u, c)) return rc;
    }
    info->n_qubits = u->n;
    info->device   = u->device;
    info->col0     = u->col0;
    info->ncols    = u->ncols;
    info->nrows    = 1ull << u->n;
    info->bytes    = u->bytes();
    info->data     = u->data;
    return QSX_OK;
}

int qsx_unitary_upload(qsx_unitary* u, uint64_t row0, uint64_t nrows, const double* host) {
    if (u == nullptr || host == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    if (nrows > (1ull << u->n) || row0 > (1ull << u->n) - nrows) return fail(QSX_ERR_BAD_ARG, "row range out of bounds");
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(u->device, &c)) return rc;
    if (row0 == 0 && nrows == (1ull << u->n)) u->lazy_identity = false;  // everything is overwritten
    if (int rc = materialize_(u, c)) return rc;
    QSX_CUDA(cudaMemcpyAsync(u->data + row0 * u->ncols, host, nrows * u->ncols * 16ull, cudaMemcpyHostToDevice, c->stream));
    QSX_CUDA(cudaStreamSynchronize(c->stream));
    return QSX_OK;
}

int qsx_unitary_download(const qsx_unitary* u, uint64_t row0, uint64_t nrows, double* host) {
    if (u == nullptr || host == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    if (nrows > (1ull << u->n) || row0 > (1ull << u->n) - nrows) return fail(QSX_ERR_BAD_ARG, "row range out of bounds");
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(u->device, &c)) return rc;
    if (int rc = materialize_(u, c)) return rc;
    QSX_CUDA(cudaMemcpyAsync(host, u->data + row0 * u->ncols, nrows * u->ncols * 16ull, cudaMemcpyDeviceToHost, c->stream));
    QSX_CUDA(cudaStreamSynchronize(c->stream));
    return QSX_OK;
}

namespace {
int copy_block(const qsx_unitary* u, uint64_t row0, uint64_t nrows, uint64_t c0, uint64_t nc, double* host, uint64_t host_ld,
               bool to_host) {
    if (u == nullptr || host == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    if (nrows > (1ull << u->n) || row0 > (1ull << u->n) - nrows) return fail(QSX_ERR_BAD_ARG, "row range out of bounds");
    if (nc > u->ncols || c0 > u->ncols - nc) return fail(QSX_ERR_BAD_ARG, "column range out of bounds (columns are shard-local)");
    if (host_ld == 0) host_ld = nc;
    if (host_ld < nc) return fail(QSX_ERR_BAD_ARG, "host leading dimension smaller than the block width");
    if (nrows == 0 || nc == 0) return QSX_OK;
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(u->device, &c)) return rc;
    if (to_host || !(nrows == (1ull << u->n) && nc == u->ncols)) {
        if (int rc = materialize_(u, c)) return rc;
    } else {
        const_cast<qsx_unitary*>(u)->lazy_identity = false;  // everything is overwritten
    }
    double2* const dev = u->data + row0 * u->ncols + c0;
    if (to_host)
        QSX_CUDA(cudaMemcpy2DAsync(host, host_ld * 16ull, dev, u->ncols * 16ull, nc * 16ull, nrows, cudaMemcpyDeviceToHost, c->stream));
    else
        QSX_CUDA(cudaMemcpy2DAsync(dev, u->ncols * 16ull, host, host_ld * 16ull, nc * 16ull, nrows, cudaMemcpyHostToDevice, c->stream));
    QSX_CUDA(cudaStreamSynchronize(c->stream));
    return QSX_OK;
}
}  // namespace

int qsx_unitary_download_block(const qsx_unitary* u, uint64_t row0, uint64_t nrows, uint64_t c0, uint64_t nc, double* host,
                               uint64_t host_ld) {
    return copy_block(u, row0, nrows, c0, nc, host, host_ld, true);
}

int qsx_unitary_upload_block(qsx_unitary* u, uint64_t row0, uint64_t nrows, uint64_t c0, uint64_t nc, const double* host,
                             uint64_t host_ld) {
    return copy_block(u, row0, nrows, c0, nc, const_cast<double*>(host), host_ld, false);
}

// ---- construction -----------------------------------------------------------------------

int qsx_plan_create(int n_qubits, uint64_t ncols, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
                    qsx_plan** out) {
    if (out == nullptr) return fail(QSX_ERR_BAD_ARG, "null out-pointer");
    *out    = nullptr;
    auto* p = new (std::nothrow) qsx_plan;
    if (p == nullptr) return fail(QSX_ERR_OOM, "host allocation failed");
    std::string msg;
    int rc;
    try {
        rc = qsx::build_plan(n_qubits, ncols, gates, n_gates, opt, p->plan, msg);
    } catch (std::bad_alloc const&) {
        rc  = QSX_ERR_OOM;
        msg = "host allocation failed while planning";
    }
    if (rc != QSX_OK) {
        delete p;
        return fail(rc, msg);
    }
    *out = p;
    return QSX_OK;
}

int qsx_plan_destroy(qsx_plan* plan) {
    if (plan == nullptr) return QSX_OK;
    for (auto& kv : plan->device_blob) {
        if (cudaSetDevice(kv.first) == cudaSuccess) cudaFree(kv.second);
    }
    delete plan;
    return QSX_OK;
}

int qsx_plan_compile(const qsx_plan* cplan, int wait) {
    if (cplan == nullptr) return fail(QSX_ERR_BAD_ARG, "null plan");
    auto* plan = const_cast<qsx_plan*>(cplan);
    {
        std::lock_guard<std::mutex> lk(plan->mu);
        if (plan->plan.cfg.jit == 0) plan->plan.cfg.jit = 1;  // an explicit request overrides the size heuristic
        request_jit_kernels(plan->plan);
    }
    if (wait)
        for (void* k : plan->plan.jit_kernels)
            if (k != nullptr) qsx::jit::wait(static_cast<qsx::jit::Kernel*>(k));
    return QSX_OK;
}

int qsx_jit_get_stats(qsx_jit_stats* out) {
    if (out == nullptr) return fail(QSX_ERR_BAD_ARG, "null stats");
    qsx::jit::Stats const st = qsx::jit::stats();
    out->requested = st.requested, out->cache_hits = st.hits, out->compiled = st.compiled, out->failed = st.failed;
    out->launches = st.launches, out->compile_seconds = st.compile_seconds, out->workers = st.workers;
    out->available = qsx::jit::available(nullptr) ? 1 : 0;
    out->disk_hits = st.disk_hits;
    return QSX_OK;
}

int qsx_plan_get_stats(const qsx_plan* plan, qsx_plan_stats* stats) {
    if (plan == nullptr || stats == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    *stats = plan->plan.stats;
    return QSX_OK;
}

int qsx_plan_dump(const qsx_plan* plan, char* buf, size_t buf_size, size_t* needed) {
    if (plan == nullptr) return fail(QSX_ERR_BAD_ARG, "null plan");
    std::string const s = qsx::dump_plan_json(plan->plan);
    if (needed != nullptr) *needed = s.size() + 1;
    if (buf != nullptr && buf_size > 0) {
        size_t const k = std::min(buf_size - 1, s.size());
        std::memcpy(buf, s.data(), k);
        buf[k] = '\0';
    }
    return QSX_OK;
}

namespace {
// enqueues the sweeps of a plan whose programs are at dblob (device); synchronises unless QSX_RUN_ASYNC
int run_sweeps(qsx::Plan& pl, qsx_unitary* u, DeviceCtx* c, unsigned char* dblob, qsx_stop_fn should_stop,
               void* stop_arg, uint32_t flags, qsx_run_stats* stats) {
    request_jit_kernels(pl);
    // QSX_RUN_ASYNC_IF_SHORT: under 64 GB of HBM traffic (~12 ms; the amount after which a run with a stop callback drains
    // the stream for the first time, below) the call returns once the sweeps are enqueued
    bool const async = (flags & QSX_RUN_ASYNC) != 0 || ((flags & QSX_RUN_ASYNC_IF_SHORT) != 0 && pl.stats.alg_bytes_per_run < 64e9);
    if (!async) QSX_CUDA(cudaEventRecord(c->ev0, c->stream));
    uint64_t launches = 0, jit_launches = 0;
    // Column blocking through the L2 (plan option l2_block_mib): all sweeps run on one block of columns before
    // the next block is touched.  Columns are independent, so any order of (block, sweep) pairs that keeps the
    // sweep order within a block is the same computation.
    uint64_t const block_cols = (pl.cfg.block_cols != 0 && pl.cfg.block_cols < u->ncols) ? pl.cfg.block_cols : u->ncols;
    bool identity_per_block   = false;
    if (u->lazy_identity && !pl.sweeps.empty()) {
        // an interpreted sweep that touches every tile can start from the synthesised identity
        qsx::DevSweep const& first = *reinterpret_cast<qsx::DevSweep const*>(pl.blob.data() + pl.sweep_offset[0]);
        // (a specialised kernel synthesises the identity for any register count and handler set)
        auto* const k0 = pl.jit_kernels.empty() ? nullptr : static_cast<qsx::jit::Kernel*>(pl.jit_kernels[0]);
        bool const jit_first = k0 != nullptr && first.dense_k == 0 && first.fixed_one == 0 &&
                               (pl.cfg.jit == 2 ? qsx::jit::wait(k0) : qsx::jit::ready(k0));
        if (jit_first || qsx::sweep_can_start_from_identity(pl.cfg.R, first)) {
            identity_per_block = true;
            u->lazy_identity   = false;
        } else {
            if (int rc = materialize_(u, c)) return rc;
            ++launches;  // (the identity kernel counts as one of this run's launches)
        }
    }
    // with a stop callback, drain the stream every few sweeps so SIGINT is honoured promptly
    // (the reference polls stop_requested() once per gate, qcir_to_tensor.cpp:174-177)
    double pending_bytes = 0.0;
    double const share   = (double)block_cols / (double)u->ncols;
    for (uint64_t cb0 = 0; cb0 < u->ncols; cb0 += block_cols) {
        for (size_t s = 0; s < pl.sweeps.size(); ++s) {
            if (should_stop != nullptr) {
                if (pending_bytes > 64e9) {  // ~10 ms of HBM traffic
                    QSX_CUDA(cudaStreamSynchronize(c->stream));
                    pending_bytes = 0.0;
                }
                if (should_stop(stop_arg)) {
                    cudaStreamSynchronize(c->stream);
                    if (stats != nullptr) stats->launches = launches;
                    if (identity_per_block) {
                        // blocks that were not reached yet hold no data: leave a defined state behind
                        u->lazy_identity = true;
                    }
                    return fail(QSX_ERR_INTERRUPTED, "interrupted by the stop callback");
                }
            }
            qsx::DevSweep const& hs = *reinterpret_cast<qsx::DevSweep const*>(pl.blob.data() + pl.sweep_offset[s]);
            bool const from_identity = identity_per_block && s == 0;
            bool launched            = false;
            auto* const k = pl.jit_kernels.empty() ? nullptr : static_cast<qsx::jit::Kernel*>(pl.jit_kernels[s]);
            if (k != nullptr && (pl.cfg.jit == 2 ? qsx::jit::wait(k) : qsx::jit::ready(k))) {
                qsx::SweepGeometry g;
                if (qsx::sweep_geometry(u->ncols, block_cols, pl.cfg.R, hs, g)) {
                    uint32_t const fl   = ((s & 1) != 0 ? 1u : 0u) | (from_identity ? 2u : 0u);
                    bool const pair     = qsx::jit::cta_threads(k) * 2 == g.threads;  // two CTAs per tile, half of the tile each
                    cudaError_t const e = qsx::jit::launch(k, u->data + cb0, g.log2_ncols, g.log2_ncb, fl, (uint32_t)(u->col0 + cb0),
                                                           pair ? g.grid * 2 : g.grid, pair ? 256 : g.threads,
                                                           pair ? g.tile_bytes / 2 : g.tile_bytes, c->stream);
                    if (e == cudaSuccess) {
                        launched = true;
                        ++jit_launches;
                    } else if (e != cudaErrorNotReady) {
                        return fail(QSX_ERR_CUDA, "specialised sweep: " + qsx::jit::last_compile_error());
                    }
                }
            }
            if (!launched) {
                if (from_identity && !qsx::sweep_can_start_from_identity(pl.cfg.R, hs))
                    return fail(QSX_ERR_CUDA, "the specialised first sweep became unavailable after the identity was skipped");
                // serpentine: every other sweep walks the tiles backwards, so it starts with what the previous
                // sweep wrote last and is still in the 126 MB L2 (n = 12, memory-bound sweeps: 77.7 -> 71.6 us)
                QSX_CUDA(qsx::launch_sweep(u->data + cb0, u->n, u->ncols, block_cols, pl.cfg.R, hs, dblob + pl.sweep_offset[s],
                                           pl.sweep_size[s], (s & 1) != 0, from_identity, u->col0 + cb0, c->stream));
            }
            pending_bytes += pl.sweeps[s].alg_bytes * share;
            ++launches;
        }
    }
    if (stats != nullptr) stats->launches = launches, stats->jit_launches = jit_launches;
    if (!async) {
        QSX_CUDA(cudaEventRecord(c->ev1, c->stream));
        QSX_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        QSX_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (stats != nullptr) stats->device_ms = ms;
    }
    return QSX_OK;
}
}  // namespace

namespace {
int run_plan(qsx_plan* plan, qsx_unitary* u, qsx_stop_fn should_stop, void* stop_arg, uint32_t flags, qsx_run_stats* stats) {
    qsx::Plan& pl = plan->plan;
    if (pl.cfg.n != u->n || pl.cfg.ncols != u->ncols) return fail(QSX_ERR_SHAPE, "plan was built for a different shard geometry");
    if (stats != nullptr) *stats = qsx_run_stats{};
    DeviceCtx* c = nullptr;
    if (int rc = get_ctx(u->device, &c)) return rc;
    if (pl.sweeps.empty()) return QSX_OK;
    // the stream order and the event pair are per device: one host thread at a time
    std::lock_guard<std::mutex> dev_lock(c->mu);
    unsigned char* dblob = nullptr;
    {
        std::lock_guard<std::mutex> lk(plan->mu);
        auto it = plan->device_blob.find(u->device);
        if (it == plan->device_blob.end()) {
            QSX_CUDA(cudaMalloc(&dblob, pl.blob.size()));
            cudaError_t e = cudaMemcpyAsync(dblob, pl.blob.data(), pl.blob.size(), cudaMemcpyHostToDevice, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // pl.blob is pageable
            if (e != cudaSuccess) {
                cudaFree(dblob);
                return cuda_fail(e, "uploading the plan");
            }
            plan->device_blob.emplace(u->device, dblob);
        } else {
            dblob = it->second;
        }
    }
    return run_sweeps(pl, u, c, dblob, should_stop, stop_arg, flags, stats);
}

void free_plan(qsx_plan* plan) {
    for (auto& kv : plan->device_blob)
        if (cudaSetDevice(kv.first) == cudaSuccess) cudaFree(kv.second);  // (cudaFree waits for the sweeps that read it)
    delete plan;
}
}  // namespace

int qsx_plan_run(const qsx_plan* cplan, qsx_unitary* u, qsx_stop_fn should_stop, void* stop_arg, uint32_t flags,
                 qsx_run_stats* stats) {
    if (cplan == nullptr || u == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return run_plan(const_cast<qsx_plan*>(cplan), u, should_stop, stop_arg, flags, stats);
}

namespace {
// ---- plan cache: `qc2ts` on a circuit that was converted before (the same operand again, a TensorMgr round trip, a
//      repeated check) finds its schedule, the uploaded programs and the specialised kernels instead of re-planning.
//      Key: geometry + options + every gate (controls, targets, qubits, matrix bytes). ----
struct PlanKey {
    uint64_t h0, h1;
    bool operator==(PlanKey const& o) const { return h0 == o.h0 && h1 == o.h1; }
};
std::mutex g_plan_cache_mu;
std::list<std::pair<PlanKey, std::shared_ptr<qsx_plan>>> g_plan_cache;  // most recently used first

PlanKey plan_key(int n, uint64_t ncols, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt) {
    uint64_t a = 0x9e3779b97f4a7c15ull ^ (uint64_t)n, b = 0xc2b2ae3d27d4eb4full ^ ncols;
    auto mix   = [&](const void* p, size_t bytes) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        size_t i               = 0;
        for (; i + 8 <= bytes; i += 8) {
            uint64_t w;
            std::memcpy(&w, c + i, 8);
            a = (a ^ w) * 0x100000001b3ull, a ^= a >> 29;
            b = (b + w) * 0xff51afd7ed558ccdull, b ^= b >> 32;
        }
        for (; i < bytes; ++i) a = (a ^ c[i]) * 0x100000001b3ull, b = (b + c[i]) * 0xff51afd7ed558ccdull;
    };
    qsx_plan_options o{};
    if (opt != nullptr) o = *opt;
    int32_t const has_opt = opt != nullptr;
    mix(&has_opt, sizeof(has_opt));
    mix(&o, sizeof(o));
    for (size_t g = 0; g < n_gates; ++g) {
        int32_t hdr[2 + QSX_MAX_GATE_QUBITS] = {};
        hdr[0] = gates[g].n_ctrls, hdr[1] = gates[g].n_targets;
        int const k = gates[g].n_ctrls + gates[g].n_targets;
        for (int q = 0; q < k && q < QSX_MAX_GATE_QUBITS; ++q) hdr[2 + q] = gates[g].qubits[q];
        mix(hdr, sizeof(hdr));
        if (gates[g].matrix != nullptr && gates[g].n_targets >= 0 && gates[g].n_targets <= 8)
            mix(gates[g].matrix, (size_t)16 << (2 * gates[g].n_targets));
    }
    return PlanKey{a ^ (uint64_t)n_gates, b};
}

size_t plan_cache_capacity() {
    static size_t const cap = [] {
        if (char const* e = std::getenv("QSX_PLAN_CACHE")) return (size_t)std::max(0, std::atoi(e));
        return (size_t)8;
    }();
    return cap;
}

int apply_impl(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt, qsx_stop_fn should_stop,
               void* stop_arg, uint32_t flags, qsx_run_stats* stats) {
    if (u == nullptr) return fail(QSX_ERR_BAD_ARG, "null unitary");
    if (n_gates > 0 && gates == nullptr) return fail(QSX_ERR_BAD_ARG, "null gate array");
    if (stats != nullptr) *stats = qsx_run_stats{};
    std::shared_ptr<qsx_plan> plan;
    PlanKey const key = plan_key(u->n, u->ncols, gates, n_gates, opt);
    size_t const cap  = plan_cache_capacity();
    if (cap > 0) {
        std::lock_guard<std::mutex> lk(g_plan_cache_mu);
        for (auto it = g_plan_cache.begin(); it != g_plan_cache.end(); ++it)
            if (it->first == key) {
                plan = it->second;
                g_plan_cache.splice(g_plan_cache.begin(), g_plan_cache, it);
                break;
            }
    }
    bool const hit = plan != nullptr;
    if (!hit) {
        plan = std::shared_ptr<qsx_plan>(new qsx_plan, free_plan);
        std::string msg;
        int rc;
        try {
            rc = qsx::build_plan(u->n, u->ncols, gates, n_gates, opt, plan->plan, msg);
        } catch (std::bad_alloc const&) {
            rc  = QSX_ERR_OOM;
            msg = "host allocation failed while planning";
        }
        if (rc != QSX_OK) return fail(rc, msg);
        if (cap > 0) {
            std::lock_guard<std::mutex> lk(g_plan_cache_mu);
            g_plan_cache.emplace_front(key, plan);
            while (g_plan_cache.size() > cap) g_plan_cache.pop_back();  // (a plan still running is kept alive by `plan`)
        }
    }
    int const rc = run_plan(plan.get(), u, should_stop, stop_arg, flags, stats);
    if (stats != nullptr) stats->plan_cache_hit = hit ? 1 : 0;
    if ((flags & (QSX_RUN_ASYNC | QSX_RUN_ASYNC_IF_SHORT)) != 0 && cap == 0) {
        // nobody else owns the plan: its programs must outlive the enqueued sweeps
        DeviceCtx* c = nullptr;
        if (get_ctx(u->device, &c) == QSX_OK) cudaStreamSynchronize(c->stream);
    }
    return rc;
}
}  // namespace

int qsx_apply_gates(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
                    qsx_stop_fn should_stop, void* stop_arg, qsx_run_stats* stats) {
    return apply_impl(u, gates, n_gates, opt, should_stop, stop_arg, 0u, stats);
}

int qsx_apply_gates_async(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt,
                          qsx_run_stats* stats) {
    return apply_impl(u, gates, n_gates, opt, nullptr, nullptr, QSX_RUN_ASYNC, stats);
}

int qsx_apply_gates_ex(qsx_unitary* u, const qsx_gate* gates, size_t n_gates, const qsx_plan_options* opt, qsx_stop_fn should_stop,
                       void* stop_arg, uint32_t flags, qsx_run_stats* stats) {
    if ((flags & ~(QSX_RUN_ASYNC | QSX_RUN_ASYNC_IF_SHORT)) != 0) return fail(QSX_ERR_BAD_ARG, "unknown run flag");
    return apply_impl(u, gates, n_gates, opt, should_stop, stop_arg, flags, stats);
}

// ---- verification -----------------------------------------------------------------------

namespace {
// Enqueues the fused reduction of every shard pair (concurrently on all devices involved), combines the per-device
// results with ONE all-reduce of 8 doubles when `combine` is set, and reads 64 bytes back from the first device.
int equiv_impl(const qsx_unitary* const* a, const qsx_unitary* const* b, int n_shards, bool combine, double out[8]) {
    if (a == nullptr || b == nullptr || out == nullptr || n_shards < 1) return fail(QSX_ERR_BAD_ARG, "null argument");
    for (int k = 0; k < n_shards; ++k) {
        if (a[k] == nullptr || b[k] == nullptr) return fail(QSX_ERR_BAD_ARG, "null shard");
        if (a[k]->n != b[k]->n || a[k]->ncols != b[k]->ncols || a[k]->col0 != b[k]->col0 || a[k]->n != a[0]->n)
            return fail(QSX_ERR_SHAPE, "the two tensors should have the same shape");
        if (a[k]->device != b[k]->device) return fail(QSX_ERR_SHAPE, "both shards of a pair must live on the same device");
    }
    // devices in order of first appearance
    std::vector<DeviceCtx*> ctx;
    std::vector<int> blocks_used;  // per device: partial blocks filled so far
    std::vector<int> local_pairs;
    std::vector<int> dev_of((size_t)n_shards, -1);
    for (int k = 0; k < n_shards; ++k) {
        size_t d = 0;
        while (d < ctx.size() && ctx[d]->device != a[k]->device) ++d;
        if (d == ctx.size()) {
            DeviceCtx* c = nullptr;
            if (int rc = get_ctx(a[k]->device, &c)) return rc;
            ctx.push_back(c), blocks_used.push_back(0), local_pairs.push_back(0);
        }
        dev_of[(size_t)k] = (int)d;
        if (++local_pairs[d] > qsx::kMaxLocalShards) return fail(QSX_ERR_BAD_ARG, "too many shard pairs on one device");
    }
    if (!combine && ctx.size() > 1) return fail(QSX_ERR_BAD_ARG, "qsx_equiv_partial takes one shard pair");
    // lock the devices in ascending id order (no lock-order inversion between concurrent callers)
    std::vector<DeviceCtx*> by_id = ctx;
    std::sort(by_id.begin(), by_id.end(), [](DeviceCtx* x, DeviceCtx* y) { return x->device < y->device; });
    std::vector<std::unique_lock<std::mutex>> locks;
    for (DeviceCtx* c : by_id) locks.emplace_back(c->mu);
    size_t const slot = qsx::equiv_partials_doubles();
    for (int k = 0; k < n_shards; ++k) {
        size_t const d = (size_t)dev_of[(size_t)k];
        DeviceCtx* c   = ctx[d];
        QSX_CUDA(cudaSetDevice(c->device));
        if (int rc = materialize_(a[k], c)) return rc;
        if (int rc = materialize_(b[k], c)) return rc;
        int nb = 0;
        QSX_CUDA(qsx::launch_equiv_partials(a[k]->data, b[k]->data, a[k]->elems(), c->scratch + (size_t)blocks_used[d] * 16, &nb, c->stream));
        blocks_used[d] += nb;
    }
    std::vector<double*> out8(ctx.size(), nullptr);
    for (size_t d = 0; d < ctx.size(); ++d) {
        DeviceCtx* c = ctx[d];
        out8[d]      = c->scratch + slot * qsx::kMaxLocalShards;
        QSX_CUDA(cudaSetDevice(c->device));
        QSX_CUDA(qsx::launch_equiv_finish(c->scratch, blocks_used[d], out8[d], c->stream));
    }
    if (combine)
        if (int rc = qsx::comm_allreduce8(ctx.data(), out8.data(), (int)ctx.size())) return rc;
    DeviceCtx* c0 = ctx[0];
    QSX_CUDA(cudaSetDevice(c0->device));
    QSX_CUDA(cudaMemcpyAsync(c0->host8, out8[0], 8 * sizeof(double), cudaMemcpyDeviceToHost, c0->stream));
    // the other devices' streams carry their half of the collective: drain them too before the scratch is reused
    for (size_t d = ctx.size(); d-- > 0;) {
        QSX_CUDA(cudaSetDevice(ctx[d]->device));
        QSX_CUDA(cudaStreamSynchronize(ctx[d]->stream));
    }
    std::memcpy(out, c0->host8, 8 * sizeof(double));
    return QSX_OK;
}
}  // namespace

int qsx_equiv_partial(const qsx_unitary* a, const qsx_unitary* b, double sums[8]) {
    if (a == nullptr || b == nullptr || sums == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    return equiv_impl(&a, &b, 1, false, sums);
}

int qsx_equiv_reduce(const qsx_unitary* const* a, const qsx_unitary* const* b, int n_shards, double sums[8]) {
    return equiv_impl(a, b, n_shards, true, sums);
}

int qsx_equiv_finish(const double sums[8], double eps, int strict, qsx_equiv_result* out) {
    if (sums == nullptr || out == nullptr) return fail(QSX_ERR_BAD_ARG, "null argument");
    std::memset(out, 0, sizeof(*out));
    // inner_product = |sum conj(a) b| ; cosine = ip(a,b) / sqrt(ip(a,a) ip(b,b))   (tensor.hpp:393-403)
    double const ip_ab = std::abs(std::complex<double>(sums[0], sums[1]));
    out->cosine        = ip_ab / std::sqrt(std::abs(sums[2]) * std::abs(sums[3]));
    bool equiv         = out->cosine >= (1 - eps);  // qtensor.hpp:414-417
    std::complex<double> const sa(sums[4], sums[5]), sb(sums[6], sums[7]);
    dvlab::Phase phase;
    if (sa == std::complex<double>(0.0, 0.0)) {
        // sum(t1) == 0: the reference divides by zero and feeds NaN to Phase (UB / endless loop,
        // SURVEY Appendix C trap 2).  Defined behaviour here: norm = inf/nan, phase undefined.
        out->norm          = (sb == std::complex<double>(0.0, 0.0)) ? std::nan("") : INFINITY;
        out->phase_rad     = std::nan("");
        out->phase_defined = 0;
    } else {
        std::complex<double> const s = sb / sa;  // global_scalar_factor, qtensor.hpp:383-385
        out->norm                    = std::abs(s);
        out->phase_rad               = std::arg(s);
        if (std::isfinite(out->phase_rad)) {
            phase              = dvlab::Phase(out->phase_rad);  // eps 1e-4 rad, phase.hpp:38-40
            out->phase_defined = 1;
        }
    }
    out->phase_num = phase.numerator();
    out->phase_den = phase.denominator();
    if (strict) {  // tensor_cmd.cpp:156-160
        if (out->norm > 1 + eps || out->norm < 1 - eps || !out->phase_defined || phase != dvlab::Phase(0)) equiv = false;
    }
    out->equivalent = equiv ? 1 : 0;
    return QSX_OK;
}

int qsx_unitary_adjoint_inplace(qsx_unitary* u) {
    if (u == nullptr) return fail(QSX_ERR_BAD_ARG, "null unitary");
    if (u->col0 != 0 || u->ncols != (1ull << u->n))
        return fail(QSX_ERR_UNSUPPORTED, "one column shard cannot be transposed alone: use qsx_unitary_adjoint_sharded");
    qsx_unitary* one[1] = {u};
    return qsx_unitary_adjoint_sharded(one, 1);
}

// Shard k holds columns [k c, (k+1) c) of the 2^n x 2^n matrix, so its rows [j c, (j+1) c) are
// the contiguous c x c block (j, k).  The adjoint swaps block (j, k) with block (k, j): the
// diagonal blocks transpose locally, every off-diagonal pair is exchanged by ONE kernel that reads
// and writes the partner block directly in the peer GPU's memory (NVLink P2P) -- the transfer and
// the transpose are the same pass, nothing is staged, no second copy of the tensor exists.
int qsx_unitary_adjoint_sharded(qsx_unitary* const* shards, int n_shards) {
    if (shards == nullptr || n_shards < 1) return fail(QSX_ERR_BAD_ARG, "no shards");
    for (int k = 0; k < n_shards; ++k)
        if (shards[k] == nullptr) return fail(QSX_ERR_BAD_ARG, "null shard");
    int const n        = shards[0]->n;
    uint64_t const dim = 1ull << n, c = shards[0]->ncols;
    if (c * (uint64_t)n_shards != dim) return fail(QSX_ERR_SHAPE, "the shards do not cover all columns");
    for (int k = 0; k < n_shards; ++k)
        if (shards[k]->n != n || shards[k]->ncols != c || shards[k]->col0 != (uint64_t)k * c)
            return fail(QSX_ERR_SHAPE, "shard k must hold columns [k c, (k+1) c)");
    std::vector<DeviceCtx*> ctx((size_t)n_shards, nullptr);
    for (int k = 0; k < n_shards; ++k)
        if (int rc = get_ctx(shards[k]->device, &ctx[(size_t)k])) return rc;
    // everything queued on any shard must be done before a peer touches it
    for (int k = 0; k < n_shards; ++k) {
        QSX_CUDA(cudaSetDevice(shards[k]->device));
        if (int rc = materialize_(shards[k], ctx[(size_t)k])) return rc;
        QSX_CUDA(cudaStreamSynchronize(ctx[(size_t)k]->stream));
    }
    for (int i = 0; i < n_shards; ++i) {
        int const di = shards[i]->device;
        QSX_CUDA(cudaSetDevice(di));
        for (int j = i; j < n_shards; ++j) {
            double2* const a = shards[i]->data + (uint64_t)j * c * c;  // block (j, i)
            if (j == i) {
                QSX_CUDA(qsx::launch_adjoint_inplace(a, c, ctx[(size_t)i]->stream));
                continue;
            }
            int const dj     = shards[j]->device;
            double2* const b = shards[j]->data + (uint64_t)i * c * c;  // block (i, j), possibly peer memory
            if (dj == di) {
                QSX_CUDA(qsx::launch_adjoint_block_pair(a, b, c, 0, ctx[(size_t)i]->stream));
                continue;
            }
            for (auto [from, to] : {std::pair<int, int>{di, dj}, std::pair<int, int>{dj, di}}) {
                int can = 0;
                QSX_CUDA(cudaDeviceCanAccessPeer(&can, from, to));
                if (!can) return fail(QSX_ERR_UNSUPPORTED, "devices " + std::to_string(from) + " and " + std::to_string(to) + " have no peer access");
                QSX_CUDA(cudaSetDevice(from));
                cudaError_t const e = cudaDeviceEnablePeerAccess(to, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(QSX_ERR_CUDA, cudaGetErrorString(e));
                cudaGetLastError();
            }
            // both owners work on the pair: each takes half of the tiles, with ITS block as the local one
            QSX_CUDA(cudaSetDevice(di));
            QSX_CUDA(qsx::launch_adjoint_block_pair(a, b, c, 1, ctx[(size_t)i]->stream));
            QSX_CUDA(cudaSetDevice(dj));
            QSX_CUDA(qsx::launch_adjoint_block_pair(b, a, c, 2, ctx[(size_t)j]->stream));
            QSX_CUDA(cudaSetDevice(di));
        }
    }
    for (int k = 0; k < n_shards; ++k) {
        QSX_CUDA(cudaSetDevice(shards[k]->device));
        QSX_CUDA(cudaStreamSynchronize(ctx[(size_t)k]->stream));
    }
    return QSX_OK;
}

}  // extern "C"
