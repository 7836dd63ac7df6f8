// comm.cpp -- the ONE collective of the path: all-reduce (SUM) of the 8 equivalence sums over the column shards
// (SURVEY 8e, 8b items 1 and 4).  NCCL is loaded at run time (dlopen "libnccl.so.2"): libqsx.so has no link-time
// dependency on it, a single-GPU host never touches it, and inside a process that already carries an NCCL (bench.py
// under torchrun) the same library instance is reused (glibc matches by soname).
//
//   one process, G devices   qsx_init(G) / first qsx_equiv_reduce over > 1 device: ncclCommInitAll on the device set
//   one process per GPU      qsx_comm_get_unique_id on rank 0, handed to every rank out of band (torchrun's store),
//                            qsx_comm_init_rank on every rank: the all-reduce then also spans the ranks
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <cstdio>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "jit.hpp"
#include "runtime_internal.hpp"

namespace qsx {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*)                                                                             = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                                   = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*)                                                    = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                            = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                                      = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)()                                                                                 = nullptr;
    ncclResult_t (*GroupEnd)()                                                                                   = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                                                                  = nullptr;
};

std::mutex g_comm_mu;
NcclApi g_nccl;
bool g_nccl_tried = false;
std::string g_nccl_why;

// device set -> communicators of one ncclCommInitAll (in the order of the key)
std::map<std::vector<int>, std::vector<ncclComm_t>> g_cliques;
// one process per GPU: this process's communicator of the job-wide clique
ncclComm_t g_rank_comm   = nullptr;
int g_rank_device        = -1;
int g_rank_world         = 1;

}  // namespace
int comm_world_size() { return g_rank_comm != nullptr ? g_rank_world : 1; }
namespace {

int load_nccl() {
    if (g_nccl.handle != nullptr) return QSX_OK;
    if (g_nccl_tried) return fail(QSX_ERR_NCCL, g_nccl_why);
    g_nccl_tried = true;
    for (char const* name : {"libnccl.so.2", "libnccl.so"}) {
        g_nccl.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.handle != nullptr) break;
    }
    if (g_nccl.handle == nullptr) {
        g_nccl_why = std::string("NCCL is not loadable: ") + dlerror();
        return fail(QSX_ERR_NCCL, g_nccl_why);
    }
    bool ok  = true;
    auto sym = [&](char const* name) {
        void* p = dlsym(g_nccl.handle, name);
        if (p == nullptr) ok = false;
        return p;
    };
    g_nccl.GetVersion     = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
    g_nccl.GetUniqueId    = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitAll    = reinterpret_cast<decltype(g_nccl.CommInitAll)>(sym("ncclCommInitAll"));
    g_nccl.CommInitRank   = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy    = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllReduce      = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.GroupStart     = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
    g_nccl.GroupEnd       = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
        dlclose(g_nccl.handle);
        g_nccl.handle = nullptr;
        g_nccl_why    = "the loaded libnccl lacks a required entry point";
        return fail(QSX_ERR_NCCL, g_nccl_why);
    }
    return QSX_OK;
}

int nccl_fail(ncclResult_t r, char const* what) {
    return fail(QSX_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString != nullptr ? g_nccl.GetErrorString(r) : "NCCL error"));
}

#define QSX_NCCL(call)                                        \
    do {                                                      \
        ncclResult_t const r__ = (call);                      \
        if (r__ != ncclSuccess) return ::qsx::nccl_fail(r__, #call); \
    } while (0)

// Communicators must be destroyed while the CUDA runtime is still alive: a process that simply returns from main with
// live communicators crashes in NCCL's teardown (seen as SIGSEGV after `quit -f` with QSX_NUM_GPUS=2).  atexit handlers
// run in reverse order of registration, and this one is registered after the runtime's own (the first CUDA call is
// long past when a communicator is created), so it runs before the runtime shuts down.
void destroy_comms_at_exit() { qsx_shutdown(); }
void register_exit_hook() {
    static bool const once = (std::atexit(destroy_comms_at_exit), true);
    (void)once;
}

// NCCL writes its version banner to STDOUT when a communicator is created (with NCCL_DEBUG=VERSION, which GPU images
// commonly export).  Stdout is the product's text interface -- `tensor equiv` must print exactly what qsyn prints -- so
// file descriptor 1 points at stderr while NCCL initialises.
struct StdoutToStderr {
    int saved = -1;
    StdoutToStderr() {
        std::fflush(stdout);
        saved = dup(1);
        if (saved >= 0) dup2(2, 1);
    }
    ~StdoutToStderr() {
        if (saved < 0) return;
        std::fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

int clique_for(std::vector<int> const& devices, std::vector<ncclComm_t>** out) {
    auto it = g_cliques.find(devices);
    if (it == g_cliques.end()) {
        if (int rc = load_nccl()) return rc;
        register_exit_hook();
        std::vector<ncclComm_t> comms(devices.size(), nullptr);
        StdoutToStderr quiet;
        QSX_NCCL(g_nccl.CommInitAll(comms.data(), (int)devices.size(), devices.data()));
        it = g_cliques.emplace(devices, std::move(comms)).first;
    }
    *out = &it->second;
    return QSX_OK;
}

}  // namespace

int comm_allreduce8(DeviceCtx* const* ctx, double* const* out8, int n_devices) {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (n_devices < 1) return fail(QSX_ERR_BAD_ARG, "no devices");
    if (g_rank_comm != nullptr && g_rank_world > 1) {
        // one process per GPU: the job-wide clique; this process contributes exactly one device
        if (n_devices != 1 || ctx[0]->device != g_rank_device)
            return fail(QSX_ERR_NCCL, "after qsx_comm_init_rank every shard of this process must live on the rank's device");
        QSX_CUDA(cudaSetDevice(g_rank_device));
        QSX_NCCL(g_nccl.AllReduce(out8[0], out8[0], 8, ncclDouble, ncclSum, g_rank_comm, ctx[0]->stream));
        return QSX_OK;
    }
    if (n_devices == 1) return QSX_OK;
    std::vector<int> devices((size_t)n_devices);
    for (int k = 0; k < n_devices; ++k) devices[(size_t)k] = ctx[k]->device;
    std::vector<ncclComm_t>* comms = nullptr;
    if (int rc = clique_for(devices, &comms)) return rc;
    QSX_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < n_devices; ++k) {
        ncclResult_t const r = g_nccl.AllReduce(out8[k], out8[k], 8, ncclDouble, ncclSum, (*comms)[(size_t)k], ctx[k]->stream);
        if (r != ncclSuccess) {
            g_nccl.GroupEnd();
            return nccl_fail(r, "ncclAllReduce");
        }
    }
    QSX_NCCL(g_nccl.GroupEnd());
    return QSX_OK;
}

}  // namespace qsx

extern "C" {

int qsx_init(int n_gpus) {
    int have = qsx::device_count_nothrow();
    if (have == 0) return qsx::fail(QSX_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (n_gpus <= 0 || n_gpus > have) n_gpus = have;
    std::vector<int> devices;
    for (int d = 0; d < n_gpus; ++d) {
        qsx::DeviceCtx* c = nullptr;
        if (int rc = qsx::get_ctx(d, &c)) return rc;
        devices.push_back(d);
    }
    // peer access between every pair (the sharded adjoint reads and writes peer memory)
    for (int a = 0; a < n_gpus; ++a) {
        for (int b = 0; b < n_gpus; ++b) {
            if (a == b) continue;
            int can = 0;
            QSX_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
            if (!can) continue;
            QSX_CUDA(cudaSetDevice(a));
            cudaError_t const e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return qsx::cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
    }
    if (n_gpus > 1) {
        std::lock_guard<std::mutex> lk(qsx::g_comm_mu);
        if (qsx::g_rank_comm == nullptr) {
            std::vector<ncclComm_t>* comms = nullptr;
            if (int rc = qsx::clique_for(devices, &comms)) return rc;
        }
    }
    QSX_CUDA(cudaSetDevice(0));
    return QSX_OK;
}

int qsx_shutdown(void) {
    qsx::jit::shutdown();  // no compiler thread may be running when the process tears its libraries down
    std::lock_guard<std::mutex> lk(qsx::g_comm_mu);
    if (qsx::g_nccl.handle != nullptr) {
        for (auto& kv : qsx::g_cliques)
            for (ncclComm_t c : kv.second)
                if (c != nullptr) qsx::g_nccl.CommDestroy(c);
        if (qsx::g_rank_comm != nullptr) qsx::g_nccl.CommDestroy(qsx::g_rank_comm);
    }
    qsx::g_cliques.clear();
    qsx::g_rank_comm = nullptr, qsx::g_rank_device = -1, qsx::g_rank_world = 1;
    return QSX_OK;
}

int qsx_comm_get_unique_id(void* id, size_t id_size) {
    if (id == nullptr || id_size < sizeof(ncclUniqueId)) return qsx::fail(QSX_ERR_BAD_ARG, "the id buffer must hold QSX_COMM_ID_BYTES bytes");
    std::lock_guard<std::mutex> lk(qsx::g_comm_mu);
    if (int rc = qsx::load_nccl()) return rc;
    ncclUniqueId uid;
    QSX_NCCL(qsx::g_nccl.GetUniqueId(&uid));
    std::memcpy(id, &uid, sizeof(uid));
    return QSX_OK;
}

int qsx_comm_init_rank(int device, int world, int rank, const void* id, size_t id_size) {
    if (id == nullptr || id_size < sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world)
        return qsx::fail(QSX_ERR_BAD_ARG, "bad communicator arguments");
    qsx::DeviceCtx* c = nullptr;
    if (int rc = qsx::get_ctx(device, &c)) return rc;
    std::lock_guard<std::mutex> lk(qsx::g_comm_mu);
    if (qsx::g_rank_comm != nullptr) return qsx::fail(QSX_ERR_BAD_ARG, "qsx_comm_init_rank was already called (qsx_shutdown first)");
    if (int rc = qsx::load_nccl()) return rc;
    qsx::register_exit_hook();
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    qsx::StdoutToStderr quiet;
    QSX_NCCL(qsx::g_nccl.CommInitRank(&qsx::g_rank_comm, world, uid, rank));
    qsx::g_rank_device = device, qsx::g_rank_world = world;
    return QSX_OK;
}

int qsx_comm_info(int* world, int* nccl_version) {
    std::lock_guard<std::mutex> lk(qsx::g_comm_mu);
    if (world != nullptr) *world = qsx::g_rank_comm != nullptr ? qsx::g_rank_world : 1;
    if (nccl_version != nullptr) {
        *nccl_version = 0;
        if (qsx::load_nccl() == QSX_OK) qsx::g_nccl.GetVersion(nccl_version);
    }
    return QSX_OK;
}

}  // extern "C"
