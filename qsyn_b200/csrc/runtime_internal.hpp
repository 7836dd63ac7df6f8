// runtime_internal.hpp -- state shared by the translation units behind include/qsx.h
// (runtime.cu: storage + plan execution, comm.cpp: NCCL, jit.cpp: specialised sweeps).  Internal.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>

#include "../../include/qsx.h"

struct qsx_unitary {
    int      device   = 0;
    int      n        = 0;
    uint64_t col0     = 0;
    uint64_t ncols    = 0;
    double2* data     = nullptr;
    // set_identity is lazy: the first sweep of the next gate stream synthesises the identity instead of
    // reading it (saves one write and one read pass); anything else that looks at the data writes it first
    bool     lazy_identity = false;
    uint64_t elems() const { return (1ull << n) * ncols; }
    uint64_t bytes() const { return elems() * 16ull; }
};

namespace qsx {

// Everything the library owns on one device.  `mu` serialises the host threads that use the device:
// the stream, the event pair and the equivalence scratch are shared by every call on the device, so a call holds
// `mu` from its first use of them to its last launch.
struct DeviceCtx {
    std::mutex   mu;
    int          device   = -1;
    cudaStream_t stream   = nullptr;
    double*      scratch  = nullptr;  // equiv partials of up to kMaxLocalShards shard pairs + 8 outputs
    double*      host8    = nullptr;  // pinned
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    // parked shard buffers (pool_alloc / pool_free): size -> (pointer, parking sequence number)
    struct Parked {
        void*    p;
        uint64_t seq;
    };
    std::mutex                    pool_mu;
    std::multimap<size_t, Parked> pool;
    size_t                        pool_bytes = 0, pool_cap = 0;
    uint64_t                      pool_seq = 0;
};
cudaError_t pool_alloc(DeviceCtx* c, void** out, size_t bytes);
void pool_free(DeviceCtx* c, void* p, size_t bytes);
void pool_trim(DeviceCtx* c);
constexpr int kMaxLocalShards = 64;  // shard pairs of one equivalence check that may share a device

int device_count_nothrow();
int fail(int code, std::string msg);            // sets the thread-local error text, returns code
int cuda_fail(cudaError_t e, char const* what);  // maps cudaErrorMemoryAllocation to QSX_ERR_OOM
int get_ctx(int device, DeviceCtx** out);        // creates the context on first use; leaves `device` current
int materialize(qsx_unitary* u, DeviceCtx* c);   // writes a lazy identity

#define QSX_CUDA(call)                                               \
    do {                                                             \
        cudaError_t const e__ = (call);                              \
        if (e__ != cudaSuccess) return ::qsx::cuda_fail(e__, #call); \
    } while (0)

// comm.cpp: all-reduce (SUM) of the 8 doubles at out8[k] (device memory of devices[k], in place) over the devices of
// this process and, when qsx_comm_init_rank was called, over the ranks of the job; enqueued on each device's stream
int comm_allreduce8(DeviceCtx* const* ctx, double* const* out8, int n_devices);
// number of processes of the job (1 unless qsx_comm_init_rank was called)
int comm_world_size();

}  // namespace qsx
