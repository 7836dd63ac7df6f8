// device_unitary.hpp -- RAII owner of a device-resident unitary: all column shards of
// U[out,in] (include/qsx.h), one shard per GPU of this process.  This is what replaces the
// xt::xarray storage inside Tensor<std::complex<double>> (reference src/tensor/tensor.hpp:38,144)
// for the n-qubit main tensor of `qc2ts`.  Value semantics like the reference's tensors:
// copy = deep copy on the device (TensorMgr copies, data_structure_manager.hpp:138-150),
// move = steal.  CUDA is reached only through the C ABI.
#pragma once

#include <array>
#include <complex>
#include <cstdlib>
#include <new>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/qsx.h"

namespace qsyn::tensor {

// One process per GPU (mpirun / torchrun style deployments): every process runs the same commands and holds the
// column block [rank 2^n / world, (rank+1) 2^n / world) of every large tensor on `device`; `tensor equiv` then
// all-reduces over the ranks (qsx_comm_init_rank).  Default: one process, QSX_NUM_GPUS devices.
struct ProcessShard {
    int rank = 0, world = 1, device = 0;
};
inline ProcessShard& process_shard() {
    static ProcessShard p;
    return p;
}

class DeviceUnitary {
public:
    DeviceUnitary() = default;
    ~DeviceUnitary() { release(); }
    DeviceUnitary(DeviceUnitary&& o) noexcept : _n(o._n), _cols(o._cols), _across_processes(o._across_processes), _shards(std::move(o._shards)) {
        o._n = 0;
        o._shards.clear();
    }
    DeviceUnitary& operator=(DeviceUnitary&& o) noexcept {
        if (this != &o) {
            release();
            _n      = o._n;
            _cols   = o._cols;
            _across_processes = o._across_processes;
            _shards = std::move(o._shards);
            o._n    = 0;
            o._shards.clear();
        }
        return *this;
    }
    DeviceUnitary(DeviceUnitary const& o) : _n(o._n), _cols(o._cols), _across_processes(o._across_processes) {
        try {
            for (auto* s : o._shards) {
                qsx_unitary* c = nullptr;
                check(qsx_unitary_clone(s, &c));
                _shards.push_back(c);
            }
        } catch (...) {
            release();  // (the destructor does not run for a constructor that throws)
            throw;
        }
    }
    DeviceUnitary& operator=(DeviceUnitary const& o) {
        if (this != &o) *this = DeviceUnitary(o);
        return *this;
    }

    bool resident() const { return !_shards.empty(); }
    int n_qubits() const { return _n; }
    size_t n_shards() const { return _shards.size(); }
    size_t dim() const { return size_t{1} << _n; }

    // number of column shards for an n-qubit unitary: QSX_NUM_GPUS (default 1) clipped to the
    // devices present and to >= 64 columns per shard
    static int shard_count(int n_qubits) {
        int want = 1;
        if (char const* e = std::getenv("QSX_NUM_GPUS")) want = std::atoi(e);
        int have = 0;
        qsx_device_count(&have);
        if (want > have) want = have;
        int g = 1;
        while (g * 2 <= want && (size_t{1} << n_qubits) / (size_t)(g * 2) >= 64) g *= 2;
        return g;
    }

    static DeviceUnitary identity(int n_qubits) {
        DeviceUnitary u = uninitialised(n_qubits);
        for (auto* s : u._shards) check(qsx_unitary_set_identity(s));
        return u;
    }

    static DeviceUnitary from_host(int n_qubits, std::complex<double> const* row_major) {
        DeviceUnitary u   = uninitialised(n_qubits);
        size_t const dim  = u.dim();
        size_t const g    = u._shards.size();
        size_t const cols = u._cols;
        // shard k = the column block [first + k cols, first + (k+1) cols) of the row-major matrix: one strided copy each
        size_t const first = u._across_processes ? (size_t)process_shard().rank * cols : 0;
        for (size_t k = 0; k < g; ++k)
            check(qsx_unitary_upload_block(u._shards[k], 0, dim, 0, cols, reinterpret_cast<double const*>(row_major + first + k * cols), dim));
        return u;
    }

    // U <- gates . U on every shard (same stream of gates, no communication)
    // returns the qsx status (QSX_OK, QSX_ERR_INTERRUPTED, QSX_ERR_UNSUPPORTED, ...)
    int apply(std::vector<qsx_gate> const& gates, qsx_stop_fn should_stop = nullptr, void* stop_arg = nullptr) {
        if (_shards.empty()) return QSX_ERR_BAD_ARG;
        if (_shards.size() == 1) {
            // a short stream (milliseconds of device time) is only enqueued: the caller's next host work -- reading and
            // lowering the other operand of a check -- overlaps it, and every later use of the tensor is ordered behind it
            // on the device's stream; a long one is waited for, with the stop callback polled between sweeps
            int const rc = qsx_apply_gates_ex(_shards[0], gates.data(), gates.size(), nullptr, should_stop, stop_arg,
                                              QSX_RUN_ASYNC_IF_SHORT, nullptr);
            if (rc == QSX_ERR_OOM) throw std::bad_alloc();
            return rc;
        }
        // several GPUs in one process: schedule once, enqueue on every device (the devices then work concurrently),
        // wait for all.  The geometry comes from _n and the shard count: asking a shard for its info would
        // materialise its lazy identity.
        uint64_t const ncols = _cols;
        qsx_plan* plan       = nullptr;
        int rc               = qsx_plan_create(_n, ncols, gates.data(), gates.size(), nullptr, &plan);
        if (rc != QSX_OK) return rc;
        for (auto* s : _shards) {
            if (should_stop != nullptr && should_stop(stop_arg)) rc = QSX_ERR_INTERRUPTED;
            if (rc == QSX_OK) rc = qsx_plan_run(plan, s, nullptr, nullptr, QSX_RUN_ASYNC, nullptr);
        }
        // shard k lives on device k (uninitialised()).  With a stop callback the wait polls it: the reference checks
        // stop_requested() once per gate (qcir_to_tensor.cpp:174-177); the enqueued sweeps cannot be recalled, but the
        // caller learns about the interruption as soon as the devices are idle and the result is discarded.
        for (size_t k = 0; k < _shards.size(); ++k) {
            int const src = qsx_device_synchronize(static_cast<int>(k));  // (several shards: shard k is on device k)
            if (rc == QSX_OK) rc = src;
        }
        if (rc == QSX_OK && should_stop != nullptr && should_stop(stop_arg)) rc = QSX_ERR_INTERRUPTED;
        qsx_plan_destroy(plan);
        if (rc == QSX_ERR_OOM) throw std::bad_alloc();
        return rc;
    }

    // {Re,Im sum conj(a)b, sum|a|^2, sum|b|^2, Re,Im sum a, Re,Im sum b} over all shards
    std::array<double, 8> equiv_sums(DeviceUnitary const& other) const {
        if (_n != other._n || _shards.size() != other._shards.size() || _cols != other._cols)
            throw std::invalid_argument("The two tensors should have the same shape");
        // every device reduces its shard pair concurrently; ONE ncclAllReduce of 8 doubles combines them.
        // A tensor that every process holds in full (too small to shard) is reduced locally.
        std::array<double, 8> tot{};
        bool const local = _shards.size() == 1 && !_across_processes;
        int const rc     = local ? qsx_equiv_partial(_shards[0], other._shards[0], tot.data())
                                 : qsx_equiv_reduce(_shards.data(), other._shards.data(), static_cast<int>(_shards.size()), tot.data());
        if (rc == QSX_ERR_SHAPE) throw std::invalid_argument("The two tensors should have the same shape");
        check(rc);
        return tot;
    }

    // rows [row0, row0+nrows) of the whole matrix (all shards), row-major nrows x 2^n: what print / write stream
    void download_rows(size_t row0, size_t nrows, std::complex<double>* out) const {
        require_local("reading a whole tensor");
        size_t const dim = this->dim(), g = _shards.size(), cols = _cols;
        for (size_t k = 0; k < g; ++k)
            check(qsx_unitary_download_block(_shards[k], row0, nrows, 0, cols, reinterpret_cast<double*>(out + k * cols), dim));
    }
    // one element U[i, j] (decomposer.hpp:64-69 style access)
    std::complex<double> element(size_t i, size_t j) const {
        require_local("element access");
        size_t const cols = _cols;
        std::complex<double> v;
        check(qsx_unitary_download_block(_shards[j / cols], i, 1, j % cols, 1, reinterpret_cast<double*>(&v), 1));
        return v;
    }

    void adjoint_inplace() {
        require_local("tensor adjoint");
        check(qsx_unitary_adjoint_sharded(_shards.data(), static_cast<int>(_shards.size())));
    }

    // whole matrix, row-major (print / write / element access; small n only)
    std::vector<std::complex<double>> download() const {
        require_local("reading a whole tensor");
        size_t const dim = this->dim(), g = _shards.size(), cols = _cols;
        std::vector<std::complex<double>> out(dim * dim);
        for (size_t k = 0; k < g; ++k)
            check(qsx_unitary_download_block(_shards[k], 0, dim, 0, cols, reinterpret_cast<double*>(out.data() + k * cols), dim));
        return out;
    }

private:
    static DeviceUnitary uninitialised(int n_qubits) {
        DeviceUnitary u;
        u._n                  = n_qubits;
        ProcessShard const ps = process_shard();
        uint64_t const dim    = uint64_t{1} << n_qubits;
        if (ps.world > 1) {
            // one process per GPU: this process's column block, or the whole (small) tensor on every process
            bool const split = (ps.world & (ps.world - 1)) == 0 && dim / (uint64_t)ps.world >= 64;
            u._cols          = split ? dim / (uint64_t)ps.world : dim;
            u._across_processes = split;
            qsx_unitary* s   = nullptr;
            check(qsx_unitary_create(ps.device, n_qubits, split ? (uint64_t)ps.rank * u._cols : 0, u._cols, &s));
            u._shards.push_back(s);
            return u;
        }
        int const g = shard_count(n_qubits);
        u._cols     = dim / (uint64_t)g;
        for (int k = 0; k < g; ++k) {
            qsx_unitary* s = nullptr;
            check(qsx_unitary_create(k, n_qubits, (uint64_t)k * u._cols, u._cols, &s));
            u._shards.push_back(s);
        }
        return u;
    }
    static void check(int rc) {
        if (rc == QSX_OK) return;
        if (rc == QSX_ERR_OOM) throw std::bad_alloc();  // -> "Memory allocation failed!!"
        throw std::runtime_error(std::string("qsx: ") + qsx_last_error());
    }
    void release() {
        for (auto* s : _shards) qsx_unitary_destroy(s);
        _shards.clear();
    }

    void require_local(char const* what) const {
        if (_across_processes)
            throw std::runtime_error(std::string(what) + " needs all columns in this process: the tensor is sharded over " +
                                     std::to_string(process_shard().world) + " processes");
    }

    int _n                 = 0;
    uint64_t _cols         = 0;      // columns per shard
    bool _across_processes = false;  // this process holds ONE column block of the tensor (ProcessShard)
    std::vector<qsx_unitary*> _shards;
};

}  // namespace qsyn::tensor
