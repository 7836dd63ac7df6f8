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

class DeviceUnitary {
public:
    DeviceUnitary() = default;
    ~DeviceUnitary() { release(); }
    DeviceUnitary(DeviceUnitary&& o) noexcept : _n(o._n), _shards(std::move(o._shards)) {
        o._n = 0;
        o._shards.clear();
    }
    DeviceUnitary& operator=(DeviceUnitary&& o) noexcept {
        if (this != &o) {
            release();
            _n      = o._n;
            _shards = std::move(o._shards);
            o._n    = 0;
            o._shards.clear();
        }
        return *this;
    }
    DeviceUnitary(DeviceUnitary const& o) : _n(o._n) {
        for (auto* s : o._shards) {
            qsx_unitary* c = nullptr;
            check(qsx_unitary_clone(s, &c));
            _shards.push_back(c);
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
        size_t const cols = dim / g;
        std::vector<std::complex<double>> tmp;
        for (size_t k = 0; k < g; ++k) {
            std::complex<double> const* src = row_major;
            if (g > 1) {
                tmp.resize(dim * cols);
                for (size_t r = 0; r < dim; ++r)
                    for (size_t c = 0; c < cols; ++c) tmp[r * cols + c] = row_major[r * dim + k * cols + c];
                src = tmp.data();
            }
            check(qsx_unitary_upload(u._shards[k], 0, dim, reinterpret_cast<double const*>(src)));
        }
        return u;
    }

    // U <- gates . U on every shard (same stream of gates, no communication)
    // returns the qsx status (QSX_OK, QSX_ERR_INTERRUPTED, QSX_ERR_UNSUPPORTED, ...)
    int apply(std::vector<qsx_gate> const& gates, qsx_stop_fn should_stop = nullptr, void* stop_arg = nullptr) {
        if (_shards.empty()) return QSX_ERR_BAD_ARG;
        if (_shards.size() == 1) {
            int const rc = qsx_apply_gates(_shards[0], gates.data(), gates.size(), nullptr, should_stop, stop_arg, nullptr);
            if (rc == QSX_ERR_OOM) throw std::bad_alloc();
            return rc;
        }
        // several GPUs in one process: schedule once, enqueue on every device, then wait for all
        qsx_unitary_info info{};
        check(qsx_unitary_get_info(_shards[0], &info));
        qsx_plan* plan = nullptr;
        int rc         = qsx_plan_create(_n, info.ncols, gates.data(), gates.size(), nullptr, &plan);
        if (rc != QSX_OK) return rc;
        for (auto* s : _shards) {
            if (should_stop != nullptr && should_stop(stop_arg)) rc = QSX_ERR_INTERRUPTED;
            if (rc == QSX_OK) rc = qsx_plan_run(plan, s, nullptr, nullptr, QSX_RUN_ASYNC, nullptr);
        }
        for (auto* s : _shards) {
            qsx_unitary_info si{};
            if (qsx_unitary_get_info(s, &si) == QSX_OK) {
                int const src = qsx_device_synchronize(si.device);
                if (rc == QSX_OK) rc = src;
            }
        }
        qsx_plan_destroy(plan);
        if (rc == QSX_ERR_OOM) throw std::bad_alloc();
        return rc;
    }

    // {Re,Im sum conj(a)b, sum|a|^2, sum|b|^2, Re,Im sum a, Re,Im sum b} over all shards
    std::array<double, 8> equiv_sums(DeviceUnitary const& other) const {
        if (_n != other._n || _shards.size() != other._shards.size())
            throw std::invalid_argument("The two tensors should have the same shape");
        std::array<double, 8> tot{};
        for (size_t k = 0; k < _shards.size(); ++k) {
            double part[8];
            int const rc = qsx_equiv_partial(_shards[k], other._shards[k], part);
            if (rc == QSX_ERR_SHAPE) throw std::invalid_argument("The two tensors should have the same shape");
            check(rc);
            for (int i = 0; i < 8; ++i) tot[i] += part[i];
        }
        return tot;
    }

    void adjoint_inplace() {
        check(qsx_unitary_adjoint_sharded(_shards.data(), static_cast<int>(_shards.size())));
    }

    // whole matrix, row-major (print / write / element access; small n only)
    std::vector<std::complex<double>> download() const {
        size_t const dim = this->dim(), g = _shards.size(), cols = dim / g;
        std::vector<std::complex<double>> out(dim * dim), tmp;
        for (size_t k = 0; k < g; ++k) {
            if (g == 1) {
                check(qsx_unitary_download(_shards[0], 0, dim, reinterpret_cast<double*>(out.data())));
            } else {
                tmp.resize(dim * cols);
                check(qsx_unitary_download(_shards[k], 0, dim, reinterpret_cast<double*>(tmp.data())));
                for (size_t r = 0; r < dim; ++r)
                    for (size_t c = 0; c < cols; ++c) out[r * dim + k * cols + c] = tmp[r * cols + c];
            }
        }
        return out;
    }

private:
    static DeviceUnitary uninitialised(int n_qubits) {
        DeviceUnitary u;
        u._n             = n_qubits;
        int const g      = shard_count(n_qubits);
        uint64_t const c = (uint64_t{1} << n_qubits) / (uint64_t)g;
        for (int k = 0; k < g; ++k) {
            qsx_unitary* s = nullptr;
            check(qsx_unitary_create(k, n_qubits, (uint64_t)k * c, c, &s));
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

    int _n = 0;
    std::vector<qsx_unitary*> _shards;
};

}  // namespace qsyn::tensor
