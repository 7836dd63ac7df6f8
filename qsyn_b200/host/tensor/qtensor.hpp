// qtensor.hpp -- class QTensor<T> : Tensor<std::complex<T>>, xtensor-free.
//
// Same interface as the reference's src/tensor/qtensor.hpp:21-107: gate / spider builders
// with the reference's exact element formulas (qtensor.hpp:121-277, 346-372), to_qtensor /
// to_matrix, file-name / procedure metadata, and the equivalence quantities
// global_scalar_factor / global_norm / global_phase / is_equivalent (qtensor.hpp:383-417).
// When both operands are device-resident unitaries the equivalence quantities come from ONE
// fused reduction on the GPU (qsx_equiv_partial) instead of 8 host passes.
#pragma once

#include <cmath>
#include <complex>
#include <ostream>
#include <string>
#include <string_view>
#include <vector>

#include "../util/phase.hpp"
#include "./tensor.hpp"

namespace qsyn::tensor {

template <typename T>
class QTensor : public Tensor<std::complex<T>> {
protected:
    using DataType     = typename Tensor<std::complex<T>>::DataType;
    using InternalType = typename Tensor<std::complex<T>>::InternalType;
    using Base         = Tensor<std::complex<T>>;

public:
    QTensor() : Base(std::complex<T>(1, 0)) {}
    QTensor(Base const& t) : Base(t) {}
    QTensor(Base&& t) : Base(std::move(t)) {}

    QTensor(detail::nested_il_t<DataType, 1> il) : Base(il) {}
    QTensor(detail::nested_il_t<DataType, 2> il) : Base(il) {}
    QTensor(detail::nested_il_t<DataType, 3> il) : Base(il) {}
    QTensor(detail::nested_il_t<DataType, 4> il) : Base(il) {}
    QTensor(detail::nested_il_t<DataType, 5> il) : Base(il) {}

    QTensor(TensorShape const& shape) : Base(shape) {}
    QTensor(TensorShape const& shape, DataType const& fill) : Base(shape, fill) {}
    QTensor(DeviceUnitary&& u, bool as_matrix)
    requires std::is_same_v<T, double>
        : Base(std::move(u), as_matrix) {}

    static QTensor<T> identity(size_t const& n_qubits);
    static QTensor<T> zspider(size_t const& arity, dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> xspider(size_t const& arity, dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> hbox(size_t const& arity, DataType const& a = -1.);
    static QTensor<T> xgate() { return {{0., 1.}, {1., 0.}}; }
    static QTensor<T> ygate() { return {{DataType(0., 0.), DataType(0., -1.)}, {DataType(0., 1.), DataType(0., 0.)}}; }
    static QTensor<T> zgate() { return {{1., 0.}, {0., -1.}}; }
    static QTensor<T> hgate() { return {{1. / std::sqrt(2), 1. / std::sqrt(2)}, {1. / std::sqrt(2), -1. / std::sqrt(2)}}; }
    static QTensor<T> rxgate(dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> rygate(dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> rzgate(dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> pxgate(dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> pygate(dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> pzgate(dvlab::Phase const& phase = dvlab::Phase(0));
    static QTensor<T> control(QTensor<T> const& gate, size_t n_ctrls = 1);

    QTensor<T> to_qtensor() const;

    template <typename U>
    friend std::complex<U> global_scalar_factor(QTensor<U> const& t1, QTensor<U> const& t2);
    template <typename U>
    friend U global_norm(QTensor<U> const& t1, QTensor<U> const& t2);
    template <typename U>
    friend dvlab::Phase global_phase(QTensor<U> const& t1, QTensor<U> const& t2);
    template <typename U>
    friend bool is_equivalent(QTensor<U> const& t1, QTensor<U> const& t2, U eps);
    template <typename U>
    friend double cosine_similarity(QTensor<U> const& t1, QTensor<U> const& t2);
    template <typename U>
    friend double inner_product(QTensor<U> const& t1, QTensor<U> const& t2);
    template <typename U>
    friend bool equivalence_sums(QTensor<U> const& t1, QTensor<U> const& t2, double (&sums)[8]);

    void set_filename(std::string const& f) { _filename = f; }
    void add_procedures(std::vector<std::string> const& ps) { _procedures.insert(_procedures.end(), ps.begin(), ps.end()); }
    void add_procedure(std::string_view p) { _procedures.emplace_back(p); }

    std::string get_filename() const { return _filename; }
    std::vector<std::string> const& get_procedures() const { return _procedures; }

    QTensor<T> to_matrix() &;
    QTensor<T> to_matrix() &&;  // moves a device-resident tensor instead of cloning it (tensor.hpp)
    QTensor<T> to_matrix(TensorAxisList const& out, TensorAxisList const& in) { return Base::to_matrix(out, in); }

    // the device-resident unitary behind this tensor (nullptr for a host tensor)
    DeviceUnitary const* device_unitary() const {
        if constexpr (std::is_same_v<T, double>) return this->_dev.resident() ? &this->_dev.unitary : nullptr;
        return nullptr;
    }
    DeviceUnitary* device_unitary() {
        if constexpr (std::is_same_v<T, double>) return this->_dev.resident() ? &this->_dev.unitary : nullptr;
        return nullptr;
    }

private:
    static DataType _nu_pow(int n) { return std::pow(2., -0.25 * n); }

    std::string _filename;
    std::vector<std::string> _procedures;
};

//------------------------------
// tensor builders (host tensors; formulas of qtensor.hpp:121-277)
//------------------------------

// n-qubit identity as a rank-2n tensor with axes [o0,i0,o1,i1,...]
template <typename T>
QTensor<T> QTensor<T>::identity(size_t const& n_qubits) {
    size_t const dim = size_t{1} << n_qubits;
    QTensor<T> t(TensorShape{dim, dim}, DataType(0));
    for (size_t i = 0; i < dim; ++i) t(i, i) = 1.;
    return t.to_qtensor();
}

template <typename T>
QTensor<T> QTensor<T>::zspider(size_t const& arity, dvlab::Phase const& phase) {
    using namespace std::literals;
    QTensor<T> t(TensorShape(arity, 2), DataType(0));
    auto const w = std::exp(std::complex<T>(0, 1) * dvlab::Phase::phase_to_floating_point<T>(phase));
    if (arity == 0) {
        t._data[0] = DataType(1.) + w;
    } else {
        t._data.front() = 1.;
        t._data.back()  = w;
    }
    t *= _nu_pow(2 - static_cast<int>(arity));
    return t;
}

template <typename T>
QTensor<T> QTensor<T>::xspider(size_t const& arity, dvlab::Phase const& phase) {
    QTensor<T> t(TensorShape(arity, 2), DataType(1));
    QTensor<T> const ket_minus({DataType(1.), DataType(-1.)});
    QTensor<T> const tmp = tensor_product_pow<DataType>(ket_minus, arity);
    auto const w         = std::exp(std::complex<T>(0, 1) * dvlab::Phase::phase_to_floating_point<T>(phase));
    for (size_t i = 0; i < t._data.size(); ++i) t._data[i] += tmp._data[i] * w;
    t /= DataType(std::pow(std::sqrt(2), arity));  // 2.0000000000000004 for arity 2, like the reference
    t *= _nu_pow(2 - static_cast<int>(arity));
    return t;
}

template <typename T>
QTensor<T> QTensor<T>::hbox(size_t const& arity, DataType const& a) {
    QTensor<T> t(TensorShape(arity, 2), DataType(1));
    t._data.back() = a;
    t *= _nu_pow(static_cast<int>(arity));
    return t;
}

template <typename T>
QTensor<T> QTensor<T>::pzgate(dvlab::Phase const& phase) { return QTensor<T>::zspider(2, phase); }
template <typename T>
QTensor<T> QTensor<T>::pxgate(dvlab::Phase const& phase) { return QTensor<T>::xspider(2, phase); }
template <typename T>
QTensor<T> QTensor<T>::pygate(dvlab::Phase const& phase) {
    auto const sdg = QTensor<T>::pzgate(dvlab::Phase(-1, 2));
    auto const px  = QTensor<T>::pxgate(phase);
    auto const s   = QTensor<T>::pzgate(dvlab::Phase(1, 2));
    return tensordot<DataType>(s, tensordot<DataType>(px, sdg, {1}, {0}), {1}, {0});
}
template <typename T>
QTensor<T> QTensor<T>::rxgate(dvlab::Phase const& phase) {
    auto t = QTensor<T>::pxgate(phase);
    t *= std::exp(std::complex<T>(0, -0.5) * dvlab::Phase::phase_to_floating_point<T>(phase));
    return t;
}
template <typename T>
QTensor<T> QTensor<T>::rygate(dvlab::Phase const& phase) {
    auto t = QTensor<T>::pygate(phase);
    t *= std::exp(std::complex<T>(0, -0.5) * dvlab::Phase::phase_to_floating_point<T>(phase));
    return t;
}
template <typename T>
QTensor<T> QTensor<T>::rzgate(dvlab::Phase const& phase) {
    auto t = QTensor<T>::pzgate(phase);
    t *= std::exp(std::complex<T>(0, -0.5) * dvlab::Phase::phase_to_floating_point<T>(phase));
    return t;
}

// 2^n x 2^n matrix -> rank-2n tensor with axes [o0,i0,o1,i1,...] (qtensor.hpp:315-330)
template <typename T>
QTensor<T> QTensor<T>::to_qtensor() const {
    assert(this->dimension() == 2);
    if constexpr (std::is_same_v<T, double>) {
        if (this->_dev.resident()) {
            QTensor<T> r(*this);
            r._dev.as_matrix = false;
            r.reset_axis_history();
            return r;
        }
    }
    auto const dim = static_cast<size_t>(std::llround(std::log2(this->_shape[0]) + std::log2(this->_shape[1])));
    assert(dim % 2 == 0);
    TensorAxisList ax;
    for (size_t i = 0; i < dim / 2; ++i) {
        ax.emplace_back(i);
        ax.emplace_back(i + dim / 2);
    }
    QTensor<T> result = *this;
    result.reshape(TensorShape(dim, 2));
    QTensor<T> out = result.transpose(ax);
    out._filename   = _filename;
    out._procedures = _procedures;
    return out;
}

// controlled gate: direct_sum(I, M), controls are the most significant qubits (qtensor.hpp:346-372)
template <typename T>
QTensor<T> QTensor<T>::control(QTensor<T> const& gate, size_t n_ctrls) {
    if (n_ctrls == 0) return gate;
    auto const dim = gate.dimension();
    assert(dim % 2 == 0);
    TensorAxisList ax;
    for (size_t i = 0; i < dim / 2; ++i) ax.emplace_back(2 * i);
    for (size_t i = 0; i < dim / 2; ++i) ax.emplace_back(2 * i + 1);
    size_t const gate_size     = size_t{1} << (dim / 2);
    size_t const identity_size = gate_size * ((size_t{1} << n_ctrls) - 1);
    QTensor<T> ident(TensorShape{identity_size, identity_size}, DataType(0));
    for (size_t i = 0; i < identity_size; ++i) ident(i, i) = 1.;
    QTensor<T> gate_matrix = gate.transpose(ax);
    gate_matrix.reshape({gate_size, gate_size});
    QTensor<T> const result = direct_sum<DataType>(ident, gate_matrix);
    return result.to_qtensor();
}

// rank-2n [o0,i0,...] -> 2^n x 2^n matrix [outputs | inputs] (qtensor.hpp:442-452)
template <typename T>
QTensor<T> QTensor<T>::to_matrix() & {
    auto const n_qubits = this->dimension() / 2;
    TensorAxisList out, in;
    for (size_t i = 0; i < n_qubits; ++i) {
        out.emplace_back(2 * i);
        in.emplace_back(2 * i + 1);
    }
    QTensor<T> r    = this->to_matrix(out, in);
    r._filename     = _filename;
    r._procedures   = _procedures;
    return r;
}

template <typename T>
QTensor<T> QTensor<T>::to_matrix() && {
    auto const n_qubits = this->dimension() / 2;
    TensorAxisList out, in;
    for (size_t i = 0; i < n_qubits; ++i) {
        out.emplace_back(2 * i);
        in.emplace_back(2 * i + 1);
    }
    std::string filename                = std::move(_filename);
    std::vector<std::string> procedures = std::move(_procedures);
    QTensor<T> r    = static_cast<Base&&>(*this).to_matrix(out, in);
    r._filename     = std::move(filename);
    r._procedures   = std::move(procedures);
    return r;
}

//------------------------------
// equivalence (qtensor.hpp:383-417, tensor.hpp:393-403)
//------------------------------

// The 8 sums {Re,Im sum conj(t1) t2, sum|t1|^2, sum|t2|^2, Re,Im sum t1, Re,Im sum t2}.
// One fused GPU pass over two device-resident unitaries; a host operand that is an n-qubit
// operator is uploaded first (QTensor::identity(n) in qcir_equiv.cpp:65, zx2ts results, tensors
// read from files).  Returns false on a shape mismatch.
template <typename U>
bool equivalence_sums(QTensor<U> const& t1, QTensor<U> const& t2, double (&sums)[8]) {
    if (t1.shape() != t2.shape()) return false;
    if constexpr (std::is_same_v<U, double>) {
        DeviceUnitary const* d1 = t1.device_unitary();
        DeviceUnitary const* d2 = t2.device_unitary();
        // a 2^n x 2^n operand (matrix, or the rank-2n tensor of an n-qubit operator) belongs on the
        // device even when both sides were built on the host (zx2ts, tensor read)
        auto qubits_of = [](QTensor<U> const& t) -> int {
            auto const sh = t.shape();
            size_t rows = 0;
            if (sh.size() == 2 && sh[0] == sh[1]) {
                rows = sh[0];
            } else if (sh.size() >= 2 && sh.size() % 2 == 0 && std::all_of(sh.begin(), sh.end(), [](size_t e) { return e == 2; })) {
                rows = size_t{1} << (sh.size() / 2);
            }
            int n = 0;
            while ((size_t{1} << n) < rows) ++n;
            return rows >= 2 && (size_t{1} << n) == rows ? n : 0;
        };
        int const nq = (d1 != nullptr) ? d1->n_qubits() : (d2 != nullptr) ? d2->n_qubits() : qubits_of(t1);
        if (nq > 0) {
            DeviceUnitary tmp1, tmp2;
            auto lift = [&](QTensor<U> const& h) {
                // host tensor -> device, in the physical [out | in] matrix layout
                QTensor<U> m = h;
                if (m.dimension() != 2) m = m.to_matrix();
                auto const v = m.host_elements();
                return DeviceUnitary::from_host(nq, v.data());
            };
            if (d1 == nullptr) {
                tmp1 = lift(t1);
                d1   = &tmp1;
            }
            if (d2 == nullptr) {
                tmp2 = lift(t2);
                d2   = &tmp2;
            }
            auto const s = d1->equiv_sums(*d2);
            for (int i = 0; i < 8; ++i) sums[i] = s[i];
            return true;
        }
    }
    // neither operand is an n-qubit operator (e.g. a 4 x 2 isometry): a handful of elements
    auto const a = t1.host_elements(), b = t2.host_elements();
    std::complex<double> ab(0, 0), sa(0, 0), sb(0, 0);
    double aa = 0, bb = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        std::complex<double> const x(a[i]), y(b[i]);
        ab += std::conj(x) * y;
        aa += std::norm(x);
        bb += std::norm(y);
        sa += x;
        sb += y;
    }
    sums[0] = ab.real(), sums[1] = ab.imag(), sums[2] = aa, sums[3] = bb;
    sums[4] = sa.real(), sums[5] = sa.imag(), sums[6] = sb.real(), sums[7] = sb.imag();
    return true;
}

template <typename U>
double inner_product(QTensor<U> const& t1, QTensor<U> const& t2) {
    double s[8];
    if (!equivalence_sums(t1, t2, s)) throw std::invalid_argument("The two tensors should have the same shape");
    return std::abs(std::complex<double>(s[0], s[1]));
}

template <typename U>
double cosine_similarity(QTensor<U> const& t1, QTensor<U> const& t2) {
    double s[8];
    if (!equivalence_sums(t1, t2, s)) throw std::invalid_argument("The two tensors should have the same shape");
    return std::abs(std::complex<double>(s[0], s[1])) / std::sqrt(std::abs(s[2]) * std::abs(s[3]));
}

// sum(t2) / sum(t1): only well defined when the cosine similarity is high (qtensor.hpp:383-385)
template <typename U>
std::complex<U> global_scalar_factor(QTensor<U> const& t1, QTensor<U> const& t2) {
    double s[8];
    if (!equivalence_sums(t1, t2, s)) throw std::invalid_argument("The two tensors should have the same shape");
    return std::complex<U>(std::complex<double>(s[6], s[7]) / std::complex<double>(s[4], s[5]));
}

template <typename U>
U global_norm(QTensor<U> const& t1, QTensor<U> const& t2) {
    return std::abs(global_scalar_factor(t1, t2));
}

template <typename U>
dvlab::Phase global_phase(QTensor<U> const& t1, QTensor<U> const& t2) {
    return dvlab::Phase(std::arg(global_scalar_factor(t1, t2)));
}

template <typename U>
bool is_equivalent(QTensor<U> const& t1, QTensor<U> const& t2, U eps = 1e-6) {
    if (t1.shape() != t2.shape()) return false;
    return cosine_similarity(t1, t2) >= (1 - eps);
}

}  // namespace qsyn::tensor
