// tensor.hpp -- class Tensor<DT>, xtensor-free.
//
// Same interface as the reference's src/tensor/tensor.hpp:34-146 for everything the
// tensor-verification path and its callers touch: construction from nested initializer
// lists / shapes, element access, shape()/dimension(), reshape/transpose/to_matrix,
// tensordot with the _axis_history bookkeeping (tensor.hpp:413-432), direct_sum,
// tensor_product_pow, adjoint[_inplace], inner_product / cosine_similarity
// (tensor.hpp:393-403), CSV read/write (tensor.hpp:326-384), and printing.
//
// Storage: small tensors (gate matrices, spider tensors, anything built on the host) live in a
// row-major std::vector.  The n-qubit main tensor produced by to_tensor(QCir) is
// DEVICE-RESIDENT (DeviceUnitary, complex<double> only): its physical layout is always the
// stored matrix U[out,in]; the two logical shapes the reference uses for it -- rank-2n
// interleaved [o0,i0,o1,i1,...] (qcir_to_tensor.cpp:208) and the 2^n x 2^n matrix
// (conversion_cmd.cpp:89) -- are a tag, so to_matrix()/to_qtensor() move no data.
// LAPACK-backed members (determinant, eigen) belong to the decomposer, which is out of scope.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <format>
#include <functional>
#include <fstream>
#include <initializer_list>
#include <iostream>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../util/logger.hpp"
#include "./device_unitary.hpp"
#include "./tensor_util.hpp"

namespace qsyn::tensor {

namespace detail {

template <typename T>
struct is_complex : std::false_type {};
template <typename T>
struct is_complex<std::complex<T>> : std::true_type {};

template <typename DT>
inline DT conj_of(DT const& v) {
    if constexpr (is_complex<DT>::value)
        return std::conj(v);
    else
        return v;
}

// nested initializer lists, depth 0..5 (xt::nested_initializer_list_t in the reference)
template <typename DT, size_t N>
struct nested_il {
    using type = std::initializer_list<typename nested_il<DT, N - 1>::type>;
};
template <typename DT>
struct nested_il<DT, 0> {
    using type = DT;
};
template <typename DT, size_t N>
using nested_il_t = typename nested_il<DT, N>::type;

template <typename DT>
void flatten(DT const& v, size_t /*depth*/, std::vector<size_t>& /*shape*/, std::vector<DT>& out) {
    out.push_back(v);
}
template <typename DT, typename L>
void flatten(std::initializer_list<L> const& il, size_t depth, std::vector<size_t>& shape, std::vector<DT>& out) {
    if (shape.size() <= depth) shape.push_back(il.size());
    if (shape[depth] != il.size()) throw std::invalid_argument("ragged initializer list");
    for (auto const& e : il) flatten<DT>(e, depth + 1, shape, out);
}

// fixed-point with 6 decimals, trailing zeros stripped (how xtensor's printer shows the
// reference's golden matrices, e.g. 0.027640 -> "0.02764", tests/conversion/sk/ref/sk.log:19-20)
inline std::string trimmed(double v) {
    std::string s = std::format("{:.6f}", v);
    while (s.size() > 1 && s.back() == '0') s.pop_back();
    if (!s.empty() && s.back() == '.') s.push_back(' ');  // xtensor prints "1." for integers
    if (s == "-0. ") s = "0. ";
    return s;
}

}  // namespace detail

// Device residency is only defined for complex<double> (the QTensor<double> main tensor)
template <typename DT>
struct DeviceSlot {
    bool resident() const { return false; }
};
template <>
struct DeviceSlot<std::complex<double>> {
    DeviceUnitary unitary;
    bool as_matrix = false;  // logical shape: false = rank-2n interleaved, true = 2^n x 2^n
    bool resident() const { return unitary.resident(); }
};

template <typename DT>
class Tensor {
protected:
    using DataType     = DT;
    using InternalType = std::vector<DT>;

public:
    Tensor(detail::nested_il_t<DT, 0> v) : _shape{}, _data{v} { reset_axis_history(); }
    Tensor(detail::nested_il_t<DT, 1> il) { _from_il(il); }
    Tensor(detail::nested_il_t<DT, 2> il) { _from_il(il); }
    Tensor(detail::nested_il_t<DT, 3> il) { _from_il(il); }
    Tensor(detail::nested_il_t<DT, 4> il) { _from_il(il); }
    Tensor(detail::nested_il_t<DT, 5> il) { _from_il(il); }

    virtual ~Tensor() = default;
    Tensor(Tensor const&)            = default;
    Tensor(Tensor&&)                 = default;
    Tensor& operator=(Tensor const&) = default;
    Tensor& operator=(Tensor&&)      = default;

    // uninitialised-value tensor of the given shape (xarray(shape) in the reference)
    Tensor(TensorShape const& shape) : _shape(shape), _data(_count(shape)) { reset_axis_history(); }
    Tensor(TensorShape const& shape, DT const& fill) : _shape(shape), _data(_count(shape), fill) { reset_axis_history(); }

    // adopt a device-resident unitary (complex<double> only)
    Tensor(DeviceUnitary&& u, bool as_matrix)
    requires std::is_same_v<DT, std::complex<double>>
    {
        _dev.unitary   = std::move(u);
        _dev.as_matrix = as_matrix;
        reset_axis_history();
    }

    bool is_device_resident() const { return _dev.resident(); }

    // Element access (tensor.hpp:156-175).  On a device-resident tensor a const access reads ONE element from the
    // device (decomposer.hpp:64-69 walks a matrix this way); asking for a writable reference brings the tensor
    // to the host first (the device copy is dropped: the caller is about to change the data).
    template <typename... Args>
    DT& operator()(Args const&... args) {
        _pull_to_host();
        return _data[_offset({static_cast<size_t>(args)...})];
    }
    template <typename... Args>
    DT const& operator()(Args const&... args) const {
        return _const_element({static_cast<size_t>(args)...});
    }
    DT& operator[](TensorIndex const& i) {
        _pull_to_host();
        return _data[_offset(i)];
    }
    DT const& operator[](TensorIndex const& i) const { return _const_element(i); }

    virtual bool operator==(Tensor<DT> const& rhs) const { return shape() == rhs.shape() && _host_view() == rhs._host_view(); }
    virtual bool operator!=(Tensor<DT> const& rhs) const { return !(*this == rhs); }

    template <typename U>
    friend std::ostream& operator<<(std::ostream& os, Tensor<U> const& t);

    virtual Tensor<DT>& operator+=(Tensor<DT> const& rhs) { return _zip(rhs, [](DT& a, DT const& b) { a += b; }); }
    virtual Tensor<DT>& operator-=(Tensor<DT> const& rhs) { return _zip(rhs, [](DT& a, DT const& b) { a -= b; }); }
    virtual Tensor<DT>& operator*=(Tensor<DT> const& rhs) { return _zip(rhs, [](DT& a, DT const& b) { a *= b; }); }
    virtual Tensor<DT>& operator/=(Tensor<DT> const& rhs) { return _zip(rhs, [](DT& a, DT const& b) { a /= b; }); }
    Tensor<DT>& operator*=(DT const& s) {
        _require_host("scalar multiply");
        for (auto& v : _data) v *= s;
        return *this;
    }
    Tensor<DT>& operator/=(DT const& s) {
        _require_host("scalar divide");
        for (auto& v : _data) v /= s;
        return *this;
    }
    friend Tensor<DT> operator+(Tensor<DT> lhs, Tensor<DT> const& rhs) { return lhs += rhs; }
    friend Tensor<DT> operator-(Tensor<DT> lhs, Tensor<DT> const& rhs) { return lhs -= rhs; }
    friend Tensor<DT> operator*(Tensor<DT> lhs, Tensor<DT> const& rhs) { return lhs *= rhs; }
    friend Tensor<DT> operator/(Tensor<DT> lhs, Tensor<DT> const& rhs) { return lhs /= rhs; }

    size_t dimension() const {
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (_dev.resident()) return _dev.as_matrix ? 2 : 2 * (size_t)_dev.unitary.n_qubits();
        }
        return _shape.size();
    }
    std::vector<size_t> shape() const {
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (_dev.resident()) {
                if (_dev.as_matrix) return {_dev.unitary.dim(), _dev.unitary.dim()};
                return std::vector<size_t>(2 * (size_t)_dev.unitary.n_qubits(), 2);
            }
        }
        return _shape;
    }

    void reset_axis_history() {
        _axis_history.clear();
        for (size_t i = 0; i < dimension(); ++i) _axis_history.emplace(i, i);
    }
    size_t get_new_axis_id(size_t const& old_id) {
        auto const it = _axis_history.find(old_id);
        return it == _axis_history.end() ? SIZE_MAX : it->second;
    }

    template <typename U>
    friend double inner_product(Tensor<U> const& t1, Tensor<U> const& t2);
    template <typename U>
    friend double cosine_similarity(Tensor<U> const& t1, Tensor<U> const& t2);
    template <typename U>
    friend Tensor<U> tensordot(Tensor<U> const& t1, Tensor<U> const& t2, TensorAxisList const& ax1, TensorAxisList const& ax2);
    template <typename U>
    friend Tensor<U> tensor_product_pow(Tensor<U> const& t, size_t n);
    template <typename U>
    friend bool is_partition(Tensor<U> const& t, TensorAxisList const& axes1, TensorAxisList const& axes2);
    template <typename U>
    friend Tensor<U> direct_sum(Tensor<U> const& t1, Tensor<U> const& t2);

    Tensor<DT> to_matrix(TensorAxisList const& row_axes, TensorAxisList const& col_axes) &;
    // on a temporary (`*t = std::move(*t).to_matrix()`, the conversion_cmd.cpp:89 pattern): a device-resident tensor is
    // re-tagged and MOVED -- the lvalue form has to clone 16 * 4^n bytes on the device to leave the original intact
    Tensor<DT> to_matrix(TensorAxisList const& row_axes, TensorAxisList const& col_axes) &&;

    void reshape(TensorShape const& shape) {
        _require_host("reshape");
        if (_count(shape) != _data.size()) throw std::invalid_argument("reshape: element count mismatch");
        _shape = shape;
    }
    Tensor<DT> transpose(TensorAxisList const& perm) const;

    void adjoint_inplace() {
        assert(dimension() == 2);
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (_dev.resident()) {
                _dev.unitary.adjoint_inplace();
                return;
            }
        }
        *this = adjoint(*this);
    }
    [[nodiscard]] friend Tensor<DT> adjoint(Tensor<DT> const& t) {
        assert(t.dimension() == 2);
        if (t.is_device_resident()) {
            Tensor<DT> c(t);
            c.adjoint_inplace();
            return c;
        }
        Tensor<DT> r(TensorShape{t._shape[1], t._shape[0]});
        for (size_t i = 0; i < t._shape[0]; ++i)
            for (size_t j = 0; j < t._shape[1]; ++j) r._data[j * t._shape[0] + i] = detail::conj_of(t._data[i * t._shape[1] + j]);
        return r;
    }

    DT trace() const {
        assert(dimension() == 2);
        auto const v = _host_view();
        auto const s = shape();
        DT acc{};
        for (size_t i = 0; i < std::min(s[0], s[1]); ++i) acc += v[i * s[1] + i];
        return acc;
    }

    bool tensor_read(std::string const& filepath);
    bool tensor_write(std::string const& filepath);

    // host copy of the elements in logical (row-major) order; downloads a device-resident matrix
    std::vector<DT> host_elements() const { return _host_view(); }

protected:
    TensorShape _shape;
    InternalType _data;
    mutable DT _element_cache{};  // last element read from a device-resident tensor
    DeviceSlot<DT> _dev;
    std::unordered_map<size_t, size_t> _axis_history;

    static size_t _count(TensorShape const& s) {
        return std::accumulate(s.begin(), s.end(), size_t{1}, std::multiplies<size_t>());
    }
    size_t _offset(std::vector<size_t> const& idx) const {
        _require_host("element access");
        assert(idx.size() == _shape.size());
        size_t off = 0;
        for (size_t a = 0; a < _shape.size(); ++a) off = off * _shape[a] + idx[a];
        return off;
    }
    void _require_host(char const* what) const {
        if (_dev.resident())
            throw std::logic_error(std::string(what) + " is not available on a device-resident tensor; use host_elements()");
    }
    void _pull_to_host() {
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (_dev.resident()) {
                auto shp  = shape();
                auto data = _host_view();
                _dev      = DeviceSlot<DT>{};
                _shape    = std::move(shp);
                _data     = std::move(data);
            }
        }
    }
    DT const& _const_element(TensorIndex const& idx) const {
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (_dev.resident()) {
                size_t r = 0, c = 0;
                if (_dev.as_matrix) {
                    if (idx.size() != 2) throw std::invalid_argument("element access: a matrix takes two indices");
                    r = idx[0], c = idx[1];
                } else {  // interleaved [o0, i0, o1, i1, ...], qubit 0 = most significant bit
                    size_t const n = (size_t)_dev.unitary.n_qubits();
                    if (idx.size() != 2 * n) throw std::invalid_argument("element access: index rank mismatch");
                    for (size_t q = 0; q < n; ++q) {
                        r = (r << 1) | (idx[2 * q] & 1);
                        c = (c << 1) | (idx[2 * q + 1] & 1);
                    }
                }
                if (r >= _dev.unitary.dim() || c >= _dev.unitary.dim()) throw std::out_of_range("element access: index out of range");
                _element_cache = _dev.unitary.element(r, c);
                return _element_cache;
            }
        }
        return _data[_offset(idx)];
    }
    template <typename IL>
    void _from_il(IL const& il) {
        detail::flatten<DT>(il, 0, _shape, _data);
        reset_axis_history();
    }
    template <typename F>
    Tensor<DT>& _zip(Tensor<DT> const& rhs, F f) {
        _require_host("element-wise arithmetic");
        rhs._require_host("element-wise arithmetic");
        if (_shape != rhs._shape) throw std::invalid_argument("element-wise arithmetic: shape mismatch");
        for (size_t i = 0; i < _data.size(); ++i) f(_data[i], rhs._data[i]);
        return *this;
    }
    // elements in logical order (matrix order for a device tensor tagged as matrix)
    std::vector<DT> _host_view() const {
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (_dev.resident()) {
                auto m = _dev.unitary.download();
                if (_dev.as_matrix) return m;
                // interleaved [o0,i0,o1,i1,...] order
                size_t const n = (size_t)_dev.unitary.n_qubits(), dim = size_t{1} << n;
                std::vector<DT> out(m.size());
                for (size_t r = 0; r < dim; ++r)
                    for (size_t c = 0; c < dim; ++c) {
                        size_t k = 0;
                        for (size_t q = 0; q < n; ++q) {
                            k = (k << 1) | ((r >> (n - 1 - q)) & 1);
                            k = (k << 1) | ((c >> (n - 1 - q)) & 1);
                        }
                        out[k] = m[r * dim + c];
                    }
                return out;
            }
        }
        return _data;
    }
};

//------------------------------
// member functions
//------------------------------

template <typename DT>
Tensor<DT> Tensor<DT>::transpose(TensorAxisList const& perm) const {
    _require_host("transpose");
    size_t const nd = _shape.size();
    if (perm.size() != nd) throw std::invalid_argument("transpose: permutation rank mismatch");
    TensorShape ns(nd);
    for (size_t a = 0; a < nd; ++a) ns[a] = _shape[perm[a]];
    Tensor<DT> r(ns);
    std::vector<size_t> stride(nd, 1);
    for (size_t a = nd; a-- > 1;) stride[a - 1] = stride[a] * _shape[a];
    std::vector<size_t> idx(nd, 0);
    for (size_t lin = 0; lin < r._data.size(); ++lin) {
        size_t src = 0;
        for (size_t a = 0; a < nd; ++a) src += idx[a] * stride[perm[a]];
        r._data[lin] = _data[src];
        for (size_t a = nd; a-- > 0;) {
            if (++idx[a] < ns[a]) break;
            idx[a] = 0;
        }
    }
    return r;
}

template <typename U>
bool is_partition(Tensor<U> const& t, TensorAxisList const& axes1, TensorAxisList const& axes2) {
    if (!is_disjoint(axes1, axes2)) return false;
    return axes1.size() + axes2.size() == t.dimension();
}

// Convert the tensor to a matrix according to the two axis lists (tensor.hpp:266-274)
template <typename DT>
Tensor<DT> Tensor<DT>::to_matrix(TensorAxisList const& row_axes, TensorAxisList const& col_axes) && {
    if constexpr (std::is_same_v<DT, std::complex<double>>) {
        if (_dev.resident() && is_partition(*this, row_axes, col_axes)) {
            bool canonical = _dev.as_matrix ? (row_axes == TensorAxisList{0} && col_axes == TensorAxisList{1}) : row_axes.size() == col_axes.size();
            for (size_t i = 0; !_dev.as_matrix && canonical && i < row_axes.size(); ++i)
                canonical = row_axes[i] == 2 * i && col_axes[i] == 2 * i + 1;
            if (canonical) {
                if (!_dev.as_matrix) {
                    _dev.as_matrix = true;
                    reset_axis_history();
                }
                return std::move(*this);
            }
        }
    }
    return static_cast<Tensor<DT>&>(*this).to_matrix(row_axes, col_axes);
}

template <typename DT>
Tensor<DT> Tensor<DT>::to_matrix(TensorAxisList const& row_axes, TensorAxisList const& col_axes) & {
    if (!is_partition(*this, row_axes, col_axes)) {
        throw std::invalid_argument("The two axis lists should partition 0~(n-1).");
    }
    if constexpr (std::is_same_v<DT, std::complex<double>>) {
        if (_dev.resident()) {
            // the device layout already IS the [outputs | inputs] matrix: only the canonical
            // partition (rows = even axes in order, cols = odd axes in order) is a no-op
            if (_dev.as_matrix) {
                if (row_axes == TensorAxisList{0} && col_axes == TensorAxisList{1}) return *this;
            } else {
                bool canonical = row_axes.size() == col_axes.size();
                for (size_t i = 0; canonical && i < row_axes.size(); ++i)
                    canonical = row_axes[i] == 2 * i && col_axes[i] == 2 * i + 1;
                if (canonical) {
                    Tensor<DT> r(*this);
                    r._dev.as_matrix = true;
                    r.reset_axis_history();
                    return r;
                }
            }
            throw std::invalid_argument("to_matrix: only the qubit-ordered [outputs | inputs] partition exists on the device");
        }
    }
    Tensor<DT> t = transpose(concat_axis_list(row_axes, col_axes));
    size_t rows  = 1, cols = 1;
    for (size_t i = 0; i < row_axes.size(); ++i) rows *= t._shape[i];
    for (size_t i = 0; i < col_axes.size(); ++i) cols *= t._shape[row_axes.size() + i];
    t._shape = {rows, cols};
    t.reset_axis_history();
    return t;
}

template <typename U>
Tensor<U> direct_sum(Tensor<U> const& t1, Tensor<U> const& t2) {
    if (t1.dimension() != 2 || t2.dimension() != 2) {
        throw std::invalid_argument("The two tensors should be 2-dimension.");
    }
    auto const s1 = t1.shape(), s2 = t2.shape();
    Tensor<U> r(TensorShape{s1[0] + s2[0], s1[1] + s2[1]}, U{});
    for (size_t i = 0; i < s1[0]; ++i)
        for (size_t j = 0; j < s1[1]; ++j) r(i, j) = t1(i, j);
    for (size_t i = 0; i < s2[0]; ++i)
        for (size_t j = 0; j < s2[1]; ++j) r(s1[0] + i, s1[1] + j) = t2(i, j);
    return r;
}

// "{{ a, b},\n { c, d}}" in the style of xtensor's printer (xio.hpp of xtensor 0.26, not part of the reference tree:
// the layout below follows its published behaviour; the only sample the reference's tests hold is the 2 x 2 matrix of
// tests/conversion/sk/ref/sk.log:19-20).  print_options defaults: more than `threshold` = 1000 elements are summarised
// with `edgeitems` = 3 leading and trailing entries per axis ("..., " inside a row, "...," on a line of its own between
// rows); a row wraps after floor(line_width = 75 / (cell width + 2)) cells.  A summarised DEVICE matrix only fetches the
// rows it prints (DeviceUnitary::download_rows), never the whole 2^n x 2^n tensor.
template <typename U>
std::ostream& operator<<(std::ostream& os, Tensor<U> const& t) {
    auto const shape = t.shape();
    size_t const nd  = shape.size();
    size_t total     = 1;
    for (size_t e : shape) total *= e;
    constexpr size_t kThreshold = 1000, kEdge = 3, kLineWidth = 75;
    bool const summarise = total > kThreshold;
    // per axis: the indices that are printed (SIZE_MAX marks the gap)
    std::vector<std::vector<size_t>> axis_idx(nd);
    for (size_t a = 0; a < nd; ++a) {
        if (summarise && shape[a] > 2 * kEdge) {
            for (size_t i = 0; i < kEdge; ++i) axis_idx[a].push_back(i);
            axis_idx[a].push_back(SIZE_MAX);
            for (size_t i = shape[a] - kEdge; i < shape[a]; ++i) axis_idx[a].push_back(i);
        } else {
            for (size_t i = 0; i < shape[a]; ++i) axis_idx[a].push_back(i);
        }
    }
    // the elements to print, in print order
    std::vector<U> vals;
    {
        std::vector<size_t> lins;
        std::vector<size_t> pos(nd, 0);
        std::function<void(size_t, size_t)> walk = [&](size_t a, size_t lin) {
            if (a == nd) {
                lins.push_back(lin);
                return;
            }
            for (size_t i : axis_idx[a])
                if (i != SIZE_MAX) walk(a + 1, lin * shape[a] + i);
        };
        walk(0, 0);
        bool fetched = false;
        if constexpr (std::is_same_v<U, std::complex<double>>) {
            if (summarise && t._dev.resident() && t._dev.as_matrix && nd == 2) {
                std::vector<U> row(shape[1]);
                size_t have = SIZE_MAX;
                for (size_t lin : lins) {
                    size_t const r = lin / shape[1], c = lin % shape[1];
                    if (r != have) {
                        t._dev.unitary.download_rows(r, 1, row.data());
                        have = r;
                    }
                    vals.push_back(row[c]);
                }
                fetched = true;
            }
        }
        if (!fetched) {
            auto const data = t._host_view();
            for (size_t lin : lins) vals.push_back(data[lin]);
        }
    }
    std::vector<std::string> cell(vals.size());
    size_t width = 0;
    if constexpr (detail::is_complex<U>::value) {
        std::vector<std::string> re(vals.size()), im(vals.size());
        size_t wr = 0, wi = 0;
        for (size_t i = 0; i < vals.size(); ++i) {
            re[i] = detail::trimmed(vals[i].real());
            if (re[i].back() == ' ') re[i].pop_back();
            std::string m = detail::trimmed(std::abs(vals[i].imag()));
            if (m.back() == ' ') m.pop_back();
            im[i] = std::string(std::signbit(vals[i].imag()) && vals[i].imag() != 0 ? "-" : "+") + m + "i";
            wr    = std::max(wr, re[i].size());
            wi    = std::max(wi, im[i].size());
        }
        for (size_t i = 0; i < vals.size(); ++i)
            cell[i] = std::string(wr - re[i].size(), ' ') + re[i] + im[i] + std::string(wi - im[i].size(), ' ');
        width = wr + wi;
    } else {
        for (size_t i = 0; i < vals.size(); ++i) {
            cell[i] = detail::trimmed(static_cast<double>(vals[i]));
            width   = std::max(width, cell[i].size());
        }
        for (auto& c : cell) c = std::string(width - c.size(), ' ') + c;
    }
    if (nd == 0) return os << (cell.empty() ? "" : cell[0]);
    size_t const line_lim = kLineWidth / (width + 2);  // cells per line in the innermost axis
    size_t next           = 0;
    std::function<void(size_t, size_t)> out = [&](size_t a, size_t blanks) {
        if (a == nd) {
            os << cell[next++];
            return;
        }
        bool const inner = a + 1 == nd;
        std::string const indents(blanks, ' ');
        size_t on_line = 0;
        auto const& ix = axis_idx[a];
        os << '{';
        for (size_t k = 0; k < ix.size(); ++k) {
            bool const last = k + 1 == ix.size();
            if (ix[k] == SIZE_MAX) {
                if (inner && line_lim != 0 && on_line >= line_lim)
                    os << " ...,";
                else if (!inner) {
                    on_line = 0;
                    os << "...,\n" << indents;
                } else {
                    os << "..., ";
                }
                continue;
            }
            if (inner && !(line_lim != 0 && on_line < line_lim)) {
                os << "\n" << indents;
                on_line = 0;
            }
            out(a + 1, blanks + 1);
            if (last) break;
            os << ',';
            ++on_line;
            if (inner) {
                if (line_lim != 0 && on_line < line_lim) os << ' ';
            } else {
                os << "\n" << indents;
            }
        }
        os << '}';
    };
    out(0, 1);
    return os;
}

template <typename DT>
bool Tensor<DT>::tensor_write(std::string const& filepath) {
    std::ofstream out_file(filepath);
    if (!out_file.is_open()) {
        spdlog::error("Failed to open file");
        return false;
    }
    auto const s = shape();
    if (s.size() != 2) return false;
    // a device-resident matrix is streamed in blocks of rows (no second full copy on the host)
    bool streamed = false;
    if constexpr (std::is_same_v<DT, std::complex<double>>) streamed = _dev.resident() && _dev.as_matrix;
    size_t const block = streamed ? std::max<size_t>(1, (size_t{1} << 22) / std::max<size_t>(s[1], 1)) : s[0];
    std::vector<DT> v;
    if (!streamed) v = _host_view();
    for (size_t row = 0; row < s[0]; row++) {
        if constexpr (std::is_same_v<DT, std::complex<double>>) {
            if (streamed && row % block == 0) {
                size_t const nrows = std::min(block, s[0] - row);
                v.resize(nrows * s[1]);
                _dev.unitary.download_rows(row, nrows, v.data());
            }
        }
        size_t const base = streamed ? (row % block) * s[1] : row * s[1];
        for (size_t col = 0; col < s[1]; col++) {
            if constexpr (detail::is_complex<DT>::value) {
                auto const re = v[base + col].real(), im = v[base + col].imag();
                // "{space if re >= 0}{re}+{|im|}j" / "{re}{im}j"  (tensor.hpp:335-339)
                out_file << (re >= 0 ? " " : "") << std::format("{}", re);
                if (im >= 0)
                    out_file << "+" << std::format("{}", std::abs(im)) << "j";
                else
                    out_file << std::format("{}", im) << "j";
            } else {
                out_file << std::format("{}", v[base + col]);
            }
            if (col != s[1] - 1) out_file << ", ";
        }
        out_file << "\n";
    }
    return true;
}

template <typename DT>
bool Tensor<DT>::tensor_read(std::string const& filepath) {
    std::ifstream in_file(filepath);
    if (!in_file.is_open()) {
        spdlog::error("Failed to open file");
        return false;
    }
    std::vector<DT> data;
    std::string line, word;
    while (std::getline(in_file, line)) {
        std::stringstream ss(line);
        while (std::getline(ss, word, ',')) {
            if constexpr (detail::is_complex<DT>::value) {
                std::stringstream ww(word);
                typename DT::value_type real = 0.0, imag = 0.0;
                char plus{}, i{};
                ww >> real >> plus >> imag >> i;
                if (plus == '-') imag = -imag;
                if (plus == 'j') {
                    imag = real;
                    real = 0;
                }
                data.push_back({real, imag});
            }
        }
    }
    auto const side = static_cast<size_t>(std::llround(std::sqrt(static_cast<double>(data.size()))));
    if (side * side != data.size()) {
        spdlog::error("The number of elements in the tensor is not a square number");
        return false;
    }
    if constexpr (std::is_same_v<DT, std::complex<double>>) _dev = DeviceSlot<DT>{};
    _shape = {side, side};
    _data  = std::move(data);
    reset_axis_history();
    return true;
}

//------------------------------
// friend functions
//------------------------------

// |sum conj(t1) * t2|  (tensor.hpp:393-398).  Host tensors only here; the device-resident case is
// handled by QTensor (one fused pass on the GPU).
template <typename U>
double inner_product(Tensor<U> const& t1, Tensor<U> const& t2) {
    if (t1.shape() != t2.shape()) {
        throw std::invalid_argument("The two tensors should have the same shape");
    }
    auto const a = t1._host_view(), b = t2._host_view();
    U acc{};
    for (size_t i = 0; i < a.size(); ++i) acc += detail::conj_of(a[i]) * b[i];
    return std::abs(acc);
}

template <typename U>
double cosine_similarity(Tensor<U> const& t1, Tensor<U> const& t2) {
    return inner_product(t1, t2) / std::sqrt(inner_product(t1, t1) * inner_product(t2, t2));
}

// tensor-dot two (host) tensors along the axes in ax1 / ax2; result axes = free axes of t1 in
// order, then free axes of t2 in order; _axis_history maps old axis ids (t2's offset by
// t1.dimension()) to the new ones (tensor.hpp:413-432)
template <typename U>
Tensor<U> tensordot(Tensor<U> const& t1, Tensor<U> const& t2, TensorAxisList const& ax1 = {}, TensorAxisList const& ax2 = {}) {
    if (ax1.size() != ax2.size()) {
        throw std::invalid_argument("The two index orders should contain the same number of indices.");
    }
    t1._require_host("tensordot");
    t2._require_host("tensordot");
    size_t const d1 = t1._shape.size(), d2 = t2._shape.size();
    TensorAxisList free1, free2;
    for (size_t i = 0; i < d1; ++i)
        if (std::find(ax1.begin(), ax1.end(), i) == ax1.end()) free1.push_back(i);
    for (size_t i = 0; i < d2; ++i)
        if (std::find(ax2.begin(), ax2.end(), i) == ax2.end()) free2.push_back(i);
    size_t m = 1, k = 1, n = 1;
    for (auto a : free1) m *= t1._shape[a];
    for (size_t i = 0; i < ax1.size(); ++i) {
        if (t1._shape[ax1[i]] != t2._shape[ax2[i]]) throw std::invalid_argument("tensordot: contracted extents differ");
        k *= t1._shape[ax1[i]];
    }
    for (auto a : free2) n *= t2._shape[a];
    Tensor<U> const a = t1.transpose(concat_axis_list(free1, ax1));  // (m, k)
    Tensor<U> const b = t2.transpose(concat_axis_list(ax2, free2));  // (k, n)
    TensorShape rs;
    for (auto x : free1) rs.push_back(t1._shape[x]);
    for (auto x : free2) rs.push_back(t2._shape[x]);
    Tensor<U> t(rs, U{});
    for (size_t i = 0; i < m; ++i)
        for (size_t p = 0; p < k; ++p) {
            U const av = a._data[i * k + p];
            for (size_t j = 0; j < n; ++j) t._data[i * n + j] += av * b._data[p * n + j];
        }
    size_t tmp_axis_id = 0;
    t._axis_history.clear();
    for (auto x : free1) t._axis_history.emplace(x, tmp_axis_id++);
    for (auto x : free2) t._axis_history.emplace(x + d1, tmp_axis_id++);
    return t;
}

template <typename U>
Tensor<U> tensor_product_pow(Tensor<U> const& t, size_t n) {
    if (n == 0) return Tensor<U>(TensorShape{}, U{1});
    if (n == 1) return t;
    auto const half = tensor_product_pow(t, n / 2);
    return n % 2 == 0 ? tensordot(half, half) : tensordot(t, tensordot(half, half));
}

template <typename U>
Tensor<U> tensor_multiply(Tensor<U> const& t1, Tensor<U> const& t2) {
    return tensordot(t1, t2, {1}, {0});
}

}  // namespace qsyn::tensor
