// rational.hpp -- dvlab::Rational for the B200 build of qsyn's tensor-verification path.
//
// API- and behaviour-compatible with the reference's src/util/rational.hpp:27-148 for
// everything the hot path uses (gate angles in, global phase out): int numerator /
// denominator, always reduced; float -> rational by a Stern-Brocot mediant search that
// returns the FIRST mediant inside [f-eps, f+eps] (rational.hpp:97-128).  Written from
// scratch for this repository; integer state instead of the reference's doubles.
#pragma once

#include <cassert>
#include <cmath>
#include <concepts>
#include <iosfwd>
#include <numeric>
#include <ostream>
#include <stdexcept>
#include <string>

namespace dvlab {

class Rational {
public:
    using IntegralType      = int;
    using FloatingPointType = double;

    constexpr Rational() = default;
    constexpr Rational(IntegralType n) : _num(n) {}
    constexpr Rational(IntegralType n, IntegralType d) : _num(n), _den(d) {
        assert(d != 0);
        reduce();
    }
    template <class T>
    requires std::floating_point<T>
    Rational(T f, T eps = 1e-4) { *this = to_rational(f, eps); }

    constexpr IntegralType numerator() const { return static_cast<IntegralType>(_num); }
    constexpr IntegralType denominator() const { return static_cast<IntegralType>(_den); }

    constexpr Rational operator+() const { return *this; }
    constexpr Rational operator-() const { return Rational(static_cast<IntegralType>(-_num), static_cast<IntegralType>(_den)); }

    constexpr Rational& operator+=(Rational const& r) {
        _num = _num * r._den + _den * r._num;
        _den = _den * r._den;
        reduce();
        return *this;
    }
    constexpr Rational& operator-=(Rational const& r) {
        _num = _num * r._den - _den * r._num;
        _den = _den * r._den;
        reduce();
        return *this;
    }
    constexpr Rational& operator*=(Rational const& r) {
        _num *= r._num;
        _den *= r._den;
        reduce();
        return *this;
    }
    constexpr Rational& operator/=(Rational const& r) {
        if (r._num == 0) throw std::overflow_error("Attempting to divide by 0");
        long long const n = _num * r._den, d = _den * r._num;
        _num = n;
        _den = d;
        reduce();
        return *this;
    }
    friend constexpr Rational operator+(Rational a, Rational const& b) { return a += b; }
    friend constexpr Rational operator-(Rational a, Rational const& b) { return a -= b; }
    friend constexpr Rational operator*(Rational a, Rational const& b) { return a *= b; }
    friend constexpr Rational operator/(Rational a, Rational const& b) { return a /= b; }

    constexpr bool operator==(Rational const& r) const { return _num == r._num && _den == r._den; }
    constexpr bool operator!=(Rational const& r) const { return !(*this == r); }
    constexpr bool operator<(Rational const& r) const { return _num * r._den < _den * r._num; }
    constexpr bool operator<=(Rational const& r) const { return *this == r || *this < r; }
    constexpr bool operator>(Rational const& r) const { return !(*this <= r); }
    constexpr bool operator>=(Rational const& r) const { return !(*this < r); }

    constexpr void reduce() {
        if (_den < 0) _num = -_num, _den = -_den;
        long long g = std::gcd(_num, _den);
        if (g == 0) g = 1;
        _num /= g;
        _den /= g;
    }

    template <class T>
    requires std::floating_point<T>
    constexpr static T rational_to_floating_point(Rational const& q) { return static_cast<T>(q._num) / static_cast<T>(q._den); }
    constexpr static float rational_to_f(Rational const& q) { return rational_to_floating_point<float>(q); }
    constexpr static double rational_to_d(Rational const& q) { return rational_to_floating_point<double>(q); }
    constexpr static long double rational_to_ld(Rational const& q) { return rational_to_floating_point<long double>(q); }

    // Stern-Brocot search; the caller must pass a finite f (the reference has UB otherwise,
    // SURVEY Appendix C trap 2) -- non-finite input throws std::domain_error here.
    template <class T>
    requires std::floating_point<T>
    static Rational to_rational(T f, T eps = 1e-4) {
        if (!std::isfinite(f)) throw std::domain_error("Rational::to_rational: non-finite value");
        auto const whole = static_cast<long long>(std::floor(f));
        f -= static_cast<T>(whole);
        long long ln = 0, ld = 1, un = 1, ud = 1, mn = 1, md = 2;
        auto const ge_lo = [&](long long n, long long d) { return (f - eps) <= static_cast<T>(n) / static_cast<T>(d); };
        auto const le_hi = [&](long long n, long long d) { return (f + eps) >= static_cast<T>(n) / static_cast<T>(d); };
        auto const make  = [&](long long n, long long d) {
            return Rational(static_cast<IntegralType>(n + whole * d), static_cast<IntegralType>(d));
        };
        if (ge_lo(ln, ld) && le_hi(ln, ld)) return make(ln, ld);
        if (ge_lo(un, ud) && le_hi(un, ud)) return make(un, ud);
        for (;;) {
            if (!ge_lo(mn, md))
                ln = mn, ld = md;
            else if (!le_hi(mn, md))
                un = mn, ud = md;
            else
                return make(mn, md);
            mn = ln + un;
            md = ld + ud;
        }
    }

    friend std::ostream& operator<<(std::ostream& os, Rational const& q) {
        os << q.numerator();
        if (q.denominator() != 1) os << '/' << q.denominator();
        return os;
    }

private:
    long long _num = 0;
    long long _den = 1;
};

}  // namespace dvlab
