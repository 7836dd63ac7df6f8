// phase.hpp -- dvlab::Phase: a rational multiple of pi, kept in (-pi, pi].
//
// API- and behaviour-compatible with the reference's src/util/phase.hpp:31-243 and
// phase.cpp:24-50 for the tensor-verification path: construction from (n,d) and from a
// floating-point angle (eps 1e-4 rad), string parsing of "3*pi/4" / "0.785" /
// "6.28*0.46", conversion to double, and the print form ("π/4", "-3π/8", "0").
// Written from scratch for this repository.
#pragma once

#include <cctype>
#include <cmath>
#include <concepts>
#include <numbers>
#include <optional>
#include <ostream>
#include <string>
#include <string_view>
#include <vector>

#include "./rational.hpp"

namespace dvlab {

template <typename T>
concept unitless = std::is_arithmetic_v<T> || std::same_as<T, Rational>;

class Phase {
public:
    using IntegralType = Rational::IntegralType;
    constexpr Phase() : _rational(0, 1) {}
    constexpr explicit Phase(IntegralType n) : _rational(n, 1) { normalize(); }
    constexpr Phase(IntegralType n, IntegralType d) : _rational(n, d) { normalize(); }
    template <class T>
    requires std::floating_point<T>
    Phase(T f, T eps = 1e-4) : _rational(f / std::numbers::pi_v<T>, eps / std::numbers::pi_v<T>) { normalize(); }

    constexpr Phase operator+() const { return *this; }
    constexpr Phase operator-() const { return Phase(-_rational.numerator(), _rational.denominator()); }
    constexpr Phase& operator+=(Phase const& rhs) {
        _rational += rhs._rational;
        normalize();
        return *this;
    }
    constexpr Phase& operator-=(Phase const& rhs) {
        _rational -= rhs._rational;
        normalize();
        return *this;
    }
    friend constexpr Phase operator+(Phase lhs, Phase const& rhs) { return lhs += rhs; }
    friend constexpr Phase operator-(Phase lhs, Phase const& rhs) { return lhs -= rhs; }

    constexpr Phase& operator*=(Rational const& rhs) {
        _rational *= rhs;
        normalize();
        return *this;
    }
    constexpr Phase& operator/=(Rational const& rhs) {
        _rational /= rhs;
        normalize();
        return *this;
    }
    friend constexpr Phase operator*(Phase lhs, Rational const& rhs) { return lhs *= rhs; }
    friend constexpr Phase operator*(Rational const& lhs, Phase rhs) { return rhs *= lhs; }
    friend constexpr Phase operator/(Phase lhs, Rational const& rhs) { return lhs /= rhs; }
    friend constexpr Rational operator/(Phase const& lhs, Phase const& rhs) { return lhs._rational / rhs._rational; }

    constexpr bool operator==(Phase const& rhs) const { return _rational == rhs._rational; }
    constexpr bool operator!=(Phase const& rhs) const { return !(*this == rhs); }

    template <class T>
    requires std::floating_point<T>
    constexpr static T phase_to_floating_point(Phase const& p) {
        return std::numbers::pi_v<T> * Rational::rational_to_floating_point<T>(p._rational);
    }
    constexpr static float phase_to_f(Phase const& p) { return phase_to_floating_point<float>(p); }
    constexpr static double phase_to_d(Phase const& p) { return phase_to_floating_point<double>(p); }
    constexpr static long double phase_to_ld(Phase const& p) { return phase_to_floating_point<long double>(p); }

    Rational get_rational() const { return _rational; }
    constexpr IntegralType numerator() const { return _rational.numerator(); }
    constexpr IntegralType denominator() const { return _rational.denominator(); }

    template <class T>
    requires std::floating_point<T>
    static Phase to_phase(T f, T eps = 1e-4) { return Phase(f, eps); }

    std::string get_ascii_string() const {
        std::string s;
        if (numerator() != 1) s += std::to_string(numerator()) + "*";
        s += "pi";
        if (denominator() != 1) s += "/" + std::to_string(denominator());
        return s;
    }
    std::string get_print_string() const {
        std::string s;
        if (numerator() == -1)
            s = "-";
        else if (numerator() != 1)
            s = std::to_string(numerator());
        if (numerator() != 0) s += "π";
        if (denominator() != 1) s += "/" + std::to_string(denominator());
        return s;
    }
    friend std::ostream& operator<<(std::ostream& os, Phase const& p) { return os << p.get_print_string(); }

    // bring the value into (-1, 1] (units of pi)
    constexpr void normalize() {
        Rational const half = _rational / Rational(2);
        // truncating division of (num - (num >= 0 ? 0 : den)) by den, as the reference does
        IntegralType const k = (half.numerator() - (half.numerator() >= 0 ? 0 : half.denominator())) / half.denominator();
        _rational -= Rational(2 * k);
        if (_rational > Rational(1)) _rational -= Rational(2);
    }

    template <class T = double>
    requires std::floating_point<T>
    static std::optional<Phase> from_string(std::string const& str) {
        Phase p;
        if (!str_to_phase<T>(str, p)) return std::nullopt;
        return p;
    }

    // "a*b/c" expression of integers, floats and the tokens pi / -pi (phase.hpp:167-233)
    template <class T = double>
    requires std::floating_point<T>
    static bool str_to_phase(std::string_view str, Phase& p) {
        std::vector<std::string> terms;
        std::vector<char> seps;
        size_t begin = 0;
        for (;;) {
            size_t const at = str.find_first_of("*/", begin);
            if (at == std::string_view::npos) {
                terms.emplace_back(str.substr(begin));
                break;
            }
            seps.push_back(str[at]);
            terms.emplace_back(str.substr(begin, at - begin));
            begin = at + 1;
        }
        if (seps.size() >= terms.size()) return false;

        IntegralType pi_power = 0, num = 1, den = 1;
        T scale               = 1.0;
        for (size_t i = 0; i < terms.size(); ++i) {
            bool const divide = i != 0 && seps[i - 1] == '/';
            std::string low   = terms[i];
            for (auto& ch : low) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
            IntegralType iv = 0;
            T fv            = 0;
            if (low == "pi" || low == "-pi") {
                if (low[0] == '-') num *= -1;
                pi_power += divide ? -1 : 1;
            } else if (parse_whole<IntegralType>(terms[i], iv)) {
                (divide ? den : num) *= iv;
            } else if (parse_whole<T>(terms[i], fv)) {
                if (divide)
                    scale /= fv;
                else
                    scale *= fv;
            } else {
                return false;
            }
        }
        Rational const approx(scale * std::pow(std::numbers::pi_v<T>, pi_power - 1), static_cast<T>(1e-4) / std::numbers::pi_v<T>);
        p = Phase(num, den) * approx;
        return true;
    }

private:
    // std::stoi / std::stod semantics + "whole string consumed" (dvlab_string.hpp:186-197)
    template <class N>
    static bool parse_whole(std::string const& s, N& out) {
        size_t used = 0;
        try {
            if constexpr (std::is_integral_v<N>)
                out = static_cast<N>(std::stoi(s, &used));
            else
                out = static_cast<N>(std::stod(s, &used));
        } catch (std::exception const&) {
            return false;
        }
        return used == s.size();
    }

    Rational _rational;
};

}  // namespace dvlab
