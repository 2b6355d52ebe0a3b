// interpolation.hpp — host-side interpolation helpers used while BUILDING device tables.
//
// Same names and numerical results as the reference (include/dxmc/interpolation.hpp:33-261):
// linear / log-log interpolation, the fixed-knot natural cubic spline behind the Compton scatter
// function, trapezoid and 20-point Gauss-Legendre quadrature. The spline additionally exposes its
// tables so Transport can flatten them for the GPU (csrc/physics.cuh scatterFactor()).
#pragma once
#include "dxmc/floating.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <concepts>
#include <iterator>
#include <numeric>
#include <vector>

namespace dxmc {

template <Floating T>
inline T interp(T x0, T x1, T y0, T y1, T x)
{
    return y0 + (y1 - y0) * (x - x0) / (x1 - x0);
}
template <Floating T, Floating U>
inline U interp(T x[2], T y[2], U xi)
{
    return y[0] + (y[1] - y[0]) * (xi - x[0]) / (x[1] - x[0]);
}

// straight line in log-log space; negative arguments propagate NaN like std::log10
template <Floating T>
inline T logloginterp(T x0, T x1, T y0, T y1, T x)
{
    const double v = std::log10(y0) + (std::log10(y1 / y0) / std::log10(x1 / x0) * std::log10(x / x0));
    return std::pow(10., v);
}
template <Floating T, Floating U>
inline T logloginterp(T x[2], T y[2], U xi)
{
    const double v = std::log10(y[0]) + (std::log10(y[1] / y[0]) / std::log10(x[1] / x[0]) * std::log10(xi / x[0]));
    return std::pow(10., v);
}

// piecewise linear table look-up, clamped at both ends
template <typename It, Floating T>
    requires std::is_same_v<typename std::iterator_traits<It>::value_type, T>
T interpolate(It xbegin, It xend, It ybegin, It yend, T xvalue)
{
    auto upper = std::upper_bound(xbegin, xend, xvalue);
    if (upper == xbegin)
        return *ybegin;
    if (upper == xend)
        return *(yend - 1);
    const auto i = std::distance(xbegin, upper);
    return interp(*(upper - 1), *upper, *(ybegin + (i - 1)), *(ybegin + i), xvalue);
}

// Natural cubic spline through N equidistant samples of `function` on [start, stop'], evaluated as
// one absolute-x cubic per interval. Note the knot step is (stop-start)/(N-2), so the last knot
// lies one step beyond `stop` — that is how the reference samples it.
template <Floating T, int N = 30>
class CubicSplineInterpolator {
public:
    template <std::regular_invocable<T> F>
        requires std::is_same_v<std::invoke_result_t<F, T>, T>
    CubicSplineInterpolator(const T start, const T stop, F function)
    {
        m_start = start;
        m_step = (stop - start) / (N - 2);
        std::array<T, N> y;
        for (std::size_t i = 0; i < N; ++i) {
            m_x[i] = m_start + m_step * i;
            y[i] = function(m_x[i]);
        }
        m_stop = m_x.back();

        // tridiagonal system for the second derivatives s (zero at both ends)
        std::array<T, N> h {}, slope {}, diag {}, rhs {};
        slope.fill(T { 1 });
        for (std::size_t i = 0; i + 1 < N; ++i) {
            h[i] = m_x[i + 1] - m_x[i];
            slope[i] = (y[i + 1] - y[i]) / h[i];
        }
        for (std::size_t i = 1; i < N; ++i) {
            diag[i] = 2 * (h[i - 1] + h[i]);
            rhs[i] = 6 * (slope[i] - slope[i - 1]);
        }
        diag[0] = diag[1];
        rhs[N - 1] = 0;
        rhs[0] = 0;
        const auto s = solveTridiagonal(h, diag, rhs);

        for (std::size_t i = 0; i + 1 < N; ++i) {
            T* c = &m_coefficients[i * 4];
            const T xa = m_x[i], xb = m_x[i + 1];
            c[0] = (s[i] * xb * xb * xb - s[i + 1] * xa * xa * xa + 6 * (y[i] * xb - y[i + 1] * xa)) / (6 * h[i]);
            c[0] += h[i] * (s[i + 1] * xa - s[i] * xb) / 6;
            c[1] = (s[i + 1] * xa * xa - s[i] * xb * xb + 2 * (y[i + 1] - y[i])) / (2 * h[i]) + h[i] * (s[i] - s[i + 1]) / 6;
            c[2] = (s[i] * xb - s[i + 1] * xa) / (2 * h[i]);
            c[3] = (s[i + 1] - s[i]) / (6 * h[i]);
        }
    }

    T operator()(const T x_val) const
    {
        const T x = std::clamp(x_val, m_start, m_stop);
        const std::size_t index = x > m_start ? static_cast<std::size_t>((x - m_start) / m_step) : 0;
        const std::size_t offset = index < N - 1 ? index * 4 : (N - 2) * 4;
        return m_coefficients[offset] + m_coefficients[offset + 1] * x + m_coefficients[offset + 2] * x * x + m_coefficients[offset + 3] * x * x * x;
    }

    // table access for the device flattening
    const std::array<T, (N - 1) * 4>& coefficients() const { return m_coefficients; }
    const std::array<T, N>& knots() const { return m_x; }
    T start() const { return m_start; }
    T step() const { return m_step; }
    T stop() const { return m_stop; }

protected:
    // Thomas algorithm; sub- and super-diagonal are both h, as in the reference's elimination
    static std::array<T, N> solveTridiagonal(const std::array<T, N>& h, std::array<T, N> diag, std::array<T, N> rhs)
    {
        for (std::size_t i = 1; i < N; ++i) {
            const T w = h[i - 1] / diag[i - 1];
            diag[i] -= w * h[i - 1];
            rhs[i] -= w * rhs[i - 1];
        }
        std::array<T, N> x;
        x[N - 1] = rhs[N - 1] / diag[N - 1];
        for (int i = N - 2; i >= 0; --i)
            x[i] = (rhs[i] - h[i] * x[i + 1]) / diag[i];
        return x;
    }

private:
    std::array<T, (N - 1) * 4> m_coefficients;
    std::array<T, N> m_x;
    T m_step = 0;
    T m_start = 0;
    T m_stop = 0;
};

template <Floating T>
std::vector<T> trapz(const std::vector<T>& f, const std::vector<T>& x)
{
    std::vector<T> integ(f.size(), 0);
    for (std::size_t i = 1; i < f.size(); ++i)
        integ[i] = integ[i - 1] + (f[i - 1] + f[i]) * T { 0.5 } * (x[i] - x[i - 1]);
    return integ;
}

namespace detail {
    // abscissae / weights of the positive half of the 20-point Gauss-Legendre rule
    inline constexpr std::array<double, 10> gaussX = { 7.6526521133497334E-02, 2.2778585114164508E-01, 3.7370608871541956E-01,
        5.1086700195082710E-01, 6.3605368072651503E-01, 7.4633190646015079E-01, 8.3911697182221882E-01, 9.1223442825132591E-01,
        9.6397192727791379E-01, 9.9312859918509492E-01 };
    inline constexpr std::array<double, 10> gaussW = { 1.5275338713072585E-01, 1.4917298647260375E-01, 1.4209610931838205E-01,
        1.3168863844917663E-01, 1.1819453196151842E-01, 1.0193011981724044E-01, 8.3276741576704749E-02, 6.2672048334109064E-02,
        4.0601429800386941E-02, 1.7614007139152118E-02 };
}

template <Floating T>
constexpr std::array<T, 20> gaussIntegrationPoints(const T start, const T stop)
{
    const T half = (stop - start) * T { 0.5 };
    const T mid = (stop + start) * T { 0.5 };
    std::array<T, 20> p;
    for (std::size_t i = 0; i < 10; ++i) {
        p[i] = static_cast<T>(detail::gaussX[i]) * half + mid;
        p[i + 10] = static_cast<T>(-detail::gaussX[i]) * half + mid;
    }
    return p;
}

template <Floating T>
constexpr T gaussIntegration(const T start, const T stop, std::array<T, 20> values)
{
    T sum { 0 };
    for (std::size_t i = 0; i < 20; ++i)
        sum = sum + static_cast<T>(detail::gaussW[i % 10]) * values[i];
    return sum * ((stop - start) * T { 0.5 });
}

template <Floating T, std::regular_invocable<T> F>
    requires std::is_same_v<std::invoke_result_t<F, T>, T>
constexpr T gaussIntegration(const T start, const T stop, const F function)
{
    const auto points = gaussIntegrationPoints(start, stop);
    std::array<T, 20> values;
    for (std::size_t i = 0; i < 20; ++i)
        values[i] = function(points[i]);
    return gaussIntegration(start, stop, values);
}
}
