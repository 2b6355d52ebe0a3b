// quadrature.hpp — cumulative trapezoid rule and the 20-point Gauss-Legendre rule used by the host-side table
// builders (reference include/dxmc/interpolation.hpp:178-261). Included by dxmc/interpolation.hpp.
#pragma once
#include "dxmc/types.hpp"

#include <array>
#include <concepts>
#include <type_traits>
#include <vector>

namespace dxmc {

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
