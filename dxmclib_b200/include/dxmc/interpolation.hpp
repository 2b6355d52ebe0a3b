// interpolation.hpp — host-side interpolation helpers used while BUILDING device tables.
//
// Same names and numerical results as the reference (include/dxmc/interpolation.hpp:33-261):
// linear / log-log interpolation and the clamped table look-up here; the fixed-knot natural cubic
// spline behind the Compton scatter function in dxmc/cubicspline.hpp; trapezoid and 20-point
// Gauss-Legendre quadrature in dxmc/quadrature.hpp. Including this header brings in all three.
#pragma once
#include "dxmc/types.hpp"

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
}

#include "dxmc/cubicspline.hpp"
#include "dxmc/quadrature.hpp"
