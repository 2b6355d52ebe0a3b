// cubicspline.hpp — the natural cubic spline behind the Compton scatter function (reference
// include/dxmc/interpolation.hpp:84-176), built on the host; it exposes its tables so that Transport can
// flatten them for the GPU (csrc/physics.cuh scatterFactor()). Included by dxmc/interpolation.hpp.
#pragma once
#include "dxmc/types.hpp"

#include <algorithm>
#include <array>
#include <concepts>
#include <type_traits>

namespace dxmc {

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
}
