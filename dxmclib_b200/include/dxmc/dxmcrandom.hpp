// dxmcrandom.hpp — host-side random number helpers and sampler BUILDERS.
//
// Public surface of the reference's include/dxmc/dxmcrandom.hpp (RandomState :37-167,
// RandomDistribution :173-279, SpecterDistribution :284-330, RITA :333-501). On the B200 path the
// tables these classes build (alias table, RITA knots) are uploaded and sampled by the kernels
// (csrc/physics.cuh sampleSpectrum / sampleFormFactor) with one counter-derived PCG32 stream per
// photon history; the host-side sampling methods are kept for API compatibility and for tests.
#pragma once
#include "dxmc/floating.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <concepts>
#include <cstdint>
#include <numeric>
#include <random>
#include <vector>

namespace dxmc {

// PCG32 (XSH-RR, 64-bit state, selectable stream)
class RandomState {
public:
    RandomState()
    {
        std::random_device d;
        std::uniform_int_distribution<std::uint64_t> dist(0);
        m_state[0] = dist(d);
        m_state[1] = dist(d);
    }
    RandomState(std::uint64_t state[2])
    {
        m_state[0] = state[0];
        m_state[1] = state[1];
    }
    RandomState(const RandomState&) = delete;
    RandomState& operator=(const RandomState&) = delete;

    // [0, 1]; the conversion of a 32 bit integer to float rounds, so exactly 1 is possible in float
    template <typename T>
    inline T randomUniform() noexcept
    {
        static_assert(std::is_floating_point_v<T>, "Uniform random number requires floating point precision");
        constexpr T scale = T { 2.32830643653869628906e-010 };
        return pcg32() * scale;
    }

    // [0, max)
    template <typename T>
    inline T randomUniform(const T max) noexcept
    {
        if constexpr (std::is_floating_point_v<T>) {
            return randomUniform<T>() * max;
        } else {
            static_assert(std::is_integral_v<T>, "Must be integral or floating point value.");
            // rejection threshold evaluated in T's own width, then truncated to 32 bits
            const std::uint32_t threshold = static_cast<std::uint32_t>(-max % max);
            for (;;) {
                const auto r = pcg32();
                if (r >= threshold)
                    return static_cast<T>(r % static_cast<std::uint32_t>(max));
            }
        }
    }

    // [min, max)
    template <typename T>
    inline T randomUniform(const T min, const T max) noexcept
    {
        if constexpr (std::is_floating_point_v<T>) {
            const T r = randomUniform<T>();
            const T range = max - min;
            return min + r * range;
        } else {
            static_assert(std::is_integral_v<T>, "Must be integral or floating point value.");
            return min + randomUniform<T>(max - min);
        }
    }

    template <std::unsigned_integral T>
    inline T randomInteger(const T max) noexcept
    {
        static_assert(sizeof(max) <= 4, "This prng only supports up to 32 bit random integers, for a capped to 32 bit random integer use randomInteger32BitCapped instead");
        const T threshold = (static_cast<T>(-max)) % max;
        for (;;) {
            const auto r = pcg32();
            if (r >= threshold)
                return r % max;
        }
    }

    template <std::unsigned_integral T>
    inline T randomInteger32BitCapped(const T max) noexcept
    {
        static_assert(sizeof(max) > 4, "This function is intended for 64 bit values or greater, use randomInteger method instead");
        const T threshold = (static_cast<T>(-max)) % max;
        for (;;) {
            const auto r = pcg32();
            if (r >= threshold)
                return r % max;
        }
    }

    inline std::uint32_t pcg32() noexcept
    {
        const std::uint64_t old = m_state[0];
        m_state[0] = old * 6364136223846793005ULL + (m_state[1] | 1);
        const std::uint32_t xorshifted = static_cast<std::uint32_t>(((old >> 18u) ^ old) >> 27u);
        const std::uint32_t rot = static_cast<std::uint32_t>(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
    }
    std::uint64_t m_state[2];
};

// Walker alias table over a discrete distribution
template <Floating T = double>
class RandomDistribution {
public:
    RandomDistribution(const std::vector<T>& weights) { buildAliasTable(weights); }
    RandomDistribution(const RandomDistribution& other)
        : m_alias(other.m_alias)
        , m_probs(other.m_probs)
    {
    }
    RandomDistribution& operator=(const RandomDistribution& other)
    {
        m_alias = other.m_alias;
        m_probs = other.m_probs;
        return *this;
    }

    std::size_t sampleIndex() { return sampleIndex(m_state); }
    std::size_t sampleIndex(RandomState& state) const
    {
        const auto r = state.randomUniform<T>();
        const auto k = state.randomUniform<std::size_t>(size());
        return r < m_probs[k] ? k : m_alias[k];
    }
    std::size_t size() const { return m_probs.size(); }
    const std::vector<std::uint64_t>& aliasingData() const { return m_alias; }
    const std::vector<T>& probabilityData() const { return m_probs; }

protected:
    RandomState m_state;

    // "squaring the histogram": pair every under-full bin with an over-full one
    void buildAliasTable(const std::vector<T>& weights)
    {
        const std::int64_t n = static_cast<std::int64_t>(weights.size());
        m_probs.assign(n, T { 0 });
        m_alias.assign(n, 0);
        const T sum = std::accumulate(weights.begin(), weights.end(), T { 0.0 });
        const T scale = weights.size() / sum;
        std::vector<T> p(n);
        for (std::int64_t i = 0; i < n; ++i)
            p[i] = weights[i] * scale;

        std::vector<std::int64_t> small, large; // used as stacks, filled from the last bin down
        small.reserve(n);
        large.reserve(n);
        for (std::int64_t i = n - 1; i >= 0; --i)
            (p[i] < T { 1.0 } ? small : large).push_back(i);

        while (!small.empty() && !large.empty()) {
            const auto s = small.back();
            const auto l = large.back();
            small.pop_back();
            large.pop_back();
            m_probs[s] = p[s];
            m_alias[s] = l;
            p[l] = p[l] + p[s] - 1;
            (p[l] < 1 ? small : large).push_back(l);
        }
        for (auto i : large)
            m_probs[i] = 1;
        for (auto i : small)
            m_probs[i] = 1;
    }

private:
    std::vector<std::uint64_t> m_alias;
    std::vector<T> m_probs;
};

// energy spectrum: alias-sampled bin, then uniform inside [E_i, E_i+1)
template <Floating T = double>
class SpecterDistribution : public RandomDistribution<T> {
public:
    SpecterDistribution(const std::vector<T>& weights, const std::vector<T>& energies)
        : RandomDistribution<T>(weights)
        , m_energies(energies)
    {
    }
    SpecterDistribution()
        : RandomDistribution<T>(std::vector<T> { 1 })
        , m_energies { 60 }
    {
    }
    T sampleValue() { return sampleValue(this->m_state); }
    T sampleValue(RandomState& state) const
    {
        const std::size_t ind = this->sampleIndex(state);
        return ind < m_energies.size() - 1 ? state.randomUniform(m_energies[ind], m_energies[ind + 1]) : m_energies[ind];
    }
    const std::vector<T>& energies() const { return m_energies; }

private:
    std::vector<T> m_energies;
};

// Rational Inverse Transform with Aliasing (PENELOPE) on an adaptive grid of N knots: numerical
// inversion of the cumulative of an analytical pdf. Used for the squared form factor.
template <Floating T, int N = 20>
class RITA {
public:
    template <std::regular_invocable<T> F>
        requires std::is_same_v<std::invoke_result_t<F, T>, T>
    RITA(const T min, const T max, F pdf)
    {
        struct Knot {
            T x, e, a, b, error;
        };
        std::size_t n = 10;
        std::vector<Knot> v(n);
        v.reserve(N);
        for (std::size_t i = 0; i < n; ++i)
            v[i] = { min + i * (max - min) / (n - 1), 0, 0, 0, -1 };

        // cumulative by Simpson on every interval, normalised by the running total; the last entry is
        // divided last so all earlier ones see the un-normalised total. The reference re-integrates every interval
        // after every insertion (dxmcrandom.hpp:384-468: 77 000 pdf evaluations per table, 90 ms for a nine-element
        // tissue); an interval's integral only depends on its end points, so it is kept and only the two halves of a
        // bisected interval are integrated anew. Same values summed in the same order: the tables keep their bits.
        std::vector<T> integral(n, T { 0 }); // integral[j]: Simpson over [x[j-1], x[j]]
        integral.reserve(N);
        for (std::size_t j = 1; j < n; ++j)
            integral[j] = simpson(v[j - 1].x, v[j].x, pdf);
        auto cumulate = [&]() {
            for (std::size_t j = 1; j < n; ++j)
                v[j].e = v[j - 1].e + integral[j];
        };
        auto normalise = [&]() {
            for (std::size_t j = 1; j < n; ++j)
                v[j].e = v[j].e / v[n - 1].e;
        };
        cumulate();
        const T total = v[n - 1].e;
        normalise();

        while (n != N) {
            for (std::size_t i = 0; i + 1 < n; ++i) {
                if (!(v[i].error < 0))
                    continue;
                const T dx = v[i + 1].x - v[i].x;
                const T de = v[i + 1].e - v[i].e;
                const T temp = de / dx;
                const T px0 = pdf(v[i].x) / total;
                const T px1 = pdf(v[i + 1].x) / total;
                v[i].b = (px0 > 0 && px1 > 0) ? 1 - temp * temp / (px0 * px1) : 0;
                v[i].a = px0 > 0 ? temp / px0 - v[i].b - 1 : 0;
                const T a = v[i].a, b = v[i].b;
                // L1 distance between the rational interpolant and the pdf on 49 interior points
                v[i].error = 0;
                for (std::size_t j = 1; j < 50; ++j) {
                    const T x = v[i].x + j * dx / 50;
                    const T t = (x - v[i].x) / dx;
                    const T f = (1 + a + b - a * t);
                    const T nn = f * (1 - std::sqrt(1 - 4 * b * t * t / (f * f))) / (2 * b * t);
                    const T p1 = (1 + a * nn + b * nn * nn);
                    const T p = p1 * p1 * de / ((1 + a + b) * (1 - b * nn * nn) * dx);
                    v[i].error += std::abs(p - pdf(x)) * dx / 50;
                }
            }
            // bisect the worst interval
            auto worst = std::max_element(v.begin(), v.end(), [](const Knot& l, const Knot& r) { return l.error < r.error; });
            const T x0 = worst->x;
            worst->error = -1;
            ++worst;
            const T x1 = worst->x;
            const auto at = static_cast<std::size_t>(std::distance(v.begin(), worst)); // the new knot's index
            v.insert(worst, Knot { x0 + (x1 - x0) / 2, 0, 0, 0, -1 });
            ++n;
            integral.insert(integral.begin() + static_cast<std::ptrdiff_t>(at), simpson(v[at - 1].x, v[at].x, pdf));
            integral[at + 1] = simpson(v[at].x, v[at + 1].x, pdf);
            cumulate();
            normalise();
        }
        for (std::size_t i = 0; i < N; ++i) {
            m_x[i] = v[i].x;
            m_e[i] = v[i].e;
            m_a[i] = v[i].a;
            m_b[i] = v[i].b;
        }
    }

    T operator()(RandomState& state) const
    {
        const auto r1 = state.randomUniform<T>();
        const std::size_t index = std::distance(m_e.cbegin(), std::upper_bound(m_e.cbegin(), m_e.cend(), r1));
        if (index == 0)
            return m_x[0];
        return invert(index - 1, r1);
    }

    // truncated to [min, maxValue] by rejection on the restricted cumulative
    T operator()(RandomState& state, const T maxValue) const
    {
        const auto ub = std::upper_bound(m_x.cbegin(), m_x.cend(), maxValue);
        const T modifier = ub != m_x.cend() ? m_e[std::distance(m_x.cbegin(), ub)] : 1;
        T res;
        do {
            const auto r1 = state.randomUniform<T>(modifier);
            const auto index = std::distance(m_e.cbegin(), std::upper_bound(m_e.cbegin(), m_e.cend(), r1)) - 1;
            res = invert(index, r1);
        } while (res > maxValue);
        return res;
    }

    // table access for the device flattening
    const std::array<T, N>& x() const { return m_x; }
    const std::array<T, N>& e() const { return m_e; }
    const std::array<T, N>& a() const { return m_a; }
    const std::array<T, N>& b() const { return m_b; }

protected:
    template <typename F>
    static T simpson(const T start, const T stop, F pdf)
    {
        const T h = (stop - start) / 50;
        T result = pdf(start) + pdf(stop);
        for (std::size_t i = 1; i < 50; ++i) {
            const T w = i % 2 == 0 ? 2 : 4;
            result += w * pdf(start + h * i);
        }
        return h * result / 3;
    }

    T invert(std::ptrdiff_t i, T r1) const
    {
        const auto v = r1 - m_e[i];
        const auto d = m_e[i + 1] - m_e[i];
        return m_x[i] + (1 + m_a[i] + m_b[i]) * d * v / (d * d + m_a[i] * d * v + m_b[i] * v * v) * (m_x[i + 1] - m_x[i]);
    }

private:
    std::array<T, N> m_x;
    std::array<T, N> m_e;
    std::array<T, N> m_b;
    std::array<T, N> m_a;
};
}
