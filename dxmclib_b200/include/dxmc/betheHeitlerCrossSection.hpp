// betheHeitlerCrossSection.hpp — bremsstrahlung of a thick tungsten target.
//
// Semi-empirical model of Poludniowski & Evans (Med. Phys. 34, 2007, 2164-2174 and 2175-2186):
// electrons entering tungsten with kinetic energy T0 are described by the Monte-Carlo derived
// joint density of depth and remaining energy (dxmc/tungsten_electron_data.hpp), scaled to T0 by
// the Thomson-Whiddington range; at every (depth, energy) the semi-relativistic Bethe-Heitler
// cross section gives the photon yield, which is attenuated along the take-off direction inside
// the anode. Function names and results follow reference
// include/dxmc/betheHeitlerCrossSection.hpp:174-411; host-only, runs once per source.
#pragma once
#include "dxmc/constants.hpp"
#include "dxmc/floating.hpp"
#include "dxmc/material.hpp"
#include "dxmc/tungsten_electron_data.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <utility>

namespace dxmc::BetheHeitlerCrossSection {

constexpr int TUNGSTEN_ATOMIC_NUMBER = 74;

template <Floating T>
constexpr T SIMULATED_ENERGY() { return T { 100.0 }; } // keV of the tabulated electron transport

template <Floating T>
constexpr T FINE_STRUCTURE_CONSTANT() { return T { 7.29735308E-03 }; }

template <Floating T>
constexpr T CLASSIC_ELECTRON_RADIUS() { return T { 2.81794092E-15 }; } // m

template <Floating T>
constexpr T PHI_BAR()
{
    return (TUNGSTEN_ATOMIC_NUMBER * TUNGSTEN_ATOMIC_NUMBER) * CLASSIC_ELECTRON_RADIUS<T>() * CLASSIC_ELECTRON_RADIUS<T>() * FINE_STRUCTURE_CONSTANT<T>();
}

namespace detail {
    // Index of the table interval used for value v: the interval that STARTS at the first knot >= v,
    // pulled back so that two knots remain (this is the reference's lower_bound convention).
    template <Floating T, std::size_t N>
    inline std::size_t intervalStart(const std::array<double, N>& knots, const T v)
    {
        std::size_t i = 0;
        while (i < N && static_cast<T>(knots[i]) < v)
            ++i;
        const std::size_t left = N - i; // knots from i to the end
        if (left < 2)
            i = N - 3 + left;
        return i;
    }

    template <typename T>
    inline T bilinear(const T q11, const T q12, const T q21, const T q22, const T x1, const T x2, const T y1, const T y2, const T x, const T y)
    {
        const auto xf1 = ((x2 - x) / (x2 - x1));
        const auto xf2 = ((x - x1) / (x2 - x1));
        const auto r1 = xf1 * q11 + xf2 * q21;
        const auto r2 = xf1 * q12 + xf2 * q22;
        return ((y2 - y) / (y2 - y1)) * r1 + ((y - y1) / (y2 - y1)) * r2;
    }

    // density of relative energy u at depth x from one of the two depth-major tables
    template <Floating T>
    inline T tableDensity(const std::array<double, tungsten::kDepths>& depths, const double (&table)[tungsten::kDepths][tungsten::kEnergies], const T uval,
        const T xval)
    {
        const T x = std::clamp(xval, static_cast<T>(depths.front()), static_cast<T>(depths.back()));
        const T u = std::clamp(uval, static_cast<T>(tungsten::relEnergy.front()), static_cast<T>(tungsten::relEnergy.back()));
        const std::size_t ix = intervalStart(depths, x);
        const std::size_t iu = intervalStart(tungsten::relEnergy, u);
        const T q11 = static_cast<T>(table[ix][iu]);
        const T q21 = static_cast<T>(table[ix + 1][iu]);
        const T q22 = static_cast<T>(table[ix + 1][iu + 1]);
        const T q12 = static_cast<T>(table[ix][iu + 1]);
        return bilinear(q11, q12, q21, q22, static_cast<T>(depths[ix]), static_cast<T>(depths[ix + 1]), static_cast<T>(tungsten::relEnergy[iu]),
            static_cast<T>(tungsten::relEnergy[iu + 1]), x, u);
    }
}

// range [mg/cm2] of an electron of T0 keV in tungsten, valid 50-150 keV
template <Floating T>
constexpr T ThomsonWiddingtonRange(const T T0) { return T { 0.0119 } * std::pow(T0, T { 1.513 }); }

// remaining fraction of T0^2 at depth x
template <Floating T>
T ThomsonWiddingtonLaw(const T x, const T tubeVoltage)
{
    const std::size_t i = detail::intervalStart(tungsten::twVoltage, tubeVoltage);
    const T t1 = static_cast<T>(tungsten::twVoltage[i]), t2 = static_cast<T>(tungsten::twVoltage[i + 1]);
    const T c1 = static_cast<T>(tungsten::twConstant[i]), c2 = static_cast<T>(tungsten::twConstant[i + 1]);
    const auto C = c1 + ((c2 - c1) / (t2 - t1)) * (tubeVoltage - t1);
    const auto twl = (tubeVoltage * tubeVoltage - C * x) / (tubeVoltage * tubeVoltage);
    return twl < 0 ? 0 : twl;
}

template <Floating T>
T numberFractionF(const T x, const T tubeVoltage)
{
    constexpr T L = 1.753;
    return std::pow(ThomsonWiddingtonLaw(x, tubeVoltage), L);
}

template <Floating T>
T numberFractionM(const T x, T tubeVoltage)
{
    constexpr auto K = T { 18.0 };
    constexpr auto Bd = T { 0.584 };
    constexpr auto B0 = T { 0.5 };
    const auto grown = 1 - std::exp(-K * x / ThomsonWiddingtonRange(tubeVoltage));
    const auto F = Bd * grown;
    const auto B = B0 + (Bd - B0) * grown;
    return numberFractionF(x, tubeVoltage) * B * (F + 1) / (1 - B * F);
}

template <Floating T>
T electronDensity_F(const T uval, const T xval) { return detail::tableDensity<T>(tungsten::depthF, tungsten::densityF, uval, xval); }

template <Floating T>
T electronDensity_M(const T uval, const T xval) { return detail::tableDensity<T>(tungsten::depthM, tungsten::densityM, uval, xval); }

template <Floating T>
T electronDensity(const T u, const T x, const T tubeVoltage)
{
    const auto f = ThomsonWiddingtonRange(SIMULATED_ENERGY<T>()) / ThomsonWiddingtonRange(tubeVoltage);
    return numberFractionF(x, tubeVoltage) * electronDensity_F(u, x * f) + numberFractionM(x, tubeVoltage) * electronDensity_M(u, x * f);
}

template <Floating T>
T tungstenFiltration(const T tungstenAtt, const T x, const T takeoffAngle)
{
    return std::exp(-tungstenAtt * x * T { 0.001 } / std::sin(takeoffAngle)); // mg/cm2 -> g/cm2
}

// differential in photon energy hv for an electron of kinetic energy Ti, with Elwert-like factor pi/pf
template <Floating T>
T betheHeitlerCrossSection(const T hv, const T Ti)
{
    constexpr T scale = (PHI_BAR<T>() * 2) / 3;
    constexpr T m = ELECTRON_REST_MASS<T>();
    const auto Ei = m + Ti;
    const auto Ef = Ei - hv;
    const auto pi2 = Ei * Ei - m * m;
    const auto pi = std::sqrt(pi2);
    const auto pf2 = Ef * Ef - m * m;
    if (pf2 <= 0)
        return 0;
    const auto pf = std::sqrt(pf2);
    const auto L = 2 * std::log((Ei * Ef + pi * pf - m * m) / (m * hv));
    const auto coulomb = pi / pf;
    return scale * (4 * Ei * Ef * L - 7 * pi * pf) / (hv * pi * pi) * coulomb;
}

// photons of energy hv leaving the anode per incident electron of T0, integrated over depth (0..14
// mg/cm2, step 0.1) and electron energy (u = 0.005..1, step 0.005); loop accumulation as the reference
template <Floating T>
T betheHeitlerSpectra(const T T0, const T hv, const T takeoffAngle)
{
    const T tungstenTotAtt = Material::getTotalAttenuation(TUNGSTEN_ATOMIC_NUMBER, hv);
    constexpr auto xmax = T { 14.0 };
    constexpr auto umax = T { 1.0 };
    constexpr auto xstep = T { 0.1 };
    constexpr auto ustep = T { 0.005 };
    if (hv <= 0)
        return 0;
    T total = 0;
    for (T x = 0; x <= xmax; x = x + xstep) {
        T atDepth = 0;
        for (T u = ustep; u <= umax; u = u + ustep)
            atDepth = atDepth + betheHeitlerCrossSection(hv, T0 * u) * electronDensity(u, x, T0) * ustep;
        total = total + atDepth * tungstenFiltration(tungstenTotAtt, x, takeoffAngle) * xstep;
    }
    return total;
}

// K-alpha / K-beta lines scaled to the bremsstrahlung produced at the line energy; the fifth slot is unused
template <Floating T>
constexpr std::array<T, 4> tungstenKLineEnergies() { return { 59.3, 58.0, 67.2, 69.1 }; }

// ... from the bremsstrahlung yields at the four line energies, wherever they were evaluated
template <Floating T>
std::array<std::pair<T, T>, 5> characteristicTungstenKedge(const std::array<T, 4>& bremsstrahlungAtLines)
{
    constexpr std::array<T, 4> lineEnergy = tungstenKLineEnergies<T>();
    constexpr std::array<T, 4> lineFraction { 0.505, 0.291, 0.162, 0.042 };
    constexpr auto P = T { 0.33 };
    constexpr auto omega_k = T { 0.94 };
    constexpr auto rk = T { 4.4 };
    std::array<std::pair<T, T>, 5> lines {};
    for (std::size_t i = 0; i < 4; ++i)
        lines[i] = { lineEnergy[i], (1 + P) * lineFraction[i] * rk * omega_k * bremsstrahlungAtLines[i] };
    return lines;
}

template <Floating T>
std::array<std::pair<T, T>, 5> characteristicTungstenKedge(const T T0, const T takeoffAngle)
{
    constexpr std::array<T, 4> lineEnergy = tungstenKLineEnergies<T>();
    std::array<T, 4> yields {};
    for (std::size_t i = 0; i < 4; ++i)
        yields[i] = betheHeitlerSpectra(T0, lineEnergy[i], takeoffAngle);
    return characteristicTungstenKedge(yields);
}
}
