// attenuationlut.hpp — all interaction look-up data of a world, built on the host.
//
// Public surface of the reference's AttenuationLut<T> (include/dxmc/attenuationlut.hpp:42-275):
// log-log attenuation fits + Woodcock majorant (AttenuationLutInterpolator), one RITA sampler of
// the squared atomic form factor per material (Rayleigh), one cubic spline of the normalised
// incoherent scatter function per material (Compton), and the 12 innermost electron shells per
// material (impulse approximation, fluorescence). Transport flattens these into dxmcb200_luts.
#pragma once
#include "dxmc/attenuationinterpolator.hpp"
#include "dxmc/constants.hpp"
#include "dxmc/dxmcrandom.hpp"
#include "dxmc/hostparallel.hpp"
#include "dxmc/interpolation.hpp"
#include "dxmc/material.hpp"
#include "dxmc/world.hpp"

#include <array>
#include <atomic>
#include <optional>
#include <thread>
#include <vector>

namespace dxmc {

namespace detail {
    // q = E / 12.3985 [1/Angstrom]: the largest momentum transfer a photon of this energy can make
    template <Floating T>
    constexpr T largestMomentumTransfer(T energy) { return energy * (1 / KEV_TO_ANGSTROM<T>()); }

    // Rayleigh: RITA sampler over q^2 in [0, q2max], q2max stepped up until F^2 < 0.001 or the kinematic limit
    template <Floating T, typename Sampler>
    Sampler formFactorSampler(const Material& medium, T maxEnergy)
    {
        const T limit = largestMomentumTransfer(maxEnergy) * largestMomentumTransfer(maxEnergy);
        T upper = 1;
        for (T ff = medium.getRayleightFormFactorSquared(upper); upper < limit && ff > T { 0.001 }; ff = medium.getRayleightFormFactorSquared(upper))
            upper += ff > T { 0.5 } ? T { 0.5 } : T { 0.1 };
        return Sampler(T { 0 }, upper, [&](T q2) -> T { return medium.getRayleightFormFactorSquared(std::sqrt(q2)); });
    }

    // Compton: spline of S(q)/Z over q in [0, qmax], qmax stepped up until S/Z > 0.999 or the kinematic limit
    template <Floating T, typename Spline>
    Spline scatterFunctionSpline(const Material& medium, T maxEnergy)
    {
        const T limit = largestMomentumTransfer(maxEnergy);
        T upper = 0.5;
        T sf = medium.getComptonNormalizedScatterFactor(upper);
        while (sf < T { 0.999 } && upper < limit) {
            sf = medium.getComptonNormalizedScatterFactor(upper);
            upper += sf < T { 0.5 } ? T { 0.5 } : T { 0.1 };
        }
        return Spline(T { 0 }, upper, [&](const T q) -> T { return medium.getComptonNormalizedScatterFactor(q); });
    }
}

template <Floating T = double>
class AttenuationLut {
public:
    using FormFactorSampler = RITA<T, 56>;
    using ScatterFunction = CubicSplineInterpolator<T, 16>;
    using ShellTable = std::array<ElectronShellConfiguration<T>, 12>;

    static constexpr T MIN_PHOTON_ENERGY() { return T { 0.5 }; }
    static constexpr T MAX_PHOTON_ENERGY() { return 2 * ELECTRON_REST_MASS<T>(); }

    // ---- scattering kinematics: q = E sin(theta/2) / 12.3985 [1/Angstrom]
    static T momentumTransferMax(T energy) { return detail::largestMomentumTransfer(energy); }
    static T momentumTransfer(T energy, T angle) { return energy * std::sin(angle * T { 0.5 }) * (1 / KEV_TO_ANGSTROM<T>()); }
    static T momentumTransferFromCos(T energy, T cosAngle)
    {
        return energy * (1 / KEV_TO_ANGSTROM<T>()) * std::sqrt(T { 0.5 } - cosAngle * T { 0.5 });
    }
    static inline T cosAngle(const T energy, const T momentumTransferSquared)
    {
        const auto invE = KEV_TO_ANGSTROM<T>() / energy;
        return 1 - 2 * momentumTransferSquared * invE * invE;
    }

    // ---- look-ups (host equivalents of csrc/physics.cuh)
    std::array<T, 3> photoComptRayAttenuation(std::size_t material, T energy) const { return m_fits(material, energy); }
    T maxTotalAttenuationInverse(T energy) const { return m_fits.maxAttenuationInverse(energy); }
    inline T comptonScatterFactor(std::size_t material, T momentumTransfer) const { return m_scatter[material](momentumTransfer); }
    T momentumTransferFromFormFactor(std::size_t material, const T momentumTransferMax, RandomState& state) const
    {
        return m_rayleigh[material](state, momentumTransferMax);
    }
    const ShellTable& electronShellConfiguration(std::size_t materialIdx) const { return m_shells[materialIdx]; }

    // ---- table access for the device flattening
    T minEnergy() const { return m_lowest; }
    T maxEnergy() const { return m_highest; }
    const AttenuationLutInterpolator<T>& attenuationData() const { return m_fits; }
    const std::vector<FormFactorSampler>& formFactorSamplers() const { return m_rayleigh; }
    const std::vector<ScatterFunction>& scatterFunctions() const { return m_scatter; }
    const std::vector<ShellTable>& electronShellConfigurations() const { return m_shells; }

    // ---- construction
    AttenuationLut() = default;
    AttenuationLut(const World<T>& world, T maxEnergy = 150, T minEnergy = 1) { generate(world, maxEnergy, minEnergy); }

    // per-material tables; material i of the vector gets table index i
    void generate(const std::vector<Material>& materials, T maxEnergy = 150, T minEnergy = 1, bool generatePhotonData = true)
    {
        m_lowest = std::max(MIN_PHOTON_ENERGY(), std::min(maxEnergy, minEnergy));
        m_highest = std::min(MAX_PHOTON_ENERGY(), std::max(maxEnergy, minEnergy));
        // The tables of different materials are independent of each other: three jobs per material on all host cores
        // (the reference builds them one after the other, attenuationlut.hpp:84-92; the values are the same).
        const std::size_t n = materials.size();
        std::vector<std::optional<FormFactorSampler>> rayleigh(n);
        std::vector<std::optional<ScatterFunction>> scatter(n);
        std::vector<ShellTable> shells(n);
        detail::parallelFor(3 * n, [&](std::size_t job) {
            const Material& medium = materials[job / 3];
            switch (job % 3) {
            case 0:
                rayleigh[job / 3].emplace(detail::formFactorSampler<T, FormFactorSampler>(medium, m_highest));
                break;
            case 1:
                scatter[job / 3].emplace(detail::scatterFunctionSpline<T, ScatterFunction>(medium, m_highest));
                break;
            default:
                shells[job / 3] = medium.template getElectronConfiguration<T>();
            }
        });
        m_rayleigh.clear();
        m_scatter.clear();
        for (std::size_t i = 0; i < n; ++i) {
            m_rayleigh.push_back(std::move(*rayleigh[i]));
            m_scatter.push_back(std::move(*scatter[i]));
        }
        // appended, never cleared: a second generate() on the same object keeps the first entries in
        // front (and thereby in use), exactly like the reference (attenuationlut.hpp:89-92)
        m_shells.insert(m_shells.end(), shells.begin(), shells.end());
        if (generatePhotonData)
            m_fits = AttenuationLutInterpolator<T>(materials, m_highest, m_lowest);
    }
    // tables for every material of a valid world; the majorant scans the world's densities on the host ...
    void generate(const World<T>& world, T maxEnergy = 150, T minEnergy = 1)
    {
        generate(world.materialMap(), maxEnergy, minEnergy, false);
        m_fits = AttenuationLutInterpolator<T>(world, m_highest, m_lowest);
    }
    // ... or takes the per-material maximum densities from the caller (Transport gets them from the device)
    void generate(const World<T>& world, const std::vector<T>& maxDensity, T maxEnergy = 150, T minEnergy = 1)
    {
        generate(world.materialMap(), maxEnergy, minEnergy, false);
        generateAttenuation(world, maxDensity);
    }
    // second half of the above, for a caller that ran generate(materials, maxEnergy, minEnergy, false) itself
    void generateAttenuation(const World<T>& world, const std::vector<T>& maxDensity)
    {
        m_fits = AttenuationLutInterpolator<T>(world, maxDensity, m_highest, m_lowest);
    }

private:
    T m_lowest = 0, m_highest = 150.0; // keV
    AttenuationLutInterpolator<T> m_fits;
    std::vector<FormFactorSampler> m_rayleigh;
    std::vector<ScatterFunction> m_scatter;
    std::vector<ShellTable> m_shells;
};
}
