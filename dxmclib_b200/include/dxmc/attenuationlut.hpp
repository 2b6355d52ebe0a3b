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
#include "dxmc/interpolation.hpp"
#include "dxmc/material.hpp"
#include "dxmc/world.hpp"

#include <array>
#include <vector>

namespace dxmc {

template <Floating T = double>
class AttenuationLut {
public:
    using FormFactorSampler = RITA<T, 56>;
    using ScatterFunction = CubicSplineInterpolator<T, 16>;

    static constexpr T MAX_PHOTON_ENERGY() { return 2 * ELECTRON_REST_MASS<T>(); }
    static constexpr T MIN_PHOTON_ENERGY() { return T { 0.5 }; }

    AttenuationLut() = default;
    AttenuationLut(const World<T>& world, T maxEnergy = 150, T minEnergy = 1) { generate(world, maxEnergy, minEnergy); }

    // tables for every material of a valid world; the majorant uses the world's densities
    void generate(const World<T>& world, T maxEnergy = 150, T minEnergy = 1)
    {
        generate(world.materialMap(), maxEnergy, minEnergy, false);
        m_attenuationData = AttenuationLutInterpolator<T>(world, m_maxEnergy, m_minEnergy);
    }

    // material i of the vector gets table index i
    void generate(const std::vector<Material>& materials, T maxEnergy = 150, T minEnergy = 1, bool generatePhotonData = true)
    {
        m_minEnergy = std::max(MIN_PHOTON_ENERGY(), std::min(maxEnergy, minEnergy));
        m_maxEnergy = std::min(MAX_PHOTON_ENERGY(), std::max(maxEnergy, minEnergy));
        buildFormFactorSamplers(materials);
        buildScatterFunctions(materials);
        // appended, never cleared: a second generate() on the same object keeps the first entries in
        // front (and thereby in use), exactly like the reference (attenuationlut.hpp:89-92)
        m_electronShellConfiguration.reserve(materials.size());
        for (const auto& m : materials)
            m_electronShellConfiguration.push_back(m.getElectronConfiguration<T>());
        if (generatePhotonData)
            m_attenuationData = AttenuationLutInterpolator<T>(materials, m_maxEnergy, m_minEnergy);
    }

    T maxTotalAttenuationInverse(T energy) const { return m_attenuationData.maxAttenuationInverse(energy); }
    std::array<T, 3> photoComptRayAttenuation(std::size_t material, T energy) const { return m_attenuationData(material, energy); }
    T momentumTransferFromFormFactor(std::size_t material, const T momentumTransferMax, RandomState& state) const
    {
        return m_formFactor[material](state, momentumTransferMax);
    }
    inline T comptonScatterFactor(std::size_t material, T momentumTransfer) const { return m_comptonScatterFactor[material](momentumTransfer); }
    const std::array<ElectronShellConfiguration<T>, 12>& electronShellConfiguration(std::size_t materialIdx) const
    {
        return m_electronShellConfiguration[materialIdx];
    }

    // q = E sin(theta/2) / 12.3985 [1/Angstrom]
    static T momentumTransfer(T energy, T angle)
    {
        constexpr T k = 1 / KEV_TO_ANGSTROM<T>();
        return energy * std::sin(angle * T { 0.5 }) * k;
    }
    static T momentumTransferFromCos(T energy, T cosAngle)
    {
        constexpr T k = 1 / KEV_TO_ANGSTROM<T>();
        return energy * k * std::sqrt(T { 0.5 } - cosAngle * T { 0.5 });
    }
    static T momentumTransferMax(T energy)
    {
        constexpr T k = 1 / KEV_TO_ANGSTROM<T>();
        return energy * k;
    }
    static inline T cosAngle(const T energy, const T momentumTransferSquared)
    {
        const auto invE = KEV_TO_ANGSTROM<T>() / energy;
        return 1 - 2 * momentumTransferSquared * invE * invE;
    }

    // table access for the device flattening
    const AttenuationLutInterpolator<T>& attenuationData() const { return m_attenuationData; }
    const std::vector<FormFactorSampler>& formFactorSamplers() const { return m_formFactor; }
    const std::vector<ScatterFunction>& scatterFunctions() const { return m_comptonScatterFactor; }
    const std::vector<std::array<ElectronShellConfiguration<T>, 12>>& electronShellConfigurations() const { return m_electronShellConfiguration; }
    T minEnergy() const { return m_minEnergy; }
    T maxEnergy() const { return m_maxEnergy; }

protected:
    // RITA over q^2 in [0, q2max] where q2max is stepped up until F^2 < 0.001 or the kinematic limit
    void buildFormFactorSamplers(const std::vector<Material>& materials)
    {
        m_formFactor.clear();
        m_formFactor.reserve(materials.size());
        const auto qmax = momentumTransferMax(m_maxEnergy);
        const auto qmaxSquared = qmax * qmax;
        for (const auto& m : materials) {
            T upper = 1;
            T ff = m.getRayleightFormFactorSquared(upper);
            while (upper < qmaxSquared && ff > T { 0.001 }) {
                upper += ff > T { 0.5 } ? T { 0.5 } : T { 0.1 };
                ff = m.getRayleightFormFactorSquared(upper);
            }
            m_formFactor.emplace_back(T { 0 }, upper, [&](T q2) -> T { return m.getRayleightFormFactorSquared(std::sqrt(q2)); });
        }
    }

    // spline over q in [0, qmax] where qmax is stepped up until S/Z > 0.999 or the kinematic limit
    void buildScatterFunctions(const std::vector<Material>& materials)
    {
        m_comptonScatterFactor.clear();
        m_comptonScatterFactor.reserve(materials.size());
        for (const auto& m : materials) {
            const T qmaxEnergy = momentumTransferMax(m_maxEnergy);
            T upper = 0.5;
            T sf = m.getComptonNormalizedScatterFactor(upper);
            while (sf < T { 0.999 } && upper < qmaxEnergy) {
                sf = m.getComptonNormalizedScatterFactor(upper);
                upper += sf < T { 0.5 } ? T { 0.5 } : T { 0.1 };
            }
            m_comptonScatterFactor.emplace_back(T { 0 }, upper, [&](const T q) -> T { return m.getComptonNormalizedScatterFactor(q); });
        }
    }

private:
    T m_minEnergy = 0;
    T m_maxEnergy = 150.0;
    std::vector<ScatterFunction> m_comptonScatterFactor;
    std::vector<FormFactorSampler> m_formFactor;
    AttenuationLutInterpolator<T> m_attenuationData;
    std::vector<std::array<ElectronShellConfiguration<T>, 12>> m_electronShellConfiguration;
};
}
