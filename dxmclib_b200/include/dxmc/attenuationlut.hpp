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

    // the same with the per-material maximum densities supplied by the caller (Transport gets them from the device)
    void generate(const World<T>& world, const std::vector<T>& maxDensity, T maxEnergy = 150, T minEnergy = 1)
    {
        generate(world.materialMap(), maxEnergy, minEnergy, false);
        generateAttenuation(world, maxDensity);
    }
    // second half of the above, for a caller that ran generate(materials, maxEnergy, minEnergy, false) itself
    void generateAttenuation(const World<T>& world, const std::vector<T>& maxDensity)
    {
        m_attenuationData = AttenuationLutInterpolator<T>(world, maxDensity, m_maxEnergy, m_minEnergy);
    }

    // material i of the vector gets table index i
    void generate(const std::vector<Material>& materials, T maxEnergy = 150, T minEnergy = 1, bool generatePhotonData = true)
    {
        m_minEnergy = std::max(MIN_PHOTON_ENERGY(), std::min(maxEnergy, minEnergy));
        m_maxEnergy = std::min(MAX_PHOTON_ENERGY(), std::max(maxEnergy, minEnergy));
        // The per-material tables are independent of each other: build them on all host cores (the reference builds
        // them one after the other, attenuationlut.hpp:84-92; the values are the same).
        std::vector<std::optional<FormFactorSampler>> samplers(materials.size());
        std::vector<std::optional<ScatterFunction>> scatter(materials.size());
        std::vector<std::array<ElectronShellConfiguration<T>, 12>> shells(materials.size());
        detail::parallelFor(materials.size() * 3, [&](std::size_t job) {
            const std::size_t i = job / 3;
            if (job % 3 == 0)
                samplers[i].emplace(buildFormFactorSampler(materials[i]));
            else if (job % 3 == 1)
                scatter[i].emplace(buildScatterFunction(materials[i]));
            else
                shells[i] = materials[i].template getElectronConfiguration<T>();
        });
        m_formFactor.clear();
        m_comptonScatterFactor.clear();
        for (std::size_t i = 0; i < materials.size(); ++i) {
            m_formFactor.push_back(std::move(*samplers[i]));
            m_comptonScatterFactor.push_back(std::move(*scatter[i]));
        }
        // appended, never cleared: a second generate() on the same object keeps the first entries in
        // front (and thereby in use), exactly like the reference (attenuationlut.hpp:89-92)
        m_electronShellConfiguration.reserve(m_electronShellConfiguration.size() + materials.size());
        for (const auto& sh : shells)
            m_electronShellConfiguration.push_back(sh);
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
    FormFactorSampler buildFormFactorSampler(const Material& m) const
    {
        const auto qmax = momentumTransferMax(m_maxEnergy);
        const auto qmaxSquared = qmax * qmax;
        T upper = 1;
        T ff = m.getRayleightFormFactorSquared(upper);
        while (upper < qmaxSquared && ff > T { 0.001 }) {
            upper += ff > T { 0.5 } ? T { 0.5 } : T { 0.1 };
            ff = m.getRayleightFormFactorSquared(upper);
        }
        return FormFactorSampler(T { 0 }, upper, [&](T q2) -> T { return m.getRayleightFormFactorSquared(std::sqrt(q2)); });
    }

    // spline over q in [0, qmax] where qmax is stepped up until S/Z > 0.999 or the kinematic limit
    ScatterFunction buildScatterFunction(const Material& m) const
    {
        const T qmaxEnergy = momentumTransferMax(m_maxEnergy);
        T upper = 0.5;
        T sf = m.getComptonNormalizedScatterFactor(upper);
        while (sf < T { 0.999 } && upper < qmaxEnergy) {
            sf = m.getComptonNormalizedScatterFactor(upper);
            upper += sf < T { 0.5 } ? T { 0.5 } : T { 0.1 };
        }
        return ScatterFunction(T { 0 }, upper, [&](const T q) -> T { return m.getComptonNormalizedScatterFactor(q); });
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
