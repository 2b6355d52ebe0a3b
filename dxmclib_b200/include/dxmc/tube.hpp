// tube.hpp — tungsten-anode x-ray tube: voltage, anode angle, added filtration -> photon spectrum.
//
// Public surface of the reference's Tube<T> (include/dxmc/tube.hpp:35-299). The spectrum feeds the
// alias table (SpecterDistribution) and the heel-effect table (HeelFilter) that Transport uploads
// to the GPU; generation itself is host-side and runs once per source.
#pragma once
#include "dxmc/betheHeitlerCrossSection.hpp"
#include "dxmc/constants.hpp"
#include "dxmc/floating.hpp"
#include "dxmc/hostparallel.hpp"
#include "dxmc/material.hpp"

#include <algorithm>
#include <cmath>
#include <execution>
#include <numeric>
#include <utility>
#include <vector>

namespace dxmc {

template <Floating T = double>
class Tube {
public:
    Tube(T tubeVoltage = 120.0, T anodeAngleDeg = 12.0, T energyResolution = 1.0)
        : m_voltage(tubeVoltage)
        , m_energyResolution(energyResolution)
    {
        setAnodeAngleDeg(anodeAngleDeg);
    }
    Tube(const Tube<T>& other)
        : m_voltage(other.m_voltage)
        , m_energyResolution(other.m_energyResolution)
        , m_anodeAngle(other.m_anodeAngle)
        , m_filtrationMaterials(other.m_filtrationMaterials)
    {
    }
    Tube& operator=(const Tube<T>& other)
    {
        m_voltage = other.m_voltage;
        m_energyResolution = other.m_energyResolution;
        m_anodeAngle = other.m_anodeAngle;
        m_filtrationMaterials = other.m_filtrationMaterials;
        m_hasCachedHVL = false;
        return *this;
    }

    static constexpr T maxVoltage() { return T { 150 }; }
    static constexpr T minVoltage() { return T { 50 }; }

    T voltage() const { return m_voltage; }
    void setVoltage(T voltage)
    {
        m_voltage = std::min(std::max(voltage, minVoltage()), maxVoltage());
        m_hasCachedHVL = false;
    }

    T anodeAngle() const { return m_anodeAngle; }
    T anodeAngleDeg() const { return m_anodeAngle * RAD_TO_DEG<T>(); }
    void setAnodeAngle(T angle)
    {
        m_anodeAngle = std::min(std::abs(angle), PI_VAL<T>() * T { 0.5 });
        m_hasCachedHVL = false;
    }
    void setAnodeAngleDeg(T angle) { setAnodeAngle(angle * DEG_TO_RAD<T>()); }

    void addFiltrationMaterial(const Material& filtrationMaterial, T mm)
    {
        m_filtrationMaterials.emplace_back(filtrationMaterial, std::abs(mm));
        m_hasCachedHVL = false;
    }
    std::vector<std::pair<Material, T>>& filtrationMaterials() { return m_filtrationMaterials; }
    const std::vector<std::pair<Material, T>>& filtrationMaterials() const { return m_filtrationMaterials; }
    void clearFiltrationMaterials() { m_filtrationMaterials.clear(); }

    void setAlFiltration(T mm) { setElementFiltration(13, "Al", mm); }
    void setCuFiltration(T mm) { setElementFiltration(29, "Cu", mm); }
    void setSnFiltration(T mm) { setElementFiltration(50, "Sn", mm); }
    T AlFiltration() const { return elementFiltration("Al"); }
    T CuFiltration() const { return elementFiltration("Cu"); }
    T SnFiltration() const { return elementFiltration("Sn"); }

    void setEnergyResolution(T energyResolution) { m_energyResolution = energyResolution; }
    T energyResolution() const { return m_energyResolution; }

    // bin energies: resolution, 2*resolution, ... <= voltage
    std::vector<T> getEnergy() const
    {
        std::vector<T> energies;
        energies.reserve(static_cast<std::size_t>(std::ceil(m_voltage / m_energyResolution)));
        for (T hv = m_energyResolution; hv <= m_voltage; hv = hv + m_energyResolution)
            energies.push_back(hv);
        return energies;
    }

    std::vector<std::pair<T, T>> getSpecter(bool normalize = true) const
    {
        const auto energies = getEnergy();
        const auto specter = getSpecter(energies, normalize);
        std::vector<std::pair<T, T>> out;
        out.reserve(specter.size());
        for (std::size_t i = 0; i < specter.size(); ++i)
            out.emplace_back(energies[i], specter[i]);
        return out;
    }
    // bremsstrahlung + tungsten K lines, filtered by the added materials, at a given take-off angle
    std::vector<T> getSpecter(const std::vector<T>& energies, const T anodeAngle, bool normalize = true) const
    {
        std::vector<T> specter(energies.size());
        // one Bethe-Heitler depth integral per energy bin, independent of each other (reference tube.hpp:191-208)
        detail::parallelFor(energies.size(),
            [&](std::size_t i) { specter[i] = BetheHeitlerCrossSection::betheHeitlerSpectra(m_voltage, energies[i], anodeAngle); });
        addCharacteristicLines(energies, specter);
        applyFiltration(energies, specter);
        if (normalize) {
            const auto sum = std::reduce(std::execution::par_unseq, specter.begin(), specter.end());
            for (auto& n : specter)
                n = n / sum;
        }
        return specter;
    }
    std::vector<T> getSpecter(const std::vector<T>& energies, bool normalize = true) const { return getSpecter(energies, m_anodeAngle, normalize); }

    T mmAlHalfValueLayer()
    {
        if (!m_hasCachedHVL) {
            m_cachedHVL = computeHalfValueLayer();
            m_hasCachedHVL = true;
        }
        return m_cachedHVL;
    }
    [[nodiscard]] T mmAlHalfValueLayer() const { return m_hasCachedHVL ? m_cachedHVL : computeHalfValueLayer(); }

protected:
    void setElementFiltration(int Z, const char* symbol, T mm)
    {
        for (auto& [material, thickness] : m_filtrationMaterials)
            if (material.name().compare(symbol) == 0) {
                thickness = std::abs(mm);
                m_hasCachedHVL = false;
                return;
            }
        addFiltrationMaterial(Material(Z), std::abs(mm));
    }
    T elementFiltration(const char* symbol) const
    {
        for (const auto& [material, thickness] : m_filtrationMaterials)
            if (material.name().compare(symbol) == 0)
                return thickness;
        return T { 0 };
    }

    // a line is added to the first bin at or above its energy when that bin is within 2 keV
    void addCharacteristicLines(const std::vector<T>& energy, std::vector<T>& specter) const
    {
        const auto lines = BetheHeitlerCrossSection::characteristicTungstenKedge(m_voltage, m_anodeAngle);
        for (const auto& [e, n] : lines) {
            const auto bin = std::lower_bound(energy.begin(), energy.end(), e);
            if (bin != energy.end() && std::abs(e - *bin) <= T { 2.0 })
                specter[std::distance(energy.begin(), bin)] += n;
        }
    }
    void applyFiltration(const std::vector<T>& energies, std::vector<T>& specter) const
    {
        for (const auto& [material, mm] : m_filtrationMaterials) {
            const T cm = mm * T { 0.1 };
            for (std::size_t i = 0; i < specter.size(); ++i) {
                const T n = specter[i];
                specter[i] = n * std::exp(-material.getTotalAttenuation(energies[i]) * material.standardDensity() * cm);
            }
        }
    }
    // fixed-point iteration x <- x + (transmission(x) - 1/2) on the aluminium thickness in cm
    T computeHalfValueLayer() const
    {
        const auto energy = getEnergy();
        const auto specter = getSpecter(energy);
        const Material al(13);
        std::vector<T> att(energy.size());
        for (std::size_t i = 0; i < energy.size(); ++i)
            att[i] = static_cast<T>(al.getTotalAttenuation(energy[i])) * static_cast<T>(al.standardDensity());
        T x { 0.5 };
        T step { 1 };
        do {
            const T g = std::transform_reduce(std::execution::par_unseq, specter.cbegin(), specter.cend(), att.cbegin(), T { 0 }, std::plus<T>(),
                [=](auto s, auto a) -> T { return s * std::exp(-a * x); });
            step = g - T { 0.5 };
            x = x + step;
        } while (std::abs(step) > T { 0.01 });
        return x * T { 10.0 };
    }

private:
    T m_voltage, m_energyResolution, m_anodeAngle;
    T m_cachedHVL = 0;
    bool m_hasCachedHVL = false;
    std::vector<std::pair<Material, T>> m_filtrationMaterials;
};
}
