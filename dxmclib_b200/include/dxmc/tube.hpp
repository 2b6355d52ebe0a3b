// tube.hpp — tungsten-anode x-ray tube: voltage, anode angle, added filtration -> photon spectrum.
//
// Public surface of the reference's Tube<T> (include/dxmc/tube.hpp:35-299). The spectrum feeds the
// alias table (SpecterDistribution) and the heel-effect table (HeelFilter) that Transport uploads
// to the GPU; generation itself is host-side, runs once per source, and spreads its energy bins over
// all host cores.
#pragma once
#include "dxmc/betheHeitlerCrossSection.hpp"
#include "dxmc/constants.hpp"
#include "dxmc/hostparallel.hpp"
#include "dxmc/material.hpp"
#include "dxmc/types.hpp"
#include "dxmcb200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <execution>
#include <numeric>
#include <optional>
#include <utility>
#include <vector>

namespace dxmc {

template <Floating T = double>
class Tube {
    using Filtration = std::vector<std::pair<Material, T>>; // material, thickness [mm]

    // the half-value layer is expensive (a full spectrum); remembered until a setter changes the tube, never copied
    struct HvlMemo {
        std::optional<T> mmAl;
        HvlMemo() = default;
        HvlMemo(const HvlMemo&) { }
        HvlMemo& operator=(const HvlMemo&)
        {
            mmAl.reset();
            return *this;
        }
    };

public:
    static constexpr T minVoltage() { return T { 50 }; }
    static constexpr T maxVoltage() { return T { 150 }; }

    Tube(T tubeVoltage = 120.0, T anodeAngleDeg = 12.0, T energyResolution = 1.0)
        : m_kV(tubeVoltage)
        , m_binWidth(energyResolution)
    {
        setAnodeAngleDeg(anodeAngleDeg);
    }

    // ---- spectrum
    // bin energies: resolution, 2*resolution, ... <= voltage
    std::vector<T> getEnergy() const
    {
        std::vector<T> bins;
        bins.reserve(static_cast<std::size_t>(std::ceil(m_kV / m_binWidth)));
        for (T hv = m_binWidth; hv <= m_kV; hv = hv + m_binWidth)
            bins.push_back(hv);
        return bins;
    }
    // bremsstrahlung + tungsten K lines, filtered by the added materials, at a given take-off angle
    std::vector<T> getSpecter(const std::vector<T>& energies, const T anodeAngle, bool normalize = true) const
    {
        std::vector<T> fluence(energies.size());
        std::array<std::pair<T, T>, 5> lines {};
        if (!bremsstrahlungOnDevice(energies, anodeAngle, fluence, lines)) {
            // one Bethe-Heitler depth integral per energy bin, independent of each other (reference tube.hpp:191-208)
            detail::parallelFor(energies.size(),
                [&](std::size_t i) { fluence[i] = BetheHeitlerCrossSection::betheHeitlerSpectra(m_kV, energies[i], anodeAngle); });
            lines = BetheHeitlerCrossSection::characteristicTungstenKedge(m_kV, m_takeOff);
        }
        // a K line goes into the first bin at or above its energy when that bin is within 2 keV
        for (const auto& [lineEnergy, lineYield] : lines) {
            const auto bin = std::lower_bound(energies.begin(), energies.end(), lineEnergy);
            if (bin != energies.end() && std::abs(lineEnergy - *bin) <= T { 2.0 })
                fluence[std::distance(energies.begin(), bin)] += lineYield;
        }
        for (const auto& [material, mm] : m_filters) {
            const T cm = mm * T { 0.1 };
            for (std::size_t i = 0; i < fluence.size(); ++i) {
                const T n = fluence[i];
                fluence[i] = n * std::exp(-material.getTotalAttenuation(energies[i]) * material.standardDensity() * cm);
            }
        }
        if (normalize) {
            const auto sum = std::reduce(std::execution::par_unseq, fluence.begin(), fluence.end());
            for (auto& n : fluence)
                n = n / sum;
        }
        return fluence;
    }
    std::vector<T> getSpecter(const std::vector<T>& energies, bool normalize = true) const { return getSpecter(energies, m_takeOff, normalize); }
    std::vector<std::pair<T, T>> getSpecter(bool normalize = true) const
    {
        const auto energies = getEnergy();
        const auto fluence = getSpecter(energies, normalize);
        std::vector<std::pair<T, T>> pairs(fluence.size());
        for (std::size_t i = 0; i < fluence.size(); ++i)
            pairs[i] = { energies[i], fluence[i] };
        return pairs;
    }

    [[nodiscard]] T mmAlHalfValueLayer() const { return m_hvl.mmAl ? *m_hvl.mmAl : computeHalfValueLayer(); }
    T mmAlHalfValueLayer()
    {
        if (!m_hvl.mmAl)
            m_hvl.mmAl = computeHalfValueLayer();
        return *m_hvl.mmAl;
    }

    // ---- added filtration
    void addFiltrationMaterial(const Material& filtrationMaterial, T mm)
    {
        m_filters.emplace_back(filtrationMaterial, std::abs(mm));
        m_hvl.mmAl.reset();
    }
    Filtration& filtrationMaterials() { return m_filters; }
    const Filtration& filtrationMaterials() const { return m_filters; }
    void clearFiltrationMaterials() { m_filters.clear(); }
    T AlFiltration() const { return elementFiltration("Al"); }
    T CuFiltration() const { return elementFiltration("Cu"); }
    T SnFiltration() const { return elementFiltration("Sn"); }
    void setAlFiltration(T mm) { setElementFiltration(13, "Al", mm); }
    void setCuFiltration(T mm) { setElementFiltration(29, "Cu", mm); }
    void setSnFiltration(T mm) { setElementFiltration(50, "Sn", mm); }

    // ---- tube settings
    T voltage() const { return m_kV; }
    void setVoltage(T voltage)
    {
        m_kV = std::clamp(voltage, minVoltage(), maxVoltage());
        m_hvl.mmAl.reset();
    }
    T anodeAngle() const { return m_takeOff; }
    T anodeAngleDeg() const { return m_takeOff * RAD_TO_DEG<T>(); }
    void setAnodeAngle(T angle)
    {
        m_takeOff = std::min(std::abs(angle), PI_VAL<T>() * T { 0.5 });
        m_hvl.mmAl.reset();
    }
    void setAnodeAngleDeg(T angle) { setAnodeAngle(angle * DEG_TO_RAD<T>()); }
    T energyResolution() const { return m_binWidth; }
    void setEnergyResolution(T energyResolution) { m_binWidth = energyResolution; }

protected:
    // The depth integrals of all bins (at anodeAngle) and of the four K lines (at the tube's own take-off angle) in one launch of
    // dxmcb200_tube_bremsstrahlung, when DXMCB200_DEVICE_SPECTRUM=1 asks for it (T = float only). false: not requested, or the
    // device path failed (said on stderr); the caller then evaluates them on the host, which is the default and the
    // reference's way (its tables are bit-identical to the reference's, the device's agree to 2e-5).
    bool bremsstrahlungOnDevice(const std::vector<T>& energies, const T anodeAngle, std::vector<T>& fluence, std::array<std::pair<T, T>, 5>& lines) const
    {
        if constexpr (!std::is_same_v<T, float>) {
            return false;
        } else {
            const char* env = std::getenv("DXMCB200_DEVICE_SPECTRUM");
            if (!env || env[0] != '1' || energies.empty())
                return false;
            const auto lineEnergy = BetheHeitlerCrossSection::tungstenKLineEnergies<float>();
            std::vector<float> hv(energies.begin(), energies.end());
            hv.insert(hv.end(), lineEnergy.begin(), lineEnergy.end());
            std::vector<float> att(hv.size());
            for (std::size_t i = 0; i < hv.size(); ++i)
                att[i] = static_cast<float>(Material::getTotalAttenuation(BetheHeitlerCrossSection::TUNGSTEN_ATOMIC_NUMBER, hv[i]));
            const float angles[2] = { anodeAngle, m_takeOff };
            std::vector<float> out(2 * hv.size());
            const int st = dxmcb200_tube_bremsstrahlung(m_kV, static_cast<std::uint32_t>(hv.size()), hv.data(), att.data(), 2, angles, out.data());
            if (st != DXMCB200_OK) {
                std::fprintf(stderr, "[dxmcb200] DXMCB200_DEVICE_SPECTRUM=1: device spectrum failed (status %d), evaluating on the host\n", st);
                return false;
            }
            const std::size_t n = energies.size();
            std::copy(out.begin(), out.begin() + static_cast<std::ptrdiff_t>(n), fluence.begin());
            std::array<float, 4> atLines {};
            for (std::size_t i = 0; i < 4; ++i)
                atLines[i] = out[hv.size() + n + i];
            lines = BetheHeitlerCrossSection::characteristicTungstenKedge(atLines);
            return true;
        }
    }
    T elementFiltration(const char* symbol) const
    {
        const auto it = std::find_if(m_filters.begin(), m_filters.end(), [&](const auto& f) { return f.first.name().compare(symbol) == 0; });
        return it != m_filters.end() ? it->second : T { 0 };
    }
    void setElementFiltration(int Z, const char* symbol, T mm)
    {
        const auto it = std::find_if(m_filters.begin(), m_filters.end(), [&](const auto& f) { return f.first.name().compare(symbol) == 0; });
        if (it == m_filters.end()) {
            addFiltrationMaterial(Material(Z), std::abs(mm));
            return;
        }
        it->second = std::abs(mm);
        m_hvl.mmAl.reset();
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
    T m_kV, m_binWidth, m_takeOff;
    Filtration m_filters;
    HvlMemo m_hvl;
};
}
