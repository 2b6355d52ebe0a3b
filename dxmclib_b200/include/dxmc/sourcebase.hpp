// sourcebase.hpp — the abstract Source and the two spectrum-free-of-tube sources.
//
// Part of the reference's source hierarchy (include/dxmc/source.hpp): Source :46-207, PencilSource :209-266,
// IsotropicSource / IsotropicCTSource :268-372. Included by dxmc/source.hpp, which is the header user code names.
#pragma once
#include "dxmc/beamfilters.hpp"
#include "dxmc/constants.hpp"
#include "dxmc/dxmcrandom.hpp"
#include "dxmc/exposure.hpp"
#include "dxmc/floating.hpp"
#include "dxmc/lowenergycorrectionmodel.hpp"
#include "dxmc/progressbar.hpp"
#include "dxmc/transport.hpp"
#include "dxmc/tube.hpp"
#include "dxmc/vectormath.hpp"
#include "dxmc/world.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <execution>
#include <memory>
#include <numeric>
#include <vector>

namespace dxmc {

template <Floating T = double>
class Source {
public:
    // not used by the library; convenience for down-casting
    enum class Type { None, CTSpiral, CTAxial, DX, CTDual, Pencil, Isotropic, IsotropicCT, CTTopogram, CBCT, Other };

    Source() = default;
    virtual ~Source() = default;

    virtual Exposure<T> getExposure(std::uint64_t i) const = 0;
    virtual T maxPhotonEnergyProduced() const { return Tube<T>::maxVoltage(); }
    virtual std::uint64_t totalExposures() const = 0;
    // factor turning energy imparted per history into absolute dose
    virtual T getCalibrationValue(LOWENERGYCORRECTION model = LOWENERGYCORRECTION::NONE, ProgressBar<T>* progress = nullptr) const = 0;
    virtual bool isValid() const = 0;
    virtual bool validate() = 0;
    virtual void updateFromWorld(const World<T>&) { }

    void setPosition(const std::array<T, 3>& position) { m_position = position; }
    void setPosition(T x, T y, T z) { m_position = { x, y, z }; }
    std::array<T, 3>& position() { return m_position; }
    const std::array<T, 3>& position() const { return m_position; }

    // x and y unit vectors of the source plane; beam direction = x cross y
    void setDirectionCosines(const std::array<T, 6>& cosines)
    {
        m_directionCosines = cosines;
        vectormath::normalize(&m_directionCosines[0]);
        vectormath::normalize(&m_directionCosines[3]);
    }
    const std::array<T, 6>& directionCosines() const { return m_directionCosines; }
    std::array<T, 6>& directionCosines() { return m_directionCosines; }

    void setHistoriesPerExposure(std::uint64_t histories) { m_historiesPerExposure = histories; }
    std::uint64_t historiesPerExposure() const { return m_historiesPerExposure; }
    Type type() const { return m_type; }

protected:
    std::array<T, 3> m_position = { 0, 0, 0 };
    std::array<T, 6> m_directionCosines = { 1, 0, 0, 0, 1, 0 };
    std::uint64_t m_historiesPerExposure = 1E6;
    Type m_type = Type::None;
};

template <Floating T = double>
class PencilSource final : public Source<T> {
public:
    PencilSource() { this->m_type = Source<T>::Type::Pencil; }

    Exposure<T> getExposure(std::uint64_t) const override
    {
        constexpr std::array<T, 2> noOpening { 0, 0 };
        Exposure<T> exposure(this->m_position, this->m_directionCosines, noOpening, this->m_historiesPerExposure);
        exposure.setMonoenergeticPhotonEnergy(m_photonEnergy);
        return exposure;
    }
    void setPhotonEnergy(T energy) { m_photonEnergy = std::clamp(energy, T { 1 }, ELECTRON_REST_MASS<T>()); }
    T photonEnergy() const { return m_photonEnergy; }
    T maxPhotonEnergyProduced() const override { return m_photonEnergy; }
    std::uint64_t totalExposures() const override { return m_totalExposures; }
    void setTotalExposures(std::uint64_t exposures)
    {
        if (exposures > 0)
            m_totalExposures = exposures;
    }
    void setAirDose(T Gycm2)
    {
        if (Gycm2 > 0.0)
            m_airDose = Gycm2;
    }
    T airDose() const { return m_airDose; }

    // air kerma of the emitted photons against the requested air dose
    T getCalibrationValue(LOWENERGYCORRECTION, ProgressBar<T>* = nullptr) const override
    {
        const Material air("Air, Dry (near sea level)");
        const T nHistories = totalExposures() * this->historiesPerExposure();
        const T mea = static_cast<T>(air.getMassEnergyAbsorbtion(m_photonEnergy));
        const T calcOutput = nHistories * m_photonEnergy * mea * KEV_TO_MJ<T>();
        return m_airDose / calcOutput;
    }
    bool isValid() const override { return true; }
    bool validate() override { return true; }

protected:
    T m_photonEnergy = 100;
    T m_airDose = 1;
    std::uint64_t m_totalExposures = 10;
};

template <Floating T = double>
class IsotropicSource : public Source<T> {
public:
    IsotropicSource()
        : m_specterDistribution(std::vector<T> { 1.0 }, std::vector<T> { 60.0 })
    {
        this->m_type = Source<T>::Type::Isotropic;
        m_maxPhotonEnergy = 60.0;
    }

    Exposure<T> getExposure(std::uint64_t) const override
    {
        return Exposure<T>(this->m_position, this->m_directionCosines, m_collimationAngles, this->m_historiesPerExposure, T { 1 }, &m_specterDistribution);
    }
    T maxPhotonEnergyProduced() const override { return m_maxPhotonEnergy; }
    void setTotalExposures(std::uint64_t nExposures) { m_totalExposures = nExposures; }
    std::uint64_t totalExposures() const override { return m_totalExposures; }
    T getCalibrationValue(LOWENERGYCORRECTION, ProgressBar<T>* = nullptr) const override { return T { 1 }; }
    bool isValid() const override { return true; }
    bool validate() override { return true; }

    void setSpecter(const std::vector<T>& weights, const std::vector<T>& energies)
    {
        m_maxPhotonEnergy = *std::max_element(energies.cbegin(), energies.cend());
        m_specterDistribution = SpecterDistribution<T>(weights, energies);
    }
    void setCollimationAngles(T x0, T x1, T y0, T y1)
    {
        constexpr T halfPi = PI_VAL<T>() / 2;
        m_collimationAngles = { std::clamp(x0, -halfPi, halfPi), std::clamp(x1, -halfPi, halfPi), std::clamp(y0, -PI_VAL<T>(), PI_VAL<T>()),
            std::clamp(y1, -PI_VAL<T>(), PI_VAL<T>()) };
    }
    void setCollimationAngles(T xRad, T yRad) { setCollimationAngles(-xRad / 2, xRad / 2, -yRad / 2, yRad / 2); }
    const std::array<T, 4>& collimationAngles() const { return m_collimationAngles; }

protected:
    std::uint64_t m_totalExposures = 1;
    std::array<T, 4> m_collimationAngles = { 0, 0, 0, 0 };
    SpecterDistribution<T> m_specterDistribution;
    T m_maxPhotonEnergy = 1.0;
};

// isotropic source stepped around the z axis, one exposure per angle
template <Floating T = double>
class IsotropicCTSource final : public IsotropicSource<T> {
public:
    IsotropicCTSource() { this->m_type = Source<T>::Type::IsotropicCT; }

    Exposure<T> getExposure(std::uint64_t exposureNumber) const override
    {
        const std::array<T, 3> axis = { 0, 0, 1 };
        const auto angle = (exposureNumber * 2 * PI_VAL<T>()) / this->m_totalExposures;
        auto cosines = this->m_directionCosines;
        auto pos = this->m_position;
        vectormath::rotate(&pos[0], axis.data(), angle);
        vectormath::rotate(&cosines[0], axis.data(), angle);
        vectormath::rotate(&cosines[3], axis.data(), angle);
        return Exposure<T>(pos, cosines, this->m_collimationAngles, this->m_historiesPerExposure, T { 1 }, &(this->m_specterDistribution));
    }
};

// tube-based source calibrated by dose-area product
}
