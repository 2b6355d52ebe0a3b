// sourcebase.hpp — Source<T> and the tube-less sources (pencil, isotropic, isotropic CT).
//
// Every source of this library is a parameter block (dxmc/sourcemodel.hpp, model::SourceParams) plus the table objects it
// refers to (spectrum, heel filter, fan filter, tube-current profile). The classes below are setters and getters over that
// block with the reference's names, defaults and clamps (reference include/dxmc/source.hpp:46-372); getExposure(i) is
// model::evaluate on the block — the same function exposureKernel runs on the device for all i at once (Transport hands
// the block to dxmcb200_generate_exposures, no per-exposure host work). A user-defined Source that overrides getExposure
// still works: Transport then calls it for every exposure (describe() returns false).
#pragma once
#include "dxmc/beamfilters.hpp"
#include "dxmc/constants.hpp"
#include "dxmc/dxmcrandom.hpp"
#include "dxmc/exposure.hpp"
#include "dxmc/floating.hpp"
#include "dxmc/lowenergycorrectionmodel.hpp"
#include "dxmc/progressbar.hpp"
#include "dxmc/sourcemodel.hpp"
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

    // the table objects one tube's exposures point to (any may be null)
    struct BeamTables {
        const SpecterDistribution<T>* specter = nullptr;
        const HeelFilter<T>* heel = nullptr;
        const BeamFilter<T>* fan = nullptr;
    };

    Source() { m_p.histories = 1000000; }
    virtual ~Source() = default;

    virtual Exposure<T> getExposure(std::uint64_t i) const
    {
        const auto p = parameters();
        model::ExposureValues<T> v;
        model::evaluate(p, tubeCurrentProfile(), i, v);
        const std::uint32_t tube = p.tubes == 2 ? static_cast<std::uint32_t>(i & 1u) : 0u;
        const auto tables = beamTables(tube);
        return Exposure<T>::fromModel(v, tables.specter, tables.heel, tables.fan);
    }
    virtual T maxPhotonEnergyProduced() const { return Tube<T>::maxVoltage(); }
    virtual std::uint64_t totalExposures() const = 0;
    // factor turning energy imparted per history into absolute dose
    virtual T getCalibrationValue(LOWENERGYCORRECTION model = LOWENERGYCORRECTION::NONE, ProgressBar<T>* progress = nullptr) const = 0;
    virtual bool isValid() const = 0;
    virtual bool validate() = 0;
    virtual void updateFromWorld(const World<T>&) { }

    // ---- device-side exposure generation: the block with the exposure count filled in, the tube-current profile it indexes
    // (aecSize entries, or null) and the tables of each tube. False: this source can only be asked exposure by exposure.
    virtual bool describe(model::SourceParams<T>& block, const T*& tubeCurrent) const
    {
        block = parameters();
        tubeCurrent = tubeCurrentProfile();
        return true;
    }
    virtual BeamTables beamTables(std::uint32_t /*tube*/) const { return {}; }

    void setPosition(const std::array<T, 3>& position) { std::copy(position.begin(), position.end(), m_p.position); }
    void setPosition(T x, T y, T z) { setPosition({ x, y, z }); }
    std::array<T, 3>& position() { return reinterpret_cast<std::array<T, 3>&>(m_p.position); }
    const std::array<T, 3>& position() const { return reinterpret_cast<const std::array<T, 3>&>(m_p.position); }

    // x and y unit vectors of the source plane; beam direction = x cross y
    void setDirectionCosines(const std::array<T, 6>& cosines)
    {
        std::copy(cosines.begin(), cosines.end(), m_p.cosines);
        vectormath::normalize(&m_p.cosines[0]);
        vectormath::normalize(&m_p.cosines[3]);
    }
    const std::array<T, 6>& directionCosines() const { return reinterpret_cast<const std::array<T, 6>&>(m_p.cosines); }
    std::array<T, 6>& directionCosines() { return reinterpret_cast<std::array<T, 6>&>(m_p.cosines); }

    void setHistoriesPerExposure(std::uint64_t histories) { m_p.histories = histories; }
    std::uint64_t historiesPerExposure() const { return m_p.histories; }
    Type type() const { return m_type; }

protected:
    model::SourceParams<T> parameters() const
    {
        auto p = m_p;
        p.exposures = totalExposures();
        return p;
    }
    virtual const T* tubeCurrentProfile() const { return nullptr; }
    // symmetric full opening angles -> the four half-plane angles of an exposure
    void setOpening(T fullX, T fullY)
    {
        m_p.collimation[0] = -fullX / 2;
        m_p.collimation[1] = fullX / 2;
        m_p.collimation[2] = -fullY / 2;
        m_p.collimation[3] = fullY / 2;
    }

    model::SourceParams<T> m_p;
    Type m_type = Type::None;
};

template <Floating T = double>
class PencilSource final : public Source<T> {
public:
    PencilSource()
    {
        this->m_type = Source<T>::Type::Pencil;
        this->setOpening(0, 0);
        this->m_p.exposures = 10;
        setPhotonEnergy(100);
    }

    void setPhotonEnergy(T energy)
    {
        m_photonEnergy = std::clamp(energy, T { 1 }, ELECTRON_REST_MASS<T>());
        this->m_p.monoEnergy = std::clamp(m_photonEnergy, T { 0 }, T { 500 }); // what an Exposure keeps of it
    }
    T photonEnergy() const { return m_photonEnergy; }
    T maxPhotonEnergyProduced() const override { return m_photonEnergy; }
    std::uint64_t totalExposures() const override { return this->m_p.exposures; }
    void setTotalExposures(std::uint64_t exposures)
    {
        if (exposures > 0)
            this->m_p.exposures = exposures;
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

private:
    T m_photonEnergy = 100;
    T m_airDose = 1;
};

template <Floating T = double>
class IsotropicSource : public Source<T> {
public:
    IsotropicSource()
        : m_specterDistribution(std::vector<T> { 1.0 }, std::vector<T> { 60.0 })
    {
        this->m_type = Source<T>::Type::Isotropic;
        this->m_p.spectrum[0] = 0;
        m_maxPhotonEnergy = 60.0;
    }

    T maxPhotonEnergyProduced() const override { return m_maxPhotonEnergy; }
    void setTotalExposures(std::uint64_t nExposures) { this->m_p.exposures = nExposures; }
    std::uint64_t totalExposures() const override { return this->m_p.exposures; }
    T getCalibrationValue(LOWENERGYCORRECTION, ProgressBar<T>* = nullptr) const override { return T { 1 }; }
    bool isValid() const override { return true; }
    bool validate() override { return true; }
    typename Source<T>::BeamTables beamTables(std::uint32_t) const override { return { &m_specterDistribution, nullptr, nullptr }; }

    void setSpecter(const std::vector<T>& weights, const std::vector<T>& energies)
    {
        m_maxPhotonEnergy = *std::max_element(energies.cbegin(), energies.cend());
        m_specterDistribution = SpecterDistribution<T>(weights, energies);
    }
    void setCollimationAngles(T x0, T x1, T y0, T y1)
    {
        constexpr T halfPi = PI_VAL<T>() / 2;
        const T limit[4] = { halfPi, halfPi, PI_VAL<T>(), PI_VAL<T>() };
        const T requested[4] = { x0, x1, y0, y1 };
        for (int k = 0; k < 4; ++k)
            this->m_p.collimation[k] = std::clamp(requested[k], -limit[k], limit[k]);
    }
    void setCollimationAngles(T xRad, T yRad) { setCollimationAngles(-xRad / 2, xRad / 2, -yRad / 2, yRad / 2); }
    const std::array<T, 4>& collimationAngles() const { return reinterpret_cast<const std::array<T, 4>&>(this->m_p.collimation); }

protected:
    SpecterDistribution<T> m_specterDistribution;
    T m_maxPhotonEnergy = 1.0;
};

// isotropic source stepped around the z axis, one exposure per angle, a full turn over all exposures
template <Floating T = double>
class IsotropicCTSource final : public IsotropicSource<T> {
public:
    IsotropicCTSource()
    {
        this->m_type = Source<T>::Type::IsotropicCT;
        this->m_p.motion = model::ORBIT;
        this->m_p.orbitFullTurn = 1;
    }
};

}
