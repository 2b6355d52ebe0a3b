// source.hpp — x-ray sources: geometry + spectrum + dose calibration, producing Exposures.
//
// Public surface of the reference's source hierarchy (include/dxmc/source.hpp:46-1702):
//   Source                       abstract base                                       :46-207
//   PencilSource                 mono-energetic pencil beam                          :209-266
//   IsotropicSource / IsotropicCTSource   user spectrum, rectangular collimation     :268-372
//   DAPSource -> DXSource, CBCTSource     tube-based radiography / cone-beam         :374-787
//   CTBaseSource -> CTSource -> CTAxialSource, CTSpiralSource                        :789-1363
//                 CTDualSource -> CTAxialDualSource, CTSpiralDualSource              :1075-1608
//   CTTopogramSource                                                                 :1610-1700
// getExposure(i) is O(1) host code; Transport evaluates it for every exposure up front and ships
// the resulting table to the GPU. CT dose calibration (ctCalibration) runs a second Transport on
// a CTDIPhantom exactly like the reference, i.e. a second pass through the same CUDA path.
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
template <Floating T = double>
class DAPSource : public Source<T> {
public:
    DAPSource()
    {
        m_fieldSize = { 100.0, 100.0 };
        applyFieldSize(m_fieldSize);
        m_tube.setAlFiltration(2.0);
        this->setDirectionCosines(zeroDirectionCosines());
    }

    Tube<T>& tube()
    {
        m_specterValid = false;
        return m_tube;
    }
    const Tube<T>& tube() const { return m_tube; }
    T maxPhotonEnergyProduced() const override { return m_tube.voltage(); }

    void setCollimationAngles(const std::array<T, 2>& angles) { applyCollimation({ std::abs(angles[0]), std::abs(angles[1]) }); }
    const std::array<T, 2>& collimationAngles() const { return m_collimationAngles; }
    void setCollimationAnglesDeg(const std::array<T, 2>& angles)
    {
        applyCollimation({ std::abs(angles[0]) * DEG_TO_RAD<T>(), std::abs(angles[1]) * DEG_TO_RAD<T>() });
    }
    const std::array<T, 2> collimationAnglesDeg() const { return { m_collimationAngles[0] * RAD_TO_DEG<T>(), m_collimationAngles[1] * RAD_TO_DEG<T>() }; }

    void setFieldSize(const std::array<T, 2>& mm) { applyFieldSize({ std::abs(mm[0]), std::abs(mm[1]) }); }
    const std::array<T, 2>& fieldSize() const { return m_fieldSize; }
    void setSourceDetectorDistance(T mm)
    {
        m_sdd = std::abs(mm);
        applyFieldSize(m_fieldSize);
    }
    T sourceDetectorDistance() const { return m_sdd; }

    // primary: rotation about z; secondary: cranio-caudal tilt; then the tube rotation about the beam
    void setSourceAngles(T primaryAngle, T secondaryAngle)
    {
        constexpr T eps = 1E-6;
        constexpr T halfPi = PI_VAL<T>() / 2;
        if (secondaryAngle > halfPi - eps)
            secondaryAngle = halfPi - eps;
        if (secondaryAngle < -halfPi + eps)
            secondaryAngle = -halfPi + eps;
        while (primaryAngle > PI_VAL<T>())
            primaryAngle -= PI_VAL<T>();
        while (primaryAngle < -PI_VAL<T>())
            primaryAngle += PI_VAL<T>();

        auto cos = zeroDirectionCosines();
        const std::array<T, 3> z = { .0, .0, 1.0 };
        const std::array<T, 3> x = { 1.0, .0, .0 };
        vectormath::rotate(cos.data(), z.data(), primaryAngle);
        vectormath::rotate(&cos[3], z.data(), primaryAngle);
        vectormath::rotate(cos.data(), x.data(), -secondaryAngle);
        vectormath::rotate(&cos[3], x.data(), -secondaryAngle);
        std::array<T, 3> beam;
        vectormath::cross(cos.data(), beam.data());
        vectormath::rotate(cos.data(), beam.data(), m_tubeRotationAngle);
        vectormath::rotate(&cos[3], beam.data(), m_tubeRotationAngle);
        this->setDirectionCosines(cos);
    }
    void setSourceAngles(const std::array<T, 2>& angles) { setSourceAngles(angles[0], angles[1]); }
    std::array<T, 2> sourceAngles() const
    {
        constexpr T eps = 1E-6;
        auto cos = this->directionCosines();
        std::array<T, 3> beam;
        vectormath::cross(cos.data(), beam.data());
        vectormath::rotate(cos.data(), beam.data(), -m_tubeRotationAngle);
        vectormath::rotate(&cos[3], beam.data(), -m_tubeRotationAngle);
        vectormath::cross(cos.data(), beam.data());
        const T xy = std::sqrt(beam[0] * beam[0] + beam[1] * beam[1]);
        if (std::abs(xy) < eps)
            return { 0, beam[2] > 0 ? -PI_VAL<T>() / 2 : PI_VAL<T>() / 2 };
        const T primary = std::asin(-beam[0]);
        const T zy = std::sqrt(beam[2] * beam[2] + beam[1] * beam[1]);
        if (std::abs(zy) < eps)
            return { primary, 0 };
        return { primary, -std::asin(beam[2] / zy) };
    }
    void setSourceAnglesDeg(T primaryAngle, T secondaryAngle) { setSourceAngles(primaryAngle * DEG_TO_RAD<T>(), secondaryAngle * DEG_TO_RAD<T>()); }
    void setSourceAnglesDeg(const std::array<T, 2>& angles) { setSourceAnglesDeg(angles[0], angles[1]); }
    std::array<T, 2> sourceAnglesDeg() const
    {
        auto a = sourceAngles();
        a[0] *= RAD_TO_DEG<T>();
        a[1] *= RAD_TO_DEG<T>();
        return a;
    }

    void setTubeRotation(T angle)
    {
        const T diff = angle - m_tubeRotationAngle;
        auto cos = this->directionCosines();
        std::array<T, 3> beam;
        vectormath::cross(cos.data(), beam.data());
        vectormath::rotate(cos.data(), beam.data(), diff);
        vectormath::rotate(&cos[3], beam.data(), diff);
        this->setDirectionCosines(cos);
        m_tubeRotationAngle = angle;
    }
    T tubeRotation() const { return m_tubeRotationAngle; }
    void setTubeRotationDeg(T angle) { setTubeRotation(angle * DEG_TO_RAD<T>()); }
    T tubeRotationDeg() const { return tubeRotation() * RAD_TO_DEG<T>(); }

    void setDap(T Gycm2)
    {
        if (Gycm2 > 0.0)
            m_dap = Gycm2;
    }
    T dap() const { return m_dap; }

    // air kerma per emitted photon from the normalised spectrum against the requested DAP
    T getCalibrationValue(LOWENERGYCORRECTION, ProgressBar<T>* = nullptr) const override
    {
        const auto specter = tube().getSpecter(true);
        const Material air("Air, Dry (near sea level)");
        T calcOutput = 0.0;
        for (const auto& [keV, weight] : specter) {
            const T massAbsorb = air.getMassEnergyAbsorbtion(keV);
            calcOutput += keV * weight * massAbsorb;
        }
        calcOutput *= this->totalExposures() * this->historiesPerExposure();
        return m_dap / calcOutput;
    }

    bool isValid() const override { return m_specterValid; }
    bool validate() override
    {
        refreshSpectrum();
        return m_specterValid;
    }
    void setModelHeelEffect(bool on) { m_modelHeelEffect = on; }
    bool modelHeelEffect() const { return m_modelHeelEffect; }

protected:
    void applyFieldSize(const std::array<T, 2>& fieldSize)
    {
        for (std::size_t i = 0; i < 2; ++i) {
            m_fieldSize[i] = fieldSize[i];
            m_collimationAngles[i] = std::atan(m_fieldSize[i] * T { 0.5 } / m_sdd) * T { 2 };
        }
        m_specterValid = false;
    }
    void applyCollimation(const std::array<T, 2>& angles)
    {
        for (std::size_t i = 0; i < 2; ++i) {
            m_collimationAngles[i] = angles[i];
            m_fieldSize[i] = std::tan(m_collimationAngles[i] / 2) * m_sdd * 2;
        }
        m_specterValid = false;
    }
    void refreshSpectrum()
    {
        if (m_specterValid)
            return;
        const auto energies = m_tube.getEnergy();
        const auto weights = m_tube.getSpecter(energies);
        m_specterDistribution = std::make_shared<SpecterDistribution<T>>(weights, energies);
        m_heelFilter = m_modelHeelEffect ? std::make_shared<HeelFilter<T>>(m_tube, m_collimationAngles[1]) : nullptr;
        m_specterValid = true;
    }
    // beam along +y, anode-cathode along z; setSourceAngles depends on this choice
    std::array<T, 6> zeroDirectionCosines() const { return { -1.0, .0, .0, .0, .0, 1.0 }; }

    T m_sdd = 1000.0;
    T m_dap = 1.0; // Gy cm2
    std::array<T, 2> m_fieldSize;
    std::array<T, 2> m_collimationAngles;
    Tube<T> m_tube;
    T m_tubeRotationAngle = 0.0;
    std::shared_ptr<SpecterDistribution<T>> m_specterDistribution;
    std::shared_ptr<HeelFilter<T>> m_heelFilter;
    bool m_modelHeelEffect = true;
    bool m_specterValid = false;
};

template <Floating T = double>
class DXSource final : public DAPSource<T> {
public:
    DXSource() { this->m_type = Source<T>::Type::DX; }

    Exposure<T> getExposure(std::uint64_t) const override
    {
        return Exposure<T>(tubePosition(), this->m_directionCosines, this->m_collimationAngles, this->m_historiesPerExposure, T { 1 },
            this->m_specterDistribution.get(), this->m_heelFilter.get());
    }
    std::uint64_t totalExposures() const override { return m_totalExposures; }
    void setTotalExposures(std::uint64_t exposures) { m_totalExposures = std::max(exposures, std::uint64_t { 1 }); }

    // focal spot: source-detector distance upstream of the reference position
    const std::array<T, 3> tubePosition() const
    {
        std::array<T, 3> beam, pos;
        vectormath::cross(this->m_directionCosines.data(), beam.data());
        for (std::size_t i = 0; i < 3; ++i)
            pos[i] = this->m_position[i] - beam[i] * this->m_sdd;
        return pos;
    }

private:
    std::uint64_t m_totalExposures = 1000;
};

// cone-beam CT: the DX geometry stepped about the y direction cosine through the isocentre
template <Floating T = double>
class CBCTSource final : public DAPSource<T> {
public:
    CBCTSource()
    {
        this->m_type = Source<T>::Type::CBCT;
        this->setSourceDetectorDistance(500.0);
    }

    const std::array<T, 3> rotationAxis() const { return { this->m_directionCosines[3], this->m_directionCosines[4], this->m_directionCosines[5] }; }

    void setSpanAngle(const T spanAngle)
    {
        m_angleSpan = std::max(spanAngle, m_angleStep);
        recount();
    }
    void setSpanAngleDeg(const T spanAngle) { setSpanAngle(spanAngle * DEG_TO_RAD<T>()); }
    const T spanAngle() const { return m_angleSpan; }
    const T spanAngleDeg() const { return m_angleSpan * RAD_TO_DEG<T>(); }
    void setStepAngle(const T stepAngle)
    {
        constexpr T minStep = PI_VAL<T>() / T { 360 };
        m_angleStep = std::max(stepAngle, minStep);
        recount();
    }
    void setStepAngleDeg(const T stepAngle) { setStepAngle(stepAngle * DEG_TO_RAD<T>()); }
    const T stepAngle() const { return m_angleStep; }
    const T stepAngleDeg() const { return m_angleStep * RAD_TO_DEG<T>(); }

    Exposure<T> getExposure(std::uint64_t i) const override
    {
        const auto angle = i * m_angleStep;
        const auto tube = tubePosition();
        const auto& iso = this->position();
        const auto axis = rotationAxis();
        std::array<T, 3> pos;
        for (std::size_t k = 0; k < 3; ++k)
            pos[k] = (tube[k] - iso[k]);
        vectormath::rotate(pos.data(), axis.data(), angle);
        for (std::size_t k = 0; k < 3; ++k)
            pos[k] += iso[k];
        auto cosines = this->m_directionCosines;
        vectormath::rotate(cosines.data(), axis.data(), angle);
        vectormath::rotate(&cosines[3], axis.data(), angle);
        return Exposure<T>(pos, cosines, this->m_collimationAngles, this->m_historiesPerExposure, T { 1 }, this->m_specterDistribution.get(),
            this->m_heelFilter.get());
    }
    std::uint64_t totalExposures() const override { return m_totalExposures; }

    const std::array<T, 3> tubePosition() const
    {
        std::array<T, 3> beam, pos;
        vectormath::cross(this->m_directionCosines.data(), beam.data());
        for (std::size_t i = 0; i < 3; ++i)
            pos[i] = this->m_position[i] - beam[i] * this->m_sdd * T { 0.5 };
        return pos;
    }

private:
    void recount() { m_totalExposures = std::max(static_cast<std::size_t>(m_angleSpan / m_angleStep), std::size_t { 2 }); }

    std::size_t m_totalExposures = 180;
    T m_angleSpan = PI_VAL<T>();
    T m_angleStep = PI_VAL<T>() / T { 180 };
};

template <Floating T>
class CTAxialSource;
template <Floating T>
class CTSpiralSource;
template <Floating T>
class CTAxialDualSource;
template <Floating T>
class CTSpiralDualSource;

// common CT state: gantry geometry, tube, bow-tie, CTDI calibration target
template <Floating T>
class CTBaseSource : public Source<T> {
public:
    CTBaseSource()
    {
        this->m_type = Source<T>::Type::None;
        m_sdd = 1190.0;
        m_collimation = 38.4;
        m_fov = 500.0;
        m_startAngle = 0.0;
        m_scanLenght = 100.0;
        tube().setAlFiltration(7.0);
        this->setDirectionCosines({ -1, 0, 0, 0, 0, 1 });
    }

    Tube<T>& tube()
    {
        m_specterValid = false;
        return m_tube;
    }
    const Tube<T>& tube() const { return m_tube; }
    virtual T maxPhotonEnergyProduced() const override { return m_tube.voltage(); }

    void setBowTieFilter(std::shared_ptr<BowTieFilter<T>> filter) { m_bowTieFilter = filter; }
    std::shared_ptr<BowTieFilter<T>> bowTieFilter() { return m_bowTieFilter; }
    const std::shared_ptr<BowTieFilter<T>> bowTieFilter() const { return m_bowTieFilter; }

    void setSourceDetectorDistance(T sdd)
    {
        m_sdd = std::abs(sdd);
        m_specterValid = false;
    }
    T sourceDetectorDistance() const { return m_sdd; }
    void setCollimation(T collimation)
    {
        m_collimation = std::abs(collimation);
        m_specterValid = false;
    }
    T collimation() const { return m_collimation; }
    void setFieldOfView(T fov) { m_fov = std::abs(fov); }
    T fieldOfView() const { return m_fov; }

    void setGantryTiltAngle(T angle) { m_gantryTiltAngle = std::clamp(angle, -PI_VAL<T>(), PI_VAL<T>()); }
    T gantryTiltAngle() const { return m_gantryTiltAngle; }
    void setGantryTiltAngleDeg(T angle) { setGantryTiltAngle(angle * DEG_TO_RAD<T>()); }
    T gantryTiltAngleDeg() const { return m_gantryTiltAngle * RAD_TO_DEG<T>(); }

    void setStartAngle(T angle) { m_startAngle = angle; }
    T startAngle() const { return m_startAngle; }
    T startAngleDeg() const { return RAD_TO_DEG<T>() * m_startAngle; }
    void setStartAngleDeg(T angle) { m_startAngle = DEG_TO_RAD<T>() * angle; }

    virtual void setScanLenght(T scanLenght) { m_scanLenght = std::abs(scanLenght); }
    T scanLenght() const { return m_scanLenght; }

    void setCtdiVol(T ctdivol)
    {
        if (ctdivol > 0.0)
            m_ctdivol = ctdivol;
    }
    T ctdiVol() const { return m_ctdivol; }
    void setCtdiPhantomDiameter(std::uint64_t mm) { m_ctdiPhantomDiameter = std::max(mm, std::uint64_t { 160 }); }
    std::uint64_t ctdiPhantomDiameter() const { return m_ctdiPhantomDiameter; }

    virtual std::uint64_t totalExposures() const override = 0;

    void setModelHeelEffect(bool on) { m_modelHeelEffect = on; }
    bool modelHeelEffect() const { return m_modelHeelEffect; }
    bool isValid() const override { return m_specterValid; }
    virtual bool validate() override
    {
        updateSpecterDistribution();
        return m_specterValid;
    }

protected:
    struct GantryFrame {
        std::array<T, 3> position;
        std::array<T, 6> cosines;
    };
    // Focal spot position and detector orientation for a gantry angle: start at (0, -sdd/2, 0), tilt the
    // rotation axis (y cosine) about x, rotate about the tilted axis, then advance along z.
    GantryFrame gantryFrame(T sdd, T angle, T zAdvance) const
    {
        GantryFrame f;
        f.position = { 0, -sdd / T { 2 }, 0 };
        f.cosines = this->m_directionCosines;
        T* rotationAxis = &f.cosines[3];
        T* otherAxis = &f.cosines[0];
        const std::array<T, 3> tiltAxis = { 1, 0, 0 };
        auto tiltCorrection = f.position;
        vectormath::rotate(tiltCorrection.data(), tiltAxis.data(), m_gantryTiltAngle);
        vectormath::rotate(rotationAxis, tiltAxis.data(), m_gantryTiltAngle);
        vectormath::rotate(otherAxis, tiltAxis.data(), m_gantryTiltAngle);
        vectormath::rotate(f.position.data(), rotationAxis, angle);
        f.position[2] += zAdvance + tiltCorrection[2];
        vectormath::rotate(otherAxis, rotationAxis, angle);
        for (std::size_t i = 0; i < 3; ++i)
            f.position[i] += this->m_position[i];
        return f;
    }
    // full fan and cone opening angles; the focal spot is sdd/2 from the isocentre
    std::array<T, 2> openingAngles(T fov, T sdd) const { return { std::atan(fov / sdd) * T { 2 }, std::atan(m_collimation / sdd) * T { 2 } }; }

    // CTDIw of one axial rotation on a CTDI phantom -> factor that scales the run to the requested CTDIvol
    template <typename U>
        requires std::is_same_v<CTAxialSource<T>, U> || std::is_same_v<CTAxialDualSource<T>, U>
    static T ctCalibration(U& sourceCopy, LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr)
    {
        T meanWeight = 0;
        for (std::uint64_t i = 0; i < sourceCopy.totalExposures(); ++i)
            meanWeight += sourceCopy.getExposure(i).beamIntensityWeight();
        meanWeight /= sourceCopy.totalExposures();

        sourceCopy.setDirectionCosines({ -1, 0, 0, 0, 0, 1 });
        sourceCopy.setPosition({ 0, 0, 0 });
        sourceCopy.setScanLenght(sourceCopy.collimation());
        sourceCopy.setUseXCareFilter(false); // organ modulation would bias the CTDI statistics

        std::size_t statCounter = CTDIPhantom<T>::ctdiMinHistories() / (sourceCopy.exposuresPerRotatition() * sourceCopy.historiesPerExposure());
        statCounter = std::max(statCounter, std::size_t { 1 });

        CTDIPhantom<T> world(sourceCopy.ctdiPhantomDiameter());
        sourceCopy.updateFromWorld(world);
        sourceCopy.setHistoriesPerExposure(sourceCopy.historiesPerExposure() * statCounter);
        sourceCopy.validate();
        if (progressBar) {
            progressBar->setPlaneNormal(ProgressBar<T>::Axis::Z);
            progressBar->setPrefixMessage("CTDI calibration ");
        }

        Transport<T> transport;
        transport.setLowEnergyCorrectionModel(model);
        const auto result = transport(world, &sourceCopy, progressBar, false);

        using Hole = typename CTDIPhantom<T>::HolePosition;
        const std::array<Hole, 5> holes = { Hole::Center, Hole::West, Hole::East, Hole::South, Hole::North };
        std::array<T, 5> dose;
        dose.fill(T { 0 });
        for (std::size_t i = 0; i < 5; ++i) {
            const auto& indices = world.holeIndices(holes[i]);
            for (const auto idx : indices)
                dose[i] += result.dose[idx];
            dose[i] /= indices.size();
        }
        const T periphery = (dose[1] + dose[2] + dose[3] + dose[4]) / T { 4 };
        T ctdiw = (dose[0] + 2 * periphery) / 3;
        ctdiw *= T { 100 } / sourceCopy.collimation();
        return sourceCopy.ctdiVol() / ctdiw / meanWeight;
    }

    virtual void updateSpecterDistribution()
    {
        if (m_specterValid)
            return;
        const auto energies = m_tube.getEnergy();
        const auto weights = m_tube.getSpecter(energies);
        m_specterDistribution = std::make_shared<SpecterDistribution<T>>(weights, energies);
        const T heelSpan = std::atan(m_collimation * T { 0.5 } / m_sdd) * T { 2.0 };
        m_heelFilter = m_modelHeelEffect ? std::make_shared<HeelFilter<T>>(m_tube, heelSpan) : nullptr;
        m_specterValid = true;
    }

    T m_sdd;
    T m_collimation;
    T m_fov;
    T m_startAngle;
    T m_scanLenght;
    T m_ctdivol = 1;
    T m_gantryTiltAngle = 0;
    std::uint64_t m_ctdiPhantomDiameter = 320;
    std::shared_ptr<BowTieFilter<T>> m_bowTieFilter;
    Tube<T> m_tube;
    std::shared_ptr<SpecterDistribution<T>> m_specterDistribution;
    std::shared_ptr<HeelFilter<T>> m_heelFilter;
    bool m_modelHeelEffect = true;
    bool m_specterValid = false;
};

// rotating CT source: angular step between exposures, tube current modulation along z (AEC) and
// around the patient (XCare)
template <Floating T>
class CTSource : public CTBaseSource<T> {
public:
    CTSource() { m_exposureAngleStep = DEG_TO_RAD<T>(); }
    virtual Exposure<T> getExposure(std::uint64_t i) const override = 0;

    void setExposureAngleStep(T angleStep) { m_exposureAngleStep = std::clamp(std::abs(angleStep), DEG_TO_RAD<T>() / 10, PI_VAL<T>() / 2); }
    T exposureAngleStep() const { return m_exposureAngleStep; }
    void setExposureAngleStepDeg(T angleStep) { setExposureAngleStep(angleStep * DEG_TO_RAD<T>()); }
    T exposureAngleStepDeg() const { return m_exposureAngleStep * RAD_TO_DEG<T>(); }

    void setAecFilter(std::shared_ptr<AECFilter<T>> filter) { m_aecFilter = filter; }
    std::shared_ptr<AECFilter<T>> aecFilter() { return m_aecFilter; }
    bool useXCareFilter() const { return m_useXCareFilter; }
    void setUseXCareFilter(bool use) { m_useXCareFilter = use; }
    XCareFilter<T>& xcareFilter() { return m_xcareFilter; }
    const XCareFilter<T>& xcareFilter() const { return m_xcareFilter; }

    virtual void updateFromWorld(const World<T>& world) override
    {
        if (m_aecFilter)
            m_aecFilter->updateFromWorld(world);
    }
    virtual std::uint64_t exposuresPerRotatition() const
    {
        constexpr T twoPi = 2 * PI_VAL<T>();
        return static_cast<std::size_t>(twoPi / m_exposureAngleStep);
    }

    T m_exposureAngleStep = RAD_TO_DEG<T>();
    std::shared_ptr<AECFilter<T>> m_aecFilter;
    XCareFilter<T> m_xcareFilter;
    bool m_useXCareFilter = false;

protected:
    // per-exposure weight from the two modulations
    T modulationWeight(T weight, const std::array<T, 3>& pos, T angle) const
    {
        if (m_aecFilter)
            weight *= m_aecFilter->sampleIntensityWeight(pos);
        if (m_useXCareFilter)
            weight *= m_xcareFilter.sampleIntensityWeight(angle);
        return weight;
    }
    std::uint64_t anglesPerRotation() const { return static_cast<std::uint64_t>(2 * PI_VAL<T>() / m_exposureAngleStep); }
};

// two tubes 90 degrees apart, exposures alternate A, B, A, B ...
template <Floating T = double>
class CTDualSource : public CTSource<T> {
public:
    CTDualSource()
    {
        this->m_type = Source<T>::Type::None;
        m_sddB = this->m_sdd;
        m_fovB = this->m_fov;
        m_startAngleB = this->m_startAngle + PI_VAL<T>() * T { 0.5 };
        m_tubeB.setAlFiltration(this->m_tube.AlFiltration());
    }

    T tubeAmas() const { return m_tubeAmas; }
    T tubeBmas() const { return m_tubeBmas; }
    void setTubeAmas(T mas)
    {
        this->m_specterValid = false;
        m_tubeAmas = std::max(T { 0.0 }, mas);
    }
    void setTubeBmas(T mas)
    {
        this->m_specterValid = false;
        m_tubeBmas = std::max(T { 0.0 }, mas);
    }
    Tube<T>& tubeB()
    {
        this->m_specterValid = false;
        return m_tubeB;
    }
    const Tube<T>& tubeB() const { return m_tubeB; }

    T maxPhotonEnergyProduced() const override { return std::max(this->m_tube.voltage(), m_tubeB.voltage()); }
    std::uint64_t exposuresPerRotatition() const override { return 2 * static_cast<std::size_t>((2 * PI_VAL<T>()) / this->m_exposureAngleStep); }
    void setBowTieFilterB(std::shared_ptr<BowTieFilter<T>> filter) { m_bowTieFilterB = filter; }
    std::shared_ptr<BowTieFilter<T>> bowTieFilterB() { return m_bowTieFilterB; }
    const std::shared_ptr<BowTieFilter<T>> bowTieFilterB() const { return m_bowTieFilterB; }
    void setSourceDetectorDistanceB(T sdd)
    {
        this->m_specterValid = false;
        m_sddB = std::abs(sdd);
    }
    T sourceDetectorDistanceB() const { return m_sddB; }
    void setFieldOfViewB(T fov) { m_fovB = std::abs(fov); }
    T fieldOfViewB() const { return m_fovB; }
    void setStartAngleB(T angle) { m_startAngleB = angle; }
    T startAngleB() const { return m_startAngleB; }
    void setStartAngleDegB(T angle) { m_startAngleB = DEG_TO_RAD<T>() * angle; }
    T startAngleDegB() const { return RAD_TO_DEG<T>() * m_startAngleB; }

    bool validate() override
    {
        updateSpecterDistribution();
        return this->m_specterValid;
    }

protected:
    struct TubeSetup {
        T sdd, startAngle, fov, weight;
        const BeamFilter<T>* bowtie;
        const SpecterDistribution<T>* specter;
        const HeelFilter<T>* heel;
    };
    TubeSetup tubeSetup(bool tubeA) const
    {
        if (tubeA)
            return { this->m_sdd, this->m_startAngle, this->m_fov, m_tubeAweight, this->m_bowTieFilter.get(), this->m_specterDistribution.get(),
                this->m_heelFilter.get() };
        return { m_sddB, m_startAngleB, m_fovB, m_tubeBweight, m_bowTieFilterB.get(), m_specterDistributionB.get(), m_heelFilterB.get() };
    }

    // both spectra normalised separately; the tubes' relative output (mAs x unnormalised yield) becomes beam weights
    void updateSpecterDistribution() override
    {
        if (this->m_specterValid)
            return;
        const auto energyA = this->m_tube.getEnergy();
        const auto energyB = m_tubeB.getEnergy();
        auto specterA = this->m_tube.getSpecter(energyA, false);
        auto specterB = m_tubeB.getSpecter(energyB, false);
        const auto sumA = std::accumulate(specterA.cbegin(), specterA.cend(), T { 0.0 });
        const auto sumB = std::accumulate(specterB.cbegin(), specterB.cend(), T { 0.0 });
        const auto weightA = m_tubeAmas * sumA;
        const auto weightB = m_tubeBmas * sumB;
        for (auto& v : specterA)
            v = v / sumA;
        for (auto& v : specterB)
            v = v / sumB;
        m_tubeAweight = weightA * T { 2 } / (weightA + weightB);
        m_tubeBweight = weightB * T { 2 } / (weightA + weightB);
        this->m_specterDistribution = std::make_shared<SpecterDistribution<T>>(specterA, energyA);
        m_specterDistributionB = std::make_shared<SpecterDistribution<T>>(specterB, energyB);
        const auto heelSpan = std::atan(this->m_collimation * T { 0.5 } / this->m_sdd) * T { 2 };
        this->m_heelFilter = std::make_shared<HeelFilter<T>>(this->m_tube, heelSpan);
        m_heelFilterB = std::make_shared<HeelFilter<T>>(m_tubeB, heelSpan);
        this->m_specterValid = true;
    }

    T m_sddB;
    T m_fovB;
    T m_startAngleB;
    T m_tubeAmas = 100.0;
    T m_tubeBmas = 100.0;
    T m_tubeBweight = -1.0;
    T m_tubeAweight = -1.0;
    std::shared_ptr<BowTieFilter<T>> m_bowTieFilterB;
    Tube<T> m_tubeB;
    std::shared_ptr<SpecterDistribution<T>> m_specterDistributionB;
    std::shared_ptr<HeelFilter<T>> m_heelFilterB;
};

template <Floating T = double>
class CTAxialSource final : public CTSource<T> {
public:
    CTAxialSource()
    {
        this->m_type = Source<T>::Type::CTAxial;
        m_step = this->m_collimation;
        this->m_scanLenght = m_step;
    }
    CTAxialSource(const CTSpiralSource<T>& other);

    Exposure<T> getExposure(std::uint64_t exposureIndex) const override
    {
        const std::uint64_t perRotation = this->anglesPerRotation();
        const std::uint64_t rotation = exposureIndex / perRotation;
        const auto angle = this->m_startAngle + this->m_exposureAngleStep * (exposureIndex - (rotation * perRotation));
        const auto frame = this->gantryFrame(this->m_sdd, angle, m_step * rotation);
        const T weight = this->modulationWeight(T { 1 }, frame.position, angle);
        return Exposure<T>(frame.position, frame.cosines, this->openingAngles(this->m_fov, this->m_sdd), this->m_historiesPerExposure, weight,
            this->m_specterDistribution.get(), this->m_heelFilter.get(), this->m_bowTieFilter.get());
    }

    void setStep(T step)
    {
        const auto absStep = std::abs(step);
        const auto nSteps = this->m_scanLenght / m_step;
        m_step = absStep > 0.01 ? absStep : 0.01;
        setScanLenght(m_step * nSteps);
    }
    T step() const { return m_step; }
    void setScanLenght(T scanLenght) override { this->m_scanLenght = std::max(m_step * std::ceil(std::abs(scanLenght) / m_step), m_step); }

    std::uint64_t totalExposures() const override
    {
        const std::uint64_t rotations = static_cast<std::uint64_t>(std::round(this->m_scanLenght / m_step));
        return static_cast<std::uint64_t>(PI_VAL<T>() * T { 2 } / this->m_exposureAngleStep) * rotations;
    }
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        auto copy = *this;
        return CTSource<T>::ctCalibration(copy, model, progressBar);
    }

private:
    T m_step;
};

template <Floating T = double>
class CTSpiralSource final : public CTSource<T> {
public:
    CTSpiralSource()
    {
        this->m_type = Source<T>::Type::CTSpiral;
        m_pitch = 1.0;
    }

    Exposure<T> getExposure(std::uint64_t exposureIndex) const override
    {
        constexpr T twoPi = T { 2 } * PI_VAL<T>();
        const auto angle = this->m_startAngle + this->m_exposureAngleStep * exposureIndex;
        const T zAdvance = (exposureIndex * this->m_exposureAngleStep) * this->m_collimation * m_pitch / twoPi;
        const auto frame = this->gantryFrame(this->m_sdd, angle, zAdvance);
        const T weight = this->modulationWeight(T { 1.0 }, frame.position, angle);
        return Exposure<T>(frame.position, frame.cosines, this->openingAngles(this->m_fov, this->m_sdd), this->m_historiesPerExposure, weight,
            this->m_specterDistribution.get(), this->m_heelFilter.get(), this->m_bowTieFilter.get());
    }

    void setPitch(T pitch) { m_pitch = std::max(T { 0.01 }, pitch); }
    T pitch() const { return m_pitch; }
    void setScanLenght(T scanLenght) override { this->m_scanLenght = std::max(std::abs(scanLenght), this->m_collimation * m_pitch * T { 0.5 }); }
    std::uint64_t totalExposures() const override
    {
        constexpr T twoPi = 2 * PI_VAL<T>();
        return static_cast<std::uint64_t>(this->m_scanLenght * twoPi / (this->m_collimation * m_pitch * this->m_exposureAngleStep));
    }
    // CTDIvol of a spiral = CTDIw / pitch: calibrate the equivalent axial scan, then scale
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        CTAxialSource<T> copy = *this;
        return this->ctCalibration(copy, model, progressBar) * m_pitch;
    }

private:
    T m_pitch;
};

template <Floating T = double>
class CTAxialDualSource final : public CTDualSource<T> {
public:
    CTAxialDualSource()
    {
        this->m_type = Source<T>::Type::CTDual;
        m_step = this->m_collimation;
        this->m_scanLenght = m_step;
    }
    CTAxialDualSource(const CTSpiralDualSource<T>& other);

    Exposure<T> getExposure(std::uint64_t exposureIndexTotal) const override
    {
        const std::uint64_t exposureIndex = exposureIndexTotal / 2;
        const auto tube = this->tubeSetup(exposureIndexTotal % 2 == 0);
        const std::uint64_t perRotation = this->anglesPerRotation();
        const std::uint64_t rotation = exposureIndex / perRotation;
        const auto angle = tube.startAngle + this->m_exposureAngleStep * (exposureIndex - (rotation * perRotation));
        // the focal-spot radius is tube A's for both tubes, as in the reference
        const auto frame = this->gantryFrame(this->m_sdd, angle, m_step * rotation);
        const T weight = this->modulationWeight(tube.weight, frame.position, angle);
        return Exposure<T>(frame.position, frame.cosines, this->openingAngles(tube.fov, tube.sdd), this->m_historiesPerExposure, weight, tube.specter,
            tube.heel, tube.bowtie);
    }

    void setStep(T step)
    {
        const auto absStep = std::abs(step);
        const auto nSteps = this->m_scanLenght / m_step;
        m_step = absStep > 0.01 ? absStep : 0.01;
        setScanLenght(m_step * nSteps);
    }
    T step() const { return m_step; }
    void setScanLenght(T scanLenght) override { this->m_scanLenght = std::max(m_step * std::ceil(std::abs(scanLenght) / m_step), m_step); }
    std::uint64_t totalExposures() const override
    {
        const std::uint64_t rotations = static_cast<std::uint64_t>(std::round(this->m_scanLenght / m_step));
        return static_cast<std::uint64_t>(PI_VAL<T>() * T { 2 } / this->m_exposureAngleStep) * rotations * 2;
    }
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        auto copy = *this;
        return this->ctCalibration(copy, model, progressBar);
    }

private:
    T m_step;
};

template <Floating T = double>
class CTSpiralDualSource final : public CTDualSource<T> {
public:
    CTSpiralDualSource()
    {
        this->m_type = Source<T>::Type::CTDual;
        m_pitch = 1.0;
    }

    Exposure<T> getExposure(std::uint64_t exposureIndexTotal) const override
    {
        constexpr T twoPi = T { 2 } * PI_VAL<T>();
        const std::uint64_t exposureIndex = exposureIndexTotal / 2;
        const auto tube = this->tubeSetup(exposureIndexTotal % 2 == 0);
        const auto angle = tube.startAngle + this->m_exposureAngleStep * exposureIndex;
        const T zAdvance = (exposureIndex * this->m_exposureAngleStep) * this->m_collimation * m_pitch / twoPi;
        const auto frame = this->gantryFrame(this->m_sdd, angle, zAdvance);
        const T weight = this->modulationWeight(tube.weight, frame.position, angle);
        return Exposure<T>(frame.position, frame.cosines, this->openingAngles(tube.fov, tube.sdd), this->m_historiesPerExposure, weight, tube.specter,
            tube.heel, tube.bowtie);
    }
    std::uint64_t totalExposures() const override
    {
        const auto single = static_cast<std::uint64_t>(this->scanLenght() * 2 * PI_VAL<T>() / (this->collimation() * pitch() * this->exposureAngleStep()));
        return single * 2;
    }
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        CTAxialDualSource<T> copy = *this;
        return CTSource<T>::ctCalibration(copy, model, progressBar) * m_pitch;
    }
    T pitch() const { return m_pitch; }
    void setPitch(T pitch) { m_pitch = std::max(T { 0.01 }, pitch); }
    void setScanLenght(T scanLenght) override
    {
        CTBaseSource<T>::setScanLenght(std::max(std::abs(scanLenght), this->collimation() * pitch() * T { 0.5 }));
    }

private:
    T m_pitch = 1.0;
};

template <Floating T>
CTAxialSource<T>::CTAxialSource(const CTSpiralSource<T>& other)
    : CTSource<T>(other)
{
    this->m_step = this->m_collimation;
    setScanLenght(other.scanLenght());
}

template <Floating T>
CTAxialDualSource<T>::CTAxialDualSource(const CTSpiralDualSource<T>& other)
    : CTDualSource<T>(other)
{
    m_step = this->m_collimation;
    setScanLenght(other.scanLenght());
}

// scout view: the tube parked at the start angle while the table moves through the scan length
template <Floating T>
class CTTopogramSource : public CTBaseSource<T> {
public:
    CTTopogramSource() { this->m_type = Source<T>::Type::CTTopogram; }

    Exposure<T> getExposure(std::uint64_t i) const override
    {
        const auto step = this->scanLenght() / (totalExposures() - 1);
        const auto frame = this->gantryFrame(this->m_sdd, this->m_startAngle, step * i);
        return Exposure<T>(frame.position, frame.cosines, this->openingAngles(this->m_fov, this->m_sdd), this->m_historiesPerExposure, T { 1 },
            this->m_specterDistribution.get(), this->m_heelFilter.get(), this->m_bowTieFilter.get());
    }
    std::uint64_t totalExposures() const override { return std::max(static_cast<std::uint64_t>(std::ceil(this->scanLenght())), std::uint64_t { 1 }); }

    // calibrated through an axial scan of equal collimation whose CTDIvol is scaled by scan length / collimation
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        CTAxialSource<T> copy;
        static_cast<CTBaseSource<T>&>(copy) = *this;
        copy.setCtdiVol(this->ctdiVol() * this->scanLenght() / this->collimation());
        copy.setScanLenght(0);
        copy.setStep(this->m_collimation);
        constexpr auto maxStep = (2 * PI_VAL<T>()) / 72;
        copy.setExposureAngleStep(std::min(2 * PI_VAL<T>() / totalExposures(), maxStep));
        const auto exposures = this->totalExposures();
        const auto factor = CTSource<T>::ctCalibration(copy, model, progressBar);
        return (factor * exposures) / copy.totalExposures();
    }
};
}
