// dapsource.hpp — tube-based projection sources calibrated by dose-area product: DAPSource and its two concrete
// forms DXSource (radiography) and CBCTSource (cone-beam rotation); reference include/dxmc/source.hpp:374-787.
// Included by dxmc/source.hpp.
#pragma once
#include "dxmc/sourcebase.hpp"

namespace dxmc {

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
}
