// dapsource.hpp — tube-based projection sources calibrated by dose-area product: DAPSource, DXSource (radiography) and
// CBCTSource (cone-beam rotation). API of reference include/dxmc/source.hpp:374-787; the state is the parameter block of
// dxmc/sourcemodel.hpp (the focal spot sits `focalOffset` upstream of Source::position along the beam; a cone-beam scan is an
// ORBIT about the y cosine), the classes are setters over it. Included by dxmc/source.hpp.
#pragma once
#include "dxmc/sourcebase.hpp"

namespace dxmc {

template <Floating T = double>
class DAPSource : public Source<T> {
public:
    DAPSource()
    {
        m_tube.setAlFiltration(2.0);
        this->m_p.spectrum[0] = 0;
        this->setDirectionCosines(beamAlongY());
        setFieldSize({ 100.0, 100.0 });
    }

    Tube<T>& tube()
    {
        m_specterValid = false;
        return m_tube;
    }
    const Tube<T>& tube() const { return m_tube; }
    T maxPhotonEnergyProduced() const override { return m_tube.voltage(); }

    // field size at the detector <-> full opening angles; either setter keeps the other quantity consistent
    void setCollimationAngles(const std::array<T, 2>& angles)
    {
        for (std::size_t k = 0; k < 2; ++k) {
            m_opening[k] = std::abs(angles[k]);
            m_fieldSize[k] = std::tan(m_opening[k] / 2) * m_sdd * 2;
        }
        openingChanged();
    }
    const std::array<T, 2>& collimationAngles() const { return m_opening; }
    void setCollimationAnglesDeg(const std::array<T, 2>& angles) { setCollimationAngles({ angles[0] * DEG_TO_RAD<T>(), angles[1] * DEG_TO_RAD<T>() }); }
    const std::array<T, 2> collimationAnglesDeg() const { return { m_opening[0] * RAD_TO_DEG<T>(), m_opening[1] * RAD_TO_DEG<T>() }; }
    void setFieldSize(const std::array<T, 2>& mm)
    {
        for (std::size_t k = 0; k < 2; ++k) {
            m_fieldSize[k] = std::abs(mm[k]);
            m_opening[k] = std::atan(m_fieldSize[k] * T { 0.5 } / m_sdd) * T { 2 };
        }
        openingChanged();
    }
    const std::array<T, 2>& fieldSize() const { return m_fieldSize; }
    void setSourceDetectorDistance(T mm)
    {
        m_sdd = std::abs(mm);
        setFieldSize(m_fieldSize);
        this->m_p.focalOffset = focalDistance();
    }
    T sourceDetectorDistance() const { return m_sdd; }

    // Orientation of the beam frame from two patient angles and the tube's own rotation: start with the beam along +y,
    // turn about z (primary), tip about x (secondary, kept short of +-90 degrees), then spin about the beam.
    void setSourceAngles(T primaryAngle, T secondaryAngle)
    {
        constexpr T margin = 1E-6;
        constexpr T quarterTurn = PI_VAL<T>() / 2;
        secondaryAngle = std::min(std::max(secondaryAngle, margin - quarterTurn), quarterTurn - margin);
        while (primaryAngle > PI_VAL<T>())
            primaryAngle -= PI_VAL<T>();
        while (primaryAngle < -PI_VAL<T>())
            primaryAngle += PI_VAL<T>();
        auto frame = beamAlongY();
        const std::array<T, 3> zAxis = { .0, .0, 1.0 }, xAxis = { 1.0, .0, .0 };
        turnFrame(frame, zAxis.data(), primaryAngle);
        turnFrame(frame, xAxis.data(), -secondaryAngle);
        spinAboutBeam(frame, m_tubeRotationAngle);
        this->setDirectionCosines(frame);
    }
    void setSourceAngles(const std::array<T, 2>& angles) { setSourceAngles(angles[0], angles[1]); }
    // the two patient angles back from the frame: undo the tube spin, read them off the beam direction
    std::array<T, 2> sourceAngles() const
    {
        constexpr T margin = 1E-6;
        auto frame = this->directionCosines();
        spinAboutBeam(frame, -m_tubeRotationAngle);
        std::array<T, 3> beam;
        vectormath::cross(frame.data(), beam.data());
        if (std::abs(std::sqrt(beam[0] * beam[0] + beam[1] * beam[1])) < margin)
            return { 0, beam[2] > 0 ? -PI_VAL<T>() / 2 : PI_VAL<T>() / 2 };
        const T primary = std::asin(-beam[0]);
        const T inYZ = std::sqrt(beam[2] * beam[2] + beam[1] * beam[1]);
        if (std::abs(inYZ) < margin)
            return { primary, 0 };
        return { primary, -std::asin(beam[2] / inYZ) };
    }
    void setSourceAnglesDeg(T primaryAngle, T secondaryAngle) { setSourceAngles(primaryAngle * DEG_TO_RAD<T>(), secondaryAngle * DEG_TO_RAD<T>()); }
    void setSourceAnglesDeg(const std::array<T, 2>& angles) { setSourceAnglesDeg(angles[0], angles[1]); }
    std::array<T, 2> sourceAnglesDeg() const
    {
        const auto a = sourceAngles();
        return { a[0] * RAD_TO_DEG<T>(), a[1] * RAD_TO_DEG<T>() };
    }

    void setTubeRotation(T angle)
    {
        auto frame = this->directionCosines();
        spinAboutBeam(frame, angle - m_tubeRotationAngle);
        this->setDirectionCosines(frame);
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

    // requested dose-area product over the air kerma the run's photons carry (normalised spectrum x mass energy absorption of air)
    T getCalibrationValue(LOWENERGYCORRECTION, ProgressBar<T>* = nullptr) const override
    {
        const Material air("Air, Dry (near sea level)");
        T kermaPerPhoton = 0.0;
        for (const auto& [keV, weight] : tube().getSpecter(true)) {
            const T absorption = air.getMassEnergyAbsorbtion(keV);
            kermaPerPhoton += keV * weight * absorption;
        }
        return m_dap / (kermaPerPhoton * (this->totalExposures() * this->historiesPerExposure()));
    }

    bool isValid() const override { return m_specterValid; }
    bool validate() override
    {
        if (!m_specterValid) {
            const auto energies = m_tube.getEnergy();
            m_specterDistribution = std::make_shared<SpecterDistribution<T>>(m_tube.getSpecter(energies), energies);
            m_heelFilter = m_modelHeelEffect ? std::make_shared<HeelFilter<T>>(m_tube, m_opening[1]) : nullptr;
            this->m_p.heel[0] = m_heelFilter ? 0 : -1;
            m_specterValid = true;
        }
        return m_specterValid;
    }
    void setModelHeelEffect(bool on) { m_modelHeelEffect = on; }
    bool modelHeelEffect() const { return m_modelHeelEffect; }
    typename Source<T>::BeamTables beamTables(std::uint32_t) const override { return { m_specterDistribution.get(), m_heelFilter.get(), nullptr }; }

protected:
    virtual T focalDistance() const { return m_sdd; } // how far upstream of Source::position the focal spot sits
    const std::array<T, 3> focalSpot() const
    {
        std::array<T, 3> beam, spot;
        vectormath::cross(this->m_p.cosines, beam.data());
        for (std::size_t k = 0; k < 3; ++k)
            spot[k] = this->m_p.position[k] - beam[k] * this->m_p.focalOffset;
        return spot;
    }
    void openingChanged()
    {
        this->setOpening(m_opening[0], m_opening[1]);
        m_specterValid = false;
    }
    static void turnFrame(std::array<T, 6>& frame, const T* axis, T angle)
    {
        vectormath::rotate(frame.data(), axis, angle);
        vectormath::rotate(frame.data() + 3, axis, angle);
    }
    static void spinAboutBeam(std::array<T, 6>& frame, T angle)
    {
        std::array<T, 3> beam;
        vectormath::cross(frame.data(), beam.data());
        turnFrame(frame, beam.data(), angle);
    }
    // anode-cathode axis along z; setSourceAngles is defined relative to this frame
    static std::array<T, 6> beamAlongY() { return { -1.0, .0, .0, .0, .0, 1.0 }; }

    T m_sdd = 1000.0;
    T m_dap = 1.0; // Gy cm2
    std::array<T, 2> m_fieldSize { 100.0, 100.0 };
    std::array<T, 2> m_opening { 0, 0 };
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
    DXSource()
    {
        this->m_type = Source<T>::Type::DX;
        this->m_p.exposures = 1000;
        this->m_p.focalOffset = this->focalDistance();
    }
    std::uint64_t totalExposures() const override { return this->m_p.exposures; }
    void setTotalExposures(std::uint64_t exposures) { this->m_p.exposures = std::max(exposures, std::uint64_t { 1 }); }
    const std::array<T, 3> tubePosition() const { return this->focalSpot(); }
};

// cone-beam CT: the radiography geometry with the focal spot half a source-detector distance from the isocentre, stepped about
// the y cosine
template <Floating T = double>
class CBCTSource final : public DAPSource<T> {
public:
    CBCTSource()
    {
        this->m_type = Source<T>::Type::CBCT;
        this->m_p.motion = model::ORBIT;
        this->setSourceDetectorDistance(500.0);
        setStepAngle(PI_VAL<T>() / T { 180 });
    }

    const std::array<T, 3> rotationAxis() const { return { this->m_p.cosines[3], this->m_p.cosines[4], this->m_p.cosines[5] }; }
    void setSpanAngle(const T spanAngle)
    {
        m_angleSpan = std::max(spanAngle, this->m_p.orbitStep);
        recount();
    }
    void setSpanAngleDeg(const T spanAngle) { setSpanAngle(spanAngle * DEG_TO_RAD<T>()); }
    const T spanAngle() const { return m_angleSpan; }
    const T spanAngleDeg() const { return m_angleSpan * RAD_TO_DEG<T>(); }
    void setStepAngle(const T stepAngle)
    {
        constexpr T finest = PI_VAL<T>() / T { 360 };
        this->m_p.orbitStep = std::max(stepAngle, finest);
        recount();
    }
    void setStepAngleDeg(const T stepAngle) { setStepAngle(stepAngle * DEG_TO_RAD<T>()); }
    const T stepAngle() const { return this->m_p.orbitStep; }
    const T stepAngleDeg() const { return this->m_p.orbitStep * RAD_TO_DEG<T>(); }
    std::uint64_t totalExposures() const override { return this->m_p.exposures; }
    const std::array<T, 3> tubePosition() const { return this->focalSpot(); }

protected:
    T focalDistance() const override { return this->m_sdd * T { 0.5 }; }

private:
    void recount() { this->m_p.exposures = std::max(static_cast<std::size_t>(m_angleSpan / this->m_p.orbitStep), std::size_t { 2 }); }

    T m_angleSpan = PI_VAL<T>();
};
}
