// ctsource.hpp — the CT source family: CTBaseSource -> CTSource -> CTAxialSource / CTSpiralSource, CTDualSource ->
// CTAxialDualSource / CTSpiralDualSource, CTTopogramSource; reference include/dxmc/source.hpp:789-1700. getExposure(i)
// is O(1) host code that Transport evaluates for every exposure up front; ctCalibration runs a second Transport on a
// CTDIPhantom like the reference does, i.e. a second pass through the same CUDA path. Included by dxmc/source.hpp.
#pragma once
#include "dxmc/sourcebase.hpp"

namespace dxmc {

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
