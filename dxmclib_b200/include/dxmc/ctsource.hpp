// ctsource.hpp — the CT source family: CTBaseSource -> CTSource -> CTAxialSource / CTSpiralSource, CTDualSource ->
// CTAxialDualSource / CTSpiralDualSource, CTTopogramSource. API of reference include/dxmc/source.hpp:789-1700.
//
// A CT scan is the GANTRY_* motion of the parameter block in dxmc/sourcemodel.hpp: per tube a focal-spot radius, field of
// view, start angle and weight; a beam width, angular step, pitch or table step, gantry tilt; the two tube-current
// modulations. The classes are setters over that block (names, defaults and clamps as in the reference); exposure i is
// model::evaluate(block, i) on the host and exposureKernel on the device. The CTDI calibration (reference :925-988) is
// a second Transport run over a CTDIPhantom, i.e. a second pass through the same CUDA path. Included by dxmc/source.hpp.
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

// gantry geometry, tube A, bow-tie, CTDI calibration target
template <Floating T>
class CTBaseSource : public Source<T> {
public:
    CTBaseSource()
    {
        auto& p = this->m_p;
        p.motion = model::GANTRY_AXIAL;
        p.sdd[0] = p.sdd[1] = 1190.0;
        p.fov[0] = p.fov[1] = 500.0;
        p.beamWidth = 38.4;
        p.scanLength = 100.0;
        p.spectrum[0] = 0;
        m_tube.setAlFiltration(7.0);
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
        this->m_p.sdd[0] = std::abs(sdd);
        m_specterValid = false;
    }
    T sourceDetectorDistance() const { return this->m_p.sdd[0]; }
    void setCollimation(T collimation)
    {
        this->m_p.beamWidth = std::abs(collimation);
        m_specterValid = false;
    }
    T collimation() const { return this->m_p.beamWidth; }
    void setFieldOfView(T fov) { this->m_p.fov[0] = std::abs(fov); }
    T fieldOfView() const { return this->m_p.fov[0]; }

    void setGantryTiltAngle(T angle) { this->m_p.tilt = std::clamp(angle, -PI_VAL<T>(), PI_VAL<T>()); }
    T gantryTiltAngle() const { return this->m_p.tilt; }
    void setGantryTiltAngleDeg(T angle) { setGantryTiltAngle(angle * DEG_TO_RAD<T>()); }
    T gantryTiltAngleDeg() const { return this->m_p.tilt * RAD_TO_DEG<T>(); }

    void setStartAngle(T angle) { this->m_p.startAngle[0] = angle; }
    T startAngle() const { return this->m_p.startAngle[0]; }
    T startAngleDeg() const { return RAD_TO_DEG<T>() * this->m_p.startAngle[0]; }
    void setStartAngleDeg(T angle) { this->m_p.startAngle[0] = DEG_TO_RAD<T>() * angle; }

    virtual void setScanLenght(T scanLenght) { this->m_p.scanLength = std::abs(scanLenght); }
    T scanLenght() const { return this->m_p.scanLength; }

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
    typename Source<T>::BeamTables beamTables(std::uint32_t) const override
    {
        return { m_specterDistribution.get(), m_heelFilter.get(), m_bowTieFilter.get() };
    }
    // take over what another CT source holds at this level (geometry of tube A, beam width, scan length, tilt, tube, bow-tie,
    // spectrum state, CTDI target) and leave the rest of this source (motion, angular step, table step, modulation) alone
    void adoptBaseOf(const CTBaseSource& other)
    {
        auto& p = this->m_p;
        const auto& q = other.m_p;
        std::copy(q.position, q.position + 3, p.position);
        std::copy(q.cosines, q.cosines + 6, p.cosines);
        p.histories = q.histories;
        p.sdd[0] = q.sdd[0];
        p.fov[0] = q.fov[0];
        p.startAngle[0] = q.startAngle[0];
        p.beamWidth = q.beamWidth;
        p.scanLength = q.scanLength;
        p.tilt = q.tilt;
        m_ctdivol = other.m_ctdivol;
        m_ctdiPhantomDiameter = other.m_ctdiPhantomDiameter;
        m_bowTieFilter = other.m_bowTieFilter;
        m_tube = other.m_tube;
        m_specterDistribution = other.m_specterDistribution;
        m_heelFilter = other.m_heelFilter;
        m_modelHeelEffect = other.m_modelHeelEffect;
        m_specterValid = other.m_specterValid;
    }
    bool describe(model::SourceParams<T>& block, const T*& tubeCurrent) const override
    {
        Source<T>::describe(block, tubeCurrent);
        const auto a = this->beamTables(0);
        block.spectrum[0] = a.specter ? 0 : -1;
        block.heel[0] = a.heel ? 0 : -1;
        block.bowtie[0] = a.fan ? 0 : -1;
        if (block.tubes == 2) {
            const auto b = this->beamTables(1);
            block.spectrum[1] = b.specter ? 1 : -1;
            block.heel[1] = b.heel ? 1 : -1;
            block.bowtie[1] = b.fan ? 1 : -1;
        }
        return true;
    }

protected:
    // Dose calibration of a rotating CT source (reference source.hpp:925-988): the scan's axial twin, centred and untilted, makes
    // one rotation over a CTDI phantom with at least CTDIPhantom::ctdiMinHistories() histories; CTDIw from the five chamber
    // bores, scaled to 100 mm / beam width, against the requested CTDIvol and the mean beam weight of the twin.
    template <typename U>
        requires std::is_same_v<CTAxialSource<T>, U> || std::is_same_v<CTAxialDualSource<T>, U>
    static T ctCalibration(U& twin, LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr)
    {
        const T meanWeight = twin.meanBeamWeight();
        twin.setDirectionCosines({ -1, 0, 0, 0, 0, 1 });
        twin.setPosition({ 0, 0, 0 });
        twin.setScanLenght(twin.collimation());
        twin.setUseXCareFilter(false); // organ modulation would bias the CTDI statistics
        const std::size_t repeats = std::max<std::size_t>(CTDIPhantom<T>::ctdiMinHistories() / (twin.exposuresPerRotatition() * twin.historiesPerExposure()), 1);
        CTDIPhantom<T> phantom(twin.ctdiPhantomDiameter());
        twin.updateFromWorld(phantom);
        twin.setHistoriesPerExposure(twin.historiesPerExposure() * repeats);
        twin.validate();
        if (progressBar) {
            progressBar->setPlaneNormal(ProgressBar<T>::Axis::Z);
            progressBar->setPrefixMessage("CTDI calibration ");
        }
        Transport<T> transport;
        transport.setLowEnergyCorrectionModel(model);
        const auto result = transport(phantom, &twin, progressBar, false);

        using Bore = typename CTDIPhantom<T>::HolePosition;
        T bore[5];
        int k = 0;
        for (const Bore where : { Bore::Center, Bore::West, Bore::East, Bore::South, Bore::North }) {
            const auto& voxels = phantom.holeIndices(where);
            T sum = 0;
            for (const auto v : voxels)
                sum += result.dose[v];
            bore[k++] = sum / voxels.size();
        }
        const T periphery = (bore[1] + bore[2] + bore[3] + bore[4]) / T { 4 };
        T ctdiw = (bore[0] + 2 * periphery) / 3;
        ctdiw *= T { 100 } / twin.collimation();
        return twin.ctdiVol() / ctdiw / meanWeight;
    }

    virtual void updateSpecterDistribution()
    {
        if (m_specterValid)
            return;
        const auto energies = m_tube.getEnergy();
        m_specterDistribution = std::make_shared<SpecterDistribution<T>>(m_tube.getSpecter(energies), energies);
        m_heelFilter = m_modelHeelEffect ? std::make_shared<HeelFilter<T>>(m_tube, heelSpan()) : nullptr;
        m_specterValid = true;
    }
    // take-off angle range the heel table covers: the cone angle of the beam width seen from the focal spot of tube A
    T heelSpan() const { return std::atan(this->m_p.beamWidth * T { 0.5 } / this->m_p.sdd[0]) * T { 2.0 }; }

    T m_ctdivol = 1;
    std::uint64_t m_ctdiPhantomDiameter = 320;
    std::shared_ptr<BowTieFilter<T>> m_bowTieFilter;
    Tube<T> m_tube;
    std::shared_ptr<SpecterDistribution<T>> m_specterDistribution;
    std::shared_ptr<HeelFilter<T>> m_heelFilter;
    bool m_modelHeelEffect = true;
    bool m_specterValid = false;
};

// rotating CT source: angular step between exposures, tube current modulation along z (AEC) and around the patient (XCare)
template <Floating T>
class CTSource : public CTBaseSource<T> {
public:
    CTSource() { this->m_p.angleStep = DEG_TO_RAD<T>(); }

    void setExposureAngleStep(T angleStep) { this->m_p.angleStep = std::clamp(std::abs(angleStep), DEG_TO_RAD<T>() / 10, PI_VAL<T>() / 2); }
    T exposureAngleStep() const { return this->m_p.angleStep; }
    void setExposureAngleStepDeg(T angleStep) { setExposureAngleStep(angleStep * DEG_TO_RAD<T>()); }
    T exposureAngleStepDeg() const { return this->m_p.angleStep * RAD_TO_DEG<T>(); }

    void setAecFilter(std::shared_ptr<AECFilter<T>> filter) { m_aecFilter = filter; }
    std::shared_ptr<AECFilter<T>> aecFilter() { return m_aecFilter; }
    bool useXCareFilter() const { return this->m_p.xcare != 0; }
    void setUseXCareFilter(bool use) { this->m_p.xcare = use ? 1 : 0; }
    XCareFilter<T>& xcareFilter() { return m_xcareFilter; }
    const XCareFilter<T>& xcareFilter() const { return m_xcareFilter; }

    virtual void updateFromWorld(const World<T>& world) override
    {
        if (m_aecFilter)
            m_aecFilter->updateFromWorld(world);
    }
    virtual std::uint64_t exposuresPerRotatition() const { return this->m_p.tubes * anglesPerRotation(); }

    // the modulation objects are the caller's to change at any time: their current values go into the block when it is read
    bool describe(model::SourceParams<T>& block, const T*& tubeCurrent) const override
    {
        CTBaseSource<T>::describe(block, tubeCurrent);
        withModulation(block);
        tubeCurrent = tubeCurrentProfile();
        return true;
    }
    Exposure<T> getExposure(std::uint64_t i) const override
    {
        auto block = this->parameters();
        withModulation(block);
        model::ExposureValues<T> v;
        model::evaluate(block, tubeCurrentProfile(), i, v);
        const auto tables = this->beamTables(block.tubes == 2 ? static_cast<std::uint32_t>(i & 1u) : 0u);
        return Exposure<T>::fromModel(v, tables.specter, tables.heel, tables.fan);
    }
    // mean beam weight over all exposures of the scan
    T meanBeamWeight() const
    {
        auto block = this->parameters();
        withModulation(block);
        model::ExposureValues<T> v;
        T sum = 0;
        for (std::uint64_t i = 0; i < block.exposures; ++i) {
            model::evaluate(block, tubeCurrentProfile(), i, v);
            sum += v.weight;
        }
        return sum / block.exposures;
    }

protected:
    const T* tubeCurrentProfile() const override { return m_aecFilter ? m_aecFilter->positionIntensity().data() : nullptr; }
    void withModulation(model::SourceParams<T>& block) const
    {
        block.xcareAngle = m_xcareFilter.filterAngle();
        block.xcareSpan = m_xcareFilter.spanAngle();
        block.xcareRamp = m_xcareFilter.rampAngle();
        block.xcareLow = m_xcareFilter.lowWeight();
        block.aecSize = 0;
        if (m_aecFilter) {
            block.aecSize = static_cast<std::uint32_t>(m_aecFilter->positionIntensity().size());
            block.aecMin = m_aecFilter->positionMin();
            block.aecMax = m_aecFilter->positionMax();
            block.aecStep = m_aecFilter->positionStep();
        }
    }
    std::uint64_t anglesPerRotation() const { return static_cast<std::uint64_t>(2 * PI_VAL<T>() / this->m_p.angleStep); }
    // table positions of a step-and-shoot scan: the scan length is a whole number of steps, at least one
    void setAxialStep(T step)
    {
        const T steps = this->m_p.scanLength / this->m_p.tableStep;
        this->m_p.tableStep = std::max(std::abs(step), T { 0.01 });
        setAxialScanLength(this->m_p.tableStep * steps);
    }
    void setAxialScanLength(T scanLenght)
    {
        this->m_p.scanLength = std::max(this->m_p.tableStep * std::ceil(std::abs(scanLenght) / this->m_p.tableStep), this->m_p.tableStep);
    }
    std::uint64_t axialExposures() const
    {
        const std::uint64_t rotations = static_cast<std::uint64_t>(std::round(this->m_p.scanLength / this->m_p.tableStep));
        return static_cast<std::uint64_t>(PI_VAL<T>() * T { 2 } / this->m_p.angleStep) * rotations * this->m_p.tubes;
    }
    std::uint64_t spiralExposures() const
    {
        constexpr T twoPi = 2 * PI_VAL<T>();
        return static_cast<std::uint64_t>(this->m_p.scanLength * twoPi / (this->m_p.beamWidth * this->m_p.pitch * this->m_p.angleStep)) * this->m_p.tubes;
    }
    void becomeAxialTwin(T scanLenght) // a spiral scan's parameters reinterpreted as step-and-shoot with step = beam width
    {
        this->m_p.motion = model::GANTRY_AXIAL;
        this->m_p.tableStep = this->m_p.beamWidth;
        setAxialScanLength(scanLenght);
    }

    std::shared_ptr<AECFilter<T>> m_aecFilter;
    XCareFilter<T> m_xcareFilter;
};

// two tubes 90 degrees apart, exposures alternate A, B, A, B ...
template <Floating T = double>
class CTDualSource : public CTSource<T> {
public:
    CTDualSource()
    {
        auto& p = this->m_p;
        p.tubes = 2;
        p.sdd[1] = p.sdd[0];
        p.fov[1] = p.fov[0];
        p.startAngle[1] = p.startAngle[0] + PI_VAL<T>() * T { 0.5 };
        p.tubeWeight[0] = p.tubeWeight[1] = -1.0;
        p.spectrum[1] = 1;
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
    void setBowTieFilterB(std::shared_ptr<BowTieFilter<T>> filter) { m_bowTieFilterB = filter; }
    std::shared_ptr<BowTieFilter<T>> bowTieFilterB() { return m_bowTieFilterB; }
    const std::shared_ptr<BowTieFilter<T>> bowTieFilterB() const { return m_bowTieFilterB; }
    void setSourceDetectorDistanceB(T sdd)
    {
        this->m_specterValid = false;
        this->m_p.sdd[1] = std::abs(sdd);
    }
    T sourceDetectorDistanceB() const { return this->m_p.sdd[1]; }
    void setFieldOfViewB(T fov) { this->m_p.fov[1] = std::abs(fov); }
    T fieldOfViewB() const { return this->m_p.fov[1]; }
    void setStartAngleB(T angle) { this->m_p.startAngle[1] = angle; }
    T startAngleB() const { return this->m_p.startAngle[1]; }
    void setStartAngleDegB(T angle) { this->m_p.startAngle[1] = DEG_TO_RAD<T>() * angle; }
    T startAngleDegB() const { return RAD_TO_DEG<T>() * this->m_p.startAngle[1]; }

    bool validate() override
    {
        updateSpecterDistribution();
        return this->m_specterValid;
    }
    typename Source<T>::BeamTables beamTables(std::uint32_t tube) const override
    {
        if (tube == 0)
            return CTSource<T>::beamTables(0);
        return { m_specterDistributionB.get(), m_heelFilterB.get(), m_bowTieFilterB.get() };
    }

protected:
    // each spectrum normalised on its own; the tubes' relative output (mAs x unnormalised yield) becomes the two beam weights,
    // which average one
    void updateSpecterDistribution() override
    {
        if (this->m_specterValid)
            return;
        T output[2];
        std::shared_ptr<SpecterDistribution<T>>* spectrum[2] = { &this->m_specterDistribution, &m_specterDistributionB };
        const Tube<T>* tubes[2] = { &this->m_tube, &m_tubeB };
        const T mas[2] = { m_tubeAmas, m_tubeBmas };
        for (int k = 0; k < 2; ++k) {
            const auto energies = tubes[k]->getEnergy();
            auto yield = tubes[k]->getSpecter(energies, false);
            const auto total = std::accumulate(yield.cbegin(), yield.cend(), T { 0.0 });
            output[k] = mas[k] * total;
            for (auto& v : yield)
                v = v / total;
            *spectrum[k] = std::make_shared<SpecterDistribution<T>>(yield, energies);
        }
        this->m_p.tubeWeight[0] = output[0] * T { 2 } / (output[0] + output[1]);
        this->m_p.tubeWeight[1] = output[1] * T { 2 } / (output[0] + output[1]);
        this->m_heelFilter = std::make_shared<HeelFilter<T>>(this->m_tube, this->heelSpan());
        m_heelFilterB = std::make_shared<HeelFilter<T>>(m_tubeB, this->heelSpan());
        this->m_specterValid = true;
    }

    T m_tubeAmas = 100.0;
    T m_tubeBmas = 100.0;
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
        this->m_p.tableStep = this->m_p.beamWidth;
        this->m_p.scanLength = this->m_p.tableStep;
    }
    CTAxialSource(const CTSpiralSource<T>& spiral);

    void setStep(T step) { this->setAxialStep(step); }
    T step() const { return this->m_p.tableStep; }
    void setScanLenght(T scanLenght) override { this->setAxialScanLength(scanLenght); }
    std::uint64_t totalExposures() const override { return this->axialExposures(); }
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        auto twin = *this;
        return CTSource<T>::ctCalibration(twin, model, progressBar);
    }
};

template <Floating T = double>
class CTSpiralSource final : public CTSource<T> {
public:
    CTSpiralSource()
    {
        this->m_type = Source<T>::Type::CTSpiral;
        this->m_p.motion = model::GANTRY_SPIRAL;
        this->m_p.pitch = 1.0;
    }

    void setPitch(T pitch) { this->m_p.pitch = std::max(T { 0.01 }, pitch); }
    T pitch() const { return this->m_p.pitch; }
    void setScanLenght(T scanLenght) override { this->m_p.scanLength = std::max(std::abs(scanLenght), this->m_p.beamWidth * this->m_p.pitch * T { 0.5 }); }
    std::uint64_t totalExposures() const override { return this->spiralExposures(); }
    // CTDIvol of a spiral = CTDIw / pitch: calibrate the axial twin, then scale
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        CTAxialSource<T> twin = *this;
        return this->ctCalibration(twin, model, progressBar) * this->m_p.pitch;
    }
};

template <Floating T = double>
class CTAxialDualSource final : public CTDualSource<T> {
public:
    CTAxialDualSource()
    {
        this->m_type = Source<T>::Type::CTDual;
        this->m_p.tableStep = this->m_p.beamWidth;
        this->m_p.scanLength = this->m_p.tableStep;
    }
    CTAxialDualSource(const CTSpiralDualSource<T>& spiral);

    void setStep(T step) { this->setAxialStep(step); }
    T step() const { return this->m_p.tableStep; }
    void setScanLenght(T scanLenght) override { this->setAxialScanLength(scanLenght); }
    std::uint64_t totalExposures() const override { return this->axialExposures(); }
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        auto twin = *this;
        return this->ctCalibration(twin, model, progressBar);
    }
};

template <Floating T = double>
class CTSpiralDualSource final : public CTDualSource<T> {
public:
    CTSpiralDualSource()
    {
        this->m_type = Source<T>::Type::CTDual;
        this->m_p.motion = model::GANTRY_SPIRAL;
        this->m_p.pitch = 1.0;
    }

    std::uint64_t totalExposures() const override { return this->spiralExposures(); }
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        CTAxialDualSource<T> twin = *this;
        return CTSource<T>::ctCalibration(twin, model, progressBar) * this->m_p.pitch;
    }
    T pitch() const { return this->m_p.pitch; }
    void setPitch(T pitch) { this->m_p.pitch = std::max(T { 0.01 }, pitch); }
    void setScanLenght(T scanLenght) override
    {
        CTBaseSource<T>::setScanLenght(std::max(std::abs(scanLenght), this->collimation() * pitch() * T { 0.5 }));
    }
};

template <Floating T>
CTAxialSource<T>::CTAxialSource(const CTSpiralSource<T>& spiral)
    : CTSource<T>(spiral)
{
    this->becomeAxialTwin(spiral.scanLenght());
}

template <Floating T>
CTAxialDualSource<T>::CTAxialDualSource(const CTSpiralDualSource<T>& spiral)
    : CTDualSource<T>(spiral)
{
    this->becomeAxialTwin(spiral.scanLenght());
}

// scout view: the tube parked at the start angle while the table moves through the scan length, one exposure per millimetre
template <Floating T>
class CTTopogramSource : public CTBaseSource<T> {
public:
    CTTopogramSource()
    {
        this->m_type = Source<T>::Type::CTTopogram;
        this->m_p.motion = model::GANTRY_TOPOGRAM;
    }

    std::uint64_t totalExposures() const override { return std::max(static_cast<std::uint64_t>(std::ceil(this->scanLenght())), std::uint64_t { 1 }); }

    // calibrated through an axial scan of equal beam width whose CTDIvol is scaled by scan length / beam width
    T getCalibrationValue(LOWENERGYCORRECTION model, ProgressBar<T>* progressBar = nullptr) const override
    {
        CTAxialSource<T> twin;
        twin.adoptBaseOf(*this);
        twin.setCtdiVol(this->ctdiVol() * this->scanLenght() / this->collimation());
        twin.setScanLenght(0);
        twin.setStep(this->collimation());
        constexpr auto coarsest = (2 * PI_VAL<T>()) / 72;
        twin.setExposureAngleStep(std::min(2 * PI_VAL<T>() / totalExposures(), coarsest));
        const auto exposures = this->totalExposures();
        const auto factor = CTSource<T>::ctCalibration(twin, model, progressBar);
        return (factor * exposures) / twin.totalExposures();
    }
};
}
