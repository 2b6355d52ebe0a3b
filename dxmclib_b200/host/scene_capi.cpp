// scene_capi.cpp — include/dxmcb200_scene.h implemented on the drop-in C++ classes in
// dxmclib_b200/include/dxmc/ (World, Material, AttenuationLut, sources, Transport). This is the
// FFI surface of the product: Python (tests/, bench.py) and any other host language reach the CUDA
// path through these calls. oracle/ref_harness.cpp implements the same header on the unmodified
// reference for comparison; nothing here touches oracle/.
#include "dxmcb200_scene.h"
#include "dxmcb200_scene_monitor.hpp"

#include "dxmc.hpp"
#include "dxmc/attenuationlut.hpp"

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>

using namespace dxmc;

struct dxs_scene {
    std::unique_ptr<World<float>> world = std::make_unique<World<float>>();
    CTDIPhantom<float>* ctdi = nullptr; // non-null when world is a CTDIPhantom
    std::unique_ptr<Source<float>> source;
    CTSource<float>* ct = nullptr;
    AttenuationLut<float> lut;
    bool lutValid = false;
    std::unique_ptr<Transport<float>> prepared; // dxs_b200_prepare .. dxs_b200_release
    std::vector<int> devices; // dxs_b200_set_devices: more than one entry spreads dxs_transport over several GPUs
};

namespace {
thread_local std::string g_lastError;

// exceptions never cross the C boundary; a device failure stays loud through its own status code
template <typename F>
int guarded(F f)
{
    try {
        return f();
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return g_lastError.rfind("dxmcb200:", 0) == 0 ? DXS_ERR_DEVICE : DXS_ERR_STATE;
    } catch (...) {
        g_lastError = "unknown exception";
        return DXS_ERR_STATE;
    }
}

void applyTube(Tube<float>& t, const dxs_tube& p)
{
    if (p.voltage > 0)
        t.setVoltage(p.voltage);
    if (p.anode_angle_deg > 0)
        t.setAnodeAngleDeg(p.anode_angle_deg);
    if (p.energy_resolution > 0)
        t.setEnergyResolution(p.energy_resolution);
    if (p.al_mm > 0)
        t.setAlFiltration(p.al_mm);
    if (p.cu_mm > 0)
        t.setCuFiltration(p.cu_mm);
    if (p.sn_mm > 0)
        t.setSnFiltration(p.sn_mm);
}

} // namespace

// ---- further source types (same text in dxmclib_b200/host/scene_capi.cpp and oracle/ref_harness.cpp: both class sets have
// the reference's API) -----------------------------------------------------------------------------------------------------
namespace {
template <typename S>
void applyCtGeometry(S& src, const dxs_ct_params* p)
{
    applyTube(src.tube(), p->tube);
    src.setPosition(p->position[0], p->position[1], p->position[2]);
    bool anyCos = false;
    for (int i = 0; i < 6; ++i)
        anyCos = anyCos || p->cosines[i] != 0;
    if (anyCos)
        src.setDirectionCosines({ p->cosines[0], p->cosines[1], p->cosines[2], p->cosines[3], p->cosines[4], p->cosines[5] });
    if (p->sdd > 0)
        src.setSourceDetectorDistance(p->sdd);
    if (p->collimation > 0)
        src.setCollimation(p->collimation);
    if (p->fov > 0)
        src.setFieldOfView(p->fov);
    src.setStartAngleDeg(p->start_angle_deg);
    src.setGantryTiltAngleDeg(p->gantry_tilt_deg);
    if (p->ctdi_vol > 0)
        src.setCtdiVol(p->ctdi_vol);
    if (p->ctdi_phantom_diameter > 0)
        src.setCtdiPhantomDiameter(p->ctdi_phantom_diameter);
    src.setModelHeelEffect(p->model_heel != 0);
    src.setHistoriesPerExposure(p->histories);
}
} // namespace

extern "C" {

const char* dxs_backend(void) { return "dxmc-b200"; }
const char* dxs_last_error(void) { return g_lastError.c_str(); }

dxs_scene* dxs_create(void) { return new (std::nothrow) dxs_scene; }
void dxs_destroy(dxs_scene* s) { delete s; }

int dxs_world_geometry(dxs_scene* s, const uint64_t dim[3], const float spacing[3], const float origin[3], const float cosines[6])
{
    if (!s || !dim || !spacing || !origin || !cosines)
        return DXS_ERR_ARG;
    return guarded([&] {
        s->world->setDimensions({ dim[0], dim[1], dim[2] });
        s->world->setSpacing({ spacing[0], spacing[1], spacing[2] });
        s->world->setOrigin({ origin[0], origin[1], origin[2] });
        s->world->setDirectionCosines({ cosines[0], cosines[1], cosines[2], cosines[3], cosines[4], cosines[5] });
        return DXS_OK;
    });
}

int dxs_world_add_material(dxs_scene* s, const char* name, double density)
{
    if (!s || !name)
        return DXS_ERR_ARG;
    return guarded([&] {
        Material m(name, "", density);
        return s->world->addMaterialToMap(m) ? DXS_OK : DXS_ERR_ARG;
    });
}

int dxs_world_add_element(dxs_scene* s, int Z)
{
    if (!s)
        return DXS_ERR_ARG;
    return guarded([&] {
        Material m(Z);
        return s->world->addMaterialToMap(m) ? DXS_OK : DXS_ERR_ARG;
    });
}

int dxs_world_arrays(dxs_scene* s, const float* density, const uint8_t* material, const uint8_t* measurement)
{
    if (!s || !density || !material)
        return DXS_ERR_ARG;
    return guarded([&] {
        const auto n = s->world->size();
        s->world->setDensityArray(std::make_shared<std::vector<float>>(density, density + n));
        s->world->setMaterialIndexArray(std::make_shared<std::vector<std::uint8_t>>(material, material + n));
        if (measurement)
            s->world->setMeasurementMapArray(std::make_shared<std::vector<std::uint8_t>>(measurement, measurement + n));
        return DXS_OK;
    });
}

int dxs_world_ctdi_phantom(dxs_scene* s, uint64_t diameter)
{
    if (!s)
        return DXS_ERR_ARG;
    return guarded([&] {
        auto p = std::make_unique<CTDIPhantom<float>>(diameter);
        s->ctdi = p.get();
        s->world = std::move(p);
        return DXS_OK;
    });
}

int dxs_world_validate(dxs_scene* s, int* valid)
{
    if (!s)
        return DXS_ERR_ARG;
    return guarded([&] {
        s->world->makeValid();
        if (valid)
            *valid = static_cast<const World<float>&>(*s->world).isValid() ? 1 : 0;
        return DXS_OK;
    });
}

int dxs_world_dimensions(dxs_scene* s, uint64_t dim[3], float spacing[3], float extent[6])
{
    if (!s)
        return DXS_ERR_ARG;
    for (int i = 0; i < 3; ++i) {
        if (dim)
            dim[i] = s->world->dimensions()[i];
        if (spacing)
            spacing[i] = s->world->spacing()[i];
    }
    if (extent)
        for (int i = 0; i < 6; ++i)
            extent[i] = s->world->matrixExtentSafe()[i];
    return DXS_OK;
}

int dxs_world_get_arrays(dxs_scene* s, float* density, uint8_t* material, uint8_t* measurement)
{
    if (!s)
        return DXS_ERR_ARG;
    const auto n = s->world->size();
    if (density && s->world->densityArray())
        std::memcpy(density, s->world->densityArray()->data(), n * sizeof(float));
    if (material && s->world->materialIndexArray())
        std::memcpy(material, s->world->materialIndexArray()->data(), n);
    if (measurement && s->world->measurementMapArray())
        std::memcpy(measurement, s->world->measurementMapArray()->data(), n);
    return DXS_OK;
}

int dxs_world_ctdi_holes(dxs_scene* s, int position, uint64_t* out, uint64_t* count)
{
    if (!s || !s->ctdi || position < 0 || position > 4)
        return DXS_ERR_ARG;
    using HP = CTDIPhantom<float>::HolePosition;
    static const HP map[5] = { HP::Center, HP::West, HP::East, HP::South, HP::North };
    const auto& idx = s->ctdi->holeIndices(map[position]);
    if (count)
        *count = idx.size();
    if (out)
        for (std::size_t i = 0; i < idx.size(); ++i)
            out[i] = idx[i];
    return DXS_OK;
}

int dxs_trace_indices(dxs_scene* s, uint64_t nRays, const float* pos, const float* dir, uint32_t nSteps, const float* steps, int64_t* outIdx,
    float* outEntry)
{
    if (!s || !pos || !dir || !outIdx || !outEntry || (nSteps && !steps))
        return DXS_ERR_ARG;
    return guarded([&] {
        s->world->makeValid();
        const World<float>& w = *s->world;
        if (!w.isValid())
            return static_cast<int>(DXS_ERR_STATE);
        dxmcb200_ctx* raw = nullptr;
        if (dxmcb200_create(dxmc::detail::currentDevice(), &raw) != DXMCB200_OK)
            throw std::runtime_error("dxmcb200: no usable CUDA device; this library has no CPU fallback");
        dxmc::detail::ContextPtr ctx(raw);
        dxmcb200_world dw {};
        for (int i = 0; i < 3; ++i) {
            dw.dim[i] = w.dimensions()[i];
            dw.spacing[i] = w.spacing()[i];
        }
        for (int i = 0; i < 6; ++i)
            dw.extent_safe[i] = w.matrixExtentSafe()[i];
        dw.density = w.densityArray()->data();
        dw.material = w.materialIndexArray()->data();
        dw.measurement = w.measurementMapArray() ? w.measurementMapArray()->data() : nullptr;
        dxmc::detail::check(ctx.get(), dxmcb200_set_world(ctx.get(), &dw), "set_world");
        dxmc::detail::check(ctx.get(), dxmcb200_trace_indices(ctx.get(), nRays, pos, dir, nSteps, steps, outIdx, outEntry), "trace_indices");
        return static_cast<int>(DXS_OK);
    });
}

static const Material* materialAt(dxs_scene* s, int idx)
{
    if (!s || idx < 0 || static_cast<std::size_t>(idx) >= s->world->materialMap().size())
        return nullptr;
    return &s->world->materialMap()[idx];
}

int dxs_material_attenuation(dxs_scene* s, int idx, double e, double out[4])
{
    const Material* m = materialAt(s, idx);
    if (!m || !out)
        return DXS_ERR_ARG;
    out[0] = m->getPhotoelectricAttenuation(e);
    out[1] = m->getComptonAttenuation(e);
    out[2] = m->getRayleightAttenuation(e);
    out[3] = m->getTotalAttenuation(e);
    return DXS_OK;
}

int dxs_material_form_factor_sq(dxs_scene* s, int idx, double q, double* out)
{
    const Material* m = materialAt(s, idx);
    if (!m || !out)
        return DXS_ERR_ARG;
    *out = m->getRayleightFormFactorSquared(q);
    return DXS_OK;
}

int dxs_material_scatter_factor(dxs_scene* s, int idx, double q, double* out)
{
    const Material* m = materialAt(s, idx);
    if (!m || !out)
        return DXS_ERR_ARG;
    *out = m->getComptonNormalizedScatterFactor(q);
    return DXS_OK;
}

int dxs_material_binding_energies(dxs_scene* s, int idx, double minValue, double* out, int* count)
{
    const Material* m = materialAt(s, idx);
    if (!m)
        return DXS_ERR_ARG;
    const auto e = m->getBindingEnergies(minValue);
    if (count)
        *count = static_cast<int>(e.size());
    if (out)
        std::copy(e.begin(), e.end(), out);
    return DXS_OK;
}

int dxs_material_shells(dxs_scene* s, int idx, double out[12 * 13])
{
    const Material* m = materialAt(s, idx);
    if (!m || !out)
        return DXS_ERR_ARG;
    const auto conf = m->getElectronConfiguration();
    for (int i = 0; i < 12; ++i) {
        double* o = out + i * 13;
        const auto& c = conf[i];
        o[0] = c.bindingEnergy;
        o[1] = c.numberElectrons;
        o[2] = c.hartreeFockOrbital_0;
        o[3] = c.photoIonizationProbability;
        o[4] = c.fluorescenceYield;
        for (int k = 0; k < 3; ++k) {
            o[5 + k] = c.fluorLineProbabilities[k];
            o[8 + k] = c.fluorLineEnergies[k];
        }
        o[11] = c.Z;
        o[12] = c.shell;
    }
    return DXS_OK;
}

int dxs_material_density(dxs_scene* s, int idx, double* out)
{
    const Material* m = materialAt(s, idx);
    if (!m || !out)
        return DXS_ERR_ARG;
    *out = m->standardDensity();
    return DXS_OK;
}

int dxs_lut_generate(dxs_scene* s, float maxEnergy)
{
    if (!s)
        return DXS_ERR_ARG;
    return guarded([&] {
        s->world->makeValid();
        if (!static_cast<const World<float>&>(*s->world).isValid())
            return static_cast<int>(DXS_ERR_STATE);
        s->lut = AttenuationLut<float>();
        s->lut.generate(*s->world, maxEnergy);
        s->lutValid = true;
        return static_cast<int>(DXS_OK);
    });
}

int dxs_lut_attenuation(dxs_scene* s, int material, float energy, float out[3])
{
    if (!s || !s->lutValid || !out)
        return DXS_ERR_STATE;
    const auto a = s->lut.photoComptRayAttenuation(material, energy);
    out[0] = a[0];
    out[1] = a[1];
    out[2] = a[2];
    return DXS_OK;
}

int dxs_lut_max_inverse(dxs_scene* s, float energy, float* out)
{
    if (!s || !s->lutValid || !out)
        return DXS_ERR_STATE;
    *out = s->lut.maxTotalAttenuationInverse(energy);
    return DXS_OK;
}

int dxs_lut_scatter_factor(dxs_scene* s, int material, float q, float* out)
{
    if (!s || !s->lutValid || !out)
        return DXS_ERR_STATE;
    *out = s->lut.comptonScatterFactor(material, q);
    return DXS_OK;
}

int dxs_lut_sample_form_factor(dxs_scene* s, int material, float qmaxSq, uint64_t seed[2], int n, float* out)
{
    if (!s || !s->lutValid || !out || !seed)
        return DXS_ERR_STATE;
    RandomState state(seed);
    for (int i = 0; i < n; ++i)
        out[i] = s->lut.momentumTransferFromFormFactor(material, qmaxSq, state);
    seed[0] = state.m_state[0];
    seed[1] = state.m_state[1];
    return DXS_OK;
}

int dxs_lut_table(dxs_scene* s, int what, float* out, uint64_t* count)
{
    if (!s)
        return DXS_ERR_STATE;
    // what >= 16: the tables of the prepared Transport (dxs_b200_prepare), which builds its majorant from the device's
    // per-material density maxima; below 16: the scene's own AttenuationLut (dxs_lut_generate, host scan)
    const bool fromTransport = what >= 16;
    if (fromTransport ? !s->prepared : !s->lutValid)
        return DXS_ERR_STATE;
    const AttenuationLut<float>& lut = fromTransport ? s->prepared->attenuationLut() : s->lut;
    what &= 15;
    std::vector<float> v;
    const auto& ip = lut.attenuationData();
    switch (what) {
    case 0:
        v = ip.knots();
        break;
    case 1:
        v = ip.coefficients();
        break;
    case 2:
        v = ip.maxCoefficients();
        break;
    case 3:
        v = { static_cast<float>(ip.linearIndex()), ip.linearStep(), ip.linearEnergy(), static_cast<float>(ip.resolution()) };
        break;
    case 4:
        for (const auto& r : lut.formFactorSamplers()) {
            v.insert(v.end(), r.x().begin(), r.x().end());
            v.insert(v.end(), r.e().begin(), r.e().end());
            v.insert(v.end(), r.a().begin(), r.a().end());
            v.insert(v.end(), r.b().begin(), r.b().end());
        }
        break;
    case 5:
        for (const auto& c : lut.scatterFunctions()) {
            v.insert(v.end(), c.coefficients().begin(), c.coefficients().end());
            v.insert(v.end(), c.knots().begin(), c.knots().end());
            v.push_back(c.step());
            v.push_back(c.start());
            v.push_back(c.stop());
        }
        break;
    default:
        return DXS_ERR_ARG;
    }
    if (count)
        *count = v.size();
    if (out)
        std::copy(v.begin(), v.end(), out);
    return DXS_OK;
}

int dxs_source_pencil(dxs_scene* s, const float pos[3], const float cosines[6], float energy, uint64_t histories, uint64_t exposures)
{
    if (!s || !pos || !cosines)
        return DXS_ERR_ARG;
    return guarded([&] {
        auto src = std::make_unique<PencilSource<float>>();
        src->setPosition(pos[0], pos[1], pos[2]);
        src->setDirectionCosines({ cosines[0], cosines[1], cosines[2], cosines[3], cosines[4], cosines[5] });
        src->setPhotonEnergy(energy);
        src->setHistoriesPerExposure(histories);
        src->setTotalExposures(exposures);
        s->ct = nullptr;
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_isotropic(dxs_scene* s, int ct, const float pos[3], const float cosines[6], const float coll[4], int n,
    const float* weights, const float* energies, uint64_t histories, uint64_t exposures)
{
    if (!s || !pos || !cosines || !coll || n < 1 || !weights || !energies)
        return DXS_ERR_ARG;
    return guarded([&] {
        std::unique_ptr<IsotropicSource<float>> src;
        if (ct)
            src = std::make_unique<IsotropicCTSource<float>>();
        else
            src = std::make_unique<IsotropicSource<float>>();
        src->setPosition(pos[0], pos[1], pos[2]);
        src->setDirectionCosines({ cosines[0], cosines[1], cosines[2], cosines[3], cosines[4], cosines[5] });
        src->setCollimationAngles(coll[0], coll[1], coll[2], coll[3]);
        src->setSpecter(std::vector<float>(weights, weights + n), std::vector<float>(energies, energies + n));
        src->setHistoriesPerExposure(histories);
        src->setTotalExposures(exposures);
        s->ct = nullptr;
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_dx(dxs_scene* s, const dxs_dx_params* p)
{
    if (!s || !p)
        return DXS_ERR_ARG;
    return guarded([&] {
        auto src = std::make_unique<DXSource<float>>();
        applyTube(src->tube(), p->tube);
        src->setPosition(p->position[0], p->position[1], p->position[2]);
        if (p->sdd > 0)
            src->setSourceDetectorDistance(p->sdd);
        if (p->field_size[0] > 0 && p->field_size[1] > 0)
            src->setFieldSize({ p->field_size[0], p->field_size[1] });
        src->setTubeRotationDeg(p->tube_rotation_deg);
        src->setSourceAnglesDeg(p->source_angles_deg[0], p->source_angles_deg[1]);
        if (p->dap > 0)
            src->setDap(p->dap);
        src->setModelHeelEffect(p->model_heel != 0);
        src->setHistoriesPerExposure(p->histories);
        src->setTotalExposures(p->exposures);
        s->ct = nullptr;
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_ct(dxs_scene* s, const dxs_ct_params* p)
{
    if (!s || !p)
        return DXS_ERR_ARG;
    return guarded([&] {
        std::unique_ptr<CTSource<float>> src;
        CTSpiralSource<float>* spiral = nullptr;
        CTAxialSource<float>* axial = nullptr;
        if (p->spiral) {
            auto sp = std::make_unique<CTSpiralSource<float>>();
            spiral = sp.get();
            src = std::move(sp);
        } else {
            auto ax = std::make_unique<CTAxialSource<float>>();
            axial = ax.get();
            src = std::move(ax);
        }
        applyTube(src->tube(), p->tube);
        src->setPosition(p->position[0], p->position[1], p->position[2]);
        bool anyCos = false;
        for (int i = 0; i < 6; ++i)
            anyCos = anyCos || p->cosines[i] != 0;
        if (anyCos)
            src->setDirectionCosines({ p->cosines[0], p->cosines[1], p->cosines[2], p->cosines[3], p->cosines[4], p->cosines[5] });
        if (p->sdd > 0)
            src->setSourceDetectorDistance(p->sdd);
        if (p->collimation > 0)
            src->setCollimation(p->collimation);
        if (p->fov > 0)
            src->setFieldOfView(p->fov);
        src->setStartAngleDeg(p->start_angle_deg);
        if (p->exposure_step_deg > 0)
            src->setExposureAngleStepDeg(p->exposure_step_deg);
        src->setGantryTiltAngleDeg(p->gantry_tilt_deg);
        if (spiral) {
            if (p->pitch > 0)
                spiral->setPitch(p->pitch);
        } else {
            if (p->step > 0)
                axial->setStep(p->step);
            else
                axial->setStep(src->collimation());
        }
        if (p->scan_length > 0)
            src->setScanLenght(p->scan_length);
        if (p->ctdi_vol > 0)
            src->setCtdiVol(p->ctdi_vol);
        if (p->ctdi_phantom_diameter > 0)
            src->setCtdiPhantomDiameter(p->ctdi_phantom_diameter);
        src->setModelHeelEffect(p->model_heel != 0);
        src->setUseXCareFilter(p->use_xcare != 0);
        if (p->use_xcare) {
            auto& x = src->xcareFilter();
            x.setFilterAngleDeg(p->xcare_filter_angle_deg);
            if (p->xcare_span_deg > 0)
                x.setSpanAngleDeg(p->xcare_span_deg);
            if (p->xcare_ramp_deg > 0)
                x.setRampAngleDeg(p->xcare_ramp_deg);
            if (p->xcare_low_weight > 0)
                x.setLowWeight(p->xcare_low_weight);
        }
        src->setHistoriesPerExposure(p->histories);
        s->ct = src.get();
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_ct_dual(dxs_scene* s, const dxs_ct_dual_params* d)
{
    if (!s || !d)
        return DXS_ERR_ARG;
    return guarded([&] {
        const dxs_ct_params* p = &d->a;
        std::unique_ptr<CTDualSource<float>> src;
        CTSpiralDualSource<float>* spiral = nullptr;
        CTAxialDualSource<float>* axial = nullptr;
        if (p->spiral) {
            auto sp = std::make_unique<CTSpiralDualSource<float>>();
            spiral = sp.get();
            src = std::move(sp);
        } else {
            auto ax = std::make_unique<CTAxialDualSource<float>>();
            axial = ax.get();
            src = std::move(ax);
        }
        applyCtGeometry(*src, p);
        applyTube(src->tubeB(), d->tube_b);
        if (d->sdd_b > 0)
            src->setSourceDetectorDistanceB(d->sdd_b);
        if (d->fov_b > 0)
            src->setFieldOfViewB(d->fov_b);
        src->setStartAngleDegB(d->start_angle_b_deg);
        if (d->mas_a > 0)
            src->setTubeAmas(d->mas_a);
        if (d->mas_b > 0)
            src->setTubeBmas(d->mas_b);
        if (p->exposure_step_deg > 0)
            src->setExposureAngleStepDeg(p->exposure_step_deg);
        if (spiral) {
            if (p->pitch > 0)
                spiral->setPitch(p->pitch);
        } else {
            axial->setStep(p->step > 0 ? p->step : src->collimation());
        }
        if (p->scan_length > 0)
            src->setScanLenght(p->scan_length);
        src->setUseXCareFilter(p->use_xcare != 0);
        if (p->use_xcare) {
            auto& x = src->xcareFilter();
            x.setFilterAngleDeg(p->xcare_filter_angle_deg);
            if (p->xcare_span_deg > 0)
                x.setSpanAngleDeg(p->xcare_span_deg);
            if (p->xcare_ramp_deg > 0)
                x.setRampAngleDeg(p->xcare_ramp_deg);
            if (p->xcare_low_weight > 0)
                x.setLowWeight(p->xcare_low_weight);
        }
        s->ct = src.get();
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_topogram(dxs_scene* s, const dxs_ct_params* p)
{
    if (!s || !p)
        return DXS_ERR_ARG;
    return guarded([&] {
        auto src = std::make_unique<CTTopogramSource<float>>();
        applyCtGeometry(*src, p);
        if (p->scan_length > 0)
            src->setScanLenght(p->scan_length);
        s->ct = nullptr;
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_cbct(dxs_scene* s, const dxs_cbct_params* c)
{
    if (!s || !c)
        return DXS_ERR_ARG;
    return guarded([&] {
        const dxs_dx_params* p = &c->dx;
        auto src = std::make_unique<CBCTSource<float>>();
        applyTube(src->tube(), p->tube);
        src->setPosition(p->position[0], p->position[1], p->position[2]);
        if (p->sdd > 0)
            src->setSourceDetectorDistance(p->sdd);
        if (p->field_size[0] > 0 && p->field_size[1] > 0)
            src->setFieldSize({ p->field_size[0], p->field_size[1] });
        src->setTubeRotationDeg(p->tube_rotation_deg);
        src->setSourceAnglesDeg(p->source_angles_deg[0], p->source_angles_deg[1]);
        if (p->dap > 0)
            src->setDap(p->dap);
        src->setModelHeelEffect(p->model_heel != 0);
        src->setHistoriesPerExposure(p->histories);
        if (c->step_deg > 0)
            src->setStepAngleDeg(c->step_deg);
        if (c->span_deg > 0)
            src->setSpanAngleDeg(c->span_deg);
        s->ct = nullptr;
        s->source = std::move(src);
        return DXS_OK;
    });
}

int dxs_source_bowtie(dxs_scene* s, int n, const float* angles, const float* weights)
{
    if (!s || !s->ct || n < 2 || !angles || !weights)
        return DXS_ERR_ARG;
    return guarded([&] {
        s->ct->setBowTieFilter(std::make_shared<BowTieFilter<float>>(std::vector<float>(angles, angles + n), std::vector<float>(weights, weights + n)));
        return DXS_OK;
    });
}

int dxs_source_aec(dxs_scene* s, int n, const float* profile)
{
    if (!s || !s->ct || n < 1 || !profile)
        return DXS_ERR_ARG;
    return guarded([&] {
        auto dens = s->world->densityArray();
        if (!dens)
            return static_cast<int>(DXS_ERR_STATE);
        s->ct->setAecFilter(std::make_shared<AECFilter<float>>(dens, s->world->spacing(), s->world->dimensions(), std::vector<float>(profile, profile + n)));
        return static_cast<int>(DXS_OK);
    });
}

int dxs_source_total_exposures(dxs_scene* s, uint64_t* n)
{
    if (!s || !s->source || !n)
        return DXS_ERR_STATE;
    *n = s->source->totalExposures();
    return DXS_OK;
}

int dxs_source_max_energy(dxs_scene* s, float* e)
{
    if (!s || !s->source || !e)
        return DXS_ERR_STATE;
    *e = s->source->maxPhotonEnergyProduced();
    return DXS_OK;
}

int dxs_source_exposure(dxs_scene* s, uint64_t i, dxs_exposure* out)
{
    if (!s || !s->source || !out)
        return DXS_ERR_STATE;
    return guarded([&] {
        s->world->makeValid();
        s->source->updateFromWorld(*s->world);
        s->source->validate();
        auto e = s->source->getExposure(i);
        e.alignToDirectionCosines(s->world->directionCosines());
        for (int k = 0; k < 3; ++k) {
            out->position[k] = e.position()[k];
            out->beam_direction[k] = e.beamDirection()[k];
        }
        for (int k = 0; k < 6; ++k)
            out->cosines[k] = e.directionCosines()[k];
        for (int k = 0; k < 4; ++k)
            out->collimation[k] = e.collimationAngles()[k];
        out->weight = e.beamIntensityWeight();
        out->mono_energy = e.monoenergeticPhotonEnergy();
        out->has_spectrum = e.specterDistribution() != nullptr;
        out->has_heel = e.heelFilter() != nullptr;
        out->has_bowtie = e.beamFilter() != nullptr;
        out->histories = e.numberOfHistories();
        return DXS_OK;
    });
}

int dxs_source_table(dxs_scene* s, int what, float* out, uint64_t* count)
{
    if (!s || !s->source || what < 0 || what > 6)
        return DXS_ERR_STATE;
    return guarded([&] {
        s->world->makeValid();
        s->source->updateFromWorld(*s->world);
        s->source->validate();
        const auto e = s->source->getExposure(0);
        std::vector<float> v;
        if (what <= 2) {
            if (const auto* sp = e.specterDistribution()) {
                if (what == 0)
                    v = sp->probabilityData();
                else if (what == 1)
                    v.assign(sp->aliasingData().begin(), sp->aliasingData().end());
                else
                    v = sp->energies();
            }
        } else if (what <= 4) {
            if (const auto* h = e.heelFilter()) {
                if (what == 3)
                    v = { h->energyStart(), h->energyStep(), static_cast<float>(h->energySize()), h->angleStart(), h->angleStep(),
                        static_cast<float>(h->angleSize()) };
                else
                    v = h->weights();
            }
        } else if (const auto* b = dynamic_cast<const BowTieFilter<float>*>(e.beamFilter())) {
            for (const auto& [angle, weight] : b->data())
                v.push_back(what == 5 ? angle : weight);
        }
        if (count)
            *count = v.size();
        if (out)
            std::copy(v.begin(), v.end(), out);
        return DXS_OK;
    });
}

int dxs_source_spectrum(dxs_scene* s, float* energies, float* weights, int* count)
{
    if (!s || !s->source)
        return DXS_ERR_STATE;
    return guarded([&] {
        std::vector<float> e, w;
        if (auto* ct = dynamic_cast<CTBaseSource<float>*>(s->source.get())) {
            e = ct->tube().getEnergy();
            w = ct->tube().getSpecter(e);
        } else if (auto* dx = dynamic_cast<DAPSource<float>*>(s->source.get())) {
            e = dx->tube().getEnergy();
            w = dx->tube().getSpecter(e);
        } else {
            return static_cast<int>(DXS_ERR_UNSUPPORTED);
        }
        if (count)
            *count = static_cast<int>(e.size());
        if (energies)
            std::copy(e.begin(), e.end(), energies);
        if (weights)
            std::copy(w.begin(), w.end(), weights);
        return static_cast<int>(DXS_OK);
    });
}

int dxs_source_calibration(dxs_scene* s, int model, float* out)
{
    if (!s || !s->source || !out)
        return DXS_ERR_STATE;
    return guarded([&] {
        s->source->validate();
        *out = s->source->getCalibrationValue(static_cast<LOWENERGYCORRECTION>(model), nullptr);
        return DXS_OK;
    });
}

int dxs_transport(dxs_scene* s, int model, int outputMode, int useCalibration, uint64_t seed, int nWorkers,
    float* dose, uint32_t* nEvents, float* variance, dxs_result_info* info)
{
    if (!s || !s->source)
        return DXS_ERR_STATE;
    return guarded([&] {
        Transport<float> tr;
        if (nWorkers > 0)
            tr.setNumberOfWorkers(nWorkers);
        tr.setLowEnergyCorrectionModel(static_cast<LOWENERGYCORRECTION>(model));
        tr.setOutputMode(outputMode == DXS_OUT_DOSE ? Transport<float>::OUTPUTMODE::DOSE : Transport<float>::OUTPUTMODE::EV_PER_HISTORY);
        detail::PhaseTrace trace;
        s->world->makeValid();
        trace("World::makeValid");
        if (seed != 0)
            tr.setSeed(seed);
        if (!s->devices.empty())
            tr.setDevices(s->devices);
        (void)nWorkers; // n_workers selects host threads / stream mode in the reference harness only
        if (s->devices.empty()) {
            // One GPU: the phases of Transport::operator() (prepare / run / collect, transport.hpp) with the result decoded and
            // downloaded STRAIGHT into the caller's arrays. operator() itself has to return a Result by value, which an FFI caller
            // would then copy once more (2.5 GB of fresh pages at 512x512x400: 0.05 s, and the source of 0.4 s outliers).
            const std::size_t n = s->world->size();
            auto fill = [&](auto* a) {
                if (a)
                    std::fill(a, a + n, 0);
            };
            std::uint64_t histories = 0;
            double seconds = 0;
            std::string units = outputMode == DXS_OUT_DOSE ? (useCalibration ? "mGy" : "keV/kg") : "eV/history";
            if (tr.prepare(*s->world, s->source.get())) {
                trace("prepare");
                tr.run<World<float>>(0, tr.preparedExposures());
                trace("run");
                units = std::string(tr.collectInto(*s->world, s->source.get(), 0, dose, nEvents, variance, useCalibration != 0));
                histories = tr.preparedHistories();
                seconds = tr.lastRunTime().count();
                tr.release();
                trace("collect into caller arrays");
            } else { // invalid world or source: the reference's all-zero Result (transport.hpp:142-151)
                fill(dose);
                fill(nEvents);
                fill(variance);
            }
            if (info) {
                info->histories = histories;
                info->seconds = seconds;
                std::memset(info->units, 0, sizeof(info->units));
                std::strncpy(info->units, units.c_str(), sizeof(info->units) - 1);
            }
            return DXS_OK;
        }
        Result<float> res = tr(*s->world, s->source.get(), nullptr, useCalibration != 0);
        trace("Transport::operator()");
        const auto n = res.dose.size();
        // the caller's arrays are usually untouched pages: copy (and fault them in) on several host threads
        const void* src[3] = { res.dose.data(), res.nEvents.data(), res.variance.data() };
        void* dst[3] = { dose, nEvents, variance };
        const std::size_t pieces = n > (std::size_t { 1 } << 22) ? 4 : 1;
        dxmc::detail::parallelFor(3 * pieces, [&](std::size_t job) {
            const std::size_t a = job / pieces, piece = job % pieces;
            if (!dst[a])
                return;
            const std::size_t begin = n * piece / pieces * 4, end = n * (piece + 1) / pieces * 4;
            std::memcpy(static_cast<char*>(dst[a]) + begin, static_cast<const char*>(src[a]) + begin, end - begin);
        });
        if (info) {
            info->histories = res.numberOfHistories;
            info->seconds = res.simulationTime.count();
            std::memset(info->units, 0, sizeof(info->units));
            std::strncpy(info->units, std::string(res.dose_units).c_str(), sizeof(info->units) - 1);
        }
        trace("copy to caller arrays");
        return DXS_OK;
    });
}

int dxs_transport_monitored(dxs_scene* s, int model, int outputMode, int useCalibration, uint64_t seed, int nWorkers, double cancelAtPercent,
    float* dose, uint32_t* nEvents, float* variance, dxs_result_info* info, dxs_progress_report* report)
{
    if (!s || !s->source)
        return DXS_ERR_STATE;
    return guarded([&] {
        Transport<float> tr;
        if (nWorkers > 0)
            tr.setNumberOfWorkers(nWorkers);
        tr.setLowEnergyCorrectionModel(static_cast<LOWENERGYCORRECTION>(model));
        tr.setOutputMode(outputMode == DXS_OUT_DOSE ? Transport<float>::OUTPUTMODE::DOSE : Transport<float>::OUTPUTMODE::EV_PER_HISTORY);
        s->world->makeValid();
        if (seed != 0)
            tr.setSeed(seed);
        if (!s->devices.empty())
            tr.setDevices(s->devices);
        Result<float> res = dxs_monitor::run<Result<float>, ProgressBar<float>>(tr, *s->world, s->source.get(), useCalibration != 0, cancelAtPercent, report);
        const auto n = res.dose.size();
        if (dose)
            std::memcpy(dose, res.dose.data(), n * sizeof(float));
        if (nEvents)
            std::memcpy(nEvents, res.nEvents.data(), n * sizeof(std::uint32_t));
        if (variance)
            std::memcpy(variance, res.variance.data(), n * sizeof(float));
        if (info) {
            info->histories = res.numberOfHistories;
            info->seconds = res.simulationTime.count();
            std::memset(info->units, 0, sizeof(info->units));
            std::strncpy(info->units, std::string(res.dose_units).c_str(), sizeof(info->units) - 1);
        }
        return DXS_OK;
    });
}

int dxs_b200_prepare(dxs_scene* s, int device, int model, uint64_t seed, uint64_t totalHistoriesAllRanks)
{
    if (!s || !s->source)
        return DXS_ERR_STATE;
    return guarded([&] {
        s->prepared = std::make_unique<Transport<float>>();
        s->prepared->setDevice(device);
        s->prepared->setLowEnergyCorrectionModel(static_cast<LOWENERGYCORRECTION>(model));
        if (seed != 0)
            s->prepared->setSeed(seed);
        s->world->makeValid();
        if (!s->prepared->prepare(*s->world, s->source.get(), totalHistoriesAllRanks)) {
            s->prepared.reset();
            return static_cast<int>(DXS_ERR_STATE);
        }
        return static_cast<int>(DXS_OK);
    });
}

int dxs_b200_run(dxs_scene* s, uint64_t expBegin, uint64_t expEnd, double* kernelMs)
{
    if (!s || !s->prepared)
        return DXS_ERR_STATE;
    return guarded([&] {
        s->prepared->run(expBegin, expEnd);
        if (kernelMs)
            dxmcb200_last_run_ms(s->prepared->context(), kernelMs);
        return DXS_OK;
    });
}

int dxs_b200_run_strided(dxs_scene* s, uint64_t expFirst, uint64_t expStride, uint64_t expCount, double* kernelMs)
{
    if (!s || !s->prepared)
        return DXS_ERR_STATE;
    return guarded([&] {
        s->prepared->runStrided(expFirst, expStride, expCount);
        if (kernelMs)
            dxmcb200_last_run_ms(s->prepared->context(), kernelMs);
        return DXS_OK;
    });
}

int dxs_b200_collect(dxs_scene* s, int outputMode, int useCalibration, uint64_t histories, float* dose, uint32_t* nEvents, float* variance,
    dxs_result_info* info)
{
    if (!s || !s->prepared)
        return DXS_ERR_STATE;
    return guarded([&] {
        auto& tr = *s->prepared;
        tr.setOutputMode(outputMode == DXS_OUT_DOSE ? Transport<float>::OUTPUTMODE::DOSE : Transport<float>::OUTPUTMODE::EV_PER_HISTORY);
        // straight into the caller's arrays: the download's host threads fault the (usually untouched) pages in
        const std::uint64_t n = histories ? histories : tr.preparedHistories();
        const std::string_view units = tr.collectInto(*s->world, s->source.get(), n, dose, nEvents, variance, useCalibration != 0);
        if (info) {
            info->histories = n;
            info->seconds = 0;
            std::memset(info->units, 0, sizeof(info->units));
            std::strncpy(info->units, std::string(units).c_str(), sizeof(info->units) - 1);
        }
        return DXS_OK;
    });
}

int dxs_b200_context(dxs_scene* s, void** ctx)
{
    if (!s || !s->prepared || !ctx)
        return DXS_ERR_STATE;
    *ctx = s->prepared->context();
    return DXS_OK;
}

int dxs_b200_set_devices(dxs_scene* s, int n, const int* devices)
{
    if (!s || n < 0 || (n > 0 && !devices))
        return DXS_ERR_ARG;
    s->devices.assign(devices, devices + n);
    return DXS_OK;
}

int dxs_b200_release(dxs_scene* s)
{
    if (!s)
        return DXS_ERR_ARG;
    s->prepared.reset();
    return DXS_OK;
}

} // extern "C"
