// ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Implements include/dxmcb200_scene.h on top of the UNMODIFIED reference headers in
// /root/reference/include (never copied into this repo) and the reference's own
// src/material.cpp, compiled where they lie by oracle/Makefile into
// oracle/_ref/libdxmc_ref.so. xraylib, which the reference needs and this image lacks, is
// supplied by oracle/xraylib_compat/xraylib.h -> dxmclib_b200/host/xrl_lite (the same data
// source the product uses), so both implementations see identical cross sections.
//
// Built with -fno-access-control so protected/private members of the reference classes
// (Transport::transport<L>, AttenuationLutInterpolator::m_coefficients, ...) can be driven
// with a seeded RandomState and dumped for bit-exact comparison.
#include "dxmcb200_scene.h"
#include "dxmcb200_scene_monitor.hpp"

#include "dxmc.hpp"

#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>

using namespace dxmc;

// Portability shim (NOT a change of behaviour): BowTieFilter<T>::normalizeData (reference
// beamfilters.hpp:88-92) calls std::reduce over pair<T,T> elements with a (double, pair) folding
// lambda. MSVC's STL, on which the reference is developed, accepts that and folds left to right;
// libstdc++ rejects it with a static_assert, so the member cannot be instantiated with g++. This
// explicit specialisation supplies the same left fold with std::accumulate.
template <>
void dxmc::BowTieFilter<float>::normalizeData()
{
    const auto mean = std::accumulate(m_data.begin(), m_data.end(), 0.0, [](auto a, auto el) { return a + el.second; }) / m_data.size();
    std::transform(m_data.begin(), m_data.end(), m_data.begin(), [=](const auto& el) { return std::make_pair(el.first, static_cast<float>(el.second / mean)); });
}

struct dxs_scene {
    std::unique_ptr<World<float>> world = std::make_unique<World<float>>();
    CTDIPhantom<float>* ctdi = nullptr; // non-null when world is a CTDIPhantom
    std::unique_ptr<Source<float>> source;
    CTSource<float>* ct = nullptr;
    AttenuationLut<float> lut;
    bool lutValid = false;
};

namespace {
template <typename F>
int guarded(F f)
{
    try {
        return f();
    } catch (...) {
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

template <int L>
void seededRun(Transport<float>& tr, const World<float>& w, const Source<float>* src, Result<float>& res, std::uint64_t seed)
{
    std::uint64_t s[2] = { seed, seed ^ 0x9E3779B97F4A7C15ULL };
    RandomState state(s);
    const auto& basis = w.directionCosines();
    const auto n = src->totalExposures();
    for (std::uint64_t i = 0; i < n; ++i) {
        auto exposure = src->getExposure(i);
        exposure.alignToDirectionCosines(basis);
        tr.template transport<L>(w, exposure, state, res);
    }
}
// The product's per-history stream keying (include/dxmcb200.h dxmcb200_history_stream), restated here so
// the reference's own sampling code can be driven with the very same random numbers as the kernels.
inline std::uint64_t mix64(std::uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline void historyStream(std::uint64_t seed, std::uint64_t exposure, std::uint64_t history, std::uint64_t out[2])
{
    const std::uint64_t golden = 0x9E3779B97F4A7C15ULL;
    const std::uint64_t s = mix64(seed + golden * (exposure + 1));
    out[0] = mix64(s + golden * (history + 1));
    out[1] = mix64(out[0] + golden) | 1ULL;
}

// Reference transport with one RandomState per history, seeded like the B200 kernels: the body of
// Transport::transport<L> (transport.hpp:729-743) with the state re-created per history. Exposures are
// spread over host threads; scoring uses the reference's own atomic adds.
template <int L>
void counterStreamRun(Transport<float>& tr, const World<float>& w, const Source<float>* src, Result<float>& res, std::uint64_t seed, unsigned nThreads)
{
    const auto& basis = w.directionCosines();
    const auto n = src->totalExposures();
    std::atomic<std::uint64_t> next { 0 };
    auto worker = [&]() {
        for (std::uint64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
            auto exposure = src->getExposure(i);
            exposure.alignToDirectionCosines(basis);
            const auto nHist = exposure.numberOfHistories();
            for (std::uint64_t h = 0; h < nHist; ++h) {
                std::uint64_t s[2];
                historyStream(seed, i, h, s);
                RandomState state(s);
                auto particle = exposure.sampleParticle(state);
                if (tr.transportParticleToWorld(w, particle))
                    tr.template woodcockParticleTracking<L>(w, particle, state, res);
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < std::max(nThreads, 1u); ++t)
        pool.emplace_back(worker);
    worker();
    for (auto& t : pool)
        t.join();
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

const char* dxs_backend(void) { return "dxmclib-reference"; }
const char* dxs_last_error(void) { return ""; }

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
        Transport<float> tr; // only its geometry members are used
        for (uint64_t r = 0; r < nRays; ++r) {
            Particle<float> p { { pos[3 * r], pos[3 * r + 1], pos[3 * r + 2] }, { dir[3 * r], dir[3 * r + 1], dir[3 * r + 2] }, 1.0f, 1.0f };
            bool in = tr.transportParticleToWorld(w, p);
            for (int k = 0; k < 3; ++k)
                outEntry[3 * r + k] = p.pos[k];
            int64_t* out = outIdx + r * (nSteps + 1);
            out[0] = (in && tr.particleInsideWorld(w, p)) ? static_cast<int64_t>(tr.indexFromPosition(p, w)) : -1;
            for (uint32_t k = 0; k < nSteps; ++k) {
                if (in) {
                    for (std::size_t i = 0; i < 3; i++)
                        p.pos[i] += p.dir[i] * steps[k];
                    in = tr.particleInsideWorld(w, p);
                }
                out[k + 1] = in ? static_cast<int64_t>(tr.indexFromPosition(p, w)) : -1;
            }
        }
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
    if (!s || !s->lutValid)
        return DXS_ERR_STATE;
    std::vector<float> v;
    const auto& ip = s->lut.m_attenuationData;
    switch (what) {
    case 0:
        v = ip.m_x;
        break;
    case 1:
        v = ip.m_coefficients;
        break;
    case 2:
        v = ip.m_maxCoefficients;
        break;
    case 3:
        v = { static_cast<float>(ip.m_linearIndex), ip.m_linearStep, ip.m_linearEnergy, static_cast<float>(ip.m_resolution) };
        break;
    case 4:
        for (const auto& r : s->lut.m_formFactor) {
            v.insert(v.end(), r.m_x.begin(), r.m_x.end());
            v.insert(v.end(), r.m_e.begin(), r.m_e.end());
            v.insert(v.end(), r.m_a.begin(), r.m_a.end());
            v.insert(v.end(), r.m_b.begin(), r.m_b.end());
        }
        break;
    case 5:
        for (const auto& c : s->lut.m_comptonScatterFactor) {
            v.insert(v.end(), c.m_coefficients.begin(), c.m_coefficients.end());
            v.insert(v.end(), c.m_x.begin(), c.m_x.end());
            v.push_back(c.m_step);
            v.push_back(c.m_start);
            v.push_back(c.m_stop);
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
        out->mono_energy = e.m_monoenergeticPhotonEnergy;
        out->has_spectrum = e.m_specterDistribution != nullptr;
        out->has_heel = e.m_heelFilter != nullptr;
        out->has_bowtie = e.m_beamFilter != nullptr;
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
            if (const auto* sp = e.m_specterDistribution) {
                if (what == 0)
                    v = sp->probabilityData();
                else if (what == 1)
                    v.assign(sp->aliasingData().begin(), sp->aliasingData().end());
                else
                    v = sp->m_energies;
            }
        } else if (what <= 4) {
            if (const auto* h = e.m_heelFilter) {
                if (what == 3)
                    v = { h->m_energyStart, h->m_energyStep, static_cast<float>(h->m_energySize), h->m_angleStart, h->m_angleStep,
                        static_cast<float>(h->m_angleSize) };
                else
                    v = h->m_weights;
            }
        } else if (const auto* b = dynamic_cast<const BowTieFilter<float>*>(e.m_beamFilter)) {
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
        s->world->makeValid();
        Result<float> res;
        const bool counterStreams = nWorkers == DXS_WORKERS_COUNTER_STREAMS;
        if (seed == 0 && !counterStreams) {
            res = tr(*s->world, s->source.get(), nullptr, useCalibration != 0);
        } else {
            // the body of Transport::operator() (transport.hpp:138-201) with the worker pool replaced
            // by one seeded worker
            const World<float>& w = *s->world;
            res = Result<float>(w.size());
            auto* src = s->source.get();
            if (w.isValid()) {
                src->updateFromWorld(w);
                src->validate();
                if (src->isValid()) {
                    res.numberOfHistories = src->historiesPerExposure() * src->totalExposures();
                    tr.m_attenuationLut.generate(w, src->maxPhotonEnergyProduced());
                    const auto t0 = std::chrono::system_clock::now();
                    if (counterStreams) {
                        const unsigned nt = std::max(std::thread::hardware_concurrency(), 1u);
                        if (model == 0)
                            counterStreamRun<0>(tr, w, src, res, seed, nt);
                        else if (model == 1)
                            counterStreamRun<1>(tr, w, src, res, seed, nt);
                        else
                            counterStreamRun<2>(tr, w, src, res, seed, nt);
                    } else if (model == 0)
                        seededRun<0>(tr, w, src, res, seed);
                    else if (model == 1)
                        seededRun<1>(tr, w, src, res, seed);
                    else
                        seededRun<2>(tr, w, src, res, seed);
                    res.simulationTime = std::chrono::system_clock::now() - t0;
                    if (outputMode == DXS_OUT_DOSE) {
                        if (useCalibration) {
                            const float cal = src->getCalibrationValue(static_cast<LOWENERGYCORRECTION>(model), nullptr);
                            tr.energyImpartedToDose(w, res, cal);
                            res.dose_units = "mGy";
                        } else {
                            tr.energyImpartedToDose(w, res);
                            res.dose_units = "keV/kg";
                        }
                    } else {
                        tr.normalizeScoring(res);
                        res.dose_units = "eV/history";
                    }
                }
            }
        }
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
        (void)seed; // the stock operator() seeds its workers from std::random_device (transport.hpp:749)
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

// B200 extensions do not exist in the reference
int dxs_b200_prepare(dxs_scene*, int, int, uint64_t, uint64_t) { return DXS_ERR_UNSUPPORTED; }
int dxs_b200_run(dxs_scene*, uint64_t, uint64_t, double*) { return DXS_ERR_UNSUPPORTED; }
int dxs_b200_run_strided(dxs_scene*, uint64_t, uint64_t, uint64_t, double*) { return DXS_ERR_UNSUPPORTED; }
int dxs_b200_collect(dxs_scene*, int, int, uint64_t, float*, uint32_t*, float*, dxs_result_info*) { return DXS_ERR_UNSUPPORTED; }
int dxs_b200_context(dxs_scene*, void**) { return DXS_ERR_UNSUPPORTED; }
int dxs_b200_release(dxs_scene*) { return DXS_ERR_UNSUPPORTED; }
int dxs_b200_set_devices(dxs_scene*, int, const int*) { return DXS_ERR_UNSUPPORTED; }

} // extern "C"
