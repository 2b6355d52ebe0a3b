// exposure.hpp — one static emitter of photon histories.
//
// Public surface of the reference's Exposure<T> (include/dxmc/exposure.hpp:36-304): position,
// direction cosines (beam = x cross y), rectangular collimation angles, beam weight, an optional
// spectrum / heel filter / fan filter, and a history count. On the B200 path an Exposure is pure
// description: Transport flattens it into a dxmcb200_exposure and the kernel's birth stage
// (csrc/physics.cuh sampleParticle) draws the photons. sampleParticle() here is the host
// equivalent, kept for API compatibility and tests.
#pragma once
#include "dxmc/beamfilters.hpp"
#include "dxmc/dxmcrandom.hpp"
#include "dxmc/sourcemodel.hpp"
#include "dxmc/types.hpp"
#include "dxmc/vectormath.hpp"

#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>

namespace dxmc {

template <Floating T = double>
class Exposure {
    using Vec3 = std::array<T, 3>;
    using Cosines = std::array<T, 6>; // x axis then y axis of the beam frame

public:
    // collimation: {x0, x1, y0, y1} half-plane angles, or {x, y} full opening angles (symmetric)
    template <std::size_t N>
        requires(N == 2 || N == 4)
    Exposure(const Vec3& position, const Cosines& directionCosines, const std::array<T, N>& collimationAngles, std::uint64_t nHistories = 1000,
        T beamIntensityWeight = 1, const SpecterDistribution<T>* specterDistribution = nullptr, const HeelFilter<T>* heelFilter = nullptr,
        const BeamFilter<T>* filter = nullptr)
        : m_nHistories(nHistories)
        , m_weight(beamIntensityWeight)
        , m_spectrum(specterDistribution)
        , m_heel(heelFilter)
        , m_fanFilter(filter)
        , m_origin(position)
    {
        setCollimationAngles(collimationAngles);
        setDirectionCosines(directionCosines);
    }

    // An exposure evaluated from a source's parameter block (dxmc/sourcemodel.hpp): the values are final (unit cosines, beam
    // direction), so nothing is normalised again; the tables are the source's objects the indices stand for.
    static Exposure fromModel(const model::ExposureValues<T>& v, const SpecterDistribution<T>* specterDistribution, const HeelFilter<T>* heelFilter,
        const BeamFilter<T>* filter)
    {
        Exposure e;
        e.m_nHistories = v.histories;
        e.m_weight = v.weight;
        e.m_monoEnergy = v.monoEnergy;
        e.m_spectrum = specterDistribution;
        e.m_heel = heelFilter;
        e.m_fanFilter = filter;
        for (std::size_t k = 0; k < 4; ++k)
            e.m_angles[k] = v.collimation[k];
        for (std::size_t k = 0; k < 3; ++k) {
            e.m_origin[k] = v.position[k];
            e.m_beam[k] = v.beam[k];
        }
        for (std::size_t k = 0; k < 6; ++k)
            e.m_frame[k] = v.cosines[k];
        return e;
    }

    // ---- the draw itself (host equivalent of the device birth stage): fan angle about the y cosine, cone angle
    // about the x cosine, energy from the spectrum (or the mono-energetic value), weight from the filters
    Particle<T> sampleParticle(RandomState& state) const noexcept
    {
        const T fan = state.randomUniform(m_angles[0], m_angles[1]);
        const T cone = state.randomUniform(m_angles[2], m_angles[3]);
        Particle<T> photon { .pos = m_origin, .dir = m_beam, .weight = m_weight };
        vectormath::rotate(photon.dir.data(), m_frame.data() + 3, fan);
        vectormath::rotate(photon.dir.data(), m_frame.data(), cone);
        photon.energy = m_spectrum ? m_spectrum->sampleValue(state) : m_monoEnergy;
        if (m_fanFilter)
            photon.weight *= m_fanFilter->sampleIntensityWeight(fan);
        if (m_heel)
            photon.weight *= m_heel->sampleIntensityWeight(cone, photon.energy);
        return photon;
    }

    // express position and orientation in the basis (x, y, x cross y) of a world
    void alignToDirectionCosines(const Cosines& worldCosines) noexcept
    {
        const T* ex = worldCosines.data();
        const T* ey = ex + 3;
        T ez[3];
        vectormath::cross(ex, ey, ez);
        for (T* v : { m_origin.data(), m_frame.data(), m_frame.data() + 3, m_beam.data() })
            vectormath::changeBasisInverse(ex, ey, ez, v);
    }

    // ---- how many, how strong, which tables (non-owning views; the source outlives its exposures)
    std::size_t numberOfHistories() const { return m_nHistories; }
    void setNumberOfHistories(std::size_t nHistories) { m_nHistories = nHistories; }
    T beamIntensityWeight() const { return m_weight; }
    void setBeamIntensityWeight(T weight) { m_weight = weight; }
    T monoenergeticPhotonEnergy() const { return m_monoEnergy; }
    void setMonoenergeticPhotonEnergy(T energy) { m_monoEnergy = std::clamp(energy, T { 0.0 }, T { 500.0 }); }
    const SpecterDistribution<T>* specterDistribution() const { return m_spectrum; }
    void setSpecterDistribution(const SpecterDistribution<T>* specter) { m_spectrum = specter; }
    const HeelFilter<T>* heelFilter() const { return m_heel; }
    void setHeelFilter(const HeelFilter<T>* filter) { m_heel = filter; }
    const BeamFilter<T>* beamFilter() const { return m_fanFilter; }
    void setBeamFilter(const BeamFilter<T>* filter) { m_fanFilter = filter; }

    // ---- collimation, stored as {x0, x1, y0, y1}
    const std::array<T, 4>& collimationAngles() const { return m_angles; }
    T collimationAngleX() const { return m_angles[1] - m_angles[0]; }
    T collimationAngleY() const { return m_angles[3] - m_angles[2]; }
    void setCollimationAngles(const std::array<T, 4>& angles) { m_angles = angles; }
    void setCollimationAngles(const T angleX, const T angleY) { m_angles = { -angleX / 2, angleX / 2, -angleY / 2, angleY / 2 }; }
    void setCollimationAngles(const std::array<T, 2>& angles) { setCollimationAngles(angles[0], angles[1]); }

    // ---- beam frame: two cosines, normalised on every change; the beam runs along their cross product
    const Cosines& directionCosines() const { return m_frame; }
    const Vec3& beamDirection() const { return m_beam; }
    void setDirectionCosines(const Cosines& cosines)
    {
        m_frame = cosines;
        vectormath::normalize(m_frame.data());
        vectormath::normalize(m_frame.data() + 3);
        vectormath::cross(m_frame.data(), m_beam.data());
    }
    void setDirectionCosines(const T cosines[6]) { setDirectionCosines(Cosines { cosines[0], cosines[1], cosines[2], cosines[3], cosines[4], cosines[5] }); }
    void setDirectionCosines(T x1, T x2, T x3, T y1, T y2, T y3) { setDirectionCosines(Cosines { x1, x2, x3, y1, y2, y3 }); }
    void setDirectionCosines(const Vec3& cosinesX, const Vec3& cosinesY)
    {
        setDirectionCosines(Cosines { cosinesX[0], cosinesX[1], cosinesX[2], cosinesY[0], cosinesY[1], cosinesY[2] });
    }

    // ---- focal spot
    const Vec3& position() const { return m_origin; }
    void setPosition(const Vec3& pos) { m_origin = pos; }
    void setPosition(const T pos[3]) { m_origin = { pos[0], pos[1], pos[2] }; }
    void setPosition(T x, T y, T z) { m_origin = { x, y, z }; }
    void setPositionZ(const T posZ) { m_origin[2] = posZ; }
    void addPosition(const Vec3& shift) { std::transform(m_origin.begin(), m_origin.end(), shift.begin(), m_origin.begin(), std::plus<T>()); }
    void subtractPosition(const Vec3& shift) { std::transform(m_origin.begin(), m_origin.end(), shift.begin(), m_origin.begin(), std::minus<T>()); }

private:
    Exposure() = default;

    std::uint64_t m_nHistories = 0;
    T m_weight = 1;
    T m_monoEnergy { 0 };
    const SpecterDistribution<T>* m_spectrum = nullptr;
    const HeelFilter<T>* m_heel = nullptr;
    const BeamFilter<T>* m_fanFilter = nullptr;
    std::array<T, 4> m_angles {};
    Vec3 m_origin {};
    Cosines m_frame {};
    Vec3 m_beam {};
};
}
