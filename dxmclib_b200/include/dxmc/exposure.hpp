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
#include "dxmc/floating.hpp"
#include "dxmc/particle.hpp"
#include "dxmc/vectormath.hpp"

#include <array>
#include <cstdint>

namespace dxmc {

template <Floating T = double>
class Exposure {
public:
    Exposure(const std::array<T, 3>& position, const std::array<T, 6>& directionCosines, const std::array<T, 4>& collimationAngles,
        std::uint64_t nHistories = 1000, T beamIntensityWeight = 1, const SpecterDistribution<T>* specterDistribution = nullptr,
        const HeelFilter<T>* heelFilter = nullptr, const BeamFilter<T>* filter = nullptr)
        : m_position(position)
        , m_directionCosines(directionCosines)
        , m_collimationAngles(collimationAngles)
        , m_beamIntensityWeight(beamIntensityWeight)
        , m_beamFilter(filter)
        , m_specterDistribution(specterDistribution)
        , m_heelFilter(heelFilter)
        , m_nHistories(nHistories)
    {
        normalizeCosines();
    }
    // symmetric collimation: full opening angles {x, y}
    Exposure(const std::array<T, 3>& position, const std::array<T, 6>& directionCosines, const std::array<T, 2>& collimationAngles,
        std::uint64_t nHistories = 1000, T beamIntensityWeight = 1, const SpecterDistribution<T>* specterDistribution = nullptr,
        const HeelFilter<T>* heelFilter = nullptr, const BeamFilter<T>* filter = nullptr)
        : m_position(position)
        , m_directionCosines(directionCosines)
        , m_beamIntensityWeight(beamIntensityWeight)
        , m_beamFilter(filter)
        , m_specterDistribution(specterDistribution)
        , m_heelFilter(heelFilter)
        , m_nHistories(nHistories)
    {
        setCollimationAngles(collimationAngles);
        normalizeCosines();
    }

    void setPosition(T x, T y, T z) { m_position = { x, y, z }; }
    void setPosition(const T pos[3]) { m_position = { pos[0], pos[1], pos[2] }; }
    void setPosition(const std::array<T, 3>& pos) { m_position = pos; }
    void setPositionZ(const T posZ) { m_position[2] = posZ; }
    const std::array<T, 3>& position() const { return m_position; }
    void addPosition(const std::array<T, 3>& pos)
    {
        for (std::size_t i = 0; i < 3; ++i)
            m_position[i] += pos[i];
    }
    void subtractPosition(const std::array<T, 3>& pos)
    {
        for (std::size_t i = 0; i < 3; ++i)
            m_position[i] -= pos[i];
    }

    void setDirectionCosines(T x1, T x2, T x3, T y1, T y2, T y3)
    {
        m_directionCosines = { x1, x2, x3, y1, y2, y3 };
        normalizeCosines();
    }
    void setDirectionCosines(const T cosines[6])
    {
        for (std::size_t i = 0; i < 6; ++i)
            m_directionCosines[i] = cosines[i];
        normalizeCosines();
    }
    void setDirectionCosines(const std::array<T, 6>& cosines)
    {
        m_directionCosines = cosines;
        normalizeCosines();
    }
    void setDirectionCosines(const std::array<T, 3>& cosinesX, const std::array<T, 3>& cosinesY)
    {
        for (std::size_t i = 0; i < 3; ++i) {
            m_directionCosines[i] = cosinesX[i];
            m_directionCosines[i + 3] = cosinesY[i];
        }
        normalizeCosines();
    }
    const std::array<T, 6>& directionCosines() const { return m_directionCosines; }
    const std::array<T, 3>& beamDirection() const { return m_beamDirection; }

    void setCollimationAngles(const std::array<T, 4>& angles) { m_collimationAngles = angles; }
    void setCollimationAngles(const std::array<T, 2>& angles) { setCollimationAngles(angles[0], angles[1]); }
    void setCollimationAngles(const T angleX, const T angleY) { m_collimationAngles = { -angleX / 2, angleX / 2, -angleY / 2, angleY / 2 }; }
    const std::array<T, 4>& collimationAngles() const { return m_collimationAngles; } // x0 x1 y0 y1
    T collimationAngleX() const { return m_collimationAngles[1] - m_collimationAngles[0]; }
    T collimationAngleY() const { return m_collimationAngles[3] - m_collimationAngles[2]; }

    void setBeamIntensityWeight(T weight) { m_beamIntensityWeight = weight; }
    T beamIntensityWeight() const { return m_beamIntensityWeight; }

    void setBeamFilter(const BeamFilter<T>* filter) { m_beamFilter = filter; }
    void setSpecterDistribution(const SpecterDistribution<T>* specter) { m_specterDistribution = specter; }
    void setHeelFilter(const HeelFilter<T>* filter) { m_heelFilter = filter; }
    // non-owning views used when the exposure is flattened for the device
    const BeamFilter<T>* beamFilter() const { return m_beamFilter; }
    const SpecterDistribution<T>* specterDistribution() const { return m_specterDistribution; }
    const HeelFilter<T>* heelFilter() const { return m_heelFilter; }

    void setMonoenergeticPhotonEnergy(T energy) { m_monoenergeticPhotonEnergy = std::clamp(energy, T { 0.0 }, T { 500.0 }); }
    T monoenergeticPhotonEnergy() const { return m_monoenergeticPhotonEnergy; }

    void setNumberOfHistories(std::size_t nHistories) { m_nHistories = nHistories; }
    std::size_t numberOfHistories() const { return m_nHistories; }

    // express position and orientation in the basis (x, y, x cross y) of a world
    void alignToDirectionCosines(const std::array<T, 6>& directionCosines) noexcept
    {
        const T* b1 = directionCosines.data();
        const T* b2 = b1 + 3;
        T b3[3];
        vectormath::cross(b1, b2, b3);
        vectormath::changeBasisInverse(b1, b2, b3, m_position.data());
        vectormath::changeBasisInverse(b1, b2, b3, m_directionCosines.data());
        vectormath::changeBasisInverse(b1, b2, b3, m_directionCosines.data() + 3);
        vectormath::changeBasisInverse(b1, b2, b3, m_beamDirection.data());
    }

    // host-side photon draw: fan angle about the y cosine, cone angle about the x cosine
    Particle<T> sampleParticle(RandomState& state) const noexcept
    {
        const T theta = state.randomUniform(m_collimationAngles[0], m_collimationAngles[1]);
        const T phi = state.randomUniform(m_collimationAngles[2], m_collimationAngles[3]);
        Particle<T> p { .pos = m_position, .dir = m_beamDirection, .weight = m_beamIntensityWeight };
        vectormath::rotate(p.dir.data(), &m_directionCosines[3], theta);
        vectormath::rotate(p.dir.data(), &m_directionCosines[0], phi);
        p.energy = m_specterDistribution ? m_specterDistribution->sampleValue(state) : m_monoenergeticPhotonEnergy;
        if (m_beamFilter)
            p.weight *= m_beamFilter->sampleIntensityWeight(theta);
        if (m_heelFilter)
            p.weight *= m_heelFilter->sampleIntensityWeight(phi, p.energy);
        return p;
    }

protected:
    void normalizeCosines()
    {
        vectormath::normalize(&m_directionCosines[0]);
        vectormath::normalize(&m_directionCosines[3]);
        vectormath::cross(m_directionCosines.data(), m_beamDirection.data());
    }

private:
    std::array<T, 3> m_position;
    std::array<T, 6> m_directionCosines;
    std::array<T, 3> m_beamDirection;
    std::array<T, 4> m_collimationAngles; // x0 x1 y0 y1
    T m_beamIntensityWeight;
    const BeamFilter<T>* m_beamFilter = nullptr;
    const SpecterDistribution<T>* m_specterDistribution = nullptr;
    const HeelFilter<T>* m_heelFilter = nullptr;
    T m_monoenergeticPhotonEnergy { 0 };
    std::uint64_t m_nHistories;
};
}
