// beamfilters.hpp — photon-weight filters of an x-ray beam.
//
// Public surface of the reference's include/dxmc/beamfilters.hpp:
//   BeamFilter   abstract weight(angle)                                           :42-69
//   BowTieFilter measured fan-angle fluence profile (per photon, on device)      :78-199
//   XCareFilter  organ-based tube current modulation (per exposure, on host)     :207-426
//   HeelFilter   anode heel effect over (angle, energy) (per photon, on device)  :436-570
//   AECFilter    tube-current profile along z from slice masses (per exposure)   :577-766
// The per-photon filters expose their tables; Transport uploads them as dxmcb200_bowtie /
// dxmcb200_heel and the kernels evaluate them in csrc/physics.cuh.
#pragma once
#include "dxmc/constants.hpp"
#include "dxmc/floating.hpp"
#include "dxmc/interpolation.hpp"
#include "dxmc/tube.hpp"
#include "dxmc/world.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <execution>
#include <memory>
#include <numeric>
#include <string>
#include <utility>
#include <vector>

namespace dxmc {

template <Floating T = double>
class BeamFilter {
public:
    virtual ~BeamFilter() = default;
    // photon weight for a fan angle [rad]; averages to one over the beam
    virtual T sampleIntensityWeight(const T angle) const = 0;
    const std::string& filterName() const { return m_filterName; }
    void setFilterName(const std::string& name) { m_filterName = name; }

private:
    std::string m_filterName;
};

template <Floating T = double>
class BowTieFilter : public BeamFilter<T> {
public:
    // symmetric profile: only |angle| is used; weights are normalised to mean one
    BowTieFilter(const std::vector<T>& angles, const std::vector<T>& weights)
    {
        if (angles.size() == weights.size()) {
            m_data.resize(angles.size());
            for (std::size_t i = 0; i < angles.size(); ++i)
                m_data[i] = { std::abs(angles[i]), weights[i] };
        }
        finish();
    }
    BowTieFilter(const std::vector<std::pair<T, T>>& angleWeightsPairs)
        : m_data(angleWeightsPairs)
    {
        for (auto& p : m_data)
            p.first = std::abs(p.first);
        finish();
    }
    BowTieFilter(const BowTieFilter& other)
        : BeamFilter<T>()
        , m_data(other.m_data)
    {
    }

    T sampleIntensityWeight(T anglePlusAndMinus) const override
    {
        const T angle = std::abs(anglePlusAndMinus);
        // bisection down to one bracketing interval [lo, hi]
        std::size_t lo = 0, hi = m_data.size() - 1;
        for (std::size_t mid = lo + (hi - lo) / 2; mid != lo; mid = lo + (hi - lo) / 2) {
            if (angle < m_data[mid].first)
                hi = mid;
            else
                lo = mid;
        }
        if (angle < m_data[lo].first)
            return m_data.front().second;
        if (angle > m_data[hi].first)
            return m_data.back().second;
        const auto& [x0, y0] = m_data[lo];
        const auto& [x1, y1] = m_data[hi];
        return y0 + (angle - x0) * (y1 - y0) / (x1 - x0);
    }

    const std::vector<std::pair<T, T>>& data() const { return m_data; }

protected:
    void finish()
    {
        std::sort(m_data.begin(), m_data.end());
        // mean accumulated in double, left to right
        double sum = 0.0;
        for (const auto& p : m_data)
            sum = sum + p.second;
        const double mean = sum / m_data.size();
        for (auto& p : m_data)
            p.second = static_cast<T>(p.second / mean);
    }

private:
    std::vector<std::pair<T, T>> m_data;
};

template <Floating T = double>
class XCareFilter : public BeamFilter<T> {
public:
    XCareFilter()
        : m_filterAngle(0)
        , m_spanAngle(120 * DEG_TO_RAD<T>())
        , m_rampAngle(20 * DEG_TO_RAD<T>())
        , m_lowWeight(T { 0.6 })
    {
    }

    T filterAngle() const { return m_filterAngle; }
    T filterAngleDeg() const { return m_filterAngle * RAD_TO_DEG<T>(); }
    void setFilterAngle(T angle)
    {
        constexpr T twoPi = T { 2 } * PI_VAL<T>();
        m_filterAngle = std::fmod(angle, twoPi);
        if (m_filterAngle < 0.0)
            m_filterAngle += twoPi;
    }
    void setFilterAngleDeg(T angle) { setFilterAngle(angle * DEG_TO_RAD<T>()); }

    T spanAngle() const { return m_spanAngle; }
    T spanAngleDeg() const { return m_spanAngle * RAD_TO_DEG<T>(); }
    void setSpanAngle(T angle)
    {
        constexpr T smallest = T { 5.0 } * DEG_TO_RAD<T>();
        if (angle > smallest && angle < PI_VAL<T>())
            m_spanAngle = angle;
    }
    void setSpanAngleDeg(T angle) { setSpanAngle(angle * DEG_TO_RAD<T>()); }

    T rampAngle() const { return m_rampAngle; }
    T rampAngleDeg() const { return m_rampAngle * RAD_TO_DEG<T>(); }
    void setRampAngle(T angle)
    {
        if (angle >= 0.0 && angle <= 0.5 * m_spanAngle)
            m_rampAngle = angle;
    }
    void setRampAngleDeg(T angle) { setRampAngle(angle * DEG_TO_RAD<T>()); }

    T lowWeight() const { return m_lowWeight; }
    void setLowWeight(T weight)
    {
        if (weight > 0.0 && weight <= 1.0)
            m_lowWeight = weight;
    }
    // weight outside the protected span such that the mean over a rotation is one
    T highWeight() const
    {
        constexpr T twoPi = T { 2 } * PI_VAL<T>();
        return (twoPi - m_spanAngle * m_lowWeight + m_lowWeight * m_rampAngle) / (twoPi - m_spanAngle + m_rampAngle);
    }

    T sampleIntensityWeight(const T angle) const override
    {
        constexpr T twoPi = T { 2 } * PI_VAL<T>();
        T a = std::fmod(angle - m_filterAngle + PI_VAL<T>(), twoPi); // protected span centred on pi
        if (a < 0)
            a += twoPi;
        const T high = highWeight();
        const T spanStart = PI_VAL<T>() - m_spanAngle * T { 0.5 };
        if (a < spanStart)
            return high;
        const T rampDownEnd = spanStart + m_rampAngle;
        if (a < rampDownEnd)
            return interp<T>(spanStart, rampDownEnd, high, m_lowWeight, a);
        const T rampUpStart = rampDownEnd + m_spanAngle - m_rampAngle;
        if (a < rampUpStart)
            return m_lowWeight;
        const T spanEnd = spanStart + m_spanAngle;
        if (a < spanEnd)
            return interp<T>(rampUpStart, spanEnd, m_lowWeight, high, a);
        return high;
    }

private:
    T m_filterAngle, m_spanAngle, m_rampAngle, m_lowWeight;
};

template <Floating T = double>
class HeelFilter {
public:
    HeelFilter(const Tube<T>& tube, const T heel_angle_span = 0.0) { update(tube, heel_angle_span); }

    // relative fluence over 5 take-off angles x (kV-10)/2 energies from the tube model, every energy row
    // normalised to mean one
    void update(const Tube<T>& tube, const T heel_angle_span = 0.0)
    {
        m_energySize = std::max(static_cast<std::size_t>((tube.voltage() - m_energyStart) / m_energyStep), std::size_t { 2 });
        m_energies.clear();
        m_energies.reserve(m_energySize);
        for (std::size_t i = 0; i < m_energySize; ++i)
            m_energies.push_back(m_energyStart + i * m_energyStep);

        const T span = std::abs(heel_angle_span);
        m_angleStart = tube.anodeAngle() < span / 2 ? -tube.anodeAngle() : -span / 2;
        m_angleStep = (span / 2 - m_angleStart) / m_angleSize;

        m_weights.assign(m_energySize * m_angleSize, T { 0 });
        for (std::size_t i = 0; i < m_angleSize; ++i) {
            const T angle = m_angleStart + i * m_angleStep + tube.anodeAngle();
            const auto specter = tube.getSpecter(m_energies, angle, false);
            for (std::size_t j = 0; j < m_energySize; ++j)
                m_weights[j * m_angleSize + i] = specter[j];
        }
        for (std::size_t j = 0; j < m_energySize; ++j) {
            T* row = &m_weights[j * m_angleSize];
            // same library reduction as the reference so the rounding of the row mean is identical
            const T mean = std::reduce(row, row + m_angleSize, 0.0) / m_angleSize;
            for (std::size_t i = 0; i < m_angleSize; ++i)
                row[i] = mean > T { 0 } ? row[i] / mean : T { 1 };
        }
    }

    // nearest energy row, linear in angle
    T sampleIntensityWeight(const T angle, const T energy) const
    {
        std::size_t e = static_cast<std::size_t>((energy - m_energyStart + T { 0.5 } * m_energyStep) / m_energyStep);
        if (e >= m_energySize)
            e = m_energySize - 1;
        if (energy < m_energyStart)
            e = 0;
        std::size_t a = static_cast<std::size_t>((angle - m_angleStart) / m_angleStep);
        if (a >= m_angleSize)
            a = m_angleSize - 1;
        if (angle < m_angleStart)
            a = 0;
        const std::size_t w = e * m_angleSize + a;
        if (a < m_angleSize - 1) {
            const T a0 = m_angleStart + m_angleStep * a;
            const T a1 = m_angleStart + m_angleStep * (a + 1);
            return interp(a0, a1, m_weights[w], m_weights[w + 1], angle);
        }
        return m_weights[w];
    }

    std::size_t energySize() const { return m_energySize; }
    std::size_t angleSize() const { return m_angleSize; }
    const std::vector<T>& weights() const { return m_weights; }
    T energyStart() const { return m_energyStart; }
    T energyStep() const { return m_energyStep; }
    T angleStart() const { return m_angleStart; }
    T angleStep() const { return m_angleStep; }
    const std::string& filterName() const { return m_filterName; }
    void setFilterName(const std::string& name) { m_filterName = name; }

private:
    T m_energyStep = 2.0;
    T m_energyStart = 10.0;
    std::size_t m_energySize = 65;
    T m_angleStep = 0.07;
    T m_angleStart = 0.07;
    std::size_t m_angleSize = 5;
    std::vector<T> m_energies;
    std::vector<T> m_weights; // [energy][angle]
    std::string m_filterName;
};

template <Floating T = double>
class AECFilter {
public:
    // exposure profile along z (one value per slice) matched to the slice masses of a density volume
    AECFilter(const std::vector<T>& densityImage, const std::array<T, 3> spacing, const std::array<std::size_t, 3> dimensions,
        const std::vector<T>& exposuremapping)
    {
        buildMassTable(densityImage.cbegin(), densityImage.cend(), spacing, dimensions, exposuremapping);
    }
    AECFilter(std::shared_ptr<std::vector<T>>& densityImage, const std::array<T, 3> spacing, const std::array<std::size_t, 3> dimensions,
        const std::vector<T>& exposuremapping)
    {
        buildMassTable(densityImage->cbegin(), densityImage->cend(), spacing, dimensions, exposuremapping);
    }
    // direct (slice mass, intensity) table; needs updateFromWorld before use
    AECFilter(const std::vector<T>& mass, const std::vector<T>& intensity)
        : m_mass(mass)
        , m_massIntensity(intensity)
    {
        m_positionIntensity.assign(1, T { 1.0 });
        m_positionStep = 1.0;
        m_positionMax = 1.0;
        m_valid = false;
    }

    T sampleIntensityWeight(const std::array<T, 3>& position) const
    {
        const T p = position[2];
        if (p < m_positionMin + m_positionStep)
            return m_positionIntensity[0];
        if (p >= m_positionMax - m_positionStep)
            return m_positionIntensity.back();
        const std::size_t ind = static_cast<std::size_t>((p - m_positionMin) / m_positionStep);
        const T x0 = ind * m_positionStep + m_positionMin;
        const T x1 = x0 + m_positionStep;
        return interp(x0, x1, m_positionIntensity[ind], m_positionIntensity[ind + 1], p);
    }

    void updateFromWorld(const World<T>& world)
    {
        buildPositionTable(world.densityArray()->cbegin(), world.densityArray()->cend(), world.spacing(), world.dimensions(), world.origin());
    }
    bool isValid() const { return m_valid; }
    // the table sampleIntensityWeight interpolates in: one intensity per slice between positionMin and positionMax
    const std::vector<T>& positionIntensity() const { return m_positionIntensity; }
    T positionMin() const { return m_positionMin; }
    T positionMax() const { return m_positionMax; }
    T positionStep() const { return m_positionStep; }
    const std::vector<T>& mass() const { return m_mass; }
    const std::vector<T>& massIntensity() const { return m_massIntensity; }
    const std::string& filterName() const { return m_filterName; }
    void setFilterName(const std::string& name) { m_filterName = name; }

protected:
    using DensIt = typename std::vector<T>::const_iterator;

    void buildMassTable(DensIt densBeg, DensIt densEnd, const std::array<T, 3>& spacing, const std::array<std::size_t, 3>& dim,
        const std::vector<T>& exposure)
    {
        m_valid = false;
        if (static_cast<std::size_t>(std::distance(densBeg, densEnd)) != dim[0] * dim[1] * dim[2] || exposure.size() != dim[2]) {
            m_mass.resize(1);
            m_massIntensity.assign(1, T { 1 });
            return;
        }
        const auto slice = dim[0] * dim[1];
        const auto voxelArea = spacing[0] * spacing[1];
        std::vector<std::pair<T, T>> massExposure(dim[2]);
        // one slice per task on all host cores; each slice mass is still accumulated in double by the same library
        // reduction the reference calls, so the table has the reference's bits
        detail::parallelFor(dim[2], [&](std::size_t k) {
            const auto sum = std::reduce(std::execution::par_unseq, densBeg + slice * k, densBeg + slice * (k + 1), 0.0);
            massExposure[k] = { static_cast<T>(sum * voxelArea), exposure[k] };
        });
        std::sort(massExposure.begin(), massExposure.end());
        const T mean = std::reduce(std::execution::par_unseq, exposure.cbegin(), exposure.cend(), T { 0.0 }) / dim[2];
        m_mass.resize(dim[2]);
        m_massIntensity.resize(dim[2]);
        for (std::size_t k = 0; k < dim[2]; ++k) {
            m_mass[k] = massExposure[k].first;
            m_massIntensity[k] = massExposure[k].second / mean;
        }
        buildPositionTable(densBeg, densEnd, spacing, dim, { 0, 0, 0 });
    }

    void buildPositionTable(DensIt densBeg, DensIt densEnd, const std::array<T, 3> spacing, const std::array<std::size_t, 3>& dim,
        const std::array<T, 3>& origin)
    {
        m_positionMin = origin[2] - spacing[2] * dim[2] * T { 0.5 };
        m_positionMax = m_positionMin + spacing[2] * dim[2];
        m_positionStep = (m_positionMax - m_positionMin) / dim[2];
        m_positionIntensity.resize(dim[2]);
        const auto slice = dim[0] * dim[1];
        const auto voxelArea = spacing[0] * spacing[1];
        detail::parallelFor(dim[2], [&](std::size_t k) {
            const T mass = std::reduce(std::execution::par_unseq, densBeg + slice * k, densBeg + slice * (k + 1), T { 0.0 }) * voxelArea;
            m_positionIntensity[k] = intensityForMass(mass);
        });
        m_valid = true;
    }

    T intensityForMass(T mass) const
    {
        const auto pos = std::upper_bound(m_mass.cbegin(), m_mass.cend(), mass);
        if (pos == m_mass.cbegin())
            return m_massIntensity.front();
        if (mass > m_mass.back() || pos == m_mass.cend())
            return m_massIntensity.back();
        const auto i = std::distance(m_mass.cbegin(), pos);
        return interp(m_mass[i - 1], m_mass[i], m_massIntensity[i - 1], m_massIntensity[i], mass);
    }

private:
    std::vector<T> m_mass;
    std::vector<T> m_massIntensity;
    T m_positionStep = 0.0;
    T m_positionMin = 0.0;
    T m_positionMax = 0.0;
    std::vector<T> m_positionIntensity;
    std::string m_filterName;
    bool m_valid = false;
};
}
