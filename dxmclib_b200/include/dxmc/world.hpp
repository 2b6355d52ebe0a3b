// world.hpp — voxel geometry container and the CTDI body/head phantom.
//
// Same public surface and numerical conventions as the reference (include/dxmc/world.hpp:37-390):
// x-fastest voxel arrays shared through shared_ptr, extents centred on the origin, "safe" extents
// pulled one ulp inwards so that a position strictly inside them always maps to a valid voxel.
// The arrays are what Transport uploads to the GPU (packed into one record per voxel there).
#pragma once
#include "dxmc/floating.hpp"
#include "dxmc/material.hpp"
#include "dxmc/vectormath.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

namespace dxmc {

template <Floating T = double>
class World {
public:
    World() { refreshExtent(); }
    virtual ~World() = default;

    void setDimensions(const std::array<std::size_t, 3>& dimensions)
    {
        m_valid = false;
        m_dimensions = dimensions;
        refreshExtent();
    }
    void setSpacing(const std::array<T, 3>& spacing)
    {
        m_valid = false;
        m_spacing = spacing;
        refreshExtent();
    }
    void setOrigin(const std::array<T, 3>& origin)
    {
        m_valid = false;
        m_origin = origin;
        refreshExtent();
    }
    void setDirectionCosines(const std::array<T, 6>& cosines)
    {
        m_valid = false;
        m_directionCosines = cosines;
        vectormath::normalize(m_directionCosines.data());
        vectormath::normalize(m_directionCosines.data() + 3);
    }

    std::size_t size() const { return m_dimensions[0] * m_dimensions[1] * m_dimensions[2]; }
    const std::array<std::size_t, 3>& dimensions() const { return m_dimensions; }
    const std::array<T, 3>& spacing() const { return m_spacing; }
    const std::array<T, 3>& origin() const { return m_origin; }
    const std::array<T, 6>& directionCosines() const { return m_directionCosines; }
    std::array<T, 3> depthDirection() const
    {
        std::array<T, 3> d;
        vectormath::cross(m_directionCosines.data(), d.data());
        return d;
    }

    void setDensityArray(std::shared_ptr<std::vector<T>> a)
    {
        m_valid = false;
        m_density = a;
    }
    std::shared_ptr<std::vector<T>> densityArray() { return m_density; }
    const std::shared_ptr<std::vector<T>> densityArray() const { return m_density; }

    void setMaterialIndexArray(std::shared_ptr<std::vector<std::uint8_t>> a)
    {
        m_valid = false;
        m_materialIndex = a;
    }
    std::shared_ptr<std::vector<std::uint8_t>> materialIndexArray() { return m_materialIndex; }
    const std::shared_ptr<std::vector<std::uint8_t>> materialIndexArray() const { return m_materialIndex; }

    // voxels flagged non-zero use forced photo-electric scoring in Transport
    void setMeasurementMapArray(std::shared_ptr<std::vector<std::uint8_t>> a)
    {
        m_valid = false;
        m_measurementMap = a;
    }
    std::shared_ptr<std::vector<std::uint8_t>> measurementMapArray() { return m_measurementMap; }
    const std::shared_ptr<std::vector<std::uint8_t>> measurementMapArray() const { return m_measurementMap; }

    const std::vector<Material>& materialMap() const { return m_materialMap; }
    bool addMaterialToMap(const Material& material)
    {
        m_valid = false;
        if (!material.isValid())
            return false;
        m_materialMap.push_back(material);
        return true;
    }
    void clearMaterialMap()
    {
        m_valid = false;
        m_materialMap.clear();
    }

    const std::array<T, 6>& matrixExtent() const { return m_extent; }
    const std::array<T, 6>& matrixExtentSafe() const { return m_extentSafe; }

    void makeValid()
    {
        if (!m_valid)
            m_valid = check();
    }
    bool isValid() const { return m_valid; }
    [[nodiscard]] bool isValid()
    {
        makeValid();
        return m_valid;
    }

private:
    void refreshExtent()
    {
        for (std::size_t i = 0; i < 3; ++i) {
            const T half = (m_dimensions[i] * m_spacing[i]) * T { 0.5 };
            const T lo = m_origin[i] - half;
            const T hi = m_origin[i] + half;
            m_extent[2 * i] = lo;
            m_extent[2 * i + 1] = hi;
            m_extentSafe[2 * i] = std::nextafter(lo, hi);
            m_extentSafe[2 * i + 1] = std::nextafter(hi, lo);
        }
    }

    bool check()
    {
        const auto n = size();
        if (n == 0 || (m_spacing[0] * m_spacing[1] * m_spacing[2]) <= T { 0 })
            return false;
        if (!m_density || !m_materialIndex)
            return false;
        if (!m_measurementMap)
            m_measurementMap = std::make_shared<std::vector<std::uint8_t>>(n, 0);
        if (m_density->size() != n || m_materialIndex->size() != n || m_measurementMap->size() != n)
            return false;
        for (const auto& m : m_materialMap)
            if (!m.isValid())
                return false;
        const auto depth = depthDirection();
        const T* x = m_directionCosines.data();
        const T* y = x + 3;
        const T skew = vectormath::dot(depth.data(), x) + vectormath::dot(depth.data(), y) + vectormath::dot(x, y);
        if (std::abs(skew) > T { 0.001 })
            return false;
        const auto [lo, hi] = std::minmax_element(m_materialIndex->cbegin(), m_materialIndex->cend());
        return *hi < m_materialMap.size() && *lo < m_materialMap.size();
    }

    std::array<T, 3> m_spacing = { 1, 1, 1 };
    std::array<T, 3> m_origin = { 0, 0, 0 };
    std::array<T, 6> m_directionCosines = { 1, 0, 0, 0, 1, 0 };
    std::array<std::size_t, 3> m_dimensions = { 0, 0, 0 };
    std::array<T, 6> m_extent = { 0, 0, 0, 0, 0, 0 };
    std::array<T, 6> m_extentSafe = { 0, 0, 0, 0, 0, 0 };
    std::shared_ptr<std::vector<T>> m_density;
    std::shared_ptr<std::vector<std::uint8_t>> m_materialIndex;
    std::shared_ptr<std::vector<std::uint8_t>> m_measurementMap;
    std::vector<Material> m_materialMap;
    bool m_valid = false;
};

// PMMA cylinder with five air-filled dosimeter bores; the bore voxels within +-50 mm of the centre
// are flagged in the measurement map (reference world.hpp:229-390).
template <Floating T = double>
class CTDIPhantom final : public World<T> {
public:
    enum class HolePosition { Center, West, East, South, North };

    CTDIPhantom(std::size_t diameter = 320) // mm
    {
        const std::array<T, 3> spacing { 1, 1, 2.5 };
        this->setDirectionCosines({ 1, 0, 0, 0, 1, 0 });
        this->setSpacing(spacing);
        this->setOrigin({ 0, 0, 0 });
        const std::size_t nxy = diameter + (diameter % 2 == 0 ? 3 : 2); // always odd
        this->setDimensions({ nxy, nxy, 60 });

        const Material air("Air, Dry (near sea level)");
        const Material pmma("Polymethyl Methacralate (Lucite, Perspex)");
        this->addMaterialToMap(air); // 0: air around the phantom
        this->addMaterialToMap(air); // 1: air inside the bores
        this->addMaterialToMap(pmma); // 2

        constexpr T boreRadius { T { 13.1 } / T { 2.0 } };
        const T boreOffset = T { 13 };
        const T radius = static_cast<T>(diameter) / T { 2 };
        const auto& dim = this->dimensions();
        const auto n = this->size();

        auto density = std::make_shared<std::vector<T>>(n, airDensity());
        auto material = std::make_shared<std::vector<std::uint8_t>>(n, 0);
        auto measurement = std::make_shared<std::vector<std::uint8_t>>(n, 0);
        this->setDensityArray(density);
        this->setMaterialIndexArray(material);
        this->setMeasurementMapArray(measurement);

        const std::array<std::size_t, 2> dim2 { dim[0], dim[1] };
        const std::array<T, 2> sp2 { spacing[0], spacing[1] };
        const std::array<T, 2> centre { dim[0] * spacing[0] * T { 0.5 }, dim[1] * spacing[1] * T { 0.5 } };
        // centre, +y, -y, +x, -x
        std::array<std::array<T, 2>, 5> bores { { { 0, 0 }, { 0, radius - boreOffset }, { 0, -radius + boreOffset }, { radius - boreOffset, 0 },
            { -radius + boreOffset, 0 } } };
        for (auto& b : bores)
            for (std::size_t i = 0; i < 2; ++i)
                b[i] += centre[i];

        const auto body = disc(dim2, sp2, centre, radius);
        std::array<std::vector<std::size_t>, 5> boreIdx;
        for (std::size_t i = 0; i < 5; ++i)
            boreIdx[i] = disc(dim2, sp2, bores[i], boreRadius);

        const std::size_t slice = dim[0] * dim[1];
        for (std::size_t k = 0; k < dim[2]; ++k) {
            const std::size_t offset = k * slice;
            for (auto idx : body) {
                (*density)[idx + offset] = pmma.standardDensity();
                (*material)[idx + offset] = 2;
            }
            for (std::size_t i = 0; i < 5; ++i)
                for (auto idx : boreIdx[i]) {
                    (*density)[idx + offset] = air.standardDensity();
                    (*material)[idx + offset] = 1;
                }
            const T sliceCentre = spacing[2] * (k - dim[2] * 0.5 + 0.5);
            if (std::abs(sliceCentre) <= 50.0)
                for (std::size_t i = 0; i < 5; ++i)
                    for (auto idx : boreIdx[i]) {
                        m_holes[i].push_back(idx + offset);
                        (*measurement)[idx + offset] = 1;
                    }
        }
        this->makeValid();
    }

    const std::vector<std::size_t>& holeIndices(HolePosition position)
    {
        switch (position) {
        case HolePosition::West:
            return m_holes[4];
        case HolePosition::East:
            return m_holes[3];
        case HolePosition::North:
            return m_holes[2];
        case HolePosition::South:
            return m_holes[1];
        default:
            return m_holes[0];
        }
    }
    static constexpr T airDensity() { return T { 0.001205 }; } // g/cm3
    static constexpr std::uint64_t ctdiMinHistories() { return 100E6; }

protected:
    // in-slice indices (x + y*nx) of the voxels whose centre lies within `radius` of `centre`
    static std::vector<std::size_t> disc(const std::array<std::size_t, 2>& dim, const std::array<T, 2>& spacing, const std::array<T, 2>& centre, T radius)
    {
        std::vector<std::size_t> out;
        const int x0 = std::max(static_cast<int>((centre[0] - radius) / spacing[0]), 0);
        const int x1 = std::min(static_cast<int>((centre[0] + radius) / spacing[0]) + 1, static_cast<int>(dim[0]));
        const int y0 = std::max(static_cast<int>((centre[1] - radius) / spacing[1]), 0);
        const int y1 = std::min(static_cast<int>((centre[1] + radius) / spacing[1]) + 1, static_cast<int>(dim[1]));
        const T r2 = radius * radius;
        for (std::size_t i = x0; i < static_cast<std::size_t>(x1); ++i) {
            const T px = centre[0] - i * spacing[0] + spacing[0] / 2;
            for (std::size_t j = y0; j < static_cast<std::size_t>(y1); ++j) {
                const T py = centre[1] - j * spacing[1] + spacing[1] / 2;
                if ((px * px + py * py) <= r2)
                    out.push_back(i + j * dim[0]);
            }
        }
        return out;
    }

private:
    std::array<std::vector<std::size_t>, 5> m_holes;
};
}
