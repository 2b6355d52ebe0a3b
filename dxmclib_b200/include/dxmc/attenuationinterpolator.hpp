// attenuationinterpolator.hpp — piecewise log-log fits of mu/rho and the Woodcock majorant.
//
// Builds, on the host and with the same arithmetic as the reference
// (include/dxmc/attenuationinterpolator.hpp:48-205), the tables the transport kernels evaluate
// per step (csrc/physics.cuh attenuation() / maxAttenuationInverse()):
//   knots       log10(E) grid: uniform in log E plus two knots (E_b - 1 eV, E_b) per shell edge
//   coefficient per material x segment x {photo, Compton, Rayleigh}: intercept b and slope a of
//               log10(mu/rho) over log10(E)
//   majorant    intercept/slope of log10(1 / max_m(rho_max,m * sum mu_m)) on the same knots
// Evaluation on the host (operator(), maxAttenuationInverse) reproduces the reference's look-up,
// including its index conventions; see DESIGN.md "quirks kept".
#pragma once
#include "dxmc/floating.hpp"
#include "dxmc/material.hpp"
#include "dxmc/world.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <iterator>
#include <numeric>
#include <thread>
#include <vector>

namespace dxmc {

template <Floating T>
class AttenuationLutInterpolator {
public:
    AttenuationLutInterpolator() = default;

    // tables for a world: per-material maximum density is taken over all voxels
    AttenuationLutInterpolator(const World<T>& world, T maxEnergy, T minEnergy)
    {
        minEnergy = std::max(T { 0.1 }, minEnergy);
        maxEnergy = std::max(maxEnergy, T { 50 });
        generate(world.materialMap(), world.densityArray()->cbegin(), world.densityArray()->cend(), world.materialIndexArray()->cbegin(), maxEnergy,
            minEnergy);
    }
    // tables for a world whose per-material maximum density is already known (computed on the device from the uploaded
    // grid, dxmcb200_material_max_density): same energy range and same tables as the constructor above, without
    // the pass over all voxels. maxDensity[i] belongs to material i; 0 for a material no voxel uses.
    AttenuationLutInterpolator(const World<T>& world, const std::vector<T>& maxDensity, T maxEnergy, T minEnergy)
    {
        minEnergy = std::max(T { 0.1 }, minEnergy);
        maxEnergy = std::max(maxEnergy, T { 50 });
        const auto& materials = world.materialMap();
        std::vector<T> dens(materials.size(), T { 0 });
        std::vector<std::uint8_t> idx(materials.size());
        for (std::size_t i = 0; i < materials.size(); ++i) {
            if (i < maxDensity.size())
                dens[i] = maxDensity[i];
            idx[i] = static_cast<std::uint8_t>(i);
        }
        generate(materials, dens.cbegin(), dens.cend(), idx.cbegin(), maxEnergy, minEnergy);
    }
    // tables for a bare material list at standard densities (note: caps the range at 50 keV)
    AttenuationLutInterpolator(const std::vector<Material>& materials, T maxEnergy, T minEnergy)
    {
        minEnergy = std::max(T { 0.1 }, minEnergy);
        maxEnergy = std::min(maxEnergy, T { 50 });
        std::vector<T> dens(materials.size());
        std::vector<std::uint8_t> idx(materials.size());
        for (std::size_t i = 0; i < materials.size(); ++i) {
            dens[i] = static_cast<T>(materials[i].standardDensity());
            idx[i] = static_cast<std::uint8_t>(i);
        }
        generate(materials, dens.cbegin(), dens.cend(), idx.cbegin(), maxEnergy, minEnergy);
    }
    template <typename DensIter, typename MatIter>
    AttenuationLutInterpolator(const std::vector<Material>& materials, const DensIter densBegin, const DensIter densEnd, const MatIter matBegin, T maxEnergy,
        T minEnergy)
    {
        minEnergy = std::max(T { 0.1 }, minEnergy);
        maxEnergy = std::min(maxEnergy, T { 50 });
        generate(materials, densBegin, densEnd, matBegin, maxEnergy, minEnergy);
    }

    template <typename DensIter, typename MatIter>
    void generate(const std::vector<Material>& materials, const DensIter densBegin, const DensIter densEnd, const MatIter matBegin, const T maxEnergy,
        const T minEnergy)
    {
        m_resolution = std::max(static_cast<std::size_t>(maxEnergy - minEnergy), std::size_t { 10 });
        const T logMin = std::log10(minEnergy);
        const T logMax = std::log10(maxEnergy);

        // shell edges of every material: a knot just below and one at the edge
        std::vector<T> edges;
        for (const auto& mat : materials)
            for (const auto e : mat.getBindingEnergies(static_cast<double>(minEnergy))) {
                const T edge = static_cast<T>(e);
                edges.push_back(std::log10(edge - shellOffset()));
                edges.push_back(std::log10(edge));
            }
        std::sort(edges.begin(), edges.end());
        edges.erase(std::unique(edges.begin(), edges.end()), edges.end());

        m_x.clear();
        m_x.reserve(m_resolution + 1 + edges.size());
        for (std::size_t i = 0; i <= m_resolution; ++i)
            m_x.push_back(logMin + (i * (logMax - logMin)) / (m_resolution));
        m_x.insert(m_x.end(), edges.begin(), edges.end());
        std::sort(m_x.begin(), m_x.end());
        m_x.erase(std::unique(m_x.begin(), m_x.end()), m_x.end());

        const std::size_t nSeg = m_x.size() - 1;
        m_coefficients.assign(materials.size() * nSeg * 6, T { 0 });
        std::vector<T> y(m_x.size());
        for (std::size_t m = 0; m < materials.size(); ++m) {
            for (std::size_t type = 0; type < 3; ++type) {
                for (std::size_t k = 0; k < m_x.size(); ++k) {
                    const T e = std::pow(T { 10 }, m_x[k]);
                    const double mu = type == 0 ? materials[m].getPhotoelectricAttenuation(e)
                        : type == 1             ? materials[m].getComptonAttenuation(e)
                                                : materials[m].getRayleightAttenuation(e);
                    y[k] = static_cast<T>(std::log10(mu));
                }
                for (std::size_t i = 0; i < nSeg; ++i) {
                    const T slope = (y[i + 1] - y[i]) / (m_x[i + 1] - m_x[i]);
                    T* c = &m_coefficients[(m * nSeg + i) * 6 + type * 2];
                    c[0] = y[i] - m_x[i] * slope;
                    c[1] = slope;
                }
            }
        }

        // from here on m_x[i] is the UPPER edge of segment i
        const T firstKnot = m_x.front();
        m_x.erase(m_x.begin());

        // above the highest shell edge the knots are uniform and the segment is found arithmetically
        if (!edges.empty()) {
            const auto pos = std::upper_bound(m_x.cbegin(), m_x.cend(), edges.back());
            if (pos == m_x.cend()) {
                m_linearIndex = m_x.size();
                m_linearStep = m_x[m_x.size() - 1] - m_x[m_x.size() - 2];
                m_linearEnergy = m_x.back();
            } else {
                m_linearIndex = std::distance(m_x.cbegin(), pos);
                m_linearStep = m_x[m_linearIndex + 1] - m_x[m_linearIndex];
                m_linearEnergy = m_x[m_linearIndex];
            }
        } else {
            m_linearIndex = 0;
            m_linearStep = m_x[1] - m_x[0];
            m_linearEnergy = m_x.front();
        }
        m_resolution = m_x.size();

        buildMajorant(materials, densBegin, densEnd, matBegin, firstKnot);
    }

    // {photo, Compton, Rayleigh} mass attenuation [cm2/g]
    std::array<T, 3> operator()(const std::size_t materialIdx, const T energy) const
    {
        const T logE = std::log10(energy);
        const std::size_t index = logE > m_linearEnergy
            ? std::min(static_cast<std::size_t>((logE - m_linearEnergy) / m_linearStep) + m_linearIndex, m_resolution - 1)
            : searchSegment(logE);
        const T* c = &m_coefficients[(materialIdx * m_resolution + index) * 6];
        std::array<T, 3> res;
        for (std::size_t i = 0; i < 3; ++i)
            res[i] = std::pow(T { 10 }, c[2 * i] + c[2 * i + 1] * logE);
        return res;
    }

    // 1 / (majorant linear attenuation) [cm]
    T maxAttenuationInverse(const T energy) const
    {
        const T logE = std::log10(energy);
        const std::size_t index = logE > m_linearEnergy ? static_cast<std::size_t>((logE - m_linearEnergy) / m_linearStep) + m_linearIndex : searchSegment(logE);
        return std::pow(T { 10 }, m_maxCoefficients[2 * index] + m_maxCoefficients[2 * index + 1] * logE);
    }

    // table access for the device flattening
    const std::vector<T>& knots() const { return m_x; }
    const std::vector<T>& coefficients() const { return m_coefficients; }
    const std::vector<T>& maxCoefficients() const { return m_maxCoefficients; }
    std::size_t resolution() const { return m_resolution; }
    std::size_t linearIndex() const { return m_linearIndex; }
    T linearStep() const { return m_linearStep; }
    T linearEnergy() const { return m_linearEnergy; }

protected:
    static constexpr T shellOffset() { return T { 0.001 }; }

    std::size_t searchSegment(const T logE) const
    {
        const auto pos = std::upper_bound(m_x.cbegin(), m_x.cend(), logE);
        return pos != m_x.cend() ? static_cast<std::size_t>(std::distance(m_x.cbegin(), pos)) : m_resolution - 1;
    }

    template <typename DensIter, typename MatIter>
    void buildMajorant(const std::vector<Material>& materials, const DensIter densBegin, const DensIter densEnd, const MatIter matBegin, const T firstKnot)
    {
        // one pass over the voxels instead of one per material, split over the host cores (maxima do not depend on
        // the order; the reference scans once per material, attenuationinterpolator.hpp:48-59)
        std::vector<T> maxDens(materials.size(), T { 0 });
        const std::size_t nVoxels = static_cast<std::size_t>(std::distance(densBegin, densEnd));
        const std::size_t nChunks = nVoxels > (std::size_t { 1 } << 22) ? std::max(1u, std::thread::hardware_concurrency()) : 1;
        std::vector<std::vector<T>> partial(nChunks, std::vector<T>(materials.size(), T { 0 }));
        std::vector<std::thread> pool;
        auto scan = [&](std::size_t chunk) {
            const std::size_t begin = nVoxels * chunk / nChunks, end = nVoxels * (chunk + 1) / nChunks;
            auto& local = partial[chunk];
            auto d = densBegin + static_cast<std::ptrdiff_t>(begin);
            auto m = matBegin + static_cast<std::ptrdiff_t>(begin);
            for (std::size_t i = begin; i < end; ++i, ++d, ++m)
                if (static_cast<std::size_t>(*m) < local.size())
                    local[*m] = std::max(local[*m], *d);
        };
        for (std::size_t chunk = 1; chunk < nChunks; ++chunk)
            pool.emplace_back(scan, chunk);
        scan(0);
        for (auto& t : pool)
            t.join();
        for (const auto& local : partial)
            for (std::size_t i = 0; i < maxDens.size(); ++i)
                maxDens[i] = std::max(maxDens[i], local[i]);

        std::vector<T> x;
        x.reserve(m_x.size() + 1);
        x.push_back(firstKnot);
        x.insert(x.end(), m_x.begin(), m_x.end());
        std::vector<T> y(x.size());
        for (std::size_t k = 0; k < x.size(); ++k) {
            T maxVal = 0;
            for (std::size_t mat = 0; mat < maxDens.size(); ++mat) {
                const auto att = this->operator()(mat, std::pow(T { 10 }, x[k]));
                maxVal = std::max(maxVal, maxDens[mat] * (((T {} + att[0]) + att[1]) + att[2]));
            }
            y[k] = std::log10(1 / maxVal);
        }
        m_maxCoefficients.resize(m_x.size() * 2);
        for (std::size_t i = 0; i + 1 < x.size(); ++i) {
            const T slope = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
            m_maxCoefficients[2 * i] = y[i] - x[i] * slope;
            m_maxCoefficients[2 * i + 1] = slope;
        }
    }

private:
    std::vector<T> m_x;
    std::vector<T> m_coefficients;
    std::vector<T> m_maxCoefficients;
    std::size_t m_resolution = 40;
    std::size_t m_linearIndex = 0;
    T m_linearStep = 0;
    T m_linearEnergy = 0;
};
}
