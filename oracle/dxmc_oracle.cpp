// dxmc_oracle.cpp — CPU RESTATEMENT of DXMClib's photon-transport hot path. TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (dxmclib_b200/) never does. It is a plain, single-threaded, scalar re-statement of the
// reference algorithm over the SAME plain-data inputs the CUDA runtime takes (include/dxmcb200.h), so a
// test can hand one set of arrays to both and compare. Every function cites the reference lines it follows.
//
// Parity status: PINNED. tests/test_oracle_pinned.py runs this restatement with the reference's own
// sequential PCG32 RandomState and requires the dose / event / energy^2 grids to be bit-identical to the
// unmodified reference (oracle/_ref/libdxmc_ref.so, seeded single worker) on the pencil-beam, isotropic
// spectrum, bow-tie/heel CT and forced-interaction (measurement map) scenes, for all three low-energy
// models. Built with -ffp-contract=off like oracle/_ref.
//
// Two RNG modes:
//   sequential   one PCG32 stream carried across histories and exposures == the reference worker loop
//                (transport.hpp:745-763 with a seeded RandomState)
//   per-history  stream re-keyed per history by dxmcb200_history_stream == what the CUDA kernels do
// Instrumentation the reference cannot expose: per-history step / look-up / interaction / scoring counters
// (the L and S of the roofline model) and voxel-index traces for fixed rays.
#include "dxmcb200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numbers>
#include <numeric>
#include <vector>

namespace {

using T = float;

constexpr T PI = std::numbers::pi_v<T>;
constexpr T ELECTRON_REST_MASS = 510.9989461f; // constants.hpp:63
constexpr T KEV_TO_ANGSTROM = 12.398520f; // constants.hpp:29
constexpr T ENERGY_CUTOFF = 1.0f, ROULETTE_THRESHOLD = 5.0f, ROULETTE_PROBABILITY = 0.8f, N_ERROR = 1.0e-9f; // transport.hpp:819-834

// ---- RandomState (dxmcrandom.hpp:37-167) ----------------------------------------------------------------
struct Rng {
    std::uint64_t s[2];
    std::uint32_t pcg32()
    {
        const std::uint64_t old = s[0];
        s[0] = old * 6364136223846793005ULL + (s[1] | 1);
        const std::uint32_t xorshifted = static_cast<std::uint32_t>(((old >> 18u) ^ old) >> 27u);
        const std::uint32_t rot = static_cast<std::uint32_t>(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
    }
    T uniform() { return pcg32() * T { 2.32830643653869628906e-010 }; } // :67-73
    T uniform(T max) { return uniform() * max; } // :85
    T uniform(T min, T max) // :108-111
    {
        const T r = uniform();
        const T range = max - min;
        return min + r * range;
    }
    std::size_t uniformIndex(std::size_t max) // :86-92, threshold in size_t arithmetic truncated to 32 bit
    {
        const std::uint32_t threshold = static_cast<std::uint32_t>(-max % max);
        for (;;) {
            const auto r = pcg32();
            if (r >= threshold)
                return static_cast<std::size_t>(r % static_cast<std::uint32_t>(max));
        }
    }
};

std::uint64_t mix64(std::uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void historyStream(std::uint64_t seed, std::uint64_t exposure, std::uint64_t history, std::uint64_t out[2])
{
    const std::uint64_t golden = 0x9E3779B97F4A7C15ULL;
    const std::uint64_t s = mix64(seed + golden * (exposure + 1));
    out[0] = mix64(s + golden * (history + 1));
    out[1] = mix64(out[0] + golden) | 1ULL;
}

// ---- vectormath (vectormath.hpp:68-99, 138-146, 197-221) -------------------------------------------------
T dot(const T a[3], const T b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
void cross(const T v1[3], const T v2[3], T res[3])
{
    res[0] = v1[1] * v2[2] - v1[2] * v2[1];
    res[1] = v1[2] * v2[0] - v1[0] * v2[2];
    res[2] = v1[0] * v2[1] - v1[1] * v2[0];
}
void rotate(T vec[3], const T axis[3], const T angle)
{
    const T sang = std::sin(angle);
    const T cang = std::cos(angle);
    const T midt = (T { 1 } - cang) * dot(vec, axis);
    T out[3];
    out[0] = cang * vec[0] + midt * axis[0] + sang * (axis[1] * vec[2] - axis[2] * vec[1]);
    out[1] = cang * vec[1] + midt * axis[1] + sang * (-axis[0] * vec[2] + axis[2] * vec[0]);
    out[2] = cang * vec[2] + midt * axis[2] + sang * (axis[0] * vec[1] - axis[1] * vec[0]);
    vec[0] = out[0];
    vec[1] = out[1];
    vec[2] = out[2];
}
int argmin3(const T vec[3])
{
    const T x = std::abs(vec[0]), y = std::abs(vec[1]), z = std::abs(vec[2]);
    return x <= y ? x <= z ? 0 : 2 : y <= z ? 1 : 2;
}
void peturb(T vec[3], const T theta, const T phi)
{
    T vec_xy[3], k[3] = { 0, 0, 0 };
    k[argmin3(vec)] = T { 1 };
    cross(vec, k, vec_xy);
    rotate(vec_xy, vec, phi);
    const T tsin = std::sin(theta);
    const T tcos = std::cos(theta);
    for (int i = 0; i < 3; ++i)
        vec[i] = vec[i] * tcos + vec_xy[i] * tsin;
}

struct Particle { // particle.hpp:30-47
    T pos[3], dir[3], energy, weight;
};

// ---- inputs (copied from the C-ABI structs) ------------------------------------------------------------------
struct Oracle {
    // world
    std::uint64_t dim[3] {};
    T spacing[3] {}, ext[6] {};
    std::vector<T> density;
    std::vector<std::uint8_t> material, measurement;
    // luts
    std::uint32_t nMat = 0, nSeg = 0, linearIndex = 0;
    T linearStep = 0, linearEnergy = 0;
    std::vector<T> knots, coeff, maxCoeff, rita, spline, shells;
    // beams
    struct Spectrum {
        std::vector<T> probs, energies;
        std::vector<std::uint32_t> alias;
    };
    struct Heel {
        T e0, de;
        std::size_t ne;
        T a0, da;
        std::size_t na;
        std::vector<T> w;
    };
    struct Bowtie {
        std::vector<T> angles, weights;
    };
    std::vector<Spectrum> spectra;
    std::vector<Heel> heels;
    std::vector<Bowtie> bowties;
    // results: the reference's own float accumulators (Result<T>, transport.hpp:42-65), plain (single thread)
    std::vector<T> dose, variance;
    std::vector<std::uint32_t> nEvents;
    // the product's scoring format next to it: 64-bit fixed point, round(e * 2^bits) (include/dxmcb200.h), so that
    // integer grids can be compared / summed across shards bit for bit
    std::vector<std::int64_t> fixedEnergy;
    std::vector<std::uint64_t> fixedEnergySq;
    int energyBits = 20, energySqBits = 10;
    // instrumentation
    dxmcb200_stats stats {};

    std::size_t nVoxels() const { return dim[0] * dim[1] * dim[2]; }

    // ---- AttenuationLutInterpolator::operator() / maxAttenuationInverse (attenuationinterpolator.hpp:207-248)
    std::size_t segment(T logEnergy, bool clamp) const
    {
        if (logEnergy > linearEnergy) {
            const std::size_t i = static_cast<std::size_t>((logEnergy - linearEnergy) / linearStep) + linearIndex;
            return clamp ? std::min(i, static_cast<std::size_t>(nSeg) - 1) : i;
        }
        const auto pos = std::upper_bound(knots.cbegin(), knots.cend(), logEnergy);
        return pos != knots.cend() ? static_cast<std::size_t>(std::distance(knots.cbegin(), pos)) : nSeg - 1;
    }
    std::array<T, 3> attenuation(std::size_t materialIdx, T energy) const
    {
        std::array<T, 3> res;
        const T logEnergy = std::log10(energy);
        const std::size_t offset = materialIdx * nSeg * 6 + segment(logEnergy, true) * 6;
        for (std::size_t i = 0; i < 3; ++i)
            res[i] = std::pow(T { 10 }, coeff[offset + 2 * i] + coeff[offset + 2 * i + 1] * logEnergy);
        return res;
    }
    T maxAttenuationInverse(T energy) const
    {
        const T logEnergy = std::log10(energy);
        const std::size_t index = std::min(segment(logEnergy, false), static_cast<std::size_t>(nSeg) - 1); // clamp = memory safety only
        return std::pow(T { 10 }, maxCoeff[2 * index] + maxCoeff[2 * index + 1] * logEnergy);
    }
    // ---- CubicSplineInterpolator::operator() (interpolation.hpp:169-176)
    T scatterFactor(std::size_t m, T x_val) const
    {
        const T* s = &spline[m * DXMCB200_SPLINE_FLOATS];
        const T start = s[60], step = s[61], stop = s[62];
        const T x = std::clamp(x_val, start, stop);
        const std::size_t index = x > start ? static_cast<std::size_t>((x - start) / step) : 0;
        const std::size_t offset = index < DXMCB200_SPLINE_N - 1 ? index * 4 : (DXMCB200_SPLINE_N - 2) * 4;
        return s[offset] + s[offset + 1] * x + s[offset + 2] * x * x + s[offset + 3] * x * x * x;
    }
    // ---- RITA::operator()(state, maxValue) (dxmcrandom.hpp:483-500)
    T sampleFormFactor(std::size_t m, T maxValue, Rng& state) const
    {
        const T* x = &rita[m * 4 * DXMCB200_RITA_N];
        const T* e = x + DXMCB200_RITA_N;
        const T* a = e + DXMCB200_RITA_N;
        const T* b = a + DXMCB200_RITA_N;
        const auto ub = std::upper_bound(x, x + DXMCB200_RITA_N, maxValue);
        const T modifier = ub != x + DXMCB200_RITA_N ? e[ub - x] : 1;
        T res;
        do {
            const auto r1 = state.uniform(modifier);
            const auto index = (std::upper_bound(e, e + DXMCB200_RITA_N, r1) - e) - 1;
            const auto v = r1 - e[index];
            const auto d = e[index + 1] - e[index];
            res = x[index] + (1 + a[index] + b[index]) * d * v / (d * d + a[index] * d * v + b[index] * v * v) * (x[index + 1] - x[index]);
        } while (res > maxValue);
        return res;
    }
    const T* shell(std::size_t m, int i) const { return &shells[(m * DXMCB200_SHELLS + i) * DXMCB200_SHELL_FLOATS]; }

    // ---- Exposure::sampleParticle (exposure.hpp:280-304) with its table look-ups
    T sampleSpectrum(const Spectrum& s, Rng& state) const // dxmcrandom.hpp:211-216, 322-326
    {
        const auto r = state.uniform();
        const auto k = state.uniformIndex(s.probs.size());
        const std::size_t ind = r < s.probs[k] ? k : s.alias[k];
        return ind < s.energies.size() - 1 ? state.uniform(s.energies[ind], s.energies[ind + 1]) : s.energies[ind];
    }
    static T bowtieWeight(const Bowtie& b, T anglePlusAndMinus) // beamfilters.hpp:150-190
    {
        const T angle = std::abs(anglePlusAndMinus);
        std::size_t first = 0, last = b.angles.size() - 1;
        std::size_t it = first + (last - first) / 2;
        while (it != first) {
            if (angle < b.angles[it]) {
                last = it;
                it = first;
            } else {
                first = it;
            }
            it += (last - first) / 2;
        }
        if (angle < b.angles[first])
            return b.weights.front();
        if (angle > b.angles[last])
            return b.weights.back();
        const T x0 = b.angles[first], x1 = b.angles[last], y0 = b.weights[first], y1 = b.weights[last];
        return y0 + (angle - x0) * (y1 - y0) / (x1 - x0);
    }
    static T heelWeight(const Heel& h, T angle, T energy) // beamfilters.hpp:514-538
    {
        const T ev = (energy - h.e0 + T { 0.5 } * h.de) / h.de;
        std::size_t e_index = ev > 0 ? static_cast<std::size_t>(ev) : 0; // the reference casts a negative float here; the guard below overrides it
        if (e_index >= h.ne)
            e_index = h.ne - 1;
        if (energy < h.e0)
            e_index = 0;
        const T av = (angle - h.a0) / h.da;
        std::size_t a_index = av > 0 ? static_cast<std::size_t>(av) : 0;
        if (a_index >= h.na)
            a_index = h.na - 1;
        if (angle < h.a0)
            a_index = 0;
        const std::size_t w_index = e_index * h.na + a_index;
        if (a_index < h.na - 1) {
            const T a0 = h.a0 + h.da * a_index;
            const T a1 = h.a0 + h.da * (a_index + 1);
            const T w0 = h.w[w_index], w1 = h.w[w_index + 1];
            return w0 + (w1 - w0) * (angle - a0) / (a1 - a0);
        }
        return h.w[w_index];
    }
    Particle sampleParticle(const dxmcb200_exposure& e, Rng& state) const
    {
        const T theta = state.uniform(e.collimation[0], e.collimation[1]);
        const T phi = state.uniform(e.collimation[2], e.collimation[3]);
        Particle p;
        for (int i = 0; i < 3; ++i) {
            p.pos[i] = e.position[i];
            p.dir[i] = e.beam_direction[i];
        }
        p.weight = e.weight;
        rotate(p.dir, &e.cosines[3], theta);
        rotate(p.dir, &e.cosines[0], phi);
        p.energy = e.spectrum >= 0 ? sampleSpectrum(spectra[e.spectrum], state) : e.mono_energy;
        if (e.bowtie >= 0)
            p.weight *= bowtieWeight(bowties[e.bowtie], theta);
        if (e.heel >= 0)
            p.weight *= heelWeight(heels[e.heel], phi, p.energy);
        return p;
    }

    // ---- geometry (transport.hpp:485-521, 702-728)
    bool inside(const T pos[3]) const
    {
        return (pos[0] > ext[0] && pos[0] < ext[1]) && (pos[1] > ext[2] && pos[1] < ext[3]) && (pos[2] > ext[4] && pos[2] < ext[5]);
    }
    std::size_t indexFromPosition(const T pos[3]) const
    {
        const std::size_t ix = static_cast<std::size_t>((pos[0] - ext[0]) / spacing[0]);
        const std::size_t iy = static_cast<std::size_t>((pos[1] - ext[2]) / spacing[1]);
        const std::size_t iz = static_cast<std::size_t>((pos[2] - ext[4]) / spacing[2]);
        return iz * dim[0] * dim[1] + iy * dim[0] + ix;
    }
    bool transportParticleToWorld(Particle& particle) const
    {
        if (inside(particle.pos))
            return true;
        auto amin = std::numeric_limits<T>::lowest();
        auto amax = std::numeric_limits<T>::max();
        for (std::size_t i = 0; i < 3; i++) {
            if (std::abs(particle.dir[i]) > N_ERROR) {
                const auto a0 = (ext[i * 2] - particle.pos[i]) / particle.dir[i];
                const auto an = (ext[i * 2 + 1] - particle.pos[i]) / particle.dir[i];
                amin = std::max(amin, std::min(a0, an));
                amax = std::min(amax, std::max(a0, an));
            }
        }
        if (amin < amax && amin > 0) {
            for (std::size_t i = 0; i < 3; i++)
                particle.pos[i] += amin * particle.dir[i];
            return true;
        }
        return false;
    }

    // ---- scoring (transport.hpp:598-609; single thread, so plain adds replace the atomic_ref adds)
    void score(std::size_t idx, T energyImparted)
    {
        dose[idx] += energyImparted;
        nEvents[idx] += 1;
        variance[idx] += energyImparted * energyImparted;
        if (std::isfinite(energyImparted)) { // the product drops a non-finite deposit (csrc/transport.cu deposit()); the float sums above keep the reference's behaviour
            fixedEnergy[idx] += std::llrint(std::ldexp(energyImparted, energyBits));
            fixedEnergySq[idx] += static_cast<std::uint64_t>(std::llrint(std::ldexp(energyImparted * energyImparted, energySqBits)));
        }
        ++stats.score_events;
    }

    // ---- photoAbsorption<L> (transport.hpp:216-263)
    template <int L>
    T photoAbsorption(Particle& particle, std::uint8_t materialIdx, Rng& state) const
    {
        const auto E = particle.energy;
        particle.energy = 0;
        if constexpr (L < 2) {
            return E;
        } else {
            std::array<T, 12> shell_probs;
            T run = 0;
            for (int i = 0; i < 12; ++i) { // transform + partial_sum
                const T* c = shell(materialIdx, i);
                const T v = E > c[0] ? c[3] : T { 0 };
                run = i == 0 ? v : run + v;
                shell_probs[i] = run;
            }
            std::size_t idx = 0;
            const auto shellIdxSample = shell_probs.back() * state.uniform();
            while (shell_probs[idx] < shellIdxSample && idx < 11)
                ++idx;
            const T* c = shell(materialIdx, static_cast<int>(idx));
            if (c[0] > E || c[0] < ENERGY_CUTOFF)
                return E;
            const auto r1 = state.uniform();
            if (r1 <= c[4]) {
                std::size_t lineIdx = 0;
                auto r3 = state.uniform() - c[5 + lineIdx];
                while (r3 > 0 && lineIdx < 2) {
                    ++lineIdx;
                    r3 -= c[5 + lineIdx];
                }
                particle.energy = c[8 + lineIdx];
                const auto theta = state.uniform(PI);
                const auto phi = state.uniform(PI + PI);
                peturb(particle.dir, theta, phi);
                return E - particle.energy;
            }
            return E;
        }
    }

    // ---- rayleightScatter<L> (transport.hpp:264-298)
    template <int L>
    void rayleighScatter(Particle& particle, std::uint8_t materialIdx, Rng& state) const
    {
        if constexpr (L == 0) {
            bool reject = true;
            T theta;
            while (reject) {
                constexpr T extreme = (4 * std::numbers::sqrt2_v<T>) / (3 * std::numbers::sqrt3_v<T>);
                const auto r1 = state.uniform(T { 0 }, extreme);
                theta = state.uniform(T { 0 }, PI);
                const auto sinang = std::sin(theta);
                reject = r1 > ((2 - sinang * sinang) * sinang);
            }
            const auto phi = state.uniform(PI + PI);
            peturb(particle.dir, theta, phi);
        } else {
            constexpr T k = 1 / KEV_TO_ANGSTROM;
            const auto qmax = particle.energy * k; // attenuationlut.hpp:187-191
            const auto qmax_squared = qmax * qmax;
            T cosAngle;
            do {
                const auto q_squared = sampleFormFactor(materialIdx, qmax_squared, state);
                const auto invE = KEV_TO_ANGSTROM / particle.energy; // attenuationlut.hpp:200-204
                cosAngle = 1 - 2 * q_squared * invE * invE;
            } while ((1 + cosAngle * cosAngle) * T { 0.5 } < state.uniform());
            const auto theta = std::acos(cosAngle);
            const auto phi = state.uniform(PI + PI);
            peturb(particle.dir, theta, phi);
        }
    }

    // ---- comptonScatterNRC (transport.hpp:342-483)
    T comptonScatterNRC(Particle& particle, std::uint8_t materialIdx, Rng& state) const
    {
        std::array<T, 12> shellProbs;
        T run = 0;
        for (int i = 0; i < 12; ++i) {
            const T* e = shell(materialIdx, i);
            const T v = e[0] < particle.energy && e[2] > 0 ? e[1] : T { 0 };
            run = i == 0 ? v : run + v;
            shellProbs[i] = run;
        }
        int shellIdx = 0;
        const auto shellIdxSample = shellProbs.back() * state.uniform();
        while (shellProbs[shellIdx] < shellIdxSample && shellIdx < 11)
            ++shellIdx;
        const T* cfg = shell(materialIdx, shellIdx);
        const auto U = cfg[0] / ELECTRON_REST_MASS;
        const auto p = std::sqrt(2 * U + U * U);
        const auto J0 = cfg[2];
        const auto k = particle.energy / ELECTRON_REST_MASS;
        if (U > k)
            return 0;
        const auto emin = 1 / (1 + 2 * k);
        const auto gmax_inv = 1 / (1 / emin + emin);
        T e, cosAngle;
        bool rejected;
        do {
            const auto r1 = state.uniform();
            e = r1 + (1 - r1) * emin;
            const auto t = (1 - e) / (k * e);
            const auto sinAngleSqr = t * (2 - t);
            cosAngle = 1 - t;
            const auto g = (1 / e + e - sinAngleSqr) * gmax_inv;
            const auto r2 = state.uniform();
            rejected = r2 > g;
            if (!rejected) {
                const auto pi = (k * (k - U) * (1 - cosAngle) - U) / std::sqrt(2 * k * (k - U) * (1 - cosAngle) + U * U);
                const auto kc = k * e;
                const auto qc = std::sqrt(k * k + kc * kc - 2 * k * kc * cosAngle);
                const auto alpha = qc * (1 + kc * (kc - k * cosAngle) / (qc * qc)) / k;
                const auto b_part = 1 + 2 * J0 * std::abs(pi);
                const auto b = (b_part * b_part + 1) / 2;
                const auto expb = std::exp(-b);
                T S;
                if (pi <= -p) {
                    S = (1 - alpha * p) * expb / 2;
                } else if (pi < p) {
                    const auto pabs = std::abs(p);
                    const auto piabs = std::abs(pi);
                    constexpr auto sp2 = 1 / (std::numbers::inv_sqrtpi_v<T> / std::numbers::sqrt2_v<T>);
                    const auto part1 = alpha * sp2 / (4 * J0);
                    constexpr auto a1 = T { 0.34802 };
                    constexpr auto a2 = T { -0.0958798 };
                    constexpr auto a3 = T { 0.7478556 };
                    const auto sqrte = std::sqrt(std::numbers::e_v<T>);
                    const auto tp = 1 / (1 + T { 0.332673 } * (1 + 2 * J0 * pabs));
                    const auto tpi = 1 / (1 + T { 0.332673 } * (1 + 2 * J0 * piabs));
                    const auto part2p = sqrte - expb * tp * (a1 + a2 * tp + a3 * tp * tp);
                    const auto part2pi = sqrte - expb * tpi * (a1 + a2 * tpi + a3 * tpi * tpi);
                    if (pi <= 0)
                        S = (1 - alpha * pi) * expb / 2 - part1 * (part2p - part2pi);
                    else
                        S = 1 - (1 - alpha * pi) * expb / 2 - part1 * (part2p - part2pi);
                } else {
                    S = 1 - (1 - alpha * p) * expb / 2;
                }
                const auto r3 = state.uniform();
                rejected = r3 > S;
                if (!rejected) {
                    T Fmax;
                    if (pi <= -p)
                        Fmax = 1 - alpha * p;
                    else if (pi >= p)
                        Fmax = 1 + alpha * p;
                    else
                        Fmax = 1 + alpha * pi;
                    const auto r4 = state.uniform();
                    const auto r_bar2 = 2 * r4 * expb;
                    T pz;
                    // the reference calls the unqualified C `sqrt` here (transport.hpp:449, 452), i.e. the double
                    // overload: `part` and the quotient are evaluated in double and only pz is narrowed to T
                    if (r_bar2 < 1) {
                        const double part = std::sqrt(static_cast<double>(1 - 2 * std::log(r_bar2)));
                        pz = static_cast<T>((1 - part) / (2 * J0) / ELECTRON_REST_MASS);
                    } else {
                        const double part = std::sqrt(static_cast<double>(1 - 2 * std::log(2 - r_bar2)));
                        pz = static_cast<T>((part - 1) / (2 * J0) / ELECTRON_REST_MASS);
                    }
                    T Fpz;
                    if (pz <= -p)
                        Fpz = 1 - alpha * p;
                    else if (pz >= p)
                        Fpz = 1 + alpha * p;
                    else
                        Fpz = 1 + alpha * pz;
                    const auto r5 = state.uniform();
                    rejected = r5 > Fpz / Fmax;
                    if (!rejected) {
                        const auto part = std::sqrt(1 - 2 * e * cosAngle + e * e * (1 - pz * pz * sinAngleSqr));
                        const auto k_bar = kc / (1 - pz * pz * e * e) * (1 - pz * pz * e * cosAngle + pz * part);
                        e = k_bar / k;
                    }
                }
            }
        } while (rejected);
        const auto theta = std::acos(cosAngle);
        const auto phi = state.uniform(PI + PI);
        peturb(particle.dir, theta, phi);
        const auto E = particle.energy;
        particle.energy *= e;
        return E - particle.energy;
    }

    // ---- comptonScatter<L> (transport.hpp:300-340)
    template <int L>
    T comptonScatter(Particle& particle, std::uint8_t materialIdx, Rng& state) const
    {
        if constexpr (L == 2) {
            return comptonScatterNRC(particle, materialIdx, state);
        } else {
            const auto E = particle.energy;
            const auto k = E / ELECTRON_REST_MASS;
            const auto emin = 1 / (1 + 2 * k);
            const auto gmax_inv = 1 / (1 / emin + emin);
            T e, cosAngle;
            bool rejected;
            do {
                const auto r1 = state.uniform();
                e = r1 + (1 - r1) * emin;
                const auto t = (1 - e) / (k * e);
                const auto sinthetasqr = t * (2 - t);
                cosAngle = 1 - t;
                const auto g = (1 / e + e - sinthetasqr) * gmax_inv;
                const auto r2 = state.uniform();
                if constexpr (L == 1) {
                    constexpr T kq = 1 / KEV_TO_ANGSTROM;
                    const auto q = E * kq * std::sqrt(T { 0.5 } - cosAngle * T { 0.5 }); // attenuationlut.hpp:175-179
                    rejected = r2 > g * scatterFactor(materialIdx, q);
                } else {
                    rejected = r2 > g;
                }
            } while (rejected);
            const auto theta = std::acos(cosAngle);
            const auto phi = state.uniform(PI + PI);
            peturb(particle.dir, theta, phi);
            particle.energy *= e;
            return E - particle.energy;
        }
    }

    // ---- computeInteractions<L> (transport.hpp:583-638)
    template <int L>
    bool computeInteractions(const std::array<T, 3>& attenuation, Particle& p, std::uint8_t matIdx, std::size_t idx, Rng& state, bool& updateMax)
    {
        const auto attPhoto = attenuation[0], attCompt = attenuation[1];
        const auto attenuationTotal = ((T { 0 } + attenuation[0]) + attenuation[1]) + attenuation[2];
        const auto r3 = state.uniform(attenuationTotal);
        if (r3 < attPhoto) {
            const auto e = photoAbsorption<L>(p, matIdx, state);
            if (p.energy < ENERGY_CUTOFF) {
                score(idx, (e + p.energy) * p.weight);
                p.energy = 0;
                return false;
            }
            score(idx, e * p.weight);
            updateMax = true;
        } else if (r3 < (attPhoto + attCompt)) {
            const auto e = comptonScatter<L>(p, matIdx, state);
            if (p.energy < ENERGY_CUTOFF) {
                score(idx, (e + p.energy) * p.weight);
                p.energy = 0;
                return false;
            }
            score(idx, e * p.weight);
            updateMax = true;
        } else {
            rayleighScatter<L>(p, matIdx, state);
        }
        return true;
    }

    // ---- computeInteractionsForced<L> (transport.hpp:523-581)
    template <int L>
    bool computeInteractionsForced(T eventProbability, const std::array<T, 3>& attenuation, Particle& p, std::uint8_t matIdx, std::size_t idx, Rng& state,
        bool& updateMax)
    {
        const auto attPhoto = attenuation[0], attCompt = attenuation[1], attRayl = attenuation[2];
        const auto attenuationTotal = ((T { 0 } + attenuation[0]) + attenuation[1]) + attenuation[2];
        const auto photoEventProbability = attPhoto / attenuationTotal;
        const auto weightCorrection = eventProbability * photoEventProbability;
        {
            auto p_forced = p;
            const auto e_forced = photoAbsorption<L>(p_forced, matIdx, state);
            if (p_forced.energy < ENERGY_CUTOFF)
                score(idx, (e_forced + p_forced.energy) * p_forced.weight * weightCorrection);
            else
                score(idx, e_forced * p_forced.weight * weightCorrection);
        }
        const auto r1 = state.uniform();
        if (r1 < eventProbability * (1 - photoEventProbability)) {
            const auto r2 = state.uniform(attCompt + attRayl);
            if (r2 < attCompt) {
                const auto e = comptonScatter<L>(p, matIdx, state);
                if (p.energy < ENERGY_CUTOFF) {
                    score(idx, (e + p.energy) * p.weight);
                    p.energy = 0;
                    return false;
                }
                score(idx, e * p.weight);
                updateMax = true;
            } else {
                rayleighScatter<L>(p, matIdx, state);
            }
        }
        p.weight *= (1 - weightCorrection);
        return true;
    }

    // ---- woodcockParticleTracking<L> (transport.hpp:640-700)
    template <int L>
    void woodcock(Particle& p, Rng& state)
    {
        T maxAttenuationInv = 0;
        bool updateMaxAttenuation = true;
        bool continueSampling = true;
        while (continueSampling) {
            if (updateMaxAttenuation) {
                maxAttenuationInv = maxAttenuationInverse(p.energy);
                updateMaxAttenuation = false;
            }
            const auto r1 = state.uniform();
            const auto stepLenght = -std::log(r1) * maxAttenuationInv * T { 10 };
            for (std::size_t i = 0; i < 3; i++)
                p.pos[i] += p.dir[i] * stepLenght;
            ++stats.steps;
            if (inside(p.pos)) {
                const std::size_t bufferIdx = indexFromPosition(p.pos);
                const auto matIdx = material[bufferIdx];
                const auto dens = density[bufferIdx];
                const auto meas = measurement.empty() ? std::uint8_t { 0 } : measurement[bufferIdx];
                ++stats.lookups;
                const auto att = attenuation(matIdx, p.energy);
                const auto attenuationTotal = (((T { 0 } + att[0]) + att[1]) + att[2]) * dens;
                const auto eventProbability = attenuationTotal * maxAttenuationInv;
                if (meas == 0) {
                    const auto r2 = state.uniform();
                    if (r2 < eventProbability) {
                        ++stats.interactions;
                        continueSampling = computeInteractions<L>(att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
                    }
                } else {
                    ++stats.interactions;
                    continueSampling = computeInteractionsForced<L>(eventProbability, att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
                }
                if (continueSampling) {
                    if (p.energy * p.weight < ROULETTE_THRESHOLD) {
                        const auto r4 = state.uniform();
                        if (r4 < ROULETTE_PROBABILITY) {
                            continueSampling = false;
                        } else {
                            constexpr T factor = T { 1 } / (T { 1 } - ROULETTE_PROBABILITY);
                            p.weight *= factor;
                        }
                    }
                }
            } else {
                continueSampling = false;
            }
        }
    }

    // =====================================================================================================
    // Empty-space traversal (NOT in the reference; restated from the product's design, DESIGN.md section 4b, so that
    // the CUDA kernels can be checked stream by stream). The voxel grid is cut into bricks of 2^k voxels per axis.
    // A brick is "air" when rho * mu_total(E) <= fAir * majorant(E) holds for all its voxels and all table energies
    // (fAir <= kAirThreshold) and it holds no measurement voxel. Photons are tracked with the reference's Woodcock
    // loop (global majorant) everywhere else. Whenever a photon stands in an air brick — at birth, after
    // transportParticleToWorld, or after a virtual collision — it is walked through the run of air bricks on its ray
    // by a parametric ray/grid traversal (Siddon 1985 / Amanatides & Woo 1987) up to the first non-air brick or the
    // edge of the grid; collision candidates on that stretch are sampled against the regional majorant
    // fAir * majorant(E) (delta tracking with any valid majorant gives the same collision density). Everything
    // that happens AT a candidate (look-up, accept / reject, forced interaction, roulette) is the reference's code.
    // =====================================================================================================
    static constexpr double kAirThreshold = 0.02;
    struct Bricks {
        bool enabled = false;
        std::uint32_t shift[3] {}, nb[3] {};
        T size[3] {}; // brick edge [mm]
        T invSize[3] {}; // 1 / size
        std::vector<std::uint8_t> air; // [nb2][nb1][nb0]
        // per octant of travel directions o = (dx<0) | (dy<0)<<1 | (dz<0)<<2 and per brick b: the edge k (in bricks, at most 255) of the
        // largest cube of air bricks that has b as its corner and extends k bricks from it in that octant's directions (bricks
        // beyond the grid count as air); 0 for non-air bricks. [8][nb2][nb1][nb0]
        std::vector<std::uint8_t> distance;
        std::vector<T> ratio; // per material: max over E of mu_total(E) / majorant(E)
        std::vector<T> brickMax; // per brick: max over voxels of rho * ratio[material]
        T fAir = 0, invFAir = 0;
        std::uint64_t nAir = 0;
    } bricks;
    std::uint64_t brickSteps = 0, walks = 0, walkCandidates = 0; // instrumentation

    // brick edge per axis: the power of two (in voxels) closest to `brickMm`; while the grid has more than kMaxBricks
    // bricks the axis with the shortest brick edge in mm is doubled (ties: z before y before x)
    static constexpr std::uint64_t kMaxBricks = 16384;
    void buildBricks(T brickMm)
    {
        Bricks& b = bricks;
        b = Bricks {};
        for (int i = 0; i < 3; ++i) {
            const double k = std::round(std::log2(static_cast<double>(brickMm) / static_cast<double>(spacing[i])));
            b.shift[i] = static_cast<std::uint32_t>(std::clamp(k, 0.0, 10.0));
        }
        for (;;) {
            std::uint64_t n = 1;
            for (int i = 0; i < 3; ++i) {
                b.nb[i] = static_cast<std::uint32_t>((dim[i] + (1ull << b.shift[i]) - 1) >> b.shift[i]);
                n *= b.nb[i];
            }
            if (n <= kMaxBricks)
                break;
            int grow = 2;
            for (int i = 1; i >= 0; --i)
                if (static_cast<double>(1u << b.shift[i]) * spacing[i] < static_cast<double>(1u << b.shift[grow]) * spacing[grow])
                    grow = i;
            ++b.shift[grow];
        }
        for (int i = 0; i < 3; ++i) {
            b.size[i] = static_cast<T>(1u << b.shift[i]) * spacing[i];
            b.invSize[i] = T { 1 } / b.size[i];
        }
        // per material: the largest mu_total(E) * majorantInverse(E) over the table. Inside a segment it is a sum of
        // exponentials in log10 E (convex), so the maximum sits at a segment end; evaluated in double.
        b.ratio.assign(nMat, T { 0 });
        for (std::size_t m = 0; m < nMat; ++m) {
            double best = 0;
            for (std::size_t k = 0; k < nSeg; ++k) {
                const double ends[2] = { k == 0 ? 0.0 : static_cast<double>(knots[k - 1]), static_cast<double>(knots[k]) };
                for (const double x : ends) {
                    double total = 0;
                    for (int i = 0; i < 3; ++i)
                        total += std::pow(10.0, static_cast<double>(coeff[(m * nSeg + k) * 6 + 2 * i]) + static_cast<double>(coeff[(m * nSeg + k) * 6 + 2 * i + 1]) * x);
                    best = std::max(best, total * std::pow(10.0, static_cast<double>(maxCoeff[2 * k]) + static_cast<double>(maxCoeff[2 * k + 1]) * x));
                }
            }
            b.ratio[m] = static_cast<T>(best);
        }
        const std::size_t nBricks = static_cast<std::size_t>(b.nb[0]) * b.nb[1] * b.nb[2];
        b.brickMax.assign(nBricks, T { 0 });
        std::vector<std::uint8_t> measured(nBricks, 0);
        for (std::size_t z = 0; z < dim[2]; ++z)
            for (std::size_t y = 0; y < dim[1]; ++y)
                for (std::size_t x = 0; x < dim[0]; ++x) {
                    const std::size_t v = (z * dim[1] + y) * dim[0] + x;
                    const std::size_t br = ((z >> b.shift[2]) * b.nb[1] + (y >> b.shift[1])) * b.nb[0] + (x >> b.shift[0]);
                    const T f = density[v] * b.ratio[material[v]];
                    if (f > b.brickMax[br]) // NaN and negative densities never win
                        b.brickMax[br] = f;
                    if (!measurement.empty() && measurement[v])
                        measured[br] = 1;
                }
        b.air.assign(nBricks, 0);
        double fAir = 0;
        for (std::size_t br = 0; br < nBricks; ++br) {
            const double f = 1.001 * static_cast<double>(b.brickMax[br]);
            if (f <= kAirThreshold && !measured[br]) {
                b.air[br] = 1;
                ++b.nAir;
                fAir = std::max(fAir, f);
            }
        }
        b.fAir = static_cast<T>(std::max(fAir, 1.0e-6));
        b.invFAir = T { 1 } / b.fAir;
        b.enabled = b.nAir > 0;
        // largest all-air cube with corner b per octant: D(b) = 1 + min over the 7 neighbours one brick further along the
        // octant's directions (a neighbour beyond the grid counts as 255), swept from the far corner of the octant backwards
        b.distance.assign(8 * nBricks, 0);
        const std::int64_t n0 = b.nb[0], n1 = b.nb[1], n2 = b.nb[2];
        for (int o = 0; o < 8; ++o) {
            const std::int64_t s0 = (o & 1) ? -1 : 1, s1 = (o & 2) ? -1 : 1, s2 = (o & 4) ? -1 : 1;
            std::uint8_t* D = b.distance.data() + static_cast<std::size_t>(o) * nBricks;
            for (std::int64_t kz = 0; kz < n2; ++kz)
                for (std::int64_t ky = 0; ky < n1; ++ky)
                    for (std::int64_t kx = 0; kx < n0; ++kx) {
                        // visit far-to-near along every axis of the octant
                        const std::int64_t x = s0 > 0 ? n0 - 1 - kx : kx, y = s1 > 0 ? n1 - 1 - ky : ky, z = s2 > 0 ? n2 - 1 - kz : kz;
                        const std::size_t at = static_cast<std::size_t>((z * n1 + y) * n0 + x);
                        if (!b.air[at])
                            continue;
                        int least = 255;
                        for (int m = 1; m < 8; ++m) {
                            const std::int64_t X = x + ((m & 1) ? s0 : 0), Y = y + ((m & 2) ? s1 : 0), Z = z + ((m & 4) ? s2 : 0);
                            if (X < 0 || Y < 0 || Z < 0 || X >= n0 || Y >= n1 || Z >= n2)
                                continue;
                            least = std::min<int>(least, D[static_cast<std::size_t>((Z * n1 + Y) * n0 + X)]);
                        }
                        D[at] = static_cast<std::uint8_t>(std::min(least + 1, 255));
                    }
        }
    }

    void voxelCoordinates(const T pos[3], std::uint32_t out[3]) const // clamped: also valid on the faces of the world
    {
        for (int i = 0; i < 3; ++i) {
            const T rel = (pos[i] - ext[2 * i]) / spacing[i];
            const std::uint64_t v = rel > 0 ? static_cast<std::uint64_t>(rel) : 0;
            out[i] = static_cast<std::uint32_t>(std::min<std::uint64_t>(v, dim[i] - 1));
        }
    }
    bool airBrick(const std::uint32_t b[3]) const { return bricks.air[(static_cast<std::size_t>(b[2]) * bricks.nb[1] + b[1]) * bricks.nb[0] + b[0]] != 0; }
    bool inAirBrick(const T pos[3]) const
    {
        std::uint32_t v[3];
        voxelCoordinates(pos, v);
        const std::uint32_t b[3] = { v[0] >> bricks.shift[0], v[1] >> bricks.shift[1], v[2] >> bricks.shift[2] };
        return airBrick(b);
    }

    // Ray parameter at which the ray leaves the run of air bricks it starts in; `exits` when it leaves the grid there. The
    // traversal is parametric like Siddon's / Amanatides & Woo's, but it does not stop at every brick face: from an air brick b
    // the ray crosses, in one step, the largest cube of air bricks that has b as its corner and opens in the ray's octant of
    // directions (edge k bricks, tabulated per octant and brick). All face parameters are taken from the starting point, so
    // nothing accumulates.
    T airRunLength(const Particle& p, bool& exits)
    {
        std::uint32_t v[3];
        voxelCoordinates(p.pos, v);
        std::int64_t b[3];
        T inv[3];
        int step[3];
        std::size_t octant = 0;
        for (int i = 0; i < 3; ++i) {
            b[i] = v[i] >> bricks.shift[i];
            if (std::abs(p.dir[i]) > N_ERROR) {
                inv[i] = T { 1 } / p.dir[i];
                step[i] = p.dir[i] > 0 ? 1 : -1;
            } else {
                inv[i] = 0;
                step[i] = 0;
            }
            if (p.dir[i] < 0)
                octant |= std::size_t { 1 } << i;
        }
        const std::size_t nBricks = static_cast<std::size_t>(bricks.nb[0]) * bricks.nb[1] * bricks.nb[2];
        const std::uint8_t* edge = bricks.distance.data() + octant * nBricks;
        exits = false;
        T travelled = 0;
        int cubes = 0;
        for (;;) {
            const std::int64_t k = edge[(static_cast<std::size_t>(b[2]) * bricks.nb[1] + static_cast<std::size_t>(b[1])) * bricks.nb[0] + static_cast<std::size_t>(b[0])];
            T t[3];
            for (int i = 0; i < 3; ++i) {
                if (step[i] != 0) {
                    const T face = ext[2 * i] + static_cast<T>(b[i] + (step[i] > 0 ? k : 1 - k)) * bricks.size[i];
                    t[i] = (face - p.pos[i]) * inv[i];
                } else {
                    t[i] = std::numeric_limits<T>::infinity();
                }
            }
            const int a = t[0] <= t[1] ? (t[0] <= t[2] ? 0 : 2) : (t[1] <= t[2] ? 1 : 2);
            if (step[a] == 0) { // zero direction: the photon never leaves
                exits = true;
                return t[a];
            }
            travelled = std::max(t[a], travelled);
            ++brickSteps;
            for (int j = 0; j < 3; ++j) {
                if (j == a) {
                    b[j] += step[j] * k;
                } else { // brick of the exit point, inside the cube by construction (the clamp absorbs rounding)
                    const T q = ((p.pos[j] + travelled * p.dir[j]) - ext[2 * j]) * bricks.invSize[j];
                    const std::int64_t c = static_cast<std::int64_t>(std::floor(q));
                    const std::int64_t far = b[j] + (p.dir[j] < 0 ? -(k - 1) : (k - 1));
                    b[j] = std::min(std::max(c, std::min(b[j], far)), std::max(b[j], far));
                }
            }
            for (int j = 0; j < 3; ++j)
                if (b[j] < 0 || b[j] >= static_cast<std::int64_t>(bricks.nb[j])) {
                    exits = true;
                    return travelled;
                }
            const std::uint32_t bb[3] = { static_cast<std::uint32_t>(b[0]), static_cast<std::uint32_t>(b[1]), static_cast<std::uint32_t>(b[2]) };
            if (!airBrick(bb))
                return travelled;
            if (++cubes >= DXMCB200_WALK_MAX_CUBES) // the run goes on, this walk does not: the photon rejoins the Woodcock steps on this face
                return travelled;
        }
    }

    // Walk a photon standing in an air brick to the end of the air run. Returns false when the history ends
    // (left the world, absorbed, roulette); `interacted` reports a real or forced interaction (the walk stops there).
    template <int L>
    bool airWalk(Particle& p, Rng& state, T maxAttenuationInv, bool& updateMaxAttenuation, bool& interacted)
    {
        interacted = false;
        ++walks;
        bool exits = false;
        T remaining = airRunLength(p, exits);
        for (;;) {
            const auto r1 = state.uniform();
            const auto stepLenght = (-std::log(r1) * maxAttenuationInv * T { 10 }) * bricks.invFAir;
            if (!(stepLenght < remaining)) { // no candidate before the end of the run
                for (std::size_t i = 0; i < 3; i++)
                    p.pos[i] += p.dir[i] * remaining;
                return !exits;
            }
            for (std::size_t i = 0; i < 3; i++)
                p.pos[i] += p.dir[i] * stepLenght;
            remaining -= stepLenght;
            ++stats.steps;
            ++walkCandidates;
            if (!inside(p.pos))
                return false;
            const std::size_t bufferIdx = indexFromPosition(p.pos);
            const auto matIdx = material[bufferIdx];
            const auto dens = density[bufferIdx];
            const auto meas = measurement.empty() ? std::uint8_t { 0 } : measurement[bufferIdx];
            ++stats.lookups;
            const auto att = attenuation(matIdx, p.energy);
            const auto attenuationTotal = (((T { 0 } + att[0]) + att[1]) + att[2]) * dens;
            const auto eventProbability = std::min((attenuationTotal * maxAttenuationInv) * bricks.invFAir, T { 1 });
            bool alive = true;
            if (meas == 0) {
                const auto r2 = state.uniform();
                if (r2 < eventProbability) {
                    ++stats.interactions;
                    interacted = true;
                    alive = computeInteractions<L>(att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
                }
            } else {
                ++stats.interactions;
                interacted = true;
                alive = computeInteractionsForced<L>(eventProbability, att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
            }
            if (alive && p.energy * p.weight < ROULETTE_THRESHOLD) {
                const auto r4 = state.uniform();
                if (r4 < ROULETTE_PROBABILITY)
                    alive = false;
                else
                    p.weight *= T { 1 } / (T { 1 } - ROULETTE_PROBABILITY);
            }
            if (!alive || interacted)
                return alive;
        }
    }

    // woodcockParticleTracking<L> with the air walk hooked in (same statements as woodcock<L> above otherwise)
    template <int L>
    void woodcockEmptySpace(Particle& p, Rng& state)
    {
        T maxAttenuationInv = maxAttenuationInverse(p.energy);
        bool updateMaxAttenuation = false;
        bool interacted = false;
        bool continueSampling = true;
        if (inAirBrick(p.pos)) // birth in (or at the face of) an air brick
            continueSampling = airWalk<L>(p, state, maxAttenuationInv, updateMaxAttenuation, interacted);
        while (continueSampling) {
            if (updateMaxAttenuation) {
                maxAttenuationInv = maxAttenuationInverse(p.energy);
                updateMaxAttenuation = false;
            }
            const auto r1 = state.uniform();
            const auto stepLenght = -std::log(r1) * maxAttenuationInv * T { 10 };
            for (std::size_t i = 0; i < 3; i++)
                p.pos[i] += p.dir[i] * stepLenght;
            ++stats.steps;
            if (!inside(p.pos))
                break;
            const std::size_t bufferIdx = indexFromPosition(p.pos);
            const auto matIdx = material[bufferIdx];
            const auto dens = density[bufferIdx];
            const auto meas = measurement.empty() ? std::uint8_t { 0 } : measurement[bufferIdx];
            ++stats.lookups;
            const auto att = attenuation(matIdx, p.energy);
            const auto attenuationTotal = (((T { 0 } + att[0]) + att[1]) + att[2]) * dens;
            const auto eventProbability = attenuationTotal * maxAttenuationInv;
            interacted = false;
            if (meas == 0) {
                const auto r2 = state.uniform();
                if (r2 < eventProbability) {
                    ++stats.interactions;
                    interacted = true;
                    continueSampling = computeInteractions<L>(att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
                }
            } else {
                ++stats.interactions;
                interacted = true;
                continueSampling = computeInteractionsForced<L>(eventProbability, att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
            }
            if (continueSampling && p.energy * p.weight < ROULETTE_THRESHOLD) {
                const auto r4 = state.uniform();
                if (r4 < ROULETTE_PROBABILITY)
                    continueSampling = false;
                else
                    p.weight *= T { 1 } / (T { 1 } - ROULETTE_PROBABILITY);
            }
            // a virtual collision in an air brick: the photon is walked to the end of the air run, then Woodcock steps resume
            if (continueSampling && !interacted && inAirBrick(p.pos)) {
                if (updateMaxAttenuation) {
                    maxAttenuationInv = maxAttenuationInverse(p.energy);
                    updateMaxAttenuation = false;
                }
                continueSampling = airWalk<L>(p, state, maxAttenuationInv, updateMaxAttenuation, interacted);
            }
        }
    }

    // ---- transport<L> over a range of exposures (transport.hpp:729-763)
    template <int L>
    void run(const dxmcb200_exposure* exposures, std::uint64_t begin, std::uint64_t end, std::uint64_t seed, bool perHistoryStreams)
    {
        Rng state;
        state.s[0] = seed;
        state.s[1] = seed ^ 0x9E3779B97F4A7C15ULL; // same seeding as oracle/ref_harness.cpp seededRun
        for (std::uint64_t i = begin; i < end; ++i) {
            const auto& e = exposures[i];
            for (std::uint64_t h = 0; h < e.histories; ++h) {
                if (perHistoryStreams)
                    historyStream(seed, i, h, state.s);
                auto particle = sampleParticle(e, state);
                ++stats.histories;
                if (transportParticleToWorld(particle)) {
                    ++stats.histories_in_world;
                    if (bricks.enabled)
                        woodcockEmptySpace<L>(particle, state);
                    else
                        woodcock<L>(particle, state);
                }
            }
        }
    }
};

} // namespace

extern "C" {

struct dxmc_oracle;

dxmc_oracle* dxmc_oracle_create() { return reinterpret_cast<dxmc_oracle*>(new Oracle); }
void dxmc_oracle_destroy(dxmc_oracle* o) { delete reinterpret_cast<Oracle*>(o); }

int dxmc_oracle_set_world(dxmc_oracle* h, const dxmcb200_world* w)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !w || !w->density || !w->material)
        return DXMCB200_ERR_ARG;
    for (int i = 0; i < 3; ++i) {
        o->dim[i] = w->dim[i];
        o->spacing[i] = w->spacing[i];
    }
    for (int i = 0; i < 6; ++i)
        o->ext[i] = w->extent_safe[i];
    const auto n = o->nVoxels();
    o->density.assign(w->density, w->density + n);
    o->material.assign(w->material, w->material + n);
    if (w->measurement)
        o->measurement.assign(w->measurement, w->measurement + n);
    else
        o->measurement.clear();
    o->dose.assign(n, 0);
    o->variance.assign(n, 0);
    o->nEvents.assign(n, 0);
    o->fixedEnergy.assign(n, 0);
    o->fixedEnergySq.assign(n, 0);
    return DXMCB200_OK;
}

int dxmc_oracle_set_fixed_point(dxmc_oracle* h, int energyBits, int energySqBits)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    o->energyBits = energyBits;
    o->energySqBits = energySqBits;
    return DXMCB200_OK;
}

int dxmc_oracle_get_fixed(dxmc_oracle* h, int64_t* energy, uint64_t* energySq)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    if (energy)
        std::memcpy(energy, o->fixedEnergy.data(), o->fixedEnergy.size() * sizeof(int64_t));
    if (energySq)
        std::memcpy(energySq, o->fixedEnergySq.data(), o->fixedEnergySq.size() * sizeof(uint64_t));
    return DXMCB200_OK;
}

int dxmc_oracle_set_luts(dxmc_oracle* h, const dxmcb200_luts* l)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !l)
        return DXMCB200_ERR_ARG;
    o->nMat = l->n_materials;
    o->nSeg = l->n_segments;
    o->linearIndex = l->linear_index;
    o->linearStep = l->linear_step;
    o->linearEnergy = l->linear_energy;
    o->knots.assign(l->knots, l->knots + l->n_segments);
    o->coeff.assign(l->coefficients, l->coefficients + static_cast<std::size_t>(l->n_materials) * l->n_segments * 6);
    o->maxCoeff.assign(l->max_coefficients, l->max_coefficients + static_cast<std::size_t>(l->n_segments) * 2);
    o->rita.assign(l->rita, l->rita + static_cast<std::size_t>(l->n_materials) * 4 * DXMCB200_RITA_N);
    o->spline.assign(l->spline, l->spline + static_cast<std::size_t>(l->n_materials) * DXMCB200_SPLINE_FLOATS);
    o->shells.assign(l->shells, l->shells + static_cast<std::size_t>(l->n_materials) * DXMCB200_SHELLS * DXMCB200_SHELL_FLOATS);
    return DXMCB200_OK;
}

int dxmc_oracle_set_beam_tables(dxmc_oracle* h, uint32_t nS, const dxmcb200_spectrum* s, uint32_t nH, const dxmcb200_heel* hl, uint32_t nB,
    const dxmcb200_bowtie* b)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    o->spectra.clear();
    o->heels.clear();
    o->bowties.clear();
    for (uint32_t i = 0; i < nS; ++i)
        o->spectra.push_back({ { s[i].probs, s[i].probs + s[i].n }, { s[i].energies, s[i].energies + s[i].n }, { s[i].alias, s[i].alias + s[i].n } });
    for (uint32_t i = 0; i < nH; ++i)
        o->heels.push_back({ hl[i].energy_start, hl[i].energy_step, hl[i].energy_size, hl[i].angle_start, hl[i].angle_step, hl[i].angle_size,
            { hl[i].weights, hl[i].weights + static_cast<std::size_t>(hl[i].energy_size) * hl[i].angle_size } });
    for (uint32_t i = 0; i < nB; ++i)
        o->bowties.push_back({ { b[i].angles, b[i].angles + b[i].n }, { b[i].weights, b[i].weights + b[i].n } });
    return DXMCB200_OK;
}

int dxmc_oracle_clear(dxmc_oracle* h)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    std::fill(o->dose.begin(), o->dose.end(), 0.0f);
    std::fill(o->variance.begin(), o->variance.end(), 0.0f);
    std::fill(o->nEvents.begin(), o->nEvents.end(), 0u);
    std::fill(o->fixedEnergy.begin(), o->fixedEnergy.end(), 0);
    std::fill(o->fixedEnergySq.begin(), o->fixedEnergySq.end(), 0u);
    o->stats = {};
    o->walks = o->brickSteps = o->walkCandidates = 0;
    return DXMCB200_OK;
}

// per_history_streams 0: one sequential PCG32 stream {seed, seed^golden} (reference worker loop);
//                     1: dxmcb200_history_stream(seed, exposure, history) per history (CUDA kernels)
int dxmc_oracle_run(dxmc_oracle* h, const dxmcb200_exposure* exposures, uint64_t begin, uint64_t end, int model, uint64_t seed, int per_history_streams)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !exposures || model < 0 || model > 2)
        return DXMCB200_ERR_ARG;
    if (model == 0)
        o->run<0>(exposures, begin, end, seed, per_history_streams != 0);
    else if (model == 1)
        o->run<1>(exposures, begin, end, seed, per_history_streams != 0);
    else
        o->run<2>(exposures, begin, end, seed, per_history_streams != 0);
    return DXMCB200_OK;
}

// tracking 0: the reference's Woodcock loop (global majorant) everywhere; 1: Woodcock + empty-space traversal through
// air bricks of about brick_mm (the product's default tracking). Call after set_world and set_luts.
int dxmc_oracle_set_tracking(dxmc_oracle* h, int tracking, float brick_mm)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || tracking < 0 || tracking > 1 || (tracking == 1 && (!(brick_mm > 0) || o->nMat == 0 || o->density.empty())))
        return DXMCB200_ERR_ARG;
    if (tracking == 1)
        o->buildBricks(brick_mm);
    else
        o->bricks = {};
    return DXMCB200_OK;
}

// the brick grid of the empty-space traversal: shift[3], nb[3], f_air, and (any pointer may be NULL) the per-material ratios
// [n_materials], the per-brick maxima [nb2*nb1*nb0] and the air flags [nb2*nb1*nb0]
int dxmc_oracle_get_bricks(dxmc_oracle* h, uint32_t shift[3], uint32_t nb[3], float* f_air, float* ratio, float* brick_max, uint8_t* air)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !shift || !nb || !f_air)
        return DXMCB200_ERR_ARG;
    for (int i = 0; i < 3; ++i) {
        shift[i] = o->bricks.shift[i];
        nb[i] = o->bricks.nb[i];
    }
    *f_air = o->bricks.enabled ? o->bricks.fAir : 0.0f;
    if (ratio)
        std::copy(o->bricks.ratio.begin(), o->bricks.ratio.end(), ratio);
    if (brick_max)
        std::copy(o->bricks.brickMax.begin(), o->bricks.brickMax.end(), brick_max);
    if (air)
        std::copy(o->bricks.air.begin(), o->bricks.air.end(), air);
    return DXMCB200_OK;
}

// per octant and brick, the edge of the largest all-air cube cornered there [8][nb2*nb1*nb0]
int dxmc_oracle_get_brick_distance(dxmc_oracle* h, uint8_t* distance)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !distance)
        return DXMCB200_ERR_ARG;
    std::copy(o->bricks.distance.begin(), o->bricks.distance.end(), distance);
    return DXMCB200_OK;
}

// instrumentation of the empty-space traversal: out[0] air walks, out[1] bricks crossed, out[2] collision candidates in air
int dxmc_oracle_get_walk_stats(dxmc_oracle* h, uint64_t out[3])
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !out)
        return DXMCB200_ERR_ARG;
    out[0] = o->walks;
    out[1] = o->brickSteps;
    out[2] = o->walkCandidates;
    return DXMCB200_OK;
}

// raw sums as the reference accumulates them (float): dose = sum e*w [keV], variance = sum (e*w)^2
int dxmc_oracle_get_raw(dxmc_oracle* h, float* dose, uint32_t* nEvents, float* variance)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    const auto n = o->nVoxels();
    if (dose)
        std::memcpy(dose, o->dose.data(), n * sizeof(float));
    if (nEvents)
        std::memcpy(nEvents, o->nEvents.data(), n * sizeof(uint32_t));
    if (variance)
        std::memcpy(variance, o->variance.data(), n * sizeof(float));
    return DXMCB200_OK;
}

int dxmc_oracle_get_stats(dxmc_oracle* h, dxmcb200_stats* s)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !s)
        return DXMCB200_ERR_ARG;
    *s = o->stats;
    return DXMCB200_OK;
}

// a9/a10 on the host
int dxmc_oracle_eval_attenuation(dxmc_oracle* h, uint64_t n, const uint8_t* material, const float* energy, float* out3, float* outMax)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) {
        const auto a = o->attenuation(material[i], energy[i]);
        out3[3 * i] = a[0];
        out3[3 * i + 1] = a[1];
        out3[3 * i + 2] = a[2];
        outMax[i] = o->maxAttenuationInverse(energy[i]);
    }
    return DXMCB200_OK;
}

// a5/a6/a7 for fixed rays and a fixed list of step lengths (same contract as dxmcb200_trace_indices)
// the air run of fixed rays, as dxmcb200_trace_air_runs reports it (include/dxmcb200.h)
int dxmc_oracle_trace_air_runs(dxmc_oracle* h, uint64_t nRays, const float* pos, const float* dir, float* outLength, uint32_t* outInfo, float* outEnd)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !o->bricks.enabled)
        return DXMCB200_ERR_STATE;
    for (uint64_t r = 0; r < nRays; ++r) {
        Particle p {};
        for (int i = 0; i < 3; ++i) {
            p.pos[i] = pos[3 * r + i];
            p.dir[i] = dir[3 * r + i];
        }
        float length = 0;
        uint32_t info = 0;
        if (o->transportParticleToWorld(p)) {
            info |= 1u << 18;
            if (o->inAirBrick(p.pos)) {
                bool exits = false;
                const auto before = o->brickSteps;
                length = o->airRunLength(p, exits);
                info |= static_cast<uint32_t>(o->brickSteps - before) | (exits ? 1u << 16 : 0u) | (1u << 17);
                for (int i = 0; i < 3; ++i)
                    p.pos[i] += p.dir[i] * length;
            }
        }
        outLength[r] = length;
        outInfo[r] = info;
        for (int i = 0; i < 3; ++i)
            outEnd[3 * r + i] = p.pos[i];
    }
    return DXMCB200_OK;
}

int dxmc_oracle_trace_indices(dxmc_oracle* h, uint64_t nRays, const float* pos, const float* dir, uint32_t nSteps, const float* steps,
    int64_t* outIdx, float* outEntry)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    for (uint64_t r = 0; r < nRays; ++r) {
        Particle p {};
        for (int i = 0; i < 3; ++i) {
            p.pos[i] = pos[3 * r + i];
            p.dir[i] = dir[3 * r + i];
        }
        bool in = o->transportParticleToWorld(p);
        for (int i = 0; i < 3; ++i)
            outEntry[3 * r + i] = p.pos[i];
        int64_t* out = outIdx + r * (nSteps + 1);
        out[0] = (in && o->inside(p.pos)) ? static_cast<int64_t>(o->indexFromPosition(p.pos)) : -1;
        for (uint32_t k = 0; k < nSteps; ++k) {
            if (in) {
                for (int i = 0; i < 3; ++i)
                    p.pos[i] += p.dir[i] * steps[k];
                in = o->inside(p.pos);
            }
            out[k + 1] = in ? static_cast<int64_t>(o->indexFromPosition(p.pos)) : -1;
        }
    }
    return DXMCB200_OK;
}

// a4 with per-history streams (same contract as dxmcb200_sample_particles)
int dxmc_oracle_sample_particles(dxmc_oracle* h, const dxmcb200_exposure* e, uint64_t exposureIndex, uint64_t seed, uint64_t n, float* out)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !e)
        return DXMCB200_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) {
        Rng state;
        historyStream(seed, exposureIndex, i, state.s);
        const auto p = o->sampleParticle(*e, state);
        float* q = out + 8 * i;
        for (int k = 0; k < 3; ++k) {
            q[k] = p.pos[k];
            q[3 + k] = p.dir[k];
        }
        q[6] = p.energy;
        q[7] = p.weight;
    }
    return DXMCB200_OK;
}

// a13-a16 (same contract as dxmcb200_sample_interaction)
int dxmc_oracle_sample_interaction(dxmc_oracle* h, int kind, int model, uint8_t material, float energy, uint64_t seed, uint64_t n, float* out)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o)
        return DXMCB200_ERR_ARG;
    for (uint64_t i = 0; i < n; ++i) {
        Rng state;
        historyStream(seed, 0, i, state.s);
        Particle p {};
        p.dir[2] = 1;
        p.energy = energy;
        p.weight = 1;
        float imparted = 0;
        auto call = [&](auto tag) {
            constexpr int L = decltype(tag)::value;
            if (kind == 0)
                imparted = o->photoAbsorption<L>(p, material, state);
            else if (kind == 1)
                imparted = o->comptonScatter<L>(p, material, state);
            else
                o->rayleighScatter<L>(p, material, state);
        };
        if (model == 0)
            call(std::integral_constant<int, 0> {});
        else if (model == 1)
            call(std::integral_constant<int, 1> {});
        else
            call(std::integral_constant<int, 2> {});
        float* q = out + 5 * i;
        q[0] = imparted;
        q[1] = p.energy;
        q[2] = p.dir[0];
        q[3] = p.dir[1];
        q[4] = p.dir[2];
    }
    return DXMCB200_OK;
}

void dxmc_oracle_history_stream(uint64_t seed, uint64_t exposure, uint64_t history, uint64_t out[2]) { historyStream(seed, exposure, history, out); }

// the host libm's float log10 — the first operation of every reference LUT look-up (attenuationinterpolator.hpp:210);
// exposed so tests can tell where it is not correctly rounded
void dxmc_oracle_log10f(uint64_t n, const float* x, float* out)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = std::log10(x[i]);
}

} // extern "C"
