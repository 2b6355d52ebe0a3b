// brick_majorant_probe.cpp — CPU experiment for DESIGN.md section 8 item 1: Woodcock tracking with REGIONAL majorants.
//
// Not product code and not part of the parity oracle: it includes the CPU restatement (oracle/dxmc_oracle.cpp) as is and
// adds ONE alternative tracking loop next to woodcock<L>. The world is cut into bricks of B^3 voxels; brick b gets a
// factor f_b <= 1 with  max_voxels_in_b(rho * mu_total(E)) <= f_b * majorant(E)  for every energy of the tables, and
// free paths inside the brick are sampled against f_b * majorant. A sampled path that would leave the brick is cut at
// the brick face and re-sampled there (delta tracking is memoryless), so the distribution of real collision sites is
// unchanged while most virtual collisions in air and soft tissue disappear. Everything else (interaction sampling,
// forced interactions, roulette after a virtual collision, scoring) is the restatement's own code.
//
//   g++ -O2 -std=c++20 -fPIC -shared -ffp-contract=off -Iinclude tools/brick_majorant_probe.cpp -o tools/libbrickprobe.so
#include "../oracle/dxmc_oracle.cpp"

namespace {

struct BrickTracker {
    Oracle& o;
    std::size_t B;
    std::size_t nb[3];
    std::vector<T> factor; // f_b
    std::uint64_t collisionSteps = 0, faceSteps = 0;

    BrickTracker(Oracle& oracle, std::size_t brick)
        : o(oracle)
        , B(brick)
    {
        for (int i = 0; i < 3; ++i)
            nb[i] = (o.dim[i] + B - 1) / B;
        const std::size_t nBricks = nb[0] * nb[1] * nb[2];
        // per brick and material: the largest density
        std::vector<T> maxDensity(nBricks * o.nMat, T { 0 });
        for (std::size_t z = 0; z < o.dim[2]; ++z)
            for (std::size_t y = 0; y < o.dim[1]; ++y)
                for (std::size_t x = 0; x < o.dim[0]; ++x) {
                    const std::size_t v = (z * o.dim[1] + y) * o.dim[0] + x;
                    const std::size_t b = ((z / B) * nb[1] + y / B) * nb[0] + x / B;
                    T& m = maxDensity[b * o.nMat + o.material[v]];
                    m = std::max(m, o.density[v]);
                }
        // The ratio rho*mu_m(E) / majorant(E) is a power law inside every table segment, so its maximum over a segment
        // sits at a segment end: evaluate just inside both ends of every segment.
        std::vector<T> energies;
        T lower = 0; // log10 of the lowest table energy (1 keV)
        for (std::size_t k = 0; k < o.nSeg; ++k) {
            const T upper = o.knots[k];
            energies.push_back(std::pow(T { 10 }, lower + (upper - lower) * T { 1e-3 }));
            energies.push_back(std::pow(T { 10 }, upper - (upper - lower) * T { 1e-3 }));
            lower = upper;
        }
        factor.assign(nBricks, T { 0 });
        for (const T e : energies) {
            const T inv = o.maxAttenuationInverse(e);
            std::vector<T> total(o.nMat);
            for (std::size_t m = 0; m < o.nMat; ++m) {
                const auto a = o.attenuation(m, e);
                total[m] = ((T { 0 } + a[0]) + a[1]) + a[2];
            }
            for (std::size_t b = 0; b < nBricks; ++b)
                for (std::size_t m = 0; m < o.nMat; ++m)
                    factor[b] = std::max(factor[b], maxDensity[b * o.nMat + m] * total[m] * inv);
        }
        for (auto& f : factor)
            f = std::min(T { 1 }, std::max(f * T { 1.001 }, T { 1e-6 }));
    }

    std::size_t brickAxis(const T pos[3], int axis) const
    {
        const T rel = (pos[axis] - o.ext[2 * axis]) / o.spacing[axis];
        const auto voxel = std::min<std::size_t>(rel > 0 ? static_cast<std::size_t>(rel) : 0, o.dim[axis] - 1);
        return voxel / B;
    }

    template <int L>
    void track(Particle& p, Rng& state)
    {
        T maxAttenuationInv = 0;
        bool updateMaxAttenuation = true;
        bool continueSampling = true;
        while (continueSampling) {
            if (updateMaxAttenuation) {
                maxAttenuationInv = o.maxAttenuationInverse(p.energy);
                updateMaxAttenuation = false;
            }
            const std::size_t ib[3] = { brickAxis(p.pos, 0), brickAxis(p.pos, 1), brickAxis(p.pos, 2) };
            const T f = factor[(ib[2] * nb[1] + ib[1]) * nb[0] + ib[0]];
            // distance (in units of |dir|) to the face through which the photon leaves the brick, and that face
            T toFace = std::numeric_limits<T>::max();
            int exitAxis = 0;
            T exitCoordinate = 0;
            for (int i = 0; i < 3; ++i) {
                if (std::abs(p.dir[i]) < T { 1e-9 })
                    continue;
                const T lo = o.ext[2 * i] + static_cast<T>(ib[i] * B) * o.spacing[i];
                const T hi = lo + static_cast<T>(B) * o.spacing[i];
                const T t = std::max(((p.dir[i] > 0 ? hi : lo) - p.pos[i]) / p.dir[i], T { 0 });
                if (t < toFace) {
                    toFace = t;
                    exitAxis = i;
                    exitCoordinate = p.dir[i] > 0 ? hi + T { 1e-3 } : lo - T { 1e-3 }; // one micrometre behind the face
                }
            }
            const auto r1 = state.uniform();
            const auto stepLenght = -std::log(r1) * maxAttenuationInv * T { 10 } / f;
            if (stepLenght >= toFace) { // no collision inside this brick: continue from just behind its face
                for (int i = 0; i < 3; ++i)
                    p.pos[i] += p.dir[i] * toFace;
                p.pos[exitAxis] = exitCoordinate;
                ++faceSteps;
                continueSampling = o.inside(p.pos);
                continue;
            }
            for (int i = 0; i < 3; ++i)
                p.pos[i] += p.dir[i] * stepLenght;
            ++collisionSteps;
            ++o.stats.steps;
            if (!o.inside(p.pos))
                break;
            const std::size_t bufferIdx = o.indexFromPosition(p.pos);
            const auto matIdx = o.material[bufferIdx];
            const auto dens = o.density[bufferIdx];
            const auto meas = o.measurement.empty() ? std::uint8_t { 0 } : o.measurement[bufferIdx];
            ++o.stats.lookups;
            const auto att = o.attenuation(matIdx, p.energy);
            const auto attenuationTotal = (((T { 0 } + att[0]) + att[1]) + att[2]) * dens;
            const auto eventProbability = attenuationTotal * maxAttenuationInv / f;
            if (meas == 0) {
                const auto r2 = state.uniform();
                if (r2 < eventProbability) {
                    ++o.stats.interactions;
                    continueSampling = o.template computeInteractions<L>(att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
                }
            } else {
                ++o.stats.interactions;
                continueSampling = o.template computeInteractionsForced<L>(eventProbability, att, p, matIdx, bufferIdx, state, updateMaxAttenuation);
            }
            if (continueSampling && p.energy * p.weight < ROULETTE_THRESHOLD) {
                const auto r4 = state.uniform();
                if (r4 < ROULETTE_PROBABILITY)
                    continueSampling = false;
                else
                    p.weight *= T { 1 } / (T { 1 } - ROULETTE_PROBABILITY);
            }
        }
    }

    template <int L>
    void run(const dxmcb200_exposure* exposures, std::uint64_t begin, std::uint64_t end, std::uint64_t seed)
    {
        Rng state;
        for (std::uint64_t i = begin; i < end; ++i)
            for (std::uint64_t h = 0; h < exposures[i].histories; ++h) {
                historyStream(seed, i, h, state.s);
                auto particle = o.sampleParticle(exposures[i], state);
                ++o.stats.histories;
                if (o.transportParticleToWorld(particle)) {
                    ++o.stats.histories_in_world;
                    track<L>(particle, state);
                }
            }
    }
};

} // namespace

// out[0] = collision steps (voxel look-ups), out[1] = brick-face steps, out[2] = mean factor over bricks
extern "C" int brick_probe_run(dxmc_oracle* h, const dxmcb200_exposure* exposures, uint64_t begin, uint64_t end, int model, uint64_t seed,
    uint64_t brickVoxels, double* out)
{
    auto* o = reinterpret_cast<Oracle*>(h);
    if (!o || !exposures || brickVoxels == 0 || model < 0 || model > 2)
        return DXMCB200_ERR_ARG;
    BrickTracker t(*o, brickVoxels);
    if (model == 0)
        t.run<0>(exposures, begin, end, seed);
    else if (model == 1)
        t.run<1>(exposures, begin, end, seed);
    else
        t.run<2>(exposures, begin, end, seed);
    if (out) {
        out[0] = static_cast<double>(t.collisionSteps);
        out[1] = static_cast<double>(t.faceSteps);
        double mean = 0;
        for (const auto f : t.factor)
            mean += f;
        out[2] = mean / static_cast<double>(t.factor.size());
    }
    return DXMCB200_OK;
}
