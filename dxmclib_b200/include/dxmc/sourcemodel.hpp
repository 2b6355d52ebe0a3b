// sourcemodel.hpp — every source of the library as ONE plain parameter block plus one function that evaluates exposure i
// from it. Compiled twice: by the host compiler (the Source classes in sourcebase.hpp / dapsource.hpp / ctsource.hpp are
// setters over this block, and their getExposure(i) is evaluate()) and by nvcc (csrc/transport.cu: exposureKernel
// evaluates all exposures of a run on the device, SURVEY 8f3, so that no per-exposure host work or upload remains).
//
// What it computes is the reference's getExposure(i) for each of its source types followed by the Exposure constructor's
// normalisation and, on request, Exposure::alignToDirectionCosines — reference include/dxmc/source.hpp:211-223 (pencil),
// :279-291 (isotropic), :355-376 (isotropic CT), :643-653, 664-673 (DX), :733-760, 767-776 (cone-beam CT), :1206-1257
// (CT axial), :1296-1340 (CT spiral), :1384-1468, 1498-1564 (dual source), :1627-1669 (topogram); exposure.hpp:59-62,
// 268-278; beamfilters.hpp:400-426 (XCare), :708-721 (AEC) — with the reference's floating-point operation order, so that
// the host evaluation in T = float gives the reference's bits. On the device the same expressions run with round-to-nearest
// intrinsics (no FMA contraction) and double-precision sin / cos / atan rounded to float, which agrees with the host to
// the last one or two units.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define DXMC_MODEL_HD __host__ __device__ inline
#else
#define DXMC_MODEL_HD inline
#include <cmath>
#endif

namespace dxmc::model {

enum Motion : std::uint32_t {
    FIXED = 0, // every exposure is the same: pencil, isotropic, radiography
    ORBIT = 1, // focal spot and beam frame turn about an axis: isotropic CT, cone-beam CT
    GANTRY_AXIAL = 2, // CT gantry, step-and-shoot
    GANTRY_SPIRAL = 3, // CT gantry, continuous table feed
    GANTRY_TOPOGRAM = 4 // CT gantry parked at the start angle, table moving
};

template <typename R>
struct SourceParams {
    std::uint32_t motion = FIXED;
    std::uint32_t tubes = 1; // gantry: 2 = dual source, exposures alternate tube A (even index), tube B (odd index)
    std::uint64_t exposures = 1;
    std::uint64_t histories = 0;
    R position[3] = { 0, 0, 0 }; // Source::position(): the focal spot (FIXED, ORBIT without focal offset) or the isocentre
    R cosines[6] = { 1, 0, 0, 0, 1, 0 };
    R collimation[4] = { 0, 0, 0, 0 }; // FIXED / ORBIT: x0 x1 y0 y1
    R monoEnergy = 0; // used when spectrum[0] < 0
    R focalOffset = 0; // FIXED / ORBIT: focal spot = position - beam * focalOffset (tube-based projection sources)
    std::int32_t spectrum[2] = { -1, -1 }, heel[2] = { -1, -1 }, bowtie[2] = { -1, -1 }; // beam-table indices per tube
    // ORBIT
    std::uint32_t orbitFullTurn = 0; // 1: angle_i = (2 i) pi / exposures about z through the origin; 0: angle_i = i * orbitStep about the y cosine through `position`
    R orbitStep = 0;
    // GANTRY
    R sdd[2] = { 0, 0 }, fov[2] = { 0, 0 }, startAngle[2] = { 0, 0 }, tubeWeight[2] = { 1, 1 };
    R beamWidth = 0; // collimation along z at the isocentre [mm]
    R angleStep = 0, pitch = 1, tableStep = 0, tilt = 0, scanLength = 0;
    // modulation of the beam weight
    std::uint32_t xcare = 0;
    R xcareAngle = 0, xcareSpan = 0, xcareRamp = 0, xcareLow = 1;
    std::uint32_t aecSize = 0; // entries of the position table handed to evaluate() (0: no AEC)
    R aecMin = 0, aecMax = 0, aecStep = 0;
    // basis of the world the exposures are expressed in (Exposure::alignToDirectionCosines); align = 0 leaves them global
    std::uint32_t align = 0;
    R worldCosines[6] = { 1, 0, 0, 0, 1, 0 };
};

// one exposure as Transport hands it to the kernels: include/dxmcb200.h dxmcb200_exposure without the padding
template <typename R>
struct ExposureValues {
    R position[3], cosines[6], beam[3], collimation[4];
    R weight, monoEnergy;
    std::int32_t spectrum, heel, bowtie;
    std::uint64_t histories;
};

namespace detail {
    // arithmetic that must not be contracted into FMAs on the device
    template <typename R>
    DXMC_MODEL_HD R mul(R a, R b)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(R) == 4)
            return __fmul_rn(a, b);
        else
            return __dmul_rn(a, b);
#else
        return a * b;
#endif
    }
    template <typename R>
    DXMC_MODEL_HD R add(R a, R b)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(R) == 4)
            return __fadd_rn(a, b);
        else
            return __dadd_rn(a, b);
#else
        return a + b;
#endif
    }
    template <typename R>
    DXMC_MODEL_HD R sub(R a, R b) { return add(a, -b); }
    template <typename R>
    DXMC_MODEL_HD R div(R a, R b)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(R) == 4)
            return __fdiv_rn(a, b);
        else
            return __ddiv_rn(a, b);
#else
        return a / b;
#endif
    }
    template <typename R>
    DXMC_MODEL_HD void sinCos(R angle, R& s, R& c)
    {
#if defined(__CUDA_ARCH__)
        double ds, dc;
        sincos(static_cast<double>(angle), &ds, &dc);
        s = static_cast<R>(ds);
        c = static_cast<R>(dc);
#else
        s = std::sin(angle);
        c = std::cos(angle);
#endif
    }
    template <typename R>
    DXMC_MODEL_HD R arcTan(R x)
    {
#if defined(__CUDA_ARCH__)
        return static_cast<R>(atan(static_cast<double>(x)));
#else
        return std::atan(x);
#endif
    }
    template <typename R>
    DXMC_MODEL_HD R squareRoot(R x)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(R) == 4)
            return __fsqrt_rn(x);
        else
            return __dsqrt_rn(x);
#else
        return std::sqrt(x);
#endif
    }
    template <typename R>
    DXMC_MODEL_HD R floatMod(R x, R y)
    {
#if defined(__CUDA_ARCH__)
        return static_cast<R>(fmod(static_cast<double>(x), static_cast<double>(y))); // exact for float operands
#else
        return std::fmod(x, y);
#endif
    }

    template <typename R>
    DXMC_MODEL_HD R dot3(const R* a, const R* b) { return add(add(mul(a[0], b[0]), mul(a[1], b[1])), mul(a[2], b[2])); }

    template <typename R>
    DXMC_MODEL_HD void cross3(const R* a, const R* b, R* out)
    {
        out[0] = sub(mul(a[1], b[2]), mul(a[2], b[1]));
        out[1] = sub(mul(a[2], b[0]), mul(a[0], b[2]));
        out[2] = sub(mul(a[0], b[1]), mul(a[1], b[0]));
    }

    template <typename R>
    DXMC_MODEL_HD void unit3(R* v)
    {
        const R inv = div(R { 1 }, squareRoot(dot3(v, v)));
        v[0] = mul(v[0], inv);
        v[1] = mul(v[1], inv);
        v[2] = mul(v[2], inv);
    }

    // turn v about the unit vector k by `angle` (Rodrigues): v cos + k (1 - cos)(v.k) + (k x v) sin
    template <typename R>
    DXMC_MODEL_HD void turn3(R* v, const R* k, R angle)
    {
        R s, c;
        sinCos(angle, s, c);
        const R along = mul(sub(R { 1 }, c), dot3(v, k));
        const R o0 = add(add(mul(c, v[0]), mul(along, k[0])), mul(s, sub(mul(k[1], v[2]), mul(k[2], v[1]))));
        const R o1 = add(add(mul(c, v[1]), mul(along, k[1])), mul(s, add(mul(-k[0], v[2]), mul(k[2], v[0]))));
        const R o2 = add(add(mul(c, v[2]), mul(along, k[2])), mul(s, sub(mul(k[0], v[1]), mul(k[1], v[0]))));
        v[0] = o0;
        v[1] = o1;
        v[2] = o2;
    }

    // components of v along three axes
    template <typename R>
    DXMC_MODEL_HD void project3(const R* e0, const R* e1, const R* e2, R* v)
    {
        const R a = dot3(e0, v), b = dot3(e1, v), c = dot3(e2, v);
        v[0] = a;
        v[1] = b;
        v[2] = c;
    }

    template <typename R>
    DXMC_MODEL_HD R lerp(R x0, R x1, R y0, R y1, R x) { return add(y0, div(mul(sub(y1, y0), sub(x, x0)), sub(x1, x0))); }

    // organ-based tube current modulation: low weight inside `span` centred on `centre`, linear ramps, expectation 1 over a turn
    template <typename R>
    DXMC_MODEL_HD R xcareWeight(const SourceParams<R>& s, R angle)
    {
        constexpr R pi = R(3.14159265358979323846);
        constexpr R twoPi = R { 2 } * pi;
        R a = floatMod(add(sub(angle, s.xcareAngle), pi), twoPi);
        if (a < 0)
            a = add(a, twoPi);
        const R high = div(add(sub(twoPi, mul(s.xcareSpan, s.xcareLow)), mul(s.xcareLow, s.xcareRamp)), add(sub(twoPi, s.xcareSpan), s.xcareRamp));
        const R begin = sub(pi, mul(s.xcareSpan, R(0.5)));
        if (a < begin)
            return high;
        const R rampDown = add(begin, s.xcareRamp);
        if (a < rampDown)
            return lerp(begin, rampDown, high, s.xcareLow, a);
        const R rampUp = sub(add(rampDown, s.xcareSpan), s.xcareRamp);
        if (a < rampUp)
            return s.xcareLow;
        const R end = add(begin, s.xcareSpan);
        if (a < end)
            return lerp(rampUp, end, s.xcareLow, high, a);
        return high;
    }

    // tube current along z: piecewise linear in a table of slice intensities
    template <typename R>
    DXMC_MODEL_HD R aecWeight(const SourceParams<R>& s, const R* table, R z)
    {
        if (z < add(s.aecMin, s.aecStep))
            return table[0];
        if (z >= sub(s.aecMax, s.aecStep))
            return table[s.aecSize - 1];
        const std::uint64_t k = static_cast<std::uint64_t>(div(sub(z, s.aecMin), s.aecStep));
        const R x0 = add(mul(static_cast<R>(k), s.aecStep), s.aecMin);
        return lerp(x0, add(x0, s.aecStep), table[k], table[k + 1], z);
    }
} // namespace detail

// Exposure i of the source described by `s`. aecTable: s.aecSize slice intensities (ignored when aecSize == 0).
template <typename R>
DXMC_MODEL_HD void evaluate(const SourceParams<R>& s, const R* aecTable, std::uint64_t i, ExposureValues<R>& e)
{
    using namespace detail;
    constexpr R pi = R(3.14159265358979323846);
    R pos[3] = { s.position[0], s.position[1], s.position[2] };
    R frame[6] = { s.cosines[0], s.cosines[1], s.cosines[2], s.cosines[3], s.cosines[4], s.cosines[5] };
    R weight = 1;
    std::uint32_t tube = 0;
    e.collimation[0] = s.collimation[0];
    e.collimation[1] = s.collimation[1];
    e.collimation[2] = s.collimation[2];
    e.collimation[3] = s.collimation[3];

    if (s.motion == FIXED || s.motion == ORBIT) {
        if (s.focalOffset != 0) { // tube-based projection source: the focal spot sits upstream of the reference point
            R beam[3];
            cross3(frame, frame + 3, beam);
            for (int k = 0; k < 3; ++k)
                pos[k] = sub(s.position[k], mul(beam[k], s.focalOffset));
        }
        if (s.motion == ORBIT) {
            if (s.orbitFullTurn) { // about z through the origin, a full turn over all exposures
                const R axis[3] = { 0, 0, 1 };
                const R angle = div(mul(static_cast<R>(i * 2), pi), static_cast<R>(s.exposures));
                turn3(pos, axis, angle);
                turn3(frame, axis, angle);
                turn3(frame + 3, axis, angle);
            } else { // about the y cosine through the reference point
                const R axis[3] = { s.cosines[3], s.cosines[4], s.cosines[5] };
                const R angle = mul(static_cast<R>(i), s.orbitStep);
                for (int k = 0; k < 3; ++k)
                    pos[k] = sub(pos[k], s.position[k]);
                turn3(pos, axis, angle);
                for (int k = 0; k < 3; ++k)
                    pos[k] = add(pos[k], s.position[k]);
                turn3(frame, axis, angle);
                turn3(frame + 3, axis, angle);
            }
        }
    } else {
        // CT gantry. Exposure index -> tube, gantry angle, table position.
        std::uint64_t index = i;
        if (s.tubes == 2) {
            tube = static_cast<std::uint32_t>(i & 1u);
            index = i >> 1;
        }
        R angle, table;
        if (s.motion == GANTRY_AXIAL) {
            const std::uint64_t perTurn = static_cast<std::uint64_t>(div(mul(R { 2 }, pi), s.angleStep));
            const std::uint64_t turn = index / perTurn;
            angle = add(s.startAngle[tube], mul(s.angleStep, static_cast<R>(index - turn * perTurn)));
            table = mul(s.tableStep, static_cast<R>(turn));
        } else if (s.motion == GANTRY_SPIRAL) {
            angle = add(s.startAngle[tube], mul(s.angleStep, static_cast<R>(index)));
            table = div(mul(mul(mul(static_cast<R>(index), s.angleStep), s.beamWidth), s.pitch), mul(R { 2 }, pi));
        } else {
            angle = s.startAngle[0];
            table = mul(div(s.scanLength, static_cast<R>(s.exposures - 1)), static_cast<R>(i));
        }
        // The focal spot starts at (0, -sdd/2, 0) (tube A's radius for both tubes); the gantry tilt turns the rotation axis
        // (the y cosine), the x cosine and a copy of the focal spot about x; the focal spot then turns about the tilted axis by
        // the gantry angle and moves along z by the table position plus the z the tilt gave the copy.
        const R tiltAxis[3] = { 1, 0, 0 };
        pos[0] = 0;
        pos[1] = div(-s.sdd[0], R { 2 });
        pos[2] = 0;
        R lifted[3] = { pos[0], pos[1], pos[2] };
        turn3(lifted, tiltAxis, s.tilt);
        turn3(frame + 3, tiltAxis, s.tilt);
        turn3(frame, tiltAxis, s.tilt);
        turn3(pos, frame + 3, angle);
        pos[2] = add(pos[2], add(table, lifted[2]));
        turn3(frame, frame + 3, angle);
        for (int k = 0; k < 3; ++k)
            pos[k] = add(pos[k], s.position[k]);
        // full fan and cone angles seen from the focal spot, sdd/2 from the isocentre
        const R fan = mul(arcTan(div(s.fov[tube], s.sdd[tube])), R { 2 });
        const R cone = mul(arcTan(div(s.beamWidth, s.sdd[tube])), R { 2 });
        e.collimation[0] = div(-fan, R { 2 });
        e.collimation[1] = div(fan, R { 2 });
        e.collimation[2] = div(-cone, R { 2 });
        e.collimation[3] = div(cone, R { 2 });
        if (s.motion != GANTRY_TOPOGRAM) {
            weight = s.tubes == 2 ? s.tubeWeight[tube] : R { 1 };
            if (s.aecSize)
                weight = mul(weight, aecWeight(s, aecTable, pos[2]));
            if (s.xcare)
                weight = mul(weight, xcareWeight(s, angle));
        }
    }

    // what the Exposure constructor does: unit cosines, beam along their cross product
    unit3(frame);
    unit3(frame + 3);
    R beam[3];
    cross3(frame, frame + 3, beam);
    if (s.align) { // components in the basis (x, y, x cross y) of the world
        R ez[3];
        cross3(s.worldCosines, s.worldCosines + 3, ez);
        project3(s.worldCosines, s.worldCosines + 3, ez, pos);
        project3(s.worldCosines, s.worldCosines + 3, ez, frame);
        project3(s.worldCosines, s.worldCosines + 3, ez, frame + 3);
        project3(s.worldCosines, s.worldCosines + 3, ez, beam);
    }
    for (int k = 0; k < 3; ++k) {
        e.position[k] = pos[k];
        e.beam[k] = beam[k];
    }
    for (int k = 0; k < 6; ++k)
        e.cosines[k] = frame[k];
    e.weight = weight;
    e.monoEnergy = s.monoEnergy;
    e.spectrum = s.spectrum[tube];
    e.heel = s.heel[tube];
    e.bowtie = s.bowtie[tube];
    e.histories = s.histories;
}

} // namespace dxmc::model
