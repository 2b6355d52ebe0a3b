// physics.cuh — device-side primitives of the photon transport hot path (sm_100a).
//
// Every function states the reference function whose arithmetic it reproduces. Float expressions
// keep the reference's operation order. The translation unit is compiled with nvcc's default FMA
// contraction (dxmclib_b200/build.py passes no -fmad=false): what must be bit-exact — geometry, voxel
// and brick indices, LUT segment and exponent arguments, the RNG — is written with explicit
// round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...), which nvcc never fuses; the interaction
// samplers may be contracted and agree with the CPU checkers statistically, not bit by bit.
#pragma once

#include "../../include/dxmcb200.h"

#include <cstdint>
#include <cuda_runtime.h>

namespace dxmcb200 {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = kPi + kPi;
constexpr float kElectronRestMass = 510.9989461f; // constants.hpp:63
constexpr float kKevToAngstrom = 12.398520f; // constants.hpp:29
constexpr float kEnergyCutoff = 1.0f; // transport.hpp:827
constexpr float kRouletteThreshold = 5.0f; // transport.hpp:823
constexpr float kRouletteProbability = 0.8f; // transport.hpp:819
constexpr float kDirEpsilon = 1.0e-9f; // transport.hpp:831

// ---- arithmetic helpers -----------------------------------------------------------------------
// trunc(fl(a / s)) for a >= 0, s > 0 — the value static_cast<size_t>(a / s) has in the reference — without
// an IEEE division on the common path. q = a * (1/s) is within ~1.5 ulp of the true quotient, so its
// truncation can only disagree with the truncation of the correctly rounded quotient when q lies within a
// few ulp of an integer; only then the exact division is evaluated. `inv` must be __frcp_rn(s).
__device__ __forceinline__ uint32_t truncDiv(float a, float s, float inv)
{
    const float q = __fmul_rn(a, inv);
    if (fabsf(__fsub_rn(q, rintf(q))) <= __fmul_rn(q, 4.8e-7f)) // within 4 ulp of an integer (rare)
        return __float2uint_rz(__fdiv_rn(a, s));
    return __float2uint_rz(q);
}

// 10^x to ~2 ulp: x*log2(10) split into a rounded product and its exact residual, MUFU.EX2 on the
// former, first-order correction with the latter
__device__ __forceinline__ float fastExp10(float x)
{
    constexpr float kLog2_10Hi = 3.3219280242919921875f;
    constexpr float kLog2_10Lo = 7.0595370e-8f;
    const float t = __fmul_rn(x, kLog2_10Hi);
    float r = __fmaf_rn(x, kLog2_10Hi, -t);
    r = __fmaf_rn(x, kLog2_10Lo, r);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return __fmaf_rn(__fmul_rn(e, r), 0.693147180559945f, e);
}

// lg2(x) by MUFU.LG2 (absolute error ~1e-7 of a mean free path in the step length, far below the float
// resolution of the position). x is a PCG draw: 0 or >= 2^-32, never denormal; lg2(0) = -inf sends the photon
// out of the world exactly like the reference's -log(0).
__device__ __forceinline__ float fastLog2(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// log10(E) correctly rounded to float (evaluated in double). It is needed once per energy change, not per step,
// and every attenuation value of the photon inherits its rounding through the log-log slope, so it is worth being
// exact: a 1-ulp error here is a ~1e-6 relative error in mu at photoelectric slopes.
__device__ __forceinline__ float log10Rounded(float energy) { return __double2float_rn(log10(static_cast<double>(energy))); }

// ---- tables as the kernels see them -------------------------------------------------------
constexpr int kSplineStride = 64; // device copy: 60 coefficients, start, step, stop, 1/step
constexpr int kCoeffStride = 8; // device copy of the log-log coefficients: one 32-byte sector per (material, segment)

struct LutView {
    uint32_t nMaterials, nSegments, linearIndex;
    float linearStep, linearEnergy, invLinearStep;
    const float* knots; // [nSegments]
    const float* coeff; // [nMaterials][nSegments][kCoeffStride]: photo {b,a}, Compton {b,a}, Rayleigh {b,a}, pad
    const float* maxCoeff; // [nSegments][2]
    const float* rita; // [nMaterials][4][56]
    const float* spline; // [nMaterials][kSplineStride]
    const float* shells; // [nMaterials][12][11]
};

struct WorldView {
    uint32_t dim[3];
    float spacing[3];
    float invSpacing[3]; // __frcp_rn(spacing)
    uint32_t exactInverse; // all three spacings are powers of two: a * invSpacing == a / spacing exactly
    float ext[6];
    const uint2* voxels; // {density bits, material | measurement<<8}; null when the grid is in palette form
    const uint8_t* palette; // palette form: index per voxel into paletteTable (grids with <= 256 distinct records)
    const uint2* paletteTable; // [256] records as in `voxels`
    uint32_t paletteNibbles; // 1: at most 16 distinct records, two 4-bit indices per byte (voxel i in byte i/2, low nibble first)
};

// Brick grid of the empty-space traversal: the voxel grid cut into bricks of 2^shift voxels per
// axis, one bit per brick: set when the brick is "air" (rho * mu_total(E) <= f_air * majorant(E) for every voxel and energy,
// no measurement voxel).
struct BrickView {
    uint32_t shift[3];
    uint32_t nb[3]; // bricks per axis
    float size[3]; // brick edge [mm]
    float invSize[3];
    float invFAir; // 1 / f_air: free paths in air bricks are the global-majorant ones times this
    uint32_t nWords; // 32-bit words of the bitmap; 0: no air bricks, plain Woodcock tracking everywhere
    const uint32_t* air; // bit b: brick b = (bz * nb[1] + by) * nb[0] + bx is air
    const uint8_t* distance; // [8][bricks]: per octant of travel directions and brick, the edge (bricks, <= 255) of the largest all-air cube cornered there; 0 for non-air bricks
};

struct SpectrumView {
    uint32_t n;
    uint32_t threshold; // rejection threshold of the bounded integer draw
    const float* probs;
    const uint32_t* alias;
    const float* energies;
};

struct HeelView {
    float energyStart, energyStep;
    uint32_t energySize;
    float angleStart, angleStep;
    uint32_t angleSize;
    const float* weights;
};

struct BowtieView {
    uint32_t n;
    const float* angles;
    const float* weights;
};

struct BeamView {
    const SpectrumView* spectra;
    const HeelView* heels;
    const BowtieView* bowties;
};

struct Photon {
    float px, py, pz;
    float dx, dy, dz;
    float energy, weight;
};

// ---- RNG: PCG32 XSH-RR, one stream per history (dxmcrandom.hpp:156-165, 67-73) -------------
__host__ __device__ inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ inline void historyStream(uint64_t seed, uint64_t exposure, uint64_t history, uint64_t& state, uint64_t& inc)
{
    const uint64_t golden = 0x9E3779B97F4A7C15ULL;
    const uint64_t s = mix64(seed + golden * (exposure + 1));
    state = mix64(s + golden * (history + 1));
    inc = mix64(state + golden) | 1ULL;
}

struct Rng {
    uint64_t state, inc;
    __device__ __forceinline__ uint32_t next()
    {
        const uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        const uint32_t xorshifted = static_cast<uint32_t>(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = static_cast<uint32_t>(old >> 59u);
        return __funnelshift_r(xorshifted, xorshifted, rot);
    }
    // u32 * 2^-32 rounded to float: can return exactly 1.0f, as the reference's does
    __device__ __forceinline__ float uniform() { return __fmul_rn(__uint2float_rn(next()), 2.32830643653869628906e-010f); }
    __device__ __forceinline__ float uniform(float maxv) { return __fmul_rn(uniform(), maxv); }
    __device__ __forceinline__ float uniform(float minv, float maxv)
    {
        const float r = uniform();
        const float range = __fsub_rn(maxv, minv);
        return __fadd_rn(minv, __fmul_rn(r, range));
    }
};

// ---- vector math (vectormath.hpp:68-99, 138-146, 197-221) -----------------------------------
__device__ __forceinline__ void rotate(float& v0, float& v1, float& v2, float a0, float a1, float a2, float angle)
{
    float sang, cang;
    sincosf(angle, &sang, &cang);
    const float midt = (1.0f - cang) * (v0 * a0 + v1 * a1 + v2 * a2);
    const float o0 = cang * v0 + midt * a0 + sang * (a1 * v2 - a2 * v1);
    const float o1 = cang * v1 + midt * a1 + sang * (-a0 * v2 + a2 * v0);
    const float o2 = cang * v2 + midt * a2 + sang * (a0 * v1 - a1 * v0);
    v0 = o0;
    v1 = o1;
    v2 = o2;
}

// scatter the direction by polar angle theta and azimuth phi; the helper axis is NOT normalised,
// exactly like the reference, so |dir| drifts below one after scatters (vectormath.hpp:197-221)
__device__ __forceinline__ void peturb(Photon& p, float theta, float phi)
{
    const float ax = fabsf(p.dx), ay = fabsf(p.dy), az = fabsf(p.dz);
    const int minInd = ax <= ay ? (ax <= az ? 0 : 2) : (ay <= az ? 1 : 2);
    const float k0 = minInd == 0 ? 1.0f : 0.0f;
    const float k1 = minInd == 1 ? 1.0f : 0.0f;
    const float k2 = minInd == 2 ? 1.0f : 0.0f;
    float x0 = p.dy * k2 - p.dz * k1;
    float x1 = p.dz * k0 - p.dx * k2;
    float x2 = p.dx * k1 - p.dy * k0;
    rotate(x0, x1, x2, p.dx, p.dy, p.dz, phi);
    float tsin, tcos;
    sincosf(theta, &tsin, &tcos);
    p.dx = p.dx * tcos + x0 * tsin;
    p.dy = p.dy * tcos + x1 * tsin;
    p.dz = p.dz * tcos + x2 * tsin;
}

// ---- geometry (transport.hpp:485-521, 702-728) ----------------------------------------------
__device__ __forceinline__ bool insideWorld(const WorldView& w, float x, float y, float z)
{
    // six compares and-ed as predicates (no short-circuit branches: the kernels call this on every step)
    return (x > w.ext[0]) & (x < w.ext[1]) & (y > w.ext[2]) & (y < w.ext[3]) & (z > w.ext[4]) & (z < w.ext[5]);
}

__device__ __forceinline__ void voxelCoords(const WorldView& w, float x, float y, float z, uint32_t& ix, uint32_t& iy, uint32_t& iz)
{
    if (w.exactInverse) { // uniform branch: scaling by a power of two is exact, so is the truncated quotient
        ix = __float2uint_rz(__fmul_rn(__fsub_rn(x, w.ext[0]), w.invSpacing[0]));
        iy = __float2uint_rz(__fmul_rn(__fsub_rn(y, w.ext[2]), w.invSpacing[1]));
        iz = __float2uint_rz(__fmul_rn(__fsub_rn(z, w.ext[4]), w.invSpacing[2]));
    } else {
        ix = truncDiv(__fsub_rn(x, w.ext[0]), w.spacing[0], w.invSpacing[0]);
        iy = truncDiv(__fsub_rn(y, w.ext[2]), w.spacing[1], w.invSpacing[1]);
        iz = truncDiv(__fsub_rn(z, w.ext[4]), w.spacing[2], w.invSpacing[2]);
    }
}

__device__ __forceinline__ uint32_t voxelIndex(const WorldView& w, float x, float y, float z)
{
    uint32_t ix, iy, iz;
    voxelCoords(w, x, y, z, ix, iy, iz);
    return (iz * w.dim[1] + iy) * w.dim[0] + ix;
}

// palette index of a voxel; random look-ups have no reuse in L1: cache them in L2 only (ld.global.cg)
__device__ __forceinline__ uint32_t paletteIndex(const WorldView& w, uint32_t voxel)
{
    if (w.paletteNibbles)
        return (static_cast<uint32_t>(__ldcg(w.palette + (voxel >> 1))) >> ((voxel & 1u) * 4u)) & 15u;
    return __ldcg(w.palette + voxel);
}

// {density bits, material | measurement << 8} of a voxel, whatever form the grid is stored in
__device__ __forceinline__ uint2 voxelRecord(const WorldView& w, uint32_t voxel)
{
    if (w.palette)
        return __ldg(w.paletteTable + paletteIndex(w, voxel));
    return __ldcg(w.voxels + voxel);
}

__device__ __forceinline__ void advance(Photon& p, float step)
{
    p.px = __fadd_rn(p.px, __fmul_rn(p.dx, step));
    p.py = __fadd_rn(p.py, __fmul_rn(p.dy, step));
    p.pz = __fadd_rn(p.pz, __fmul_rn(p.dz, step));
}

// slab-method entry into the world box; false when the ray misses (transport.hpp:702-728)
__device__ __forceinline__ bool transportToWorld(const WorldView& w, Photon& p)
{
    if (insideWorld(w, p.px, p.py, p.pz))
        return true;
    float amin = -3.402823466e+38f;
    float amax = 3.402823466e+38f;
    const float pos[3] = { p.px, p.py, p.pz };
    const float dir[3] = { p.dx, p.dy, p.dz };
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (fabsf(dir[i]) > kDirEpsilon) {
            const float a0 = __fdiv_rn(__fsub_rn(w.ext[i * 2], pos[i]), dir[i]);
            const float an = __fdiv_rn(__fsub_rn(w.ext[i * 2 + 1], pos[i]), dir[i]);
            amin = fmaxf(amin, fminf(a0, an));
            amax = fminf(amax, fmaxf(a0, an));
        }
    }
    if (amin < amax && amin > 0.0f) {
        p.px = __fadd_rn(p.px, __fmul_rn(amin, p.dx));
        p.py = __fadd_rn(p.py, __fmul_rn(amin, p.dy));
        p.pz = __fadd_rn(p.pz, __fmul_rn(amin, p.dz));
        return true;
    }
    return false;
}

// ---- empty-space traversal: brick look-up and ray / brick-grid traversal ----------------------
__device__ __forceinline__ uint32_t brickOfVoxel(const BrickView& b, uint32_t ix, uint32_t iy, uint32_t iz)
{
    return ((iz >> b.shift[2]) * b.nb[1] + (iy >> b.shift[1])) * b.nb[0] + (ix >> b.shift[0]);
}

// voxel coordinate along one axis, clamped into the grid: also defined for points on (or a rounding error outside) the
// faces of the world, where photons stand after transportToWorld
__device__ __forceinline__ uint32_t axisVoxelClamped(const WorldView& w, int axis, float x)
{
    const float rel = __fsub_rn(x, w.ext[2 * axis]);
    if (!(rel > 0.0f))
        return 0u;
    uint32_t v;
    if (w.exactInverse) // a uniform branch, not a select: the other side is the costlier one
        v = __float2uint_rz(__fmul_rn(rel, w.invSpacing[axis]));
    else
        v = truncDiv(rel, w.spacing[axis], w.invSpacing[axis]);
    return min(v, w.dim[axis] - 1u);
}

__device__ __forceinline__ bool airBit(const uint32_t* __restrict__ bitmap, uint32_t brick) { return (__ldg(bitmap + (brick >> 5)) >> (brick & 31u)) & 1u; }

__device__ __forceinline__ bool inAirBrick(const WorldView& w, const BrickView& b, float x, float y, float z)
{
    const uint32_t ix = axisVoxelClamped(w, 0, x), iy = axisVoxelClamped(w, 1, y), iz = axisVoxelClamped(w, 2, z);
    return airBit(b.air, brickOfVoxel(b, ix, iy, iz));
}

// Ray parameter at which the photon's ray leaves the run of air bricks it starts in (or has crossed DXMCB200_WALK_MAX_CUBES all-air
// cubes of it); `exits`: the ray leaves the grid there.
// A parametric ray / grid traversal (Siddon 1985, Amanatides & Woo 1987) over the brick grid that does not stop at every
// brick face: from an air brick b the ray crosses, in one step, the largest cube of air bricks that has b as its corner and
// opens in the ray's octant of directions (edge k bricks, tabulated per octant and brick: 2-3 steps per walk on the bench
// phantom instead of 13 face crossings; every step is a dependent table look-up, and that latency is what the walk costs).
// All face parameters are taken from the starting point, so nothing accumulates. Round-to-nearest intrinsics throughout: the
// CPU restatement (oracle/dxmc_oracle.cpp, airRunLength) computes the same bits.
__device__ __forceinline__ float airRunLength(const WorldView& w, const BrickView& b, const Photon& p, bool& exits, uint32_t& crossed)
{
    crossed = 0;
    const float pos[3] = { p.px, p.py, p.pz };
    const float dir[3] = { p.dx, p.dy, p.dz };
    int brick[3], step[3];
    float inv[3];
    uint32_t octant = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        brick[i] = static_cast<int>(axisVoxelClamped(w, i, pos[i]) >> b.shift[i]);
        if (fabsf(dir[i]) > kDirEpsilon) {
            inv[i] = __frcp_rn(dir[i]); // the correctly rounded reciprocal = 1.0f / dir[i]
            step[i] = dir[i] > 0.0f ? 1 : -1;
        } else {
            inv[i] = 0.0f;
            step[i] = 0;
        }
        if (dir[i] < 0.0f)
            octant |= 1u << i;
    }
    const uint8_t* const edge = b.distance + octant * (b.nb[0] * b.nb[1] * b.nb[2]);
    exits = false;
    float travelled = 0.0f;
    for (;;) {
        const int k = static_cast<int>(__ldg(edge + (static_cast<uint32_t>(brick[2]) * b.nb[1] + static_cast<uint32_t>(brick[1])) * b.nb[0] + static_cast<uint32_t>(brick[0])));
        float t[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float face = __fadd_rn(w.ext[2 * i], __fmul_rn(static_cast<float>(brick[i] + (step[i] > 0 ? k : 1 - k)), b.size[i]));
            t[i] = step[i] != 0 ? __fmul_rn(__fsub_rn(face, pos[i]), inv[i]) : __int_as_float(0x7f800000);
        }
        const int a = t[0] <= t[1] ? (t[0] <= t[2] ? 0 : 2) : (t[1] <= t[2] ? 1 : 2);
        const float ta = a == 0 ? t[0] : a == 1 ? t[1] : t[2];
        if ((a == 0 ? step[0] : a == 1 ? step[1] : step[2]) == 0) { // zero direction: the photon never leaves
            exits = true;
            return ta;
        }
        travelled = fmaxf(ta, travelled);
        ++crossed;
        bool outside = false;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            int c;
            if (j == a) {
                c = brick[j] + step[j] * k;
            } else { // brick of the exit point, inside the cube by construction (the clamp absorbs rounding)
                const float q = __fmul_rn(__fsub_rn(__fadd_rn(pos[j], __fmul_rn(travelled, dir[j])), w.ext[2 * j]), b.invSize[j]);
                const int far = brick[j] + (dir[j] < 0.0f ? -(k - 1) : (k - 1));
                c = min(max(__float2int_rd(q), min(brick[j], far)), max(brick[j], far));
            }
            brick[j] = c;
            outside = outside || c < 0 || c >= static_cast<int>(b.nb[j]);
        }
        if (outside) {
            exits = true;
            return travelled;
        }
        if (!airBit(b.air, (static_cast<uint32_t>(brick[2]) * b.nb[1] + static_cast<uint32_t>(brick[1])) * b.nb[0] + static_cast<uint32_t>(brick[0])))
            return travelled;
        if (crossed >= DXMCB200_WALK_MAX_CUBES) // the run goes on, this walk does not: the photon rejoins the Woodcock steps on this face
            return travelled;
    }
}

// ---- attenuation tables (attenuationinterpolator.hpp:207-248) -------------------------------
// first index with knots[i] > v, or n when none (std::upper_bound)
__device__ __forceinline__ uint32_t upperBound(const float* __restrict__ a, uint32_t n, float v)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (v < a[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ uint32_t segmentIndex(const LutView& l, float logE, bool clampLinear)
{
    if (logE > l.linearEnergy) {
        const uint32_t i = truncDiv(__fsub_rn(logE, l.linearEnergy), l.linearStep, l.invLinearStep) + l.linearIndex;
        return clampLinear ? min(i, l.nSegments - 1) : i;
    }
    const uint32_t pos = upperBound(l.knots, l.nSegments, logE);
    return pos != l.nSegments ? pos : l.nSegments - 1;
}

// photo, Compton, Rayleigh mass attenuation of one material at log10(E)
// the same with the segment already known (it only changes with the photon's energy)
__device__ __forceinline__ void attenuationAt(const LutView& l, uint32_t material, uint32_t segment, float logE, float& photo, float& compton,
    float& rayleigh)
{
    // one 256-bit load (sm_100 LDG.E.256) of the 32-byte record: photo {b, a}, Compton {b, a}, Rayleigh {b, a}, pad.
    // Lanes of a warp sit in different (material, segment) records, so every load instruction costs one L1 pass per
    // distinct line; ncu showed the former 128-bit + 64-bit pair taking 2 x 12.8 passes per warp-step.
    const float* c = l.coeff + (material * l.nSegments + segment) * kCoeffStride;
    float pb, pa, cb, ca, rb, ra, pad0, pad1;
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(pb), "=f"(pa), "=f"(cb), "=f"(ca), "=f"(rb), "=f"(ra), "=f"(pad0), "=f"(pad1)
        : "l"(c));
    (void)pad0, (void)pad1;
    photo = fastExp10(__fadd_rn(pb, __fmul_rn(pa, logE)));
    compton = fastExp10(__fadd_rn(cb, __fmul_rn(ca, logE)));
    rayleigh = fastExp10(__fadd_rn(rb, __fmul_rn(ra, logE)));
}

__device__ __forceinline__ void attenuation(const LutView& l, uint32_t material, float logE, float& photo, float& compton, float& rayleigh)
{
    attenuationAt(l, material, segmentIndex(l, logE, true), logE, photo, compton, rayleigh);
}

// inverse of the Woodcock majorant; the linear branch is not clamped in the reference either
__device__ __forceinline__ float maxAttenuationInverse(const LutView& l, float logE)
{
    uint32_t index = segmentIndex(l, logE, false);
    index = min(index, l.nSegments - 1); // memory safety only; never binds for E <= max table energy
    const float2 c = __ldg(reinterpret_cast<const float2*>(l.maxCoeff) + index);
    return fastExp10(__fadd_rn(c.x, __fmul_rn(c.y, logE)));
}

// Compton scatter function S(q)/Z, cubic spline (interpolation.hpp:169-176)
__device__ __forceinline__ float scatterFactor(const LutView& l, uint32_t material, float q)
{
    const float* s = l.spline + material * kSplineStride;
    const float4 lim = __ldg(reinterpret_cast<const float4*>(s + 60)); // start, step, stop, 1/step
    const float start = lim.x, stop = lim.z;
    const float x = fminf(fmaxf(q, start), stop);
    const uint32_t index = x > start ? truncDiv(__fsub_rn(x, start), lim.y, lim.w) : 0u;
    const uint32_t offset = index < DXMCB200_SPLINE_N - 1 ? index * 4 : (DXMCB200_SPLINE_N - 2) * 4;
    const float4 c = __ldg(reinterpret_cast<const float4*>(s + offset));
    // c0 + c1*x + c2*x*x + c3*x*x*x, left to right
    float r = __fadd_rn(c.x, __fmul_rn(c.y, x));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(c.z, x), x));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(c.w, x), x), x));
    return r;
}

// q^2 sampled from the squared form factor, truncated at maxValue (dxmcrandom.hpp:483-500)
__device__ __forceinline__ float sampleFormFactor(const LutView& l, uint32_t material, float maxValue, Rng& rng)
{
    const float* x = l.rita + material * (4 * DXMCB200_RITA_N);
    const float* e = x + DXMCB200_RITA_N;
    const float* a = e + DXMCB200_RITA_N;
    const float* b = a + DXMCB200_RITA_N;
    const uint32_t maxIndex = upperBound(x, DXMCB200_RITA_N, maxValue);
    const float modifier = maxIndex != DXMCB200_RITA_N ? e[maxIndex] : 1.0f;
    float res;
    do {
        const float r1 = rng.uniform(modifier);
        const int index = static_cast<int>(upperBound(e, DXMCB200_RITA_N, r1)) - 1;
        const float v = r1 - e[index];
        const float d = e[index + 1] - e[index];
        const float ai = a[index], bi = b[index];
        res = x[index] + (1.0f + ai + bi) * d * v / (d * d + ai * d * v + bi * v * v) * (x[index + 1] - x[index]);
    } while (res > maxValue);
    return res;
}

// ---- beam sampling (exposure.hpp:280-304, dxmcrandom.hpp:211-216, 322-326, beamfilters.hpp) ---
__device__ __forceinline__ float sampleSpectrum(const SpectrumView& s, Rng& rng)
{
    const float r = rng.uniform();
    uint32_t k;
    for (;;) { // bounded integer by rejection, threshold as the reference computes it
        const uint32_t u = rng.next();
        if (u >= s.threshold) {
            k = u % s.n;
            break;
        }
    }
    const uint32_t ind = r < s.probs[k] ? k : s.alias[k];
    return ind < s.n - 1 ? rng.uniform(s.energies[ind], s.energies[ind + 1]) : s.energies[ind];
}

__device__ __forceinline__ float bowtieWeight(const BowtieView& b, float anglePlusMinus)
{
    const float angle = fabsf(anglePlusMinus);
    uint32_t first = 0, last = b.n - 1;
    uint32_t it = first + (last - first) / 2;
    while (it != first) { // the reference's bisection (beamfilters.hpp:158-171)
        if (angle < b.angles[it])
            last = it;
        else
            first = it;
        it = first + (last - first) / 2;
    }
    if (angle < b.angles[first])
        return b.weights[0];
    if (angle > b.angles[last])
        return b.weights[b.n - 1];
    const float x0 = b.angles[first], x1 = b.angles[last];
    const float y0 = b.weights[first], y1 = b.weights[last];
    return y0 + (angle - x0) * (y1 - y0) / (x1 - x0);
}

__device__ __forceinline__ float heelWeight(const HeelView& h, float angle, float energy)
{
    // size_t casts of negative floats are undefined in the reference; the two guards below override them
    uint32_t eIndex = __float2uint_rz((energy - h.energyStart + 0.5f * h.energyStep) / h.energyStep);
    if (eIndex >= h.energySize)
        eIndex = h.energySize - 1;
    if (energy < h.energyStart)
        eIndex = 0;
    uint32_t aIndex = __float2uint_rz((angle - h.angleStart) / h.angleStep);
    if (aIndex >= h.angleSize)
        aIndex = h.angleSize - 1;
    if (angle < h.angleStart)
        aIndex = 0;
    const uint32_t wIndex = eIndex * h.angleSize + aIndex;
    if (aIndex < h.angleSize - 1) {
        const float a0 = h.angleStart + h.angleStep * aIndex;
        const float a1 = h.angleStart + h.angleStep * (aIndex + 1);
        const float w0 = h.weights[wIndex];
        const float w1 = h.weights[wIndex + 1];
        return w0 + (w1 - w0) * (angle - a0) / (a1 - a0);
    }
    return h.weights[wIndex];
}

__device__ __forceinline__ Photon sampleParticle(const dxmcb200_exposure& e, const BeamView& beams, Rng& rng)
{
    const float theta = rng.uniform(e.collimation[0], e.collimation[1]);
    const float phi = rng.uniform(e.collimation[2], e.collimation[3]);
    Photon p;
    p.px = e.position[0];
    p.py = e.position[1];
    p.pz = e.position[2];
    p.dx = e.beam_direction[0];
    p.dy = e.beam_direction[1];
    p.dz = e.beam_direction[2];
    p.weight = e.weight;
    rotate(p.dx, p.dy, p.dz, e.cosines[3], e.cosines[4], e.cosines[5], theta);
    rotate(p.dx, p.dy, p.dz, e.cosines[0], e.cosines[1], e.cosines[2], phi);
    p.energy = e.spectrum >= 0 ? sampleSpectrum(beams.spectra[e.spectrum], rng) : e.mono_energy;
    if (e.bowtie >= 0)
        p.weight *= bowtieWeight(beams.bowties[e.bowtie], theta);
    if (e.heel >= 0)
        p.weight *= heelWeight(beams.heels[e.heel], phi, p.energy);
    return p;
}

// ---- interactions (transport.hpp:216-483) ------------------------------------------------------
struct ShellView {
    const float* s;
    __device__ __forceinline__ float binding(int i) const { return s[i * DXMCB200_SHELL_FLOATS + 0]; }
    __device__ __forceinline__ float nElectrons(int i) const { return s[i * DXMCB200_SHELL_FLOATS + 1]; }
    __device__ __forceinline__ float j0(int i) const { return s[i * DXMCB200_SHELL_FLOATS + 2]; }
    __device__ __forceinline__ float photoProb(int i) const { return s[i * DXMCB200_SHELL_FLOATS + 3]; }
    __device__ __forceinline__ float yield(int i) const { return s[i * DXMCB200_SHELL_FLOATS + 4]; }
    __device__ __forceinline__ float lineProb(int i, int k) const { return s[i * DXMCB200_SHELL_FLOATS + 5 + k]; }
    __device__ __forceinline__ float lineEnergy(int i, int k) const { return s[i * DXMCB200_SHELL_FLOATS + 8 + k]; }
};

// The three interaction samplers come in two forms. The *Deferred form does everything but the change of direction: it
// reports the polar angle and leaves the azimuth draw and the rotation (peturb) to the caller, so that a kernel whose lanes
// sit in different channels runs that common tail ONCE for all of them instead of once per channel at a few lanes each
// (interactKernel: the Rayleigh copy of the tail ran at 2 of 32 lanes and cost 10 % of the kernel's instructions). The draw
// order is unchanged: the azimuth is the last draw of every channel in the reference too.
struct Deflection {
    bool scattered = false;
    float theta = 0.0f;
};
__device__ __forceinline__ void deflect(Photon& p, const Deflection& d, Rng& rng)
{
    if (d.scattered) {
        const float phi = rng.uniform(kTwoPi);
        peturb(p, d.theta, phi);
    }
}

// returns energy imparted locally; p.energy is 0 or the fluorescence line energy afterwards
template <int L>
__device__ __forceinline__ float photoAbsorptionDeferred(const LutView& l, Photon& p, uint32_t material, Rng& rng, Deflection& d)
{
    const float E = p.energy;
    p.energy = 0.0f;
    if constexpr (L < 2) {
        return E;
    } else {
        const ShellView sh { l.shells + material * (DXMCB200_SHELLS * DXMCB200_SHELL_FLOATS) };
        float cum[DXMCB200_SHELLS];
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < DXMCB200_SHELLS; ++i) {
            acc += E > sh.binding(i) ? sh.photoProb(i) : 0.0f;
            cum[i] = acc;
        }
        const float sample = cum[DXMCB200_SHELLS - 1] * rng.uniform();
        int idx = 0;
        while (cum[idx] < sample && idx < DXMCB200_SHELLS - 1)
            ++idx;
        if (sh.binding(idx) > E || sh.binding(idx) < kEnergyCutoff)
            return E;
        const float r1 = rng.uniform();
        if (r1 <= sh.yield(idx)) {
            int line = 0;
            float r3 = rng.uniform() - sh.lineProb(idx, line);
            while (r3 > 0.0f && line < 2) {
                ++line;
                r3 -= sh.lineProb(idx, line);
            }
            p.energy = sh.lineEnergy(idx, line);
            d.theta = rng.uniform(kPi);
            d.scattered = true;
            return E - p.energy;
        }
        return E;
    }
}
template <int L>
__device__ __forceinline__ float photoAbsorption(const LutView& l, Photon& p, uint32_t material, Rng& rng)
{
    Deflection d;
    const float e = photoAbsorptionDeferred<L>(l, p, material, rng, d);
    deflect(p, d, rng);
    return e;
}

template <int L>
__device__ __forceinline__ void rayleighScatterDeferred(const LutView& l, const Photon& p, uint32_t material, Rng& rng, Deflection& d)
{
    float theta;
    if constexpr (L == 0) {
        bool reject = true;
        while (reject) {
            constexpr float extreme = (4.0f * 1.41421356237309504880f) / (3.0f * 1.73205080756887729353f);
            const float r1 = rng.uniform(0.0f, extreme);
            theta = rng.uniform(0.0f, kPi);
            const float sinang = sinf(theta);
            reject = r1 > ((2.0f - sinang * sinang) * sinang);
        }
    } else {
        constexpr float kInv = 1.0f / kKevToAngstrom;
        const float qmax = p.energy * kInv;
        const float qmaxSquared = qmax * qmax;
        float cosAngle;
        do {
            const float qSquared = sampleFormFactor(l, material, qmaxSquared, rng);
            const float invE = kKevToAngstrom / p.energy;
            cosAngle = 1.0f - 2.0f * qSquared * invE * invE;
        } while ((1.0f + cosAngle * cosAngle) * 0.5f < rng.uniform());
        theta = acosf(cosAngle);
    }
    d.theta = theta;
    d.scattered = true;
}
template <int L>
__device__ __forceinline__ void rayleighScatter(const LutView& l, Photon& p, uint32_t material, Rng& rng)
{
    Deflection d;
    rayleighScatterDeferred<L>(l, p, material, rng, d);
    deflect(p, d, rng);
}

// EGSnrc-style impulse approximation with Doppler broadening (transport.hpp:342-483)
__device__ __noinline__ float comptonScatterIA(const LutView& l, Photon& p, uint32_t material, Rng& rng, Deflection& d)
{
    const ShellView sh { l.shells + material * (DXMCB200_SHELLS * DXMCB200_SHELL_FLOATS) };
    float cum[DXMCB200_SHELLS];
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < DXMCB200_SHELLS; ++i) {
        acc += (sh.binding(i) < p.energy && sh.j0(i) > 0.0f) ? sh.nElectrons(i) : 0.0f;
        cum[i] = acc;
    }
    int shellIdx = 0;
    const float shellSample = cum[DXMCB200_SHELLS - 1] * rng.uniform();
    while (cum[shellIdx] < shellSample && shellIdx < DXMCB200_SHELLS - 1)
        ++shellIdx;

    const float U = sh.binding(shellIdx) / kElectronRestMass;
    const float pb = sqrtf(2.0f * U + U * U);
    const float J0 = sh.j0(shellIdx);
    const float k = p.energy / kElectronRestMass;
    if (U > k)
        return 0.0f;

    const float emin = 1.0f / (1.0f + 2.0f * k);
    const float gmaxInv = 1.0f / (1.0f / emin + emin);
    float e, cosAngle;
    bool rejected;
    do {
        const float r1 = rng.uniform();
        e = r1 + (1.0f - r1) * emin;
        const float t = (1.0f - e) / (k * e);
        const float sinAngleSqr = t * (2.0f - t);
        cosAngle = 1.0f - t;
        const float g = (1.0f / e + e - sinAngleSqr) * gmaxInv;
        const float r2 = rng.uniform();
        rejected = r2 > g;
        if (!rejected) {
            const float pi = (k * (k - U) * (1.0f - cosAngle) - U) / sqrtf(2.0f * k * (k - U) * (1.0f - cosAngle) + U * U);
            const float kc = k * e;
            const float qc = sqrtf(k * k + kc * kc - 2.0f * k * kc * cosAngle);
            const float alpha = qc * (1.0f + kc * (kc - k * cosAngle) / (qc * qc)) / k;
            const float bPart = 1.0f + 2.0f * J0 * fabsf(pi);
            const float b = (bPart * bPart + 1.0f) / 2.0f;
            const float expb = expf(-b);

            float S;
            if (pi <= -pb) {
                S = (1.0f - alpha * pb) * expb / 2.0f;
            } else if (pi < pb) {
                const float pabs = fabsf(pb);
                const float piabs = fabsf(pi);
                constexpr float sp2 = 1.0f / (0.564189583547756286948f / 1.41421356237309504880f);
                const float part1 = alpha * sp2 / (4.0f * J0);
                constexpr float a1 = 0.34802f, a2 = -0.0958798f, a3 = 0.7478556f;
                const float sqrte = sqrtf(2.71828182845904523536f);
                const float tp = 1.0f / (1.0f + 0.332673f * (1.0f + 2.0f * J0 * pabs));
                const float tpi = 1.0f / (1.0f + 0.332673f * (1.0f + 2.0f * J0 * piabs));
                const float part2p = sqrte - expb * tp * (a1 + a2 * tp + a3 * tp * tp);
                const float part2pi = sqrte - expb * tpi * (a1 + a2 * tpi + a3 * tpi * tpi);
                if (pi <= 0.0f)
                    S = (1.0f - alpha * pi) * expb / 2.0f - part1 * (part2p - part2pi);
                else
                    S = 1.0f - (1.0f - alpha * pi) * expb / 2.0f - part1 * (part2p - part2pi);
            } else {
                S = 1.0f - (1.0f - alpha * pb) * expb / 2.0f;
            }
            const float r3 = rng.uniform();
            rejected = r3 > S;
            if (!rejected) {
                float Fmax;
                if (pi <= -pb)
                    Fmax = 1.0f - alpha * pb;
                else if (pi >= pb)
                    Fmax = 1.0f + alpha * pb;
                else
                    Fmax = 1.0f + alpha * pi;
                const float r4 = rng.uniform();
                const float rBar2 = 2.0f * r4 * expb;
                float pz;
                if (rBar2 < 1.0f) {
                    const float part = sqrtf(1.0f - 2.0f * logf(rBar2));
                    pz = (1.0f - part) / (2.0f * J0) / kElectronRestMass;
                } else {
                    const float part = sqrtf(1.0f - 2.0f * logf(2.0f - rBar2));
                    pz = (part - 1.0f) / (2.0f * J0) / kElectronRestMass;
                }
                float Fpz;
                if (pz <= -pb)
                    Fpz = 1.0f - alpha * pb;
                else if (pz >= pb)
                    Fpz = 1.0f + alpha * pb;
                else
                    Fpz = 1.0f + alpha * pz;
                const float r5 = rng.uniform();
                rejected = r5 > Fpz / Fmax;
                if (!rejected) {
                    const float part = sqrtf(1.0f - 2.0f * e * cosAngle + e * e * (1.0f - pz * pz * sinAngleSqr));
                    const float kBar = kc / (1.0f - pz * pz * e * e) * (1.0f - pz * pz * e * cosAngle + pz * part);
                    e = kBar / k;
                }
            }
        }
    } while (rejected);
    d.theta = acosf(cosAngle);
    d.scattered = true;
    const float E = p.energy;
    p.energy *= e;
    return E - p.energy;
}

// Klein-Nishina rejection sampling, optionally weighted by the scatter function (transport.hpp:300-340): at most maxTrials
// trials of the reference's loop (maxTrials < 0: until one is accepted). Every trial starts from fresh draws, so a caller may
// stop after a few rejections and resume later with the same random stream: the draws a history sees do not change. Returns
// whether a trial was accepted; then `e` is the energy fraction kept by the photon and `cosAngle` the polar cosine.
template <int L>
__device__ __forceinline__ bool comptonTrials(const LutView& l, float E, uint32_t material, Rng& rng, int maxTrials, float& e, float& cosAngle)
{
    static_assert(L < 2, "the impulse-approximation sampler draws its shell before the loop: comptonScatterIA");
    const float k = E / kElectronRestMass;
    const float emin = 1.0f / (1.0f + 2.0f * k);
    const float gmaxInv = 1.0f / (1.0f / emin + emin);
    for (int trial = 0; maxTrials < 0 || trial < maxTrials; ++trial) {
        const float r1 = rng.uniform();
        e = r1 + (1.0f - r1) * emin;
        const float t = (1.0f - e) / (k * e);
        const float sinthetasqr = t * (2.0f - t);
        cosAngle = 1.0f - t;
        const float g = (1.0f / e + e - sinthetasqr) * gmaxInv;
        const float r2 = rng.uniform();
        bool rejected;
        if constexpr (L == 1) {
            constexpr float kInv = 1.0f / kKevToAngstrom;
            const float q = E * kInv * sqrtf(0.5f - cosAngle * 0.5f);
            rejected = r2 > g * scatterFactor(l, material, q);
        } else {
            rejected = r2 > g;
        }
        if (!rejected)
            return true;
    }
    return false;
}

template <int L>
__device__ __forceinline__ float comptonScatterDeferred(const LutView& l, Photon& p, uint32_t material, Rng& rng, Deflection& d)
{
    if constexpr (L == 2) {
        return comptonScatterIA(l, p, material, rng, d);
    } else {
        const float E = p.energy;
        float e, cosAngle;
        comptonTrials<L>(l, E, material, rng, -1, e, cosAngle);
        d.theta = acosf(cosAngle);
        d.scattered = true;
        p.energy *= e;
        return E - p.energy;
    }
}
template <int L>
__device__ __forceinline__ float comptonScatter(const LutView& l, Photon& p, uint32_t material, Rng& rng)
{
    Deflection d;
    const float e = comptonScatterDeferred<L>(l, p, material, rng, d);
    deflect(p, d, rng);
    return e;
}

} // namespace dxmcb200
