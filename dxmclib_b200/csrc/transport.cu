// transport.cu — sm_100a photon-transport kernels and the C ABI runtime around them.
//
// Replaces the reference's worker threads (transport.hpp:729-778) by a wavefront of three kernels that hand
// photons to each other through record buffers in HBM. Every kernel keeps all 32 lanes of a warp on ONE kind of
// work; the divergent stages of a history (birth, stepping, interaction) never share a warp:
//
//   exposureKernel   (a0) Source::getExposure(i) for all i from one parameter block (dxmc/sourcemodel.hpp), SURVEY 8f3.
//   generateKernel   (a) Exposure::sampleParticle + transportParticleToWorld (exposure.hpp:280-304,
//                    transport.hpp:702-728, 733-741), one history per thread, and the first air walk of photons born into an
//                    air brick; photons that reach the dense part of the grid become 64-byte photon records carrying their own
//                    counter-derived PCG32 stream, log10(E), the LUT segment and the Woodcock majorant.
//   transportKernel  (b) Woodcock delta tracking (transport.hpp:640-700), persistent grid. A lane steps one photon
//                    until it leaves the world, is killed by Russian roulette, a real / forced interaction is due, or it
//                    stands in an air brick after a virtual collision; events go to the event buffer as 64-byte records, air-bound
//                    photons to the air-walk buffer, and empty lanes take the next records of the wave straight from global
//                    memory (tiles of 256 records per atomic, prefetched into L2).
//   airWalkKernel    (b') empty-space traversal: photons that a Woodcock step left in an "air" brick
//                    are walked through the run of air bricks on their ray (Siddon / Amanatides-Woo style traversal of the
//                    brick grid in whole all-air cubes) against the regional majorant of air, then rejoin the next wave. Not
//                    part of the reference's algorithm; dxmcb200_set_tracking(ctx, 0) switches it off and leaves the
//                    reference's Woodcock loop everywhere.
//   interactKernel   (c)+(d) computeInteractions[Forced] (transport.hpp:523-638): photoelectric / Compton /
//                    Rayleigh sampling against the LUTs, 64-bit fixed-point scoring, Russian roulette; surviving
//                    photons go to the NEXT wave's photon buffer.
//
// Output slots are claimed by warps a tile at a time and ahead of need (TileWriter), so no atomic's round trip stalls a warp;
// slots a warp claimed and did not use are dead markers. One wave = births + survivors of the previous wave, at most
// `waveRecords` photons; the host tops every wave up with new births until all histories are issued and then drains. Two
// such pipelines run on two streams. Scoring is integer atomics and every history carries its own random stream, so results
// do not depend on the wave size, scheduling or GPU partition.
#include "hostio.cuh"
#include "physics.cuh"
#include "spectrum.cuh"
#include "../include/dxmc/sourcemodel.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <new>
#include <string>
#include <vector>

using namespace dxmcb200;

namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

enum LaneState : uint32_t { DEAD = 0, STEP = 1, EVENT = 2, EXHAUSTED = 3, AIRBORNE = 4 };

struct Counters {
    unsigned long long histories, inWorld, steps, lookups, interactions, scores, airWalks, bricksCrossed;
};

// one in-flight photon between kernels: 4 x 16 bytes
struct alignas(16) PhotonRecord {
    float4 posE; // px py pz energy
    float4 dirW; // dx dy dz weight
    uint4 rng; // state lo/hi, increment lo/hi
    float4 lut; // log10(E), 1/majorant, LUT segment (bits), event probability (event records only)
};

// a photon with an interaction due: 4 x 16 bytes. The photon record's majorant is left out (interactKernel recomputes it in the
// few cases the energy does not change) and the LUT segment shares a word with the voxel's material word, which has 17 bits.
struct alignas(16) EventRecord {
    float4 posE; // px py pz energy
    float4 dirW; // dx dy dz weight
    uint4 rng; // state lo/hi, increment lo/hi
    uint4 where; // voxel index, material | measurement << 8 | in-air-brick << 16 | LUT segment << 17 (kNoEvent marks an unused slot),
                 // log10(E) bits, event probability bits
};
constexpr unsigned kSegmentShift = 17;
constexpr uint32_t kNoEvent = 0xffffffffu;
constexpr uint32_t kInAirBrick = 0x10000u; // bit 16 of a voxel record's second word: the voxel lies in an air brick (flagRecordsKernel)

// Every buffer is split into kShards regions with their own append / claim cursors on separate 128-byte lines:
// tens of thousands of warps appending through ONE counter serialise in a single L2 atomic unit (measured: the
// interaction kernel spent 4 of its 4.4 ms queueing on it); spread over 64 addresses the cost disappears.
constexpr unsigned kShards = 64;
struct alignas(128) ShardCursor {
    unsigned int stored; // slots appended to the region
    unsigned int live; // photons among them (interactKernel claims slots in tiles; what a tile does not use is marked dead)
    unsigned int taken; // slots of the region claimed by the consumer
    unsigned int pad[29];
};
struct WaveCursors {
    ShardCursor photons[2][kShards];
    ShardCursor events[kShards];
    ShardCursor air[kShards];
    ShardCursor overflow; // .stored != 0: a region was too small, the run is invalid
};

struct KernelParams {
    WorldView world;
    LutView lut;
    BeamView beams;
    BrickView bricks;
    const dxmcb200_exposure* exposures; // absolute indexing
    const uint64_t* prefix; // [nExp+1] cumulative histories of the run's exposure range
    uint64_t expBegin; // first exposure of the run's range
    uint64_t expStride; // the range holds exposures expBegin + k * expStride (multi-GPU interleaving), k < nExp
    uint32_t nExp;
    uint32_t uniformHistories; // >0: every exposure of the range has this many histories (< 2^31)
    uint64_t chunkBegin; // first history generateKernel makes, counted from the start of the range
    uint32_t chunkCount; // histories generateKernel makes
    uint32_t chunkFirstExposure; // uniform case: exposure (relative to expBegin) holding chunkBegin ...
    uint32_t chunkFirstOffset; // ... and chunkBegin's history index inside it
    uint32_t refillBatch; // lanes that must be empty before a warp stops stepping to re-fill
    uint64_t seed;
    PhotonRecord* photonsIn; // wave being transported: kShards regions of photonRegion records
    ShardCursor* inCursors;
    PhotonRecord* photonsOut; // next wave: generateKernel and interactKernel append here
    ShardCursor* outCursors;
    EventRecord* events; // kShards regions of eventRegion slots, claimed in tiles of kEventTile
    ShardCursor* eventCursors;
    PhotonRecord* airborne; // photons a Woodcock step left in an air brick: kShards regions of photonRegion records
    ShardCursor* airCursors;
    unsigned int* overflow;
    uint32_t photonRegion, eventRegion;
    unsigned long long* acc; // [nVoxels][4]
    Counters* counters;
    float energyScale, energySqScale;
};

__device__ __forceinline__ unsigned long long warpSum(uint32_t v)
{
    unsigned long long s = v;
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(kFull, s, o);
    return s;
}

// Append one record per lane in `mask` to the warp's shard region of a buffer; returns the slot index inside the buffer
// (kNoSlot for lanes outside the mask, or when the region is full, which invalidates the run).
constexpr size_t kNoSlot = ~static_cast<size_t>(0);
__device__ __forceinline__ size_t appendSlots(ShardCursor* cursors, uint32_t region, unsigned int* overflow, unsigned mask, unsigned lane)
{
    const unsigned shard = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) % kShards;
    unsigned base = 0;
    const int leader = __ffs(mask) - 1;
    const unsigned n = __popc(mask);
    if (static_cast<int>(lane) == leader) // stored += n and live += n with one 64-bit atomic
        base = static_cast<unsigned>(atomicAdd(reinterpret_cast<unsigned long long*>(&cursors[shard].stored), (static_cast<unsigned long long>(n) << 32) | n));
    base = __shfl_sync(kFull, base, leader);
    if (base + n > region) {
        if (static_cast<int>(lane) == leader)
            atomicExch(overflow, 1u);
        return kNoSlot;
    }
    if (!((mask >> lane) & 1u))
        return kNoSlot;
    return static_cast<size_t>(shard) * region + base + __popc(mask & ((1u << lane) - 1u));
}

// the photons of the lanes in `mask` go to the output wave
__device__ __forceinline__ PhotonRecord* appendPhotons(const KernelParams& P, unsigned mask, unsigned lane)
{
    const size_t slot = appendSlots(P.outCursors, P.photonRegion, P.overflow, mask, lane);
    return slot == kNoSlot ? nullptr : P.photonsOut + slot;
}

// everything about a photon that only changes with its energy: log10(E) correctly rounded, the LUT segment it
// falls in, the inverse Woodcock majorant (attenuationinterpolator.hpp:207-248)
__device__ __forceinline__ void energyDependent(const LutView& lut, float energy, float& logE, uint32_t& seg, float& maxAttInv)
{
    logE = log10Rounded(energy);
    seg = segmentIndex(lut, logE, true);
    maxAttInv = maxAttenuationInverse(lut, logE);
}

// ---- photon and event records <-> registers -----------------------------------------------------------
// Records are written once and read once, tens of GB per wave pair: they go through L2 with the streaming (evict-first)
// policy so that they do not push the voxel grid and the accumulator lines out of it.
#ifndef DXMCB200_STREAM_RECORDS
#define DXMCB200_STREAM_RECORDS 1
#endif
__device__ __forceinline__ void recStore(float4* at, float4 v)
{
#if DXMCB200_STREAM_RECORDS
    __stcs(at, v);
#else
    *at = v;
#endif
}
__device__ __forceinline__ void recStore(uint4* at, uint4 v)
{
#if DXMCB200_STREAM_RECORDS
    __stcs(at, v);
#else
    *at = v;
#endif
}
__device__ __forceinline__ float4 recLoad(const float4* at)
{
#if DXMCB200_STREAM_RECORDS
    return __ldcs(at);
#else
    return *at;
#endif
}
__device__ __forceinline__ uint4 recLoad(const uint4* at)
{
#if DXMCB200_STREAM_RECORDS
    return __ldcs(at);
#else
    return *at;
#endif
}
__device__ __forceinline__ void storePhoton(PhotonRecord* r, const Photon& p, const Rng& rng, float logE, float maxAttInv, uint32_t seg, float extra)
{
    recStore(&r->posE, make_float4(p.px, p.py, p.pz, p.energy));
    recStore(&r->dirW, make_float4(p.dx, p.dy, p.dz, p.weight));
    recStore(&r->rng, make_uint4(static_cast<uint32_t>(rng.state), static_cast<uint32_t>(rng.state >> 32), static_cast<uint32_t>(rng.inc),
                          static_cast<uint32_t>(rng.inc >> 32)));
    recStore(&r->lut, make_float4(logE, maxAttInv, __uint_as_float(seg), extra));
}

__device__ __forceinline__ void loadPhoton(const PhotonRecord* r, Photon& p, Rng& rng, float& logE, float& maxAttInv, uint32_t& seg, float& extra)
{
    const float4 a = recLoad(&r->posE), b = recLoad(&r->dirW), d = recLoad(&r->lut);
    const uint4 c = recLoad(&r->rng);
    p.px = a.x, p.py = a.y, p.pz = a.z, p.energy = a.w;
    p.dx = b.x, p.dy = b.y, p.dz = b.z, p.weight = b.w;
    rng.state = (static_cast<uint64_t>(c.y) << 32) | c.x;
    rng.inc = (static_cast<uint64_t>(c.w) << 32) | c.z;
    logE = d.x, maxAttInv = d.y, seg = __float_as_uint(d.z), extra = d.w;
}

// dead marker: energy 0 in a whole record of zeros (a consumer loads all four parts of a record before it looks at the energy)
__device__ __forceinline__ void storeDeadPhoton(PhotonRecord* r)
{
    recStore(&r->posE, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    recStore(&r->dirW, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    recStore(&r->rng, make_uint4(0u, 0u, 0u, 0u));
    recStore(&r->lut, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
}

__device__ __forceinline__ void storeEvent(EventRecord* e, const Photon& p, const Rng& rng, float logE, uint32_t seg, uint32_t voxel, uint32_t material,
    float eventProbability)
{
    recStore(&e->posE, make_float4(p.px, p.py, p.pz, p.energy));
    recStore(&e->dirW, make_float4(p.dx, p.dy, p.dz, p.weight));
    recStore(&e->rng, make_uint4(static_cast<uint32_t>(rng.state), static_cast<uint32_t>(rng.state >> 32), static_cast<uint32_t>(rng.inc),
                          static_cast<uint32_t>(rng.inc >> 32)));
    recStore(&e->where, make_uint4(voxel, material | (seg << kSegmentShift), __float_as_uint(logE), __float_as_uint(eventProbability)));
}

// -ln(r) * maxAttInv * 10 (transport.hpp:655-657) is evaluated as lg2(r) * kStepScale * maxAttInv; cm -> mm is the factor 10
constexpr float kStepScale = -6.931471805599453f;

// ---- (b') empty-space traversal ------------------------------------------------------------------------
// A photon standing in an air brick is walked to the end of the run of air bricks on its ray. Collision candidates on
// that stretch are sampled against the regional majorant f_air * majorant(E) (free paths are the global-majorant ones
// times 1/f_air); what happens AT a candidate is the reference's step body (transport.hpp:659-693): look-up, accept /
// reject (or forced interaction in a measurement voxel), Russian roulette after a virtual collision.
enum WalkOutcome : uint32_t {
    WALK_DENSE = 0, // the photon stands on the face of a non-air brick: back to Woodcock steps
    WALK_GONE = 1, // left the world or lost Russian roulette
    WALK_EVENT = 2 // an interaction is due at the photon's position
};
struct WalkEvent {
    uint32_t voxel = 0, material = 0;
    float eventProbability = 0.0f;
};

template <bool kStats>
__device__ __forceinline__ uint32_t airWalk(const KernelParams& P, Photon& p, Rng& rng, float logE, uint32_t seg, float maxAttInv, WalkEvent& ev, uint32_t& cSteps,
    uint32_t& cLookups, uint32_t& cBricks)
{
    bool exits = false;
    uint32_t crossed = 0;
    float remaining = airRunLength(P.world, P.bricks, p, exits, crossed);
    if constexpr (kStats)
        cBricks += crossed;
    bool lowWeight = p.energy * p.weight < kRouletteThreshold;
    for (;;) {
        const float r1 = rng.uniform();
        const float s = __fmul_rn(__fmul_rn(__fmul_rn(fastLog2(r1), kStepScale), maxAttInv), P.bricks.invFAir);
        if (!(s < remaining)) { // no candidate before the end of the run
            advance(p, remaining);
            return exits ? WALK_GONE : WALK_DENSE;
        }
        advance(p, s);
        remaining = __fsub_rn(remaining, s);
        if constexpr (kStats)
            ++cSteps;
        if (!insideWorld(P.world, p.px, p.py, p.pz))
            return WALK_GONE;
        const uint32_t voxel = voxelIndex(P.world, p.px, p.py, p.pz);
        const uint2 rec = voxelRecord(P.world, voxel);
        if constexpr (kStats)
            ++cLookups;
        float aP, aC, aR;
        attenuationAt(P.lut, rec.y & 0xffu, seg, logE, aP, aC, aR);
        const float attTotal = ((aP + aC) + aR) * __uint_as_float(rec.x);
        const float eventProbability = fminf(__fmul_rn(__fmul_rn(attTotal, maxAttInv), P.bricks.invFAir), 1.0f);
        bool event = (rec.y & 0xff00u) != 0;
        if (!event)
            event = rng.uniform() < eventProbability;
        if (event) {
            ev.voxel = voxel;
            ev.material = rec.y;
            ev.eventProbability = eventProbability;
            return WALK_EVENT;
        }
        if (lowWeight) { // Russian roulette after a virtual collision (transport.hpp:684-693)
            const float r4 = rng.uniform();
            if (r4 < kRouletteProbability)
                return WALK_GONE;
            constexpr float factor = 1.0f / (1.0f - kRouletteProbability);
            p.weight *= factor;
            lowWeight = p.energy * p.weight < kRouletteThreshold;
        }
    }
}

// hand the outcome of a walk on: photons on the face of a non-air brick join `photons` / `photonCursors`, photons with an
// interaction due get an event record (claimed with an exact count: the consumer skips nothing but kNoEvent slots)
__device__ __forceinline__ void emitWalked(const KernelParams& P, PhotonRecord* photons, ShardCursor* photonCursors, uint32_t outcome, bool valid, const Photon& p,
    const Rng& rng, float logE, float maxAttInv, uint32_t seg, const WalkEvent& ev, unsigned lane)
{
    const unsigned denseMask = __ballot_sync(kFull, valid && outcome == WALK_DENSE);
    const unsigned eventMask = __ballot_sync(kFull, valid && outcome == WALK_EVENT);
    if (denseMask) {
        const size_t slot = appendSlots(photonCursors, P.photonRegion, P.overflow, denseMask, lane);
        if (slot != kNoSlot)
            storePhoton(photons + slot, p, rng, logE, maxAttInv, seg, 0.0f);
    }
    if (eventMask) {
        const size_t slot = appendSlots(P.eventCursors, P.eventRegion, P.overflow, eventMask, lane);
        if (slot != kNoSlot) {
            storeEvent(P.events + slot, p, rng, logE, seg, ev.voxel, ev.material, ev.eventProbability);
        }
    }
}

// ---- (a) exposure-to-photon generation -------------------------------------------------------------
#ifndef DXMCB200_WALK_MINBLOCKS
#define DXMCB200_WALK_MINBLOCKS 6
#endif
template <bool kStats, bool kAir>
__global__ void __launch_bounds__(kThreads, DXMCB200_WALK_MINBLOCKS) generateKernel(const __grid_constant__ KernelParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    uint32_t cHist = 0, cWorld = 0, cSteps = 0, cLookups = 0, cBricks = 0, cWalks = 0;
    const uint32_t rounded = (P.chunkCount + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < rounded; i += gridDim.x * kThreads) {
        bool keep = false;
        Photon p {};
        Rng rng { 0, 1 };
        float logE = 0.0f, maxAttInv = 0.0f;
        uint32_t seg = 0;
        uint32_t outcome = WALK_DENSE;
        WalkEvent ev;
        if (i < P.chunkCount) {
            uint32_t e;
            uint64_t history;
            if (P.uniformHistories) {
                const uint32_t t = i + P.chunkFirstOffset;
                const uint32_t q = t / P.uniformHistories;
                e = P.chunkFirstExposure + q;
                history = t - q * P.uniformHistories;
            } else { // exposure owning history g: last prefix entry <= g
                const uint64_t g = P.chunkBegin + i;
                uint32_t lo = 0, hi = P.nExp;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(P.prefix + mid) <= g)
                        lo = mid;
                    else
                        hi = mid;
                }
                e = lo;
                history = g - __ldg(P.prefix + lo);
            }
            const uint64_t exposure = P.expBegin + e * P.expStride;
            historyStream(P.seed, exposure, history, rng.state, rng.inc);
            p = sampleParticle(P.exposures[exposure], P.beams, rng);
            keep = transportToWorld(P.world, p);
            if constexpr (kStats) {
                ++cHist;
                cWorld += keep ? 1u : 0u;
            }
            if (keep) {
                energyDependent(P.lut, p.energy, logE, seg, maxAttInv);
                if constexpr (kAir) { // born into (or onto the face of) an air brick: walk to the first non-air brick
                    if (inAirBrick(P.world, P.bricks, p.px, p.py, p.pz)) {
                        outcome = airWalk<kStats>(P, p, rng, logE, seg, maxAttInv, ev, cSteps, cLookups, cBricks);
                        if constexpr (kStats)
                            ++cWalks;
                    }
                }
            }
        }
        if constexpr (kAir) {
            emitWalked(P, P.photonsOut, P.outCursors, outcome, keep, p, rng, logE, maxAttInv, seg, ev, lane);
        } else {
            const unsigned keepMask = __ballot_sync(kFull, keep);
            if (keepMask == 0)
                continue;
            PhotonRecord* r = appendPhotons(P, keepMask, lane);
            if (r)
                storePhoton(r, p, rng, logE, maxAttInv, seg, 0.0f);
        }
    }
    if constexpr (kStats) {
        const unsigned long long h = warpSum(cHist), w = warpSum(cWorld), st = warpSum(cSteps), l = warpSum(cLookups), b = warpSum(cBricks),
                                 wk = warpSum(cWalks);
        if (lane == 0) {
            atomicAdd(&P.counters->histories, h);
            atomicAdd(&P.counters->inWorld, w);
            if constexpr (kAir) {
                atomicAdd(&P.counters->steps, st);
                atomicAdd(&P.counters->lookups, l);
                atomicAdd(&P.counters->bricksCrossed, b);
                atomicAdd(&P.counters->airWalks, wk);
            }
        }
    }
}

// photons the transport kernel left in air bricks: walk them, survivors join the NEXT wave
template <bool kStats>
__global__ void __launch_bounds__(kThreads, DXMCB200_WALK_MINBLOCKS) airWalkKernel(const __grid_constant__ KernelParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    uint32_t cSteps = 0, cLookups = 0, cBricks = 0, cWalks = 0;
    const unsigned blocksPerShard = gridDim.x / kShards; // the grid is a multiple of kShards
    const unsigned shard = blockIdx.x % kShards;
    const unsigned nSlots = min(P.airCursors[shard].stored, P.photonRegion);
    const PhotonRecord* const region = P.airborne + static_cast<size_t>(shard) * P.photonRegion;
    for (unsigned base = (blockIdx.x / kShards) * kThreads; base < nSlots; base += blocksPerShard * kThreads) {
        const unsigned i = base + threadIdx.x;
        Photon p {};
        Rng rng { 0, 1 };
        float logE = 0.0f, maxAttInv = 0.0f, extra = 0.0f;
        uint32_t seg = 0;
        uint32_t outcome = WALK_GONE;
        WalkEvent ev;
        bool valid = i < nSlots;
        if (valid) {
            loadPhoton(region + i, p, rng, logE, maxAttInv, seg, extra);
            valid = p.energy > 0.0f; // energy 0: an unused slot of a transportKernel tile
        }
        if (valid) {
            outcome = airWalk<kStats>(P, p, rng, logE, seg, maxAttInv, ev, cSteps, cLookups, cBricks);
            if constexpr (kStats)
                ++cWalks;
        }
        emitWalked(P, P.photonsOut, P.outCursors, outcome, valid, p, rng, logE, maxAttInv, seg, ev, lane);
    }
    if constexpr (kStats) {
        const unsigned long long st = warpSum(cSteps), l = warpSum(cLookups), b = warpSum(cBricks), wk = warpSum(cWalks);
        if (lane == 0) {
            atomicAdd(&P.counters->steps, st);
            atomicAdd(&P.counters->lookups, l);
            atomicAdd(&P.counters->bricksCrossed, b);
            atomicAdd(&P.counters->airWalks, wk);
        }
    }
}

// ---- (d) scoring: warp-aggregated 64-bit fixed point (replaces safeValueAdd, transport.hpp:208-214) ----
// A lane's deposit waits in a ScoreSlot until the warp is converged again; then the warp scores together: lanes are
// grouped by voxel (match.any), every group adds up its fixed-point energies, squares and event count with shuffles
// and its first lane issues the three 64-bit atomics. A pencil beam puts most lanes of a warp into the same few
// voxels (BASELINE config #1: up to 32 same-address atomics per warp become one; measured +2.5 % histories/s). In a CT
// scan groups of one are the rule and the grouping only costs (measured -2 %), so the runtime picks the aggregating
// kernel variant for narrow beams only (runRange). Integer sums: the grids are the same bits either way.
constexpr uint32_t kNoVoxel = 0xffffffffu;
struct ScoreSlot {
    uint32_t voxel = kNoVoxel;
    float energy = 0.0f; // energy imparted x weight [keV]
    __device__ __forceinline__ void set(uint32_t v, float e)
    {
        voxel = v;
        energy = e;
    }
};

// one lane, one deposit, three atomics: what a CT scan wants (lanes of a warp sit in different voxels)
__device__ __forceinline__ void scoreNow(const KernelParams& P, uint32_t voxel, float energyImparted)
{
    const long long fe = __float2ll_rn(energyImparted * P.energyScale);
    const unsigned long long fe2 = __float2ull_rn((energyImparted * energyImparted) * P.energySqScale);
    unsigned long long* a = P.acc + static_cast<size_t>(voxel) * 4;
    atomicAdd(a + 0, static_cast<unsigned long long>(fe));
    atomicAdd(a + 1, fe2);
    atomicAdd(a + 2, 1ULL);
}

// kAggregate: park the deposit for scoreWarp; otherwise score at once
// A deposit that is not a finite number is dropped: the impulse-approximation sampler inherits from the reference a 2^-32-per-draw
// path to NaN (transport.hpp:441-455: a uniform draw of exactly 0 makes log(0) = -inf, then sqrt(-inf)), which in the reference
// poisons the voxel's float sum and in a fixed-point sum would read as -2^63.
template <bool kAggregate>
__device__ __forceinline__ void deposit(const KernelParams& P, ScoreSlot& slot, uint32_t voxel, float energyImparted)
{
    if (!(fabsf(energyImparted) <= 3.0e38f))
        return;
    if constexpr (kAggregate)
        slot.set(voxel, energyImparted);
    else
        scoreNow(P, voxel, energyImparted);
}

// called by all 32 lanes of a converged warp; lanes without a deposit pass an empty slot
__device__ __forceinline__ void scoreWarp(const KernelParams& P, const ScoreSlot& slot, unsigned lane)
{
    const unsigned group = __match_any_sync(kFull, slot.voxel);
    if (slot.voxel == kNoVoxel)
        return;
    long long fe = __float2ll_rn(slot.energy * P.energyScale);
    unsigned long long fe2 = __float2ull_rn((slot.energy * slot.energy) * P.energySqScale);
    unsigned long long events = 1ULL;
    const unsigned leader = static_cast<unsigned>(__ffs(group) - 1);
    if (group != (1u << lane)) { // several lanes in this voxel: the lanes of the group (and only they) add up
        const long long myFe = fe;
        const unsigned long long myFe2 = fe2;
        fe = 0;
        fe2 = 0;
        for (unsigned m = group; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            fe += __shfl_sync(group, myFe, src);
            fe2 += __shfl_sync(group, myFe2, src);
        }
        events = static_cast<unsigned long long>(__popc(group));
    }
    if (lane == leader) {
        unsigned long long* a = P.acc + static_cast<size_t>(slot.voxel) * 4;
        atomicAdd(a + 0, static_cast<unsigned long long>(fe));
        atomicAdd(a + 1, fe2);
        atomicAdd(a + 2, events);
    }
}

struct Pending { // what an INTERACT lane needs from the step that found the event
    float attPhoto, attCompton, attRayleigh;
    float eventProbability;
    uint32_t voxel;
    uint32_t material; // bits 0-7 material, bits 8-15 measurement flag (forced interaction when non-zero)
};

// computeInteractionsForced (transport.hpp:523-581)
template <int L, bool kStats, bool kAggregate>
__device__ __forceinline__ bool interactForced(const KernelParams& P, Photon& p, const Pending& pe, Rng& rng, bool& energyChanged, uint32_t& nScores,
    ScoreSlot& score, ScoreSlot& forcedScore, Deflection& turn)
{
    const uint32_t mat = pe.material & 0xffu;
    const float attTotal = ((0.0f + pe.attPhoto) + pe.attCompton) + pe.attRayleigh;
    const float photoEventProbability = pe.attPhoto / attTotal;
    const float weightCorrection = pe.eventProbability * photoEventProbability;
    {
        Photon forced = p; // the forced photo-absorption acts on a copy: its fluorescence photon is never followed, only its draws count
        Deflection unused;
        const float eForced = photoAbsorptionDeferred<L>(P.lut, forced, mat, rng, unused);
        if (unused.scattered)
            (void)rng.uniform(kTwoPi);
        if constexpr (kStats)
            ++nScores;
        if (forced.energy < kEnergyCutoff)
            deposit<kAggregate>(P, forcedScore, pe.voxel, (eForced + forced.energy) * forced.weight * weightCorrection);
        else
            deposit<kAggregate>(P, forcedScore, pe.voxel, eForced * forced.weight * weightCorrection);
    }
    const float r1 = rng.uniform();
    if (r1 < pe.eventProbability * (1.0f - photoEventProbability)) {
        const float r2 = rng.uniform(pe.attCompton + pe.attRayleigh);
        if (r2 < pe.attCompton) {
            const float e = comptonScatterDeferred<L>(P.lut, p, mat, rng, turn);
            if constexpr (kStats)
                ++nScores;
            if (p.energy < kEnergyCutoff) {
                deposit<kAggregate>(P, score, pe.voxel, (e + p.energy) * p.weight);
                p.energy = 0.0f;
                turn.scattered = false;
                return false;
            }
            deposit<kAggregate>(P, score, pe.voxel, e * p.weight);
            energyChanged = true;
        } else {
            rayleighScatterDeferred<L>(P.lut, p, mat, rng, turn);
        }
    }
    p.weight *= (1.0f - weightCorrection);
    return true;
}

// ---- record hand-over between the kernels ------------------------------------------------------------
constexpr unsigned kTile = 256; // photon records a warp of transportKernel claims with one atomic
constexpr unsigned kEventTile = 64; // event slots a warp claims with one atomic
constexpr unsigned kSurvivorTile = 64; // next-wave photon slots a warp of interactKernel claims with one atomic
constexpr unsigned kAirTile = 64; // air-walk slots a warp of transportKernel claims with one atomic

// A warp's private window into its shard region of an output buffer. Slots are claimed a tile at a time, and AHEAD of need:
// `prepare` issues the atomic as soon as the current tile might not take one more slot per lane, `take` first looks at its
// result when the slots are actually needed, typically a trip of the caller's loop later, so the atomic's round trip through
// L2 never stalls the warp (ncu: 14 % of interactKernel's stall samples sat on it when every trip claimed its exact count).
// What a warp has claimed but not used by the end of the kernel is marked dead (`finish`), and the consumer skips it.
// All members are warp-uniform except `claimed`, which only lane 0 holds.
struct TileWriter {
    unsigned pos = 0, end = 0; // unused slots of the current tile, as indices inside the shard region
    unsigned claimed = 0; // base of the tile claimed ahead
    bool ahead = false;

    __device__ __forceinline__ void prepare(ShardCursor* cursor, unsigned tile, unsigned lane)
    {
        if (!ahead && end - pos < 32u) {
            if (lane == 0)
                claimed = atomicAdd(&cursor->stored, tile);
            ahead = true;
        }
    }
    // one slot for every lane in `mask` (at most 32 - a prepare() came first); returns this lane's slot inside the region
    __device__ __forceinline__ unsigned take(unsigned mask, unsigned lane, unsigned tile, unsigned region, unsigned int* overflow)
    {
        const unsigned n = __popc(mask), room = end - pos, first = pos;
        unsigned fresh = 0;
        if (n > room) { // move on to the tile claimed ahead
            fresh = __shfl_sync(kFull, claimed, 0);
            ahead = false;
            if (fresh + tile > region) { // region full: the run is flagged invalid, stores stay in bounds
                if (lane == 0)
                    atomicExch(overflow, 1u);
                fresh = 0;
            }
            pos = fresh + (n - room);
            end = fresh + tile;
        } else {
            pos += n;
        }
        const unsigned rank = __popc(mask & ((1u << lane) - 1u));
        return rank < room ? first + rank : fresh + (rank - room);
    }
    // calls mark(slot) for every slot the warp claimed and did not use
    template <typename Mark>
    __device__ __forceinline__ void finish(unsigned lane, unsigned tile, unsigned region, Mark mark)
    {
        for (unsigned slot = pos + lane; slot < end; slot += 32)
            mark(slot);
        const unsigned spare = __shfl_sync(kFull, claimed, 0);
        if (ahead && spare + tile <= region)
            for (unsigned slot = spare + lane; slot < spare + tile; slot += 32)
                mark(slot);
    }
};

// ---- (b) Woodcock delta tracking (transport.hpp:640-700) -------------------------------------------
constexpr unsigned kMaxBrickWords = 512; // the brick grid holds at most 16384 bricks (rule shared with the CPU restatement)

#ifndef DXMCB200_TK_MINBLOCKS
#define DXMCB200_TK_MINBLOCKS 6
#endif
// Persistent grid; a lane steps one photon until it leaves the world, loses Russian roulette, or a real / forced interaction
// (or, with the empty-space traversal, an air walk) is due. When `refillBatch` lanes are empty the warp services them together:
// lanes with an interaction due write an event record, lanes bound for the air walk a photon record (both through TileWriters),
// and all empty lanes take the next records of the wave, straight from global memory: the warp claims tiles of kTile
// consecutive records with one atomic, asks for the whole tile in L2 at once (prefetch), and the lanes' 64-byte loads of
// consecutive records coalesce. With the empty-space traversal a photon stays for 3 steps on average, so the service code
// weighs as much as the step itself, and this form of it takes a third of the instructions of the staged one it replaced (a
// per-warp cp.async ring in shared memory, which hid the load latency but cost 190 warp instructions per warp-step in
// bookkeeping); the latency is left to the other warps, and the shared memory it frees goes back to L1.
// kNibbleGrid: the voxel grid is a 4-bit palette and the spacings are powers of two (segmentation phantoms at 1 mm, the bench
// workload): the three warp-uniform tests "palette form? 4-bit? exact inverse spacing?" leave the step loop (10 of 224 warp
// instructions per step). false: any grid, tested at run time.
template <bool kStats, bool kAir, bool kNibbleGrid>
__global__ void __launch_bounds__(kThreads, DXMCB200_TK_MINBLOCKS) transportKernel(const __grid_constant__ KernelParams P)
{
    __shared__ uint2 sPalette[256];

    const unsigned lane = threadIdx.x & 31u;
    const unsigned laneLt = (1u << lane) - 1u;
    const unsigned myShard = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) % kShards;
    const bool paletteForm = kNibbleGrid || P.world.palette != nullptr;
    if (paletteForm)
        sPalette[threadIdx.x] = P.world.paletteTable[threadIdx.x];
    __syncthreads();
    // shared-state-space address of the table: the look-up below is one LEA + LDS instead of a generic-window address
    const unsigned paletteBase = static_cast<unsigned>(__cvta_generic_to_shared(sPalette));

    Rng rng { 0, 1 };
    Photon p {};
    float logE = 0.0f, maxAttInv = 0.0f, eventProbability = 0.0f;
    uint32_t seg = 0, voxel = 0, material = 0;
    bool lowWeight = false; // E*w below the Russian-roulette threshold (transport.hpp:684)
    uint32_t state = DEAD;
    uint32_t cSteps = 0, cLookups = 0;

    // input (warp-uniform): records [inPos, inEnd) of the tile in hand; tiles come from the warp's own shard first, then round the others
    unsigned inPos = 0, inEnd = 0;
    unsigned inShard = myShard, shardsTried = 0;
    TileWriter events, airborne;
    unsigned exhaustedMask = 0;
    unsigned refillAt = P.refillBatch; // empty lanes that trigger a service: refillBatch + the exhausted ones

    // claim the next tile of the wave; false when every shard is drained
    auto claimTile = [&]() {
        while (shardsTried < kShards) {
            const unsigned n = min(P.inCursors[inShard].stored, P.photonRegion); // filled by completed kernels
            unsigned t = n;
            if (lane == 0 && n)
                t = atomicAdd(&P.inCursors[inShard].taken, kTile);
            t = __shfl_sync(kFull, t, 0);
            if (t < n) {
                inPos = inShard * P.photonRegion + t;
                inEnd = inShard * P.photonRegion + min(t + kTile, n);
                // the whole tile on its way into L2: one 128-byte line per two records
                const char* lines = reinterpret_cast<const char*>(P.photonsIn + inPos);
                for (unsigned k = lane; 2 * k < inEnd - inPos; k += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(lines + 128u * k));
                return true;
            }
            inShard = (inShard + 1) % kShards;
            ++shardsTried;
        }
        return false;
    };

    for (;;) {
        // ---- one Woodcock step (transport.hpp:655-682)
        if (state == STEP) {
            const float r1 = rng.uniform();
            advance(p, (fastLog2(r1) * kStepScale) * maxAttInv);
            if constexpr (kStats)
                ++cSteps;
            if (!insideWorld(P.world, p.px, p.py, p.pz)) {
                state = DEAD;
            } else {
                uint2 rec;
                // random look-ups have no reuse in L1: cache them in L2 only and leave L1 to the LUT coefficients
                if constexpr (kNibbleGrid) {
                    const uint32_t ix = __float2uint_rz(__fmul_rn(__fsub_rn(p.px, P.world.ext[0]), P.world.invSpacing[0]));
                    const uint32_t iy = __float2uint_rz(__fmul_rn(__fsub_rn(p.py, P.world.ext[2]), P.world.invSpacing[1]));
                    const uint32_t iz = __float2uint_rz(__fmul_rn(__fsub_rn(p.pz, P.world.ext[4]), P.world.invSpacing[2]));
                    voxel = (iz * P.world.dim[1] + iy) * P.world.dim[0] + ix;
                    const unsigned index = (static_cast<uint32_t>(__ldcg(P.world.palette + (voxel >> 1))) >> ((voxel & 1u) * 4u)) & 15u;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rec.x), "=r"(rec.y) : "r"(paletteBase + 8u * index));
                } else {
                    voxel = voxelIndex(P.world, p.px, p.py, p.pz);
                    if (paletteForm) {
                        const unsigned slot = paletteBase + 8u * paletteIndex(P.world, voxel);
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rec.x), "=r"(rec.y) : "r"(slot));
                    } else
                        rec = __ldcg(P.world.voxels + voxel);
                }
                const float density = __uint_as_float(rec.x);
                material = rec.y;
                if constexpr (kStats)
                    ++cLookups;
                float aP, aC, aR;
                attenuationAt(P.lut, material & 0xffu, seg, logE, aP, aC, aR);
                const float attTotal = ((aP + aC) + aR) * density; // the reference's 0 + aP is aP exactly
                eventProbability = attTotal * maxAttInv;
                bool event = (material & 0xff00u) != 0; // measurement voxel: forced interaction, no draw
                if (!event)
                    event = rng.uniform() < eventProbability;
                if (event) {
                    state = EVENT;
                } else {
                    if (lowWeight) { // Russian roulette after a virtual collision (transport.hpp:684-693)
                        const float r4 = rng.uniform();
                        if (r4 < kRouletteProbability) {
                            state = DEAD;
                        } else {
                            constexpr float factor = 1.0f / (1.0f - kRouletteProbability);
                            p.weight *= factor;
                            lowWeight = p.energy * p.weight < kRouletteThreshold;
                        }
                    }
                    if constexpr (kAir) { // a virtual collision in an air brick (bit 16 of the voxel record): the photon leaves for the air walk
                        if (state == STEP && (material & kInAirBrick))
                            state = AIRBORNE;
                    }
                }
            }
        }
        const unsigned notStepping = __ballot_sync(kFull, state != STEP); // lanes with an event due, dead or exhausted
        if (static_cast<unsigned>(__popc(notStepping)) < refillAt && notStepping != kFull)
            continue; // keep stepping until enough lanes are empty

        // ---- service. Tiles for this round's (or the next one's) records are claimed first, nothing waits for them yet
        events.prepare(P.eventCursors + myShard, kEventTile, lane);
        if constexpr (kAir)
            airborne.prepare(P.airCursors + myShard, kAirTile, lane);
        // lanes with an interaction due write an event record and become empty
        const unsigned eventMask = __ballot_sync(kFull, state == EVENT);
        if (eventMask) {
            const unsigned slot = events.take(eventMask, lane, kEventTile, P.eventRegion, P.overflow);
            if (state == EVENT) {
                storeEvent(P.events + static_cast<size_t>(myShard) * P.eventRegion + slot, p, rng, logE, seg, voxel, material, eventProbability);
                state = DEAD;
            }
        }
        if constexpr (kAir) { // lanes whose photon stands in an air brick hand it to the air walk
            const unsigned airMask = __ballot_sync(kFull, state == AIRBORNE);
            if (airMask) {
                const unsigned slot = airborne.take(airMask, lane, kAirTile, P.photonRegion, P.overflow);
                if (state == AIRBORNE) {
                    storePhoton(P.airborne + static_cast<size_t>(myShard) * P.photonRegion + slot, p, rng, logE, maxAttInv, seg, 0.0f);
                    state = DEAD;
                }
            }
        }
        // empty lanes take the next records of the wave (again when a tile ends half-way, or a record is a dead marker)
        unsigned deadMask = __ballot_sync(kFull, state == DEAD);
        while (deadMask) {
            if (inPos == inEnd && !claimTile()) { // drained: nothing left to hand out in this wave
                if (state == DEAD)
                    state = EXHAUSTED;
                exhaustedMask |= deadMask;
                refillAt = P.refillBatch + static_cast<unsigned>(__popc(exhaustedMask));
                break;
            }
            const unsigned rank = __popc(deadMask & laneLt);
            if (state == DEAD && rank < inEnd - inPos) {
                float unused;
                loadPhoton(P.photonsIn + inPos + rank, p, rng, logE, maxAttInv, seg, unused);
                lowWeight = p.energy * p.weight < kRouletteThreshold;
                state = p.energy > 0.0f ? STEP : DEAD; // energy 0: an unused slot of an interactKernel tile
            }
            inPos = min(inEnd, inPos + static_cast<unsigned>(__popc(deadMask)));
            deadMask = __ballot_sync(kFull, state == DEAD);
        }
        if (exhaustedMask == kFull)
            break;
    }
    // dead markers in what is left of the warp's tiles
    events.finish(lane, kEventTile, P.eventRegion, [&](unsigned slot) {
        recStore(&P.events[static_cast<size_t>(myShard) * P.eventRegion + slot].where, make_uint4(0u, kNoEvent, 0u, 0u));
    });
    if constexpr (kAir)
        airborne.finish(lane, kAirTile, P.photonRegion, [&](unsigned slot) {
            storeDeadPhoton(&P.airborne[static_cast<size_t>(myShard) * P.photonRegion + slot]);
        });

    if constexpr (kStats) {
        const unsigned long long s = warpSum(cSteps), l = warpSum(cLookups);
        if (lane == 0) {
            atomicAdd(&P.counters->steps, s);
            atomicAdd(&P.counters->lookups, l);
        }
    }
}

// ---- (c) interactions + (d) scoring: one event per thread ------------------------------------------
// computeInteractions (transport.hpp:583-638). The channels cost very different amounts of work per event (ncu: the RITA search of
// the Rayleigh sampler runs at 2.3 of 32 lanes, the Klein-Nishina loop at 9), but regrouping events by channel has not paid in any
// form tried: a block-level sort and a per-warp Rayleigh queue in shared memory (round 1), and handing Rayleigh events and
// Compton events with two rejected trials on to channel-pure follow-up passes through lists in HBM (round 2: 2.77e9 -> 2.64e9
// histories/s; Rayleigh alone: 2.66e9). The kernel is bound by the latency of its loads and atomics, not by issue slots, and
// every extra pass adds a load-compute-atomic-store chain per event it touches. Staging each thread's next event through 64 bytes
// of shared memory with cp.async does hide the event load (this kernel 191 -> 171 ms per 1e9 histories), but the 96 KB of shared
// memory per SM come out of the L1 of every co-resident kernel: 3.04e9 -> 2.83e9 overall (profiles/README.md).
template <int L, bool kStats, bool kAggregate>
__global__ void __launch_bounds__(kThreads, 6) interactKernel(const __grid_constant__ KernelParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    uint32_t cInter = 0, cScores = 0;
    // block b works on event shard b % kShards together with the other blocks of the same residue
    const unsigned blocksPerShard = gridDim.x / kShards; // the grid is a multiple of kShards
    const unsigned shard = blockIdx.x % kShards;
    // transportKernel claims slots in tiles of kEventTile, the walk kernels claim exact counts: any number of slots. The trip
    // count is the same for all lanes of a warp (the loop body holds full-mask ballots and shuffles); slots past the end read
    // as unused.
    const unsigned nSlots = min(P.eventCursors[shard].stored, P.eventRegion);
    const EventRecord* const region = P.events + static_cast<size_t>(shard) * P.eventRegion;
    // survivors go to the next wave through a TileWriter: its atomic is issued at the start of a trip and hides behind the sampling
    const unsigned outShard = (blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) % kShards;
    TileWriter out;
    unsigned survivors = 0; // photons this warp has stored
    for (unsigned base = (blockIdx.x / kShards) * kThreads; base < nSlots; base += blocksPerShard * kThreads) {
        const unsigned i = base + threadIdx.x;
        const EventRecord* e = region + min(i, nSlots - 1u);
        const uint4 where = i < nSlots ? recLoad(&e->where) : make_uint4(0u, kNoEvent, 0u, 0u);
        if (__any_sync(kFull, where.y != kNoEvent))
            out.prepare(P.outCursors + outShard, kSurvivorTile, lane);
        bool alive = false; // the photon goes on to the next wave
        bool done = false; // the interaction has been sampled (false: no event in this slot)
        bool energyChanged = false;
        Photon p {};
        Rng rng { 0, 1 };
        float logE = 0.0f, maxAttInv = 0.0f;
        uint32_t seg = 0;
        Deflection turn;
        ScoreSlot score, forcedScore; // deposits of this event, scored once the warp has reconverged
        const uint32_t mat = where.y & 0xffu;
        const uint32_t flags = where.y & ((1u << kSegmentShift) - 1u); // material | measurement << 8 | in-air-brick << 16
        // energy `imparted` was given up by the photon in a photoelectric or Compton event (transport.hpp:598-625)
        auto afterEnergyLoss = [&](float imparted) {
            if constexpr (kStats)
                ++cScores;
            if (p.energy < kEnergyCutoff) { // absorbed here: what is left of the energy stays in the voxel, nothing to turn
                deposit<kAggregate>(P, score, where.x, (imparted + p.energy) * p.weight);
                p.energy = 0.0f;
                turn.scattered = false;
                alive = false;
            } else {
                deposit<kAggregate>(P, score, where.x, imparted * p.weight);
                energyChanged = true;
                alive = true;
            }
            done = true;
        };
        if (where.y != kNoEvent) {
            Pending pe;
            {
                const float4 a = recLoad(&e->posE), b = recLoad(&e->dirW);
                const uint4 c = recLoad(&e->rng);
                p.px = a.x, p.py = a.y, p.pz = a.z, p.energy = a.w;
                p.dx = b.x, p.dy = b.y, p.dz = b.z, p.weight = b.w;
                rng.state = (static_cast<uint64_t>(c.y) << 32) | c.x;
                rng.inc = (static_cast<uint64_t>(c.w) << 32) | c.z;
            }
            logE = __uint_as_float(where.z);
            seg = where.y >> kSegmentShift;
            pe.eventProbability = __uint_as_float(where.w);
            pe.voxel = where.x;
            pe.material = flags;
            attenuationAt(P.lut, mat, seg, logE, pe.attPhoto, pe.attCompton, pe.attRayleigh);
            if (flags & 0xff00u) {
                alive = interactForced<L, kStats, kAggregate>(P, p, pe, rng, energyChanged, cScores, score, forcedScore, turn);
                done = true;
            } else { // the channel draw of computeInteractions (transport.hpp:590-596)
                const float attTotal = ((0.0f + pe.attPhoto) + pe.attCompton) + pe.attRayleigh;
                const float r3 = rng.uniform(attTotal);
                if (r3 < pe.attPhoto) {
                    afterEnergyLoss(photoAbsorptionDeferred<L>(P.lut, p, mat, rng, turn));
                } else if (r3 < (pe.attPhoto + pe.attCompton)) {
                    afterEnergyLoss(comptonScatterDeferred<L>(P.lut, p, mat, rng, turn));
                } else {
                    rayleighScatterDeferred<L>(P.lut, p, mat, rng, turn);
                    alive = true;
                    done = true;
                }
            }
        }
        if (done) {
            if (!(p.energy <= 3.0e38f)) // NaN or inf from the sampler (see deposit): the history ends here
                alive = false;
            deflect(p, turn, rng); // one azimuth draw + rotation for whichever channel scattered
            if constexpr (kStats)
                ++cInter;
            // Russian roulette (transport.hpp:684-693)
            if (alive && p.energy * p.weight < kRouletteThreshold) {
                const float r4 = rng.uniform();
                if (r4 < kRouletteProbability) {
                    alive = false;
                } else {
                    constexpr float factor = 1.0f / (1.0f - kRouletteProbability);
                    p.weight *= factor;
                }
            }
            if (alive) { // the next wave's record carries the majorant of the photon's energy
                if (energyChanged)
                    energyDependent(P.lut, p.energy, logE, seg, maxAttInv);
                else
                    maxAttInv = maxAttenuationInverse(P.lut, logE);
            }
        }
        if constexpr (kAggregate) {
            scoreWarp(P, score, lane);
            if (__any_sync(kFull, forcedScore.voxel != kNoVoxel)) // forced interactions only (measurement voxels)
                scoreWarp(P, forcedScore, lane);
        }
        const unsigned aliveMask = __ballot_sync(kFull, alive);
        if (aliveMask == 0)
            continue;
        survivors += __popc(aliveMask);
        const unsigned slot = out.take(aliveMask, lane, kSurvivorTile, P.photonRegion, P.overflow);
        if (alive)
            storePhoton(P.photonsOut + static_cast<size_t>(outShard) * P.photonRegion + slot, p, rng, logE, maxAttInv, seg, 0.0f);
    }
    if (lane == 0 && survivors)
        atomicAdd(&P.outCursors[outShard].live, survivors);
    out.finish(lane, kSurvivorTile, P.photonRegion, [&](unsigned slot) { // dead markers (energy 0): transportKernel drops them at the re-fill
        storeDeadPhoton(&P.photonsOut[static_cast<size_t>(outShard) * P.photonRegion + slot]);
    });
    if constexpr (kStats) {
        const unsigned long long i = warpSum(cInter), sc = warpSum(cScores);
        if (lane == 0) {
            atomicAdd(&P.counters->interactions, i);
            atomicAdd(&P.counters->scores, sc);
        }
    }
}

// ---- exposure table on the device: exposure i of a source is a pure function of its parameter block ----
static_assert(sizeof(dxmcb200_source_params) == sizeof(dxmc::model::SourceParams<float>), "dxmcb200_source_params mirrors model::SourceParams<float>");
static_assert(offsetof(dxmcb200_source_params, world_cosines) == offsetof(dxmc::model::SourceParams<float>, worldCosines), "layout of the parameter block");
static_assert(offsetof(dxmcb200_source_params, sdd) == offsetof(dxmc::model::SourceParams<float>, sdd), "layout of the parameter block");
static_assert(offsetof(dxmcb200_source_params, xcare_angle) == offsetof(dxmc::model::SourceParams<float>, xcareAngle), "layout of the parameter block");

__global__ void exposureKernel(const __grid_constant__ dxmc::model::SourceParams<float> source, const float* __restrict__ tubeCurrent, uint64_t n,
    dxmcb200_exposure* __restrict__ out)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    dxmc::model::ExposureValues<float> v;
    dxmc::model::evaluate(source, tubeCurrent, i, v);
    dxmcb200_exposure e;
    for (int k = 0; k < 3; ++k) {
        e.position[k] = v.position[k];
        e.beam_direction[k] = v.beam[k];
    }
    for (int k = 0; k < 6; ++k)
        e.cosines[k] = v.cosines[k];
    for (int k = 0; k < 4; ++k)
        e.collimation[k] = v.collimation[k];
    e.weight = v.weight;
    e.mono_energy = v.monoEnergy;
    e.spectrum = v.spectrum;
    e.heel = v.heel;
    e.bowtie = v.bowtie;
    e.reserved = 0;
    e.histories = v.histories;
    out[i] = e;
}

// reset the cursors of one buffer between waves
__global__ void resetCursorsKernel(ShardCursor* cursors, int storedToo)
{
    {
        cursors[threadIdx.x].taken = 0;
        if (storedToo) {
            cursors[threadIdx.x].stored = 0;
            cursors[threadIdx.x].live = 0;
        }
    }
}

// ---- small kernels -----------------------------------------------------------------------------
__global__ void packVoxelsKernel(const float* __restrict__ density, const uint8_t* __restrict__ material,
    const uint8_t* __restrict__ measurement, uint2* __restrict__ out, uint64_t n)
{
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t m = material[i] | (measurement ? (static_cast<uint32_t>(measurement[i]) << 8) : 0u);
        out[i] = make_uint2(__float_as_uint(density[i]), m);
    }
}

// ---- palette form of the voxel grid ---------------------------------------------------------------
// Segmentation phantoms hold a handful of distinct {density, material, measurement} records. When there are at
// most 256 of them the grid is stored as one byte per voxel plus a 256-entry record table: lossless, 8x less
// look-up traffic, and at 512x512x400 small enough to stay resident in the 126 MB L2.
constexpr unsigned kPaletteSlots = 4096; // open-addressing hash table of 64-bit record keys
constexpr unsigned long long kEmptyKey = ~0ULL;

__device__ __forceinline__ unsigned long long voxelKey(const float* density, const uint8_t* material, const uint8_t* measurement, uint64_t i)
{
    const uint32_t m = material[i] | (measurement ? (static_cast<uint32_t>(measurement[i]) << 8) : 0u);
    return (static_cast<unsigned long long>(m) << 32) | __float_as_uint(density[i]);
}

__device__ __forceinline__ unsigned paletteHash(unsigned long long key)
{
    return static_cast<unsigned>((key * 0x9E3779B97F4A7C15ULL) >> 40) & (kPaletteSlots - 1);
}

// pass 1: collect the distinct records; *distinct > 256 (or a full table) means "no palette"
__global__ void paletteCollectKernel(const float* __restrict__ density, const uint8_t* __restrict__ material,
    const uint8_t* __restrict__ measurement, uint64_t n, unsigned long long* table, unsigned* distinct)
{
    const uint64_t threads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t per = (n + threads - 1) / threads;
    const uint64_t begin = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) * per;
    const uint64_t end = begin + per < n ? begin + per : n;
    unsigned long long last = kEmptyKey;
    for (uint64_t i = begin; i < end; ++i) { // a thread walks a contiguous run, so repeated records cost nothing
        const unsigned long long key = voxelKey(density, material, measurement, i);
        if (key == last)
            continue;
        last = key;
        if (*reinterpret_cast<volatile unsigned*>(distinct) > 256u)
            return;
        unsigned h = paletteHash(key);
        bool placed = false;
        for (unsigned probe = 0; probe < kPaletteSlots && !placed; ++probe) {
            const unsigned long long old = atomicCAS(table + h, kEmptyKey, key);
            if (old == kEmptyKey) {
                atomicAdd(distinct, 1u);
                placed = true;
            } else if (old == key) {
                placed = true;
            }
            h = (h + 1) & (kPaletteSlots - 1);
        }
        if (!placed)
            atomicAdd(distinct, 1000u);
    }
}

// pass 2 (one block): number the occupied slots
__global__ void paletteNumberKernel(const unsigned long long* table, unsigned* slotIndex, uint2* paletteTable)
{
    __shared__ unsigned next;
    if (threadIdx.x == 0)
        next = 0;
    __syncthreads();
    for (unsigned s = threadIdx.x; s < kPaletteSlots; s += blockDim.x) {
        const unsigned long long key = table[s];
        if (key != kEmptyKey) {
            const unsigned idx = atomicAdd(&next, 1u);
            slotIndex[s] = idx;
            if (idx < 256u)
                paletteTable[idx] = make_uint2(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32));
        }
    }
}

// pass 3: one byte per voxel (or, with at most 16 distinct records, one byte per two voxels)
__global__ void paletteIndexKernel(const float* __restrict__ density, const uint8_t* __restrict__ material,
    const uint8_t* __restrict__ measurement, uint64_t n, const unsigned long long* __restrict__ table, const unsigned* __restrict__ slotIndex,
    uint8_t* __restrict__ out, int nibbles)
{
    auto indexOf = [&](uint64_t i) {
        const unsigned long long key = voxelKey(density, material, measurement, i);
        unsigned h = paletteHash(key);
        while (table[h] != key)
            h = (h + 1) & (kPaletteSlots - 1);
        return slotIndex[h];
    };
    const uint64_t items = nibbles ? (n + 1) / 2 : n; // output bytes
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < items; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        if (nibbles) {
            const unsigned lo = indexOf(2 * i);
            const unsigned hi = 2 * i + 1 < n ? indexOf(2 * i + 1) : 0u;
            out[i] = static_cast<uint8_t>(lo | (hi << 4));
        } else {
            out[i] = static_cast<uint8_t>(indexOf(i));
        }
    }
}

// Brick grid of the empty-space traversal: per brick, the maximum over its voxels of density * ratio[material] (ratio[m] =
// the largest mu_total,m(E) / majorant(E) over the table energies, computed on the host), and whether it holds a measurement
// voxel. One coalesced pass over the grid; a warp covers 32 consecutive voxels of a row, i.e. one or two bricks: lanes are
// grouped by brick (match.any) and each group issues one atomicMax. Positive floats order like their bit patterns; negative
// and NaN products never win (they count as 0, like the oracle's `f > max` test).
__global__ void brickMaxKernel(WorldView w, BrickView b, const float* __restrict__ ratio, uint64_t n, unsigned* __restrict__ brickMaxBits,
    unsigned* __restrict__ brickMeasured)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t trips = (n + stride - 1) / stride; // the same for every thread: the warp stays converged
    const uint64_t plane = static_cast<uint64_t>(w.dim[0]) * w.dim[1];
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    for (uint64_t t = 0; t < trips; ++t, i += stride) {
        unsigned brick = 0xffffffffu, bits = 0u, measured = 0u;
        if (i < n) {
            const uint32_t iz = static_cast<uint32_t>(i / plane);
            const uint32_t rem = static_cast<uint32_t>(i - iz * plane);
            const uint32_t iy = rem / w.dim[0], ix = rem - iy * w.dim[0];
            const uint2 r = voxelRecord(w, static_cast<uint32_t>(i));
            const float f = __fmul_rn(__uint_as_float(r.x), ratio[r.y & 0xffu]);
            bits = f > 0.0f ? __float_as_uint(f) : 0u;
            measured = (r.y & 0xff00u) ? 1u : 0u;
            brick = brickOfVoxel(b, ix, iy, iz);
        }
        const unsigned group = __match_any_sync(kFull, brick);
        const unsigned best = __reduce_max_sync(group, bits);
        const unsigned any = __reduce_or_sync(group, measured);
        if (brick != 0xffffffffu && lane == static_cast<unsigned>(__ffs(group) - 1)) {
            atomicMax(brickMaxBits + brick, best);
            if (any)
                atomicOr(brickMeasured + brick, 1u);
        }
    }
}

// ---- "this voxel lies in an air brick" as bit 16 of the voxel record --------------------------------------
// transportKernel hands a photon to the air walk after a virtual collision in an air brick. Testing the brick bitmap for that
// cost a brick index and a shared-memory look-up on 80 % of all steps (7 % of the kernel's instructions); with the answer in
// the record the look-up brings it along. Record grids get the bit per voxel. Palette grids get, for every entry that can
// occur in an air brick, a second entry with the bit set, and the voxels inside air bricks are re-indexed to it (the tables
// `baseOf` / `flaggedOf` map an index to its plain and its flagged entry, so the pass can be repeated for another brick grid).

__device__ __forceinline__ bool voxelInAirBrick(const WorldView& w, const BrickView& b, const uint32_t* __restrict__ bitmap, uint32_t voxel)
{
    const uint32_t plane = w.dim[0] * w.dim[1];
    const uint32_t iz = voxel / plane, rem = voxel - iz * plane;
    const uint32_t iy = rem / w.dim[0], ix = rem - iy * w.dim[0];
    const uint32_t brick = brickOfVoxel(b, ix, iy, iz);
    return (bitmap[brick >> 5] >> (brick & 31u)) & 1u;
}

// bitmap == nullptr clears the flags
__global__ void flagRecordsKernel(WorldView w, BrickView b, const uint32_t* __restrict__ bitmap, uint64_t n, uint2* __restrict__ voxels)
{
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        uint2 r = voxels[i];
        const uint32_t y = (r.y & ~kInAirBrick) | (bitmap && voxelInAirBrick(w, b, bitmap, static_cast<uint32_t>(i)) ? kInAirBrick : 0u);
        if (y != r.y) {
            r.y = y;
            voxels[i] = r;
        }
    }
}

// remap = baseOf[256] followed by flaggedOf[256]
__global__ void flagPaletteKernel(WorldView w, BrickView b, const uint32_t* __restrict__ bitmap, uint64_t n, const uint8_t* __restrict__ remap,
    uint8_t* __restrict__ palette, int nibbles)
{
    __shared__ uint8_t sRemap[512];
    for (unsigned k = threadIdx.x; k < 512; k += blockDim.x)
        sRemap[k] = remap[k];
    __syncthreads();
    auto mapped = [&](uint64_t voxel, unsigned index) -> unsigned {
        const unsigned plain = sRemap[index];
        return bitmap && voxelInAirBrick(w, b, bitmap, static_cast<uint32_t>(voxel)) ? sRemap[256 + plain] : plain;
    };
    const uint64_t items = nibbles ? (n + 1) / 2 : n; // bytes
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < items; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const unsigned old = palette[i];
        unsigned now;
        if (nibbles) {
            const unsigned lo = mapped(2 * i, old & 15u);
            const unsigned hi = 2 * i + 1 < n ? mapped(2 * i + 1, old >> 4) : 0u;
            now = lo | (hi << 4);
        } else {
            now = mapped(i, old);
        }
        if (now != old)
            palette[i] = static_cast<uint8_t>(now);
    }
}

// a palette grid that has outgrown its form: 4-bit indices to bytes, bytes to 8-byte records
__global__ void expandNibblesKernel(const uint8_t* __restrict__ nibbles, uint64_t n, uint8_t* __restrict__ bytes)
{
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        bytes[i] = (nibbles[i >> 1] >> ((i & 1u) * 4u)) & 15u;
}
__global__ void expandPaletteKernel(const uint8_t* __restrict__ palette, const uint2* __restrict__ table, uint64_t n, uint2* __restrict__ voxels)
{
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        voxels[i] = table[palette[i]];
}

// Per-material maximum density of the grid: the input of the Woodcock majorant (attenuationinterpolator.hpp:48-59,
// where it is one transform_reduce(init 0, max) over all voxels per material). Positive floats order like their bit
// patterns; a negative, -0.0f or NaN density (nothing upstream rejects them) counts as 0, exactly like std::max(0, x)
// leaves the reference's running maximum untouched. Per trip a warp groups its lanes by material (match.any), takes each
// group's maximum (redux.sync) and issues ONE shared-memory atomicMax per distinct material; block maxima go out with
// one global atomicMax per material.
__global__ void maxDensityKernel(const uint2* __restrict__ records, uint64_t n, unsigned* __restrict__ maxBits)
{
    __shared__ unsigned sMax[257]; // slot 256: lanes past the end of the grid
    for (unsigned k = threadIdx.x; k < 257; k += blockDim.x)
        sMax[k] = 0u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    const uint64_t trips = (n + stride - 1) / stride; // the same for every thread: the warp stays converged
    uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    for (uint64_t t = 0; t < trips; ++t, i += stride) {
        unsigned material = 256u, bits = 0u;
        if (i < n) {
            const uint2 r = records[i];
            material = r.y & 0xffu;
            bits = __uint_as_float(r.x) > 0.0f ? r.x : 0u; // sign bit or NaN: ignored
        }
        const unsigned group = __match_any_sync(kFull, material);
        const unsigned best = __reduce_max_sync(group, bits);
        if (lane == static_cast<unsigned>(__ffs(group) - 1))
            atomicMax(&sMax[material], best);
    }
    __syncthreads();
    if (threadIdx.x < 256 && sMax[threadIdx.x])
        atomicMax(maxBits + threadIdx.x, sMax[threadIdx.x]);
}

// normalizeScoring / energyImpartedToDose (transport.hpp:780-816) fused with the fixed-point decode
__global__ void resultKernel(const unsigned long long* __restrict__ acc, const uint2* __restrict__ voxels, const uint8_t* __restrict__ palette,
    const uint2* __restrict__ paletteTable, int paletteNibbles, uint64_t first, uint64_t n, int mode,
    float energyLsb, float energySqLsb, uint64_t histories, float calibration, float voxelVolume, float* __restrict__ dose,
    uint32_t* __restrict__ nEvents, float* __restrict__ variance)
{
    const float hInv = 1.0f / static_cast<float>(histories - 1);
    const float hdInv = 1.0e3f / static_cast<float>(histories);
    const float hvInv = 1.0e6f / static_cast<float>(histories);
    // voxels [first, first + n): the output arrays are indexed from the start of that range
    for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t i = first + k;
        const ulonglong4 a = reinterpret_cast<const ulonglong4*>(acc)[i];
        const float e = static_cast<float>(static_cast<double>(static_cast<long long>(a.x)) * static_cast<double>(energyLsb));
        const float e2 = static_cast<float>(static_cast<double>(a.y) * static_cast<double>(energySqLsb));
        float d = e, v = e2;
        if (mode == 0) {
            d = e * hdInv;
            v = (e2 * hvInv - d * d) * hInv;
        } else if (mode == 1) {
            const unsigned slot = !palette ? 0u : paletteNibbles ? (palette[i >> 1] >> ((i & 1) * 4)) & 15u : palette[i];
            const float de = __uint_as_float(palette ? paletteTable[slot].x : voxels[i].x);
            const float voxelMass = de * voxelVolume * 0.001f;
            const float factor = calibration / voxelMass;
            d = de > 0.0f ? e * factor : 0.0f;
            v = de > 0.0f ? e2 * factor * factor : 0.0f;
        }
        if (dose)
            dose[k] = d;
        if (variance)
            variance[k] = v;
        if (nEvents)
            nEvents[k] = static_cast<uint32_t>(a.z);
    }
}

__global__ void rawKernel(const unsigned long long* __restrict__ acc, uint64_t n, long long* energy, unsigned long long* energySq,
    unsigned long long* events)
{
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const ulonglong4 a = reinterpret_cast<const ulonglong4*>(acc)[i];
        if (energy)
            energy[i] = static_cast<long long>(a.x);
        if (energySq)
            energySq[i] = a.y;
        if (events)
            events[i] = a.z;
    }
}

__global__ void evalAttenuationKernel(LutView lut, uint64_t n, const uint8_t* material, const float* energy, float* out3, float* outMax)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    const float logE = log10Rounded(energy[i]);
    float a, b, c;
    attenuation(lut, material[i], logE, a, b, c);
    out3[3 * i + 0] = a;
    out3[3 * i + 1] = b;
    out3[3 * i + 2] = c;
    outMax[i] = maxAttenuationInverse(lut, logE);
}

__global__ void traceIndicesKernel(WorldView w, uint64_t nRays, const float* pos, const float* dir, uint32_t nSteps, const float* steps,
    long long* outIdx, float* outEntry)
{
    const uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= nRays)
        return;
    Photon p {};
    p.px = pos[3 * r];
    p.py = pos[3 * r + 1];
    p.pz = pos[3 * r + 2];
    p.dx = dir[3 * r];
    p.dy = dir[3 * r + 1];
    p.dz = dir[3 * r + 2];
    long long* o = outIdx + r * (nSteps + 1);
    bool in = transportToWorld(w, p);
    outEntry[3 * r] = p.px;
    outEntry[3 * r + 1] = p.py;
    outEntry[3 * r + 2] = p.pz;
    // the entry point sits on the safe extent; the reference only indexes after the first step
    o[0] = (in && insideWorld(w, p.px, p.py, p.pz)) ? static_cast<long long>(voxelIndex(w, p.px, p.py, p.pz)) : -1;
    for (uint32_t k = 0; k < nSteps; ++k) {
        if (in) {
            advance(p, steps[k]);
            in = insideWorld(w, p.px, p.py, p.pz);
        }
        o[k + 1] = in ? static_cast<long long>(voxelIndex(w, p.px, p.py, p.pz)) : -1;
    }
}

__global__ void sampleParticlesKernel(dxmcb200_exposure e, BeamView beams, uint64_t exposureIndex, uint64_t seed, uint64_t n, float* out)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    Rng rng;
    historyStream(seed, exposureIndex, i, rng.state, rng.inc);
    const Photon p = sampleParticle(e, beams, rng);
    float* o = out + 8 * i;
    o[0] = p.px;
    o[1] = p.py;
    o[2] = p.pz;
    o[3] = p.dx;
    o[4] = p.dy;
    o[5] = p.dz;
    o[6] = p.energy;
    o[7] = p.weight;
}

// test hook of the empty-space traversal: the air run of fixed rays (entry into the world first, like a birth)
__global__ void traceAirRunsKernel(WorldView w, BrickView b, uint64_t nRays, const float* pos, const float* dir, float* outLength, uint32_t* outInfo,
    float* outEnd)
{
    const uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (r >= nRays)
        return;
    Photon p {};
    p.px = pos[3 * r], p.py = pos[3 * r + 1], p.pz = pos[3 * r + 2];
    p.dx = dir[3 * r], p.dy = dir[3 * r + 1], p.dz = dir[3 * r + 2];
    float length = 0.0f;
    uint32_t info = 0;
    if (transportToWorld(w, p)) {
        info |= 1u << 18; // reaches the world
        if (inAirBrick(w, b, p.px, p.py, p.pz)) {
            bool exits = false;
            uint32_t crossed = 0;
            length = airRunLength(w, b, p, exits, crossed);
            info |= crossed | (exits ? 1u << 16 : 0u) | (1u << 17); // cubes crossed, leaves the grid, started in an air brick
            advance(p, length);
        }
    }
    outLength[r] = length;
    outInfo[r] = info;
    outEnd[3 * r] = p.px, outEnd[3 * r + 1] = p.py, outEnd[3 * r + 2] = p.pz;
}

template <int L>
__global__ void sampleInteractionKernel(LutView lut, int kind, uint8_t material, float energy, uint64_t seed, uint64_t n, float* out)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i >= n)
        return;
    Rng rng;
    historyStream(seed, 0, i, rng.state, rng.inc);
    Photon p {};
    p.dz = 1.0f;
    p.energy = energy;
    p.weight = 1.0f;
    float imparted = 0.0f;
    if (kind == 0)
        imparted = photoAbsorption<L>(lut, p, material, rng);
    else if (kind == 1)
        imparted = comptonScatter<L>(lut, p, material, rng);
    else
        rayleighScatter<L>(lut, p, material, rng);
    float* o = out + 5 * i;
    o[0] = imparted;
    o[1] = p.energy;
    o[2] = p.dx;
    o[3] = p.dy;
    o[4] = p.dz;
}

} // namespace

// ================================================================================================
// runtime
// ================================================================================================
constexpr int kMaxPipes = 4;
// accumulator records beyond the grid (always zero): a reduce-scatter over up to 64 GPUs needs equal slices
constexpr uint64_t kAccPadding = 64;

struct dxmcb200_ctx {
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    std::string error;

    // world
    WorldView world {};
    uint64_t nVoxels = 0;
    uint2* dVoxels = nullptr; // 8-byte records, or null when the grid is in palette form
    uint8_t* dPalette = nullptr; // palette form: one byte per voxel ...
    uint2* dPaletteTable = nullptr; // ... into this 256-entry record table
    bool allowPalette = true;
    bool allowNibbles = true;
    unsigned paletteCount = 0; // distinct records of the grid = plain entries of the palette table
    std::vector<uint2> hPaletteBase; // host copy of the plain entries
    bool voxelsFlagged = false; // the grid carries air-brick flags (bit 16 of the records) of some brick grid
    uint8_t* dPaletteRemap = nullptr; // baseOf[256], flaggedOf[256] of the flagged palette
    // tracking: 1 = Woodcock + empty-space traversal through air bricks (default), 0 = the reference's Woodcock loop with the
    // global majorant everywhere (DXMCB200_TRACKING, dxmcb200_set_tracking)
    int tracking = 1;
    float brickMm = 16.0f; // target brick edge (DXMCB200_BRICK_MM)
    bool bricksValid = false;
    BrickView bricks {};
    uint32_t* dBrickBits = nullptr;
    uint8_t* dBrickDistance = nullptr;
    std::vector<float> hRatio, hBrickMax; // what the brick grid was built from, for dxmcb200_get_bricks
    std::vector<uint8_t> hAir, hDistance;
    float fAir = 0.0f;
    std::vector<float> hKnots, hCoeff, hMaxCoeff; // host copies of the attenuation fits (brick classification)
    // interaction / air-walk kernels are sized for all block slots of an SM rather than their pipeline's share: they are not persistent,
    // surplus blocks just queue, and the other pipeline's persistent kernels leave slots free at their tails (measured +3.6 %;
    // DXMCB200_INTERACT_FULL=0 for the former sizing)
    bool shardedFullOccupancy = true;
    int aggregateScores = -1; // warp-aggregated scoring: -1 automatic (narrow beams), 0 never, 1 always (DXMCB200_AGGREGATE)
    bool aggregateThisRun = false;
    unsigned long long* dAcc = nullptr;

    // luts
    LutView lut {};
    float* dLutBlob = nullptr;

    // beams
    BeamView beams {};
    void* dBeamBlob = nullptr;

    // exposures
    dxmcb200_exposure* dExposures = nullptr;
    uint64_t nExposuresResident = 0;
    std::vector<dxmcb200_exposure> hExposures; // host copy of the resident table (for prefix sums)
    uint64_t* dPrefix = nullptr;
    uint64_t prefixCapacity = 0;

    // Independent wave pipelines on their own streams (two by default; DXMCB200_PIPES=1..4 for experiments: three add
    // < 0.5 %, four leave one block per SM and kernel and lose 18 %): while one is in its (issue-bound) transport kernel
    // another runs its (latency-bound, divergent) interaction and generation kernels on the same SMs. Each owns two
    // photon buffers (ping-pong), an event buffer and its cursors.
    struct Pipe {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        cudaEvent_t mark[5] = { nullptr, nullptr, nullptr, nullptr, nullptr }; // around generate / transport / air walk / interact of the wave in flight
        bool generated = false;
        PhotonRecord* dPhotons[2] = { nullptr, nullptr };
        PhotonRecord* dAir = nullptr; // photons left in air bricks by the transport kernel (empty-space traversal)
        EventRecord* dEvents = nullptr;
        WaveCursors* dCursors = nullptr;
        WaveCursors* hCursors = nullptr; // pinned host copy for the per-wave read-back
        int cur = 0;
        unsigned survivors = 0; // slots of the next wave's buffer that hold records (a few of them dead markers)
        unsigned alive = 0; // photons among them
        bool pending = false;
    } pipes[kMaxPipes];
    int nPipes = 2;
    int persistentBlocks = 0; // blocks per SM of a persistent kernel; 0: occupancy / pipelines
    double kernelMs[4] = { 0, 0, 0, 0 }; // summed device time of generate / transport / air walk / interact launches since clear
    uint64_t kernelLaunches[4] = { 0, 0, 0, 0 };
    uint64_t photonRegion = 0, eventRegion = 0; // slots per shard region of the photon / event buffers
    Counters* dCounters = nullptr;

    int energyBits = 20, energySqBits = 10;
    bool collectStats = false;
    uint32_t waveRecords = 1u << 26; // photons per wave: 6.4 GB per photon buffer, 8.1 GB of event records, x2 pipelines (measured at 1e10
                                     // histories: 2^25: 2.38e9, 2^26: 2.44e9, 2^27: 2.45e9 histories/s; a third pipeline adds < 0.5 %)
    // empty lanes that make a warp stop stepping and re-fill. With the reference's tracking (32 steps per history) 4: 2.17e9, 8: 2.20e9,
    // 12: 2.18e9 histories/s; with the empty-space traversal a photon segment is 3 steps long and the service code weighs more:
    // 8: 2.59e9, 16: 2.66e9
    uint32_t refillBatch = 0; // 0: 16 with the empty-space traversal, 8 without

    double lastRunMs = 0, totalMs = 0;
    uint64_t launches = 0;
};

namespace {

#define CU_CHECK(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            (ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                     \
            return DXMCB200_ERR_CUDA;                                                              \
        }                                                                                          \
    } while (0)

// Host -> device copy that has LANDED when it returns. cudaMemcpy from pageable memory may return while the DMA from its staging
// buffer is still in flight, and only the legacy stream orders later work behind it; the context's streams are non-blocking, so a
// kernel launched right after such a copy could read the old contents. Copying on the context's stream and waiting for it is
// ordered and complete.
cudaError_t uploadNow(const dxmcb200_ctx* c, void* dst, const void* src, size_t bytes)
{
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
    return e != cudaSuccess ? e : cudaStreamSynchronize(c->stream);
}

int gridFor(const dxmcb200_ctx* c, uint64_t n)
{
    const uint64_t blocks = (n + 255) / 256;
    return static_cast<int>(std::min<uint64_t>(blocks, static_cast<uint64_t>(c->smCount) * 16));
}

template <typename T>
T* advancePtr(char*& cursor, size_t count)
{
    // 16-byte aligned carve-out from a blob
    size_t addr = reinterpret_cast<size_t>(cursor);
    addr = (addr + 15) & ~static_cast<size_t>(15);
    T* p = reinterpret_cast<T*>(addr);
    cursor = reinterpret_cast<char*>(addr + count * sizeof(T));
    return p;
}

// persistent grid: a whole number of resident CTAs per SM, never more lanes than work items
// with n pipelines every kernel takes 1/n of an SM's block slots so that kernels of all pipelines are co-resident
int blocksPerSmFor(const dxmcb200_ctx* c, int occupancy)
{
    if (c->persistentBlocks > 0) // experiment (DXMCB200_PERSIST_BLOCKS): a fixed share, whatever the number of pipelines
        return std::min(occupancy, c->persistentBlocks);
    return std::max(1, occupancy / std::max(1, c->nPipes));
}

template <typename K>
cudaError_t launchPersistent(const dxmcb200_ctx* c, cudaStream_t stream, K kernel, const KernelParams& P, uint64_t items)
{
    int blocksPerSm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, kernel, kThreads, 0);
    if (e != cudaSuccess)
        return e;
    uint64_t blocks = static_cast<uint64_t>(c->smCount) * blocksPerSmFor(c, blocksPerSm);
    blocks = std::max<uint64_t>(1, std::min(blocks, (items + kThreads - 1) / kThreads));
    kernel<<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(P);
    return cudaGetLastError();
}

// interactKernel assigns block b to event shard b % kShards, so its grid is a multiple of kShards
template <typename K>
cudaError_t launchSharded(const dxmcb200_ctx* c, cudaStream_t stream, K kernel, const KernelParams& P, uint64_t items)
{
    int blocksPerSm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, kernel, kThreads, 0);
    if (e != cudaSuccess)
        return e;
    uint64_t blocks = static_cast<uint64_t>(c->smCount) * (c->shardedFullOccupancy ? blocksPerSm : blocksPerSmFor(c, blocksPerSm));
    blocks = std::min(blocks, (items + kThreads - 1) / kThreads);
    blocks = std::max<uint64_t>(1, blocks / kShards) * kShards;
    kernel<<<static_cast<unsigned>(blocks), kThreads, 0, stream>>>(P);
    return cudaGetLastError();
}

template <int L>
cudaError_t launchInteract(const dxmcb200_ctx* c, cudaStream_t stream, const KernelParams& P, uint64_t items)
{
    if (c->aggregateThisRun)
        return c->collectStats ? launchSharded(c, stream, interactKernel<L, true, true>, P, items)
                               : launchSharded(c, stream, interactKernel<L, false, true>, P, items);
    return c->collectStats ? launchSharded(c, stream, interactKernel<L, true, false>, P, items)
                           : launchSharded(c, stream, interactKernel<L, false, false>, P, items);
}

template <bool kAir, bool kNibbleGrid>
cudaError_t launchTransportForm(const dxmcb200_ctx* c, cudaStream_t stream, const KernelParams& P, uint64_t items)
{
    return c->collectStats ? launchPersistent(c, stream, transportKernel<true, kAir, kNibbleGrid>, P, items)
                           : launchPersistent(c, stream, transportKernel<false, kAir, kNibbleGrid>, P, items);
}
cudaError_t launchTransport(const dxmcb200_ctx* c, cudaStream_t stream, const KernelParams& P, uint64_t items, bool air)
{
    const bool nibbleGrid = P.world.palette && P.world.paletteNibbles && P.world.exactInverse;
    if (air)
        return nibbleGrid ? launchTransportForm<true, true>(c, stream, P, items) : launchTransportForm<true, false>(c, stream, P, items);
    return nibbleGrid ? launchTransportForm<false, true>(c, stream, P, items) : launchTransportForm<false, false>(c, stream, P, items);
}

unsigned maxTransportBlocks(const dxmcb200_ctx* c)
{
    int occ[4] = { 0, 0, 0, 0 };
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], transportKernel<false, false, false>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], transportKernel<true, false, false>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], transportKernel<false, true, true>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[3], transportKernel<true, true, true>, kThreads, 0);
    return static_cast<unsigned>(c->smCount) * static_cast<unsigned>(std::max({ occ[0], occ[1], occ[2], occ[3], 1 }));
}

// Layout of the brick grid (shared rule with the CPU restatement, oracle/dxmc_oracle.cpp buildBricks): the brick edge per axis
// is the power of two (in voxels) closest to brickMm; while the grid has more than kMaxBricks bricks the axis with the
// shortest brick edge in mm is doubled (ties: z before y before x).
constexpr uint64_t kMaxBricks = static_cast<uint64_t>(kMaxBrickWords) * 32;
constexpr double kAirThreshold = 0.02;
inline float __uint_as_float_host(uint32_t bits)
{
    float f;
    std::memcpy(&f, &bits, sizeof f);
    return f;
}

void brickLayout(const uint64_t dim[3], const float spacing[3], float brickMm, uint32_t shift[3], uint32_t nb[3])
{
    for (int i = 0; i < 3; ++i) {
        const double k = std::round(std::log2(static_cast<double>(brickMm) / static_cast<double>(spacing[i])));
        shift[i] = static_cast<uint32_t>(std::clamp(k, 0.0, 10.0));
    }
    for (;;) {
        uint64_t n = 1;
        for (int i = 0; i < 3; ++i) {
            nb[i] = static_cast<uint32_t>((dim[i] + (1ull << shift[i]) - 1) >> shift[i]);
            n *= nb[i];
        }
        if (n <= kMaxBricks)
            break;
        int grow = 2;
        for (int i = 1; i >= 0; --i)
            if (static_cast<double>(1u << shift[i]) * spacing[i] < static_cast<double>(1u << shift[grow]) * spacing[grow])
                grow = i;
        ++shift[grow];
    }
}

// Write the air-brick flags of brick grid `b` into the voxel records (b == nullptr: clear them); see flagRecordsKernel.
// `ratio` as in ensureBricks. A palette grid whose flagged entries no longer fit its form is widened: 4-bit indices to bytes,
// bytes to 8-byte records.
int applyAirFlags(dxmcb200_ctx* c, const BrickView* b, const std::vector<float>& ratio)
{
    const uint64_t n = c->nVoxels;
    const BrickView none {};
    if (c->voxelsFlagged) { // flags of a former brick grid: every voxel back to its plain record
        if (c->dPalette)
            flagPaletteKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->world, none, nullptr, n, c->dPaletteRemap, c->dPalette, static_cast<int>(c->world.paletteNibbles));
        else
            flagRecordsKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->world, none, nullptr, n, c->dVoxels);
        CU_CHECK(c, cudaGetLastError());
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
        c->voxelsFlagged = false;
    }
    if (!b)
        return DXMCB200_OK;
    if (c->dPalette) {
        // plain entries that can occur in an air brick get a flagged twin behind the plain ones
        std::vector<uint8_t> remap(512, 0);
        std::vector<uint2> table(c->hPaletteBase);
        unsigned total = c->paletteCount;
        for (unsigned i = 0; i < c->paletteCount; ++i) {
            remap[i] = static_cast<uint8_t>(i);
            remap[256 + i] = static_cast<uint8_t>(i);
            const uint2 r = table[i];
            const float f = __uint_as_float_host(r.x) * ratio[r.y & 0xffu];
            const bool airLike = !(r.y & 0xff00u) && !(1.001 * static_cast<double>(f > 0.0f ? f : 0.0f) > kAirThreshold);
            if (!airLike)
                continue;
            if (total < 256) {
                table[total] = make_uint2(r.x, r.y | kInAirBrick);
                remap[total] = static_cast<uint8_t>(i);
                remap[256 + i] = static_cast<uint8_t>(total);
            }
            ++total;
        }
        if (total > 256) { // no room for the twins: 8-byte records from here on
            CU_CHECK(c, hostio::poolAlloc(c->device, &c->dVoxels, n * sizeof(uint2)));
            if (c->world.paletteNibbles) {
                uint8_t* bytes = nullptr;
                CU_CHECK(c, hostio::poolAlloc(c->device, &bytes, n));
                expandNibblesKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->dPalette, n, bytes);
                expandPaletteKernel<<<gridFor(c, n), 256, 0, c->stream>>>(bytes, c->dPaletteTable, n, c->dVoxels);
                CU_CHECK(c, cudaStreamSynchronize(c->stream));
                hostio::poolFree(bytes);
            } else {
                expandPaletteKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->dPalette, c->dPaletteTable, n, c->dVoxels);
                CU_CHECK(c, cudaStreamSynchronize(c->stream));
            }
            hostio::poolFree(c->dPalette);
            cudaFree(c->dPaletteTable);
            c->dPalette = nullptr;
            c->dPaletteTable = nullptr;
            c->world.palette = nullptr;
            c->world.paletteTable = nullptr;
            c->world.paletteNibbles = 0;
            c->world.voxels = c->dVoxels;
        } else {
            if (c->world.paletteNibbles && total > 16) { // 4-bit indices no longer do
                uint8_t* bytes = nullptr;
                CU_CHECK(c, hostio::poolAlloc(c->device, &bytes, n));
                expandNibblesKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->dPalette, n, bytes);
                CU_CHECK(c, cudaStreamSynchronize(c->stream));
                hostio::poolFree(c->dPalette);
                c->dPalette = bytes;
                c->world.palette = bytes;
                c->world.paletteNibbles = 0;
            }
            if (!c->dPaletteRemap)
                CU_CHECK(c, cudaMalloc(&c->dPaletteRemap, 512));
            CU_CHECK(c, cudaMemcpyAsync(c->dPaletteRemap, remap.data(), 512, cudaMemcpyHostToDevice, c->stream));
            CU_CHECK(c, cudaMemcpyAsync(c->dPaletteTable, table.data(), 256 * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
            flagPaletteKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->world, *b, b->air, n, c->dPaletteRemap, c->dPalette, static_cast<int>(c->world.paletteNibbles));
            CU_CHECK(c, cudaGetLastError());
            CU_CHECK(c, cudaStreamSynchronize(c->stream)); // `remap` and `table` live on this stack frame
            c->voxelsFlagged = true;
            return DXMCB200_OK;
        }
    }
    flagRecordsKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->world, *b, b->air, n, c->dVoxels);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    c->voxelsFlagged = true;
    return DXMCB200_OK;
}

// Classify the bricks of the uploaded grid against the uploaded attenuation fits (once per world / table pair).
int ensureBricks(dxmcb200_ctx* c)
{
    if (c->bricksValid)
        return DXMCB200_OK;
    c->bricks = BrickView {};
    c->fAir = 0.0f;
    c->hRatio.clear();
    c->hBrickMax.clear();
    c->hAir.clear();
    c->hDistance.clear();
    if (c->tracking == 0) {
        const int st = applyAirFlags(c, nullptr, {});
        if (st != DXMCB200_OK)
            return st;
        c->bricksValid = true;
        return DXMCB200_OK;
    }
    const uint32_t nMat = c->lut.nMaterials, nSeg = c->lut.nSegments;
    // per material: the largest mu_total(E) * majorantInverse(E) over the table. Inside a segment it is a sum of exponentials
    // in log10 E (convex), so the maximum sits at a segment end; evaluated in double from the float fits.
    std::vector<float> ratio(256, 0.0f);
    for (uint32_t m = 0; m < nMat; ++m) {
        double best = 0;
        for (uint32_t k = 0; k < nSeg; ++k) {
            const double ends[2] = { k == 0 ? 0.0 : static_cast<double>(c->hKnots[k - 1]), static_cast<double>(c->hKnots[k]) };
            const float* fit = c->hCoeff.data() + (static_cast<size_t>(m) * nSeg + k) * 6;
            for (const double x : ends) {
                double total = 0;
                for (int i = 0; i < 3; ++i)
                    total += std::pow(10.0, static_cast<double>(fit[2 * i]) + static_cast<double>(fit[2 * i + 1]) * x);
                best = std::max(best, total * std::pow(10.0, static_cast<double>(c->hMaxCoeff[2 * k]) + static_cast<double>(c->hMaxCoeff[2 * k + 1]) * x));
            }
        }
        ratio[m] = static_cast<float>(best);
    }
    BrickView b {};
    const uint64_t dim[3] = { c->world.dim[0], c->world.dim[1], c->world.dim[2] };
    brickLayout(dim, c->world.spacing, c->brickMm, b.shift, b.nb);
    for (int i = 0; i < 3; ++i) {
        b.size[i] = static_cast<float>(1u << b.shift[i]) * c->world.spacing[i];
        b.invSize[i] = 1.0f / b.size[i];
    }
    const size_t nBricks = static_cast<size_t>(b.nb[0]) * b.nb[1] * b.nb[2];
    float* dRatio = nullptr;
    unsigned* dMax = nullptr; // [nBricks] maxima, then [nBricks] measurement flags
    CU_CHECK(c, cudaMalloc(&dRatio, 256 * sizeof(float)));
    cudaError_t e = cudaMalloc(&dMax, 2 * nBricks * sizeof(unsigned));
    std::vector<unsigned> host(2 * nBricks, 0u);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dRatio, ratio.data(), 256 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(dMax, 0, 2 * nBricks * sizeof(unsigned), c->stream);
    if (e == cudaSuccess) {
        brickMaxKernel<<<gridFor(c, c->nVoxels), 256, 0, c->stream>>>(c->world, b, dRatio, c->nVoxels, dMax, dMax + nBricks);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(host.data(), dMax, host.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(dRatio);
    cudaFree(dMax);
    CU_CHECK(c, e);
    c->hRatio.assign(ratio.begin(), ratio.begin() + nMat);
    c->hBrickMax.resize(nBricks);
    std::memcpy(c->hBrickMax.data(), host.data(), nBricks * sizeof(float));
    c->hAir.assign(nBricks, 0);
    std::vector<uint32_t> bitmap((nBricks + 31) / 32, 0u);
    double fAir = 0;
    size_t nAir = 0;
    for (size_t k = 0; k < nBricks; ++k) {
        const double f = 1.001 * static_cast<double>(c->hBrickMax[k]);
        if (f <= kAirThreshold && !host[nBricks + k]) {
            c->hAir[k] = 1;
            bitmap[k >> 5] |= 1u << (k & 31);
            fAir = std::max(fAir, f);
            ++nAir;
        }
    }
    // Per octant of travel directions o = (dx<0) | (dy<0)<<1 | (dz<0)<<2 and per brick b: the edge k (bricks, at most 255) of the
    // largest cube of air bricks that has b as its corner and opens in that octant's directions (bricks beyond the grid count as
    // air); 0 for non-air bricks. D(b) = 1 + min over the 7 neighbours one brick further along the octant's directions, swept
    // from the far corner of the octant backwards (same rule as oracle/dxmc_oracle.cpp buildBricks).
    c->hDistance.assign(8 * nBricks, 0);
    {
        const int64_t n0 = b.nb[0], n1 = b.nb[1], n2 = b.nb[2];
        for (int o = 0; o < 8; ++o) {
            const int64_t s0 = (o & 1) ? -1 : 1, s1 = (o & 2) ? -1 : 1, s2 = (o & 4) ? -1 : 1;
            uint8_t* D = c->hDistance.data() + static_cast<size_t>(o) * nBricks;
            for (int64_t kz = 0; kz < n2; ++kz)
                for (int64_t ky = 0; ky < n1; ++ky)
                    for (int64_t kx = 0; kx < n0; ++kx) {
                        const int64_t x = s0 > 0 ? n0 - 1 - kx : kx, y = s1 > 0 ? n1 - 1 - ky : ky, z = s2 > 0 ? n2 - 1 - kz : kz;
                        const size_t at = static_cast<size_t>((z * n1 + y) * n0 + x);
                        if (!c->hAir[at])
                            continue;
                        int least = 255;
                        for (int m = 1; m < 8; ++m) {
                            const int64_t X = x + ((m & 1) ? s0 : 0), Y = y + ((m & 2) ? s1 : 0), Z = z + ((m & 4) ? s2 : 0);
                            if (X < 0 || Y < 0 || Z < 0 || X >= n0 || Y >= n1 || Z >= n2)
                                continue;
                            least = std::min<int>(least, D[static_cast<size_t>((Z * n1 + Y) * n0 + X)]);
                        }
                        D[at] = static_cast<uint8_t>(std::min(least + 1, 255));
                    }
        }
    }
    if (nAir > 0) {
        c->fAir = static_cast<float>(std::max(fAir, 1.0e-6));
        b.invFAir = 1.0f / c->fAir;
        b.nWords = static_cast<uint32_t>(bitmap.size());
    cudaFree(c->dBrickBits);
        cudaFree(c->dBrickDistance);
        c->dBrickBits = nullptr;
        c->dBrickDistance = nullptr;
        CU_CHECK(c, cudaMalloc(&c->dBrickBits, bitmap.size() * sizeof(uint32_t)));
        CU_CHECK(c, uploadNow(c, c->dBrickBits, bitmap.data(), bitmap.size() * sizeof(uint32_t)));
        CU_CHECK(c, cudaMalloc(&c->dBrickDistance, 8 * nBricks));
        CU_CHECK(c, uploadNow(c, c->dBrickDistance, c->hDistance.data(), 8 * nBricks));
        b.air = c->dBrickBits;
        b.distance = c->dBrickDistance;
    }
    c->bricks = b; // nWords == 0: no air bricks, the kernels without the traversal run
    {
        const int st = applyAirFlags(c, nAir > 0 ? &b : nullptr, ratio);
        if (st != DXMCB200_OK)
            return st;
    }
    c->bricksValid = true;
    return DXMCB200_OK;
}

// transports exposures expBegin + k * stride, k in [0, nExp)
int runRange(dxmcb200_ctx* c, const dxmcb200_exposure* hostExposures, const dxmcb200_exposure* devExposures, uint64_t expBegin,
    uint64_t nExp, uint64_t stride, int model, uint64_t seed, const volatile int* cancel, dxmcb200_progress_cb cb, void* user)
{
    if ((!c->dVoxels && !c->dPalette) || !c->dLutBlob)
        return DXMCB200_ERR_STATE;
    if (model < 0 || model > 2 || stride == 0)
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    c->lastRunMs = 0;
    if (nExp == 0)
        return DXMCB200_OK;
    const char* traceEnv = std::getenv("DXMCB200_TRACE");
    const bool trace = traceEnv && traceEnv[0] == '1';
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace)
            return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dxmcb200]     run: %-24s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t0).count());
        t0 = now;
    };
    {
        const int st = ensureBricks(c);
        if (st != DXMCB200_OK)
            return st;
    }
    lap("brick grid + air flags");
    const bool air = c->bricks.nWords > 0; // empty-space traversal on and the grid has air bricks
    if (nExp >= (1ULL << 32)) {
        c->error = "more than 2^32 exposures in one run";
        return DXMCB200_ERR_ARG;
    }

    // cumulative histories of the range; histories are numbered 0..total-1 across it
    std::vector<uint64_t> prefix(nExp + 1, 0);
    bool uniform = hostExposures[expBegin].histories > 0 && hostExposures[expBegin].histories < (1ULL << 31);
    for (uint64_t e = 0; e < nExp; ++e) {
        prefix[e + 1] = prefix[e] + hostExposures[expBegin + e * stride].histories;
        uniform = uniform && hostExposures[expBegin + e * stride].histories == hostExposures[expBegin].histories;
    }
    // Warp-aggregated scoring pays when the lanes of a warp deposit in the same few voxels, i.e. for pencil-like beams:
    // every exposure of the run collimated to less than 0.02 rad (about 1 degree) in both directions.
    bool narrow = true;
    for (uint64_t e = 0; e < nExp && narrow; ++e) {
        const float* col = hostExposures[expBegin + e * stride].collimation;
        narrow = std::fabs(col[1] - col[0]) < 0.02f && std::fabs(col[3] - col[2]) < 0.02f;
    }
    c->aggregateThisRun = c->aggregateScores < 0 ? narrow : c->aggregateScores != 0;
    const uint64_t total = prefix.back();
    if (total == 0) {
        if (cb)
            cb(nExp, user);
        return DXMCB200_OK;
    }
    if (prefix.size() > c->prefixCapacity) {
        if (c->dPrefix)
            cudaFree(c->dPrefix);
        c->dPrefix = nullptr;
        c->prefixCapacity = std::max<uint64_t>(prefix.size(), 4096);
        CU_CHECK(c, cudaMalloc(&c->dPrefix, c->prefixCapacity * sizeof(uint64_t)));
    }
    CU_CHECK(c, cudaMemcpyAsync(c->dPrefix, prefix.data(), prefix.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));

    // wave buffers sized for this run (kept for the next one when large enough): kShards regions each, with 50 %
    // head-room over an even split (shards fill evenly: every warp appends to shard `global warp index % kShards`)
    const int nPipes = total >= 4 * static_cast<uint64_t>(c->waveRecords) ? c->nPipes : 1; // small runs: one pipeline
    const uint64_t wave = std::max<uint64_t>(std::min<uint64_t>(c->waveRecords, total), 1024);
    const uint64_t warpsPerShard = (static_cast<uint64_t>(maxTransportBlocks(c)) * (kThreads / 32) + kShards - 1) / kShards;
    // + the tiles interactKernel's warps hold open (two per warp, at most 8 blocks of 8 warps per SM)
    const uint64_t tileSlack = (static_cast<uint64_t>(c->smCount) * 64 / kShards + 1) * 2 * kSurvivorTile;
    const uint64_t photonRegion = ((wave + wave / 2) / kShards + 1024 + tileSlack + 15) & ~15ULL;
    const uint64_t eventRegion = (photonRegion + 2 * warpsPerShard * kEventTile + kEventTile - 1) / kEventTile * kEventTile; // a warp holds up to two tiles
    if (photonRegion > c->photonRegion || eventRegion > c->eventRegion) {
        for (auto& pipe : c->pipes) {
            hostio::poolFree(pipe.dPhotons[0]);
            hostio::poolFree(pipe.dPhotons[1]);
            hostio::poolFree(pipe.dEvents);
            hostio::poolFree(pipe.dAir);
            pipe.dPhotons[0] = pipe.dPhotons[1] = pipe.dAir = nullptr;
            pipe.dEvents = nullptr;
        }
        c->photonRegion = std::max(photonRegion, c->photonRegion);
        c->eventRegion = std::max(eventRegion, c->eventRegion);
    }
    for (int i = 0; i < nPipes; ++i) {
        auto& pipe = c->pipes[i];
        if (!pipe.dEvents) {
            CU_CHECK(c, hostio::poolAlloc(c->device, &pipe.dPhotons[0], c->photonRegion * kShards * sizeof(PhotonRecord)));
            CU_CHECK(c, hostio::poolAlloc(c->device, &pipe.dPhotons[1], c->photonRegion * kShards * sizeof(PhotonRecord)));
            CU_CHECK(c, hostio::poolAlloc(c->device, &pipe.dEvents, c->eventRegion * kShards * sizeof(EventRecord)));
        }
        if (air && !pipe.dAir)
            CU_CHECK(c, hostio::poolAlloc(c->device, &pipe.dAir, c->photonRegion * kShards * sizeof(PhotonRecord)));
    }

    KernelParams base {};
    base.world = c->world;
    base.lut = c->lut;
    base.beams = c->beams;
    base.bricks = c->bricks;
    base.exposures = devExposures;
    base.prefix = c->dPrefix;
    base.expBegin = expBegin;
    base.expStride = stride;
    base.nExp = static_cast<uint32_t>(nExp);
    base.uniformHistories = uniform ? static_cast<uint32_t>(hostExposures[expBegin].histories) : 0u;
    base.refillBatch = c->refillBatch ? c->refillBatch : (air ? 16u : 8u);
    base.seed = seed;
    base.photonRegion = static_cast<uint32_t>(c->photonRegion);
    base.eventRegion = static_cast<uint32_t>(c->eventRegion);
    base.acc = c->dAcc;
    base.counters = c->dCounters;
    base.energyScale = std::ldexp(1.0f, c->energyBits);
    base.energySqScale = std::ldexp(1.0f, c->energySqBits);

    uint64_t issuedHistories = 0; // births generated so far
    uint64_t expDone = 0; // exposures of the range whose histories have all been issued, for the progress callback
    const int savedPipes = c->nPipes;
    c->nPipes = nPipes; // grid sizing of the launch helpers
    struct Restore {
        dxmcb200_ctx* c;
        int n;
        ~Restore() { c->nPipes = n; }
    } restore { c, savedPipes };

    // one wave of pipeline `pipe`: top up with births, transport, interact, read the survivor count back
    auto enqueueWave = [&](dxmcb200_ctx::Pipe& pipe) -> int {
        KernelParams P = base;
        const int cur = pipe.cur, nxt = cur ^ 1;
        P.events = pipe.dEvents;
        P.eventCursors = pipe.dCursors->events;
        P.airborne = pipe.dAir;
        P.airCursors = pipe.dCursors->air;
        P.overflow = &pipe.dCursors->overflow.stored;
        // every kernel of the wave may append events (air walks included), so the event cursors are reset first
        resetCursorsKernel<<<1, kShards, 0, pipe.stream>>>(pipe.dCursors->events, 1);
        if (air)
            resetCursorsKernel<<<1, kShards, 0, pipe.stream>>>(pipe.dCursors->air, 1);
        // (a) births
        const uint64_t births = std::min<uint64_t>(wave - std::min<uint64_t>(pipe.alive, wave), total - issuedHistories);
        P.photonsOut = pipe.dPhotons[cur];
        P.outCursors = pipe.dCursors->photons[cur];
        CU_CHECK(c, cudaEventRecord(pipe.mark[0], pipe.stream));
        pipe.generated = births > 0;
        if (births > 0) {
            P.chunkBegin = issuedHistories;
            P.chunkCount = static_cast<uint32_t>(births);
            if (uniform) {
                P.chunkFirstExposure = static_cast<uint32_t>(issuedHistories / P.uniformHistories);
                P.chunkFirstOffset = static_cast<uint32_t>(issuedHistories % P.uniformHistories);
            }
            if (air)
                CU_CHECK(c, c->collectStats ? launchPersistent(c, pipe.stream, generateKernel<true, true>, P, births)
                                            : launchPersistent(c, pipe.stream, generateKernel<false, true>, P, births));
            else
                CU_CHECK(c, c->collectStats ? launchPersistent(c, pipe.stream, generateKernel<true, false>, P, births)
                                            : launchPersistent(c, pipe.stream, generateKernel<false, false>, P, births));
            issuedHistories += births;
            ++c->launches;
        }
        // (b) step every photon of the wave to its next event
        const uint64_t records = pipe.survivors + births; // upper bound: births that miss the world store nothing
        P.photonsIn = pipe.dPhotons[cur];
        P.inCursors = pipe.dCursors->photons[cur];
        resetCursorsKernel<<<1, kShards, 0, pipe.stream>>>(pipe.dCursors->photons[cur], 0);
        CU_CHECK(c, cudaEventRecord(pipe.mark[1], pipe.stream));
        CU_CHECK(c, launchTransport(c, pipe.stream, P, records, air));
        CU_CHECK(c, cudaEventRecord(pipe.mark[2], pipe.stream));
        // (b') photons left in air bricks take the air walk; (c)+(d) interactions and scoring; the survivors of both open the next wave
        P.photonsOut = pipe.dPhotons[nxt];
        P.outCursors = pipe.dCursors->photons[nxt];
        resetCursorsKernel<<<1, kShards, 0, pipe.stream>>>(pipe.dCursors->photons[nxt], 1);
        if (air) {
            CU_CHECK(c, c->collectStats ? launchSharded(c, pipe.stream, airWalkKernel<true>, P, records)
                                        : launchSharded(c, pipe.stream, airWalkKernel<false>, P, records));
            ++c->launches;
        }
        CU_CHECK(c, cudaEventRecord(pipe.mark[3], pipe.stream));
        CU_CHECK(c, model == 0 ? launchInteract<0>(c, pipe.stream, P, records)
                               : model == 1 ? launchInteract<1>(c, pipe.stream, P, records) : launchInteract<2>(c, pipe.stream, P, records));
        CU_CHECK(c, cudaEventRecord(pipe.mark[4], pipe.stream));
        c->launches += 2;
        CU_CHECK(c, cudaMemcpyAsync(pipe.hCursors->photons[nxt], pipe.dCursors->photons[nxt], sizeof(ShardCursor) * kShards, cudaMemcpyDeviceToHost,
                        pipe.stream));
        CU_CHECK(c, cudaMemcpyAsync(&pipe.hCursors->overflow, &pipe.dCursors->overflow, sizeof(ShardCursor), cudaMemcpyDeviceToHost, pipe.stream));
        CU_CHECK(c, cudaEventRecord(pipe.done, pipe.stream));
        pipe.pending = true;
        return DXMCB200_OK;
    };
    // wait for the wave in flight on `pipe` and pick up its survivor count
    auto completeWave = [&](dxmcb200_ctx::Pipe& pipe) -> int {
        CU_CHECK(c, cudaEventSynchronize(pipe.done));
        pipe.pending = false;
        const int nxt = pipe.cur ^ 1;
        for (int k = 0; k < 4; ++k) { // device time of the wave's kernels (mark[1] is recorded after the tiny cursor resets)
            float ms = 0;
            if ((k > 0 || pipe.generated) && (k != 2 || air) && cudaEventElapsedTime(&ms, pipe.mark[k], pipe.mark[k + 1]) == cudaSuccess) {
                c->kernelMs[k] += ms;
                ++c->kernelLaunches[k];
            }
        }
        if (pipe.hCursors->overflow.stored) {
            c->error = "wave buffer region overflow";
            return DXMCB200_ERR_STATE;
        }
        uint64_t slots = 0, alive = 0;
        for (unsigned k = 0; k < kShards; ++k) {
            slots += pipe.hCursors->photons[nxt][k].stored;
            alive += pipe.hCursors->photons[nxt][k].live;
        }
        if (alive > wave) {
            c->error = "wave buffer overflow";
            return DXMCB200_ERR_STATE;
        }
        pipe.survivors = static_cast<unsigned>(slots);
        pipe.alive = static_cast<unsigned>(alive);
        pipe.cur = nxt;
        return DXMCB200_OK;
    };

    CU_CHECK(c, cudaStreamSynchronize(c->stream)); // uploads issued on the ctx stream are visible to both pipelines
    lap("wave buffers, prefix");
    for (int i = 0; i < nPipes; ++i) {
        auto& pipe = c->pipes[i];
        pipe.cur = 0;
        pipe.survivors = 0;
        pipe.alive = 0;
        pipe.pending = false;
        CU_CHECK(c, cudaMemsetAsync(pipe.dCursors, 0, sizeof(WaveCursors), pipe.stream));
    }
    CU_CHECK(c, cudaEventRecord(c->evStart, c->pipes[0].stream));
    int status = DXMCB200_OK;
    int turn = 0; // pipelines are served round robin, so the host always waits on the older wave
    for (;;) {
        auto& pipe = c->pipes[turn];
        if (pipe.pending) {
            status = completeWave(pipe);
            if (status != DXMCB200_OK)
                break;
            if (cb) {
                while (expDone + 1 < nExp && prefix[expDone + 1] <= issuedHistories)
                    ++expDone; // the last exposure is reported when everything has drained
                cb(expDone, user); // every wave, also without news: the callee polls its cancel request here
            }
        }
        if (cancel && *cancel) {
            status = DXMCB200_ERR_CANCELLED;
            break;
        }
        if (issuedHistories < total || pipe.alive > 0) { // dead markers alone (survivors > 0, alive == 0) are not worth a wave
            status = enqueueWave(pipe);
            if (status != DXMCB200_OK)
                break;
        }
        bool any = false;
        for (int i = 0; i < nPipes; ++i)
            any = any || c->pipes[i].pending;
        if (!any)
            break;
        turn = (turn + 1) % nPipes;
    }
    for (int i = 0; i < nPipes; ++i)
        cudaStreamSynchronize(c->pipes[i].stream);
    if (status != DXMCB200_OK)
        return status;
    CU_CHECK(c, cudaEventRecord(c->evStop, c->pipes[0].stream));
    CU_CHECK(c, cudaStreamSynchronize(c->pipes[0].stream));
    float ms = 0;
    CU_CHECK(c, cudaEventElapsedTime(&ms, c->evStart, c->evStop));
    lap("waves (host wall)");
    if (trace)
        std::fprintf(stderr, "[dxmcb200]     run: %-24s %8.1f ms\n", "waves (device events)", static_cast<double>(ms));
    c->lastRunMs = ms;
    c->totalMs += ms;
    if (cb)
        cb(nExp, user);
    return DXMCB200_OK;
}

} // namespace

extern "C" {

int dxmcb200_device_count(int* count)
{
    if (!count)
        return DXMCB200_ERR_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        *count = 0;
        return DXMCB200_ERR_NO_DEVICE;
    }
    *count = n;
    return DXMCB200_OK;
}

int dxmcb200_create(int device, dxmcb200_ctx** out)
{
    if (!out)
        return DXMCB200_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
        return DXMCB200_ERR_NO_DEVICE; // there is deliberately no CPU fallback
    auto* c = new (std::nothrow) dxmcb200_ctx;
    if (!c)
        return DXMCB200_ERR_STATE;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetAttribute(&c->smCount, cudaDevAttrMultiProcessorCount, device) != cudaSuccess
        || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->evStart) != cudaSuccess
        || cudaEventCreate(&c->evStop) != cudaSuccess
        || cudaMalloc(&c->dCounters, sizeof(Counters)) != cudaSuccess || cudaMemsetAsync(c->dCounters, 0, sizeof(Counters), c->stream) != cudaSuccess) {
        dxmcb200_destroy(c);
        return DXMCB200_ERR_CUDA;
    }
    c->pipes[0].stream = c->stream;
    if (const char* env = std::getenv("DXMCB200_PIPES"))
        c->nPipes = std::clamp(std::atoi(env), 1, kMaxPipes);
    for (int i = 0; i < c->nPipes; ++i) { // only the pipelines that will run get streams, events and (pinned) cursor blocks
        auto& pipe = c->pipes[i];
        if ((i > 0 && cudaStreamCreateWithFlags(&pipe.stream, cudaStreamNonBlocking) != cudaSuccess)
            || cudaEventCreateWithFlags(&pipe.done, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&pipe.mark[0]) != cudaSuccess
            || cudaEventCreate(&pipe.mark[1]) != cudaSuccess || cudaEventCreate(&pipe.mark[2]) != cudaSuccess || cudaEventCreate(&pipe.mark[3]) != cudaSuccess
            || cudaEventCreate(&pipe.mark[4]) != cudaSuccess
            || cudaMalloc(&pipe.dCursors, sizeof(WaveCursors)) != cudaSuccess || cudaMallocHost(&pipe.hCursors, sizeof(WaveCursors)) != cudaSuccess) {
            dxmcb200_destroy(c);
            return DXMCB200_ERR_CUDA;
        }
    }
    const char* stats = std::getenv("DXMCB200_STATS");
    c->collectStats = stats && stats[0] == '1';
    if (const char* env = std::getenv("DXMCB200_PALETTE")) { // 0: 8-byte records, 8: byte indices only, else automatic
        c->allowPalette = env[0] != '0';
        c->allowNibbles = env[0] != '8';
    }
    if (const char* env = std::getenv("DXMCB200_AGGREGATE"))
        c->aggregateScores = std::clamp(std::atoi(env), -1, 1);
    if (const char* env = std::getenv("DXMCB200_PERSIST_BLOCKS"))
        c->persistentBlocks = std::clamp(std::atoi(env), 0, 8);
    if (const char* env = std::getenv("DXMCB200_INTERACT_FULL"))
        c->shardedFullOccupancy = env[0] == '1';
    if (const char* env = std::getenv("DXMCB200_TRACKING")) // 0: the reference's Woodcock loop everywhere, 1: + empty-space traversal
        c->tracking = std::clamp(std::atoi(env), 0, 1);
    if (const char* env = std::getenv("DXMCB200_BRICK_MM")) {
        const double mm = std::atof(env);
        if (mm > 0)
            c->brickMm = static_cast<float>(mm);
    }
    if (const char* env = std::getenv("DXMCB200_BATCH")) { // experiments: <refill batch>[,<log2 wave records>]
        int r = 8, lg = 26;
        std::sscanf(env, "%d,%d", &r, &lg);
        c->refillBatch = static_cast<uint32_t>(std::clamp(r, 1, 32));
        c->waveRecords = 1u << std::clamp(lg, 10, 28);
    }
    *out = c;
    return DXMCB200_OK;
}

void dxmcb200_destroy(dxmcb200_ctx* c)
{
    if (!c)
        return;
    const char* traceEnv = std::getenv("DXMCB200_TRACE");
    const bool trace = traceEnv && traceEnv[0] == '1';
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace)
            return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dxmcb200]     destroy: %-18s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t0).count());
        t0 = now;
    };
    cudaSetDevice(c->device);
    // nothing may still be running on the blocks that go back to the pool
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    for (int i = 1; i < kMaxPipes; ++i)
        if (c->pipes[i].stream)
            cudaStreamSynchronize(c->pipes[i].stream);
    lap("stream sync");
    hostio::poolFree(c->dVoxels);
    hostio::poolFree(c->dPalette);
    cudaFree(c->dPaletteTable);
    hostio::poolFree(c->dAcc);
    cudaFree(c->dLutBlob);
    cudaFree(c->dBeamBlob);
    cudaFree(c->dExposures);
    cudaFree(c->dPrefix);
    cudaFree(c->dPaletteRemap);
    cudaFree(c->dBrickBits);
    cudaFree(c->dBrickDistance);
    lap("world, tables");
    for (int i = 0; i < kMaxPipes; ++i) {
        auto& pipe = c->pipes[i];
        cudaFree(pipe.dCursors);
        cudaFreeHost(pipe.hCursors);
        lap("cursors");
        hostio::poolFree(pipe.dPhotons[0]);
        hostio::poolFree(pipe.dPhotons[1]);
        hostio::poolFree(pipe.dEvents);
        hostio::poolFree(pipe.dAir);
        if (pipe.done)
            cudaEventDestroy(pipe.done);
        for (auto& m : pipe.mark)
            if (m)
                cudaEventDestroy(m);
        if (i > 0 && pipe.stream)
            cudaStreamDestroy(pipe.stream); // pipes[0] runs on the ctx stream
        lap("wave buffers");
    }
    cudaFree(c->dCounters);
    if (c->evStart)
        cudaEventDestroy(c->evStart);
    if (c->evStop)
        cudaEventDestroy(c->evStop);
    if (c->stream)
        cudaStreamDestroy(c->stream);
    delete c;
    lap("events, streams");
}

const char* dxmcb200_last_error(dxmcb200_ctx* c) { return c ? c->error.c_str() : "null context"; }

int dxmcb200_trim_pool(int device)
{
    hostio::Pool::instance().trim(device);
    return DXMCB200_OK;
}

int dxmcb200_set_pool_limit(uint64_t bytes)
{
    hostio::poolLimit().store(static_cast<size_t>(bytes));
    if (bytes == 0)
        hostio::Pool::instance().trim(-1);
    return DXMCB200_OK;
}

int dxmcb200_material_max_density(dxmcb200_ctx* c, uint32_t nMaterials, float* out)
{
    if (!c || !out || nMaterials == 0 || nMaterials > 256 || (!c->dVoxels && !c->dPalette))
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    unsigned* dMax = nullptr;
    CU_CHECK(c, cudaMalloc(&dMax, 256 * sizeof(unsigned)));
    CU_CHECK(c, cudaMemsetAsync(dMax, 0, 256 * sizeof(unsigned), c->stream));
    if (c->dPalette) // every distinct {density, material} record of the grid is in the 256-entry table already
        maxDensityKernel<<<1, 256, 0, c->stream>>>(c->dPaletteTable, 256, dMax);
    else
        maxDensityKernel<<<gridFor(c, c->nVoxels), 256, 0, c->stream>>>(c->dVoxels, c->nVoxels, dMax);
    unsigned bits[256];
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(bits, dMax, sizeof(bits), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(dMax);
    CU_CHECK(c, e);
    for (uint32_t m = 0; m < nMaterials; ++m)
        std::memcpy(out + m, bits + m, sizeof(float));
    return DXMCB200_OK;
}

int dxmcb200_set_world(dxmcb200_ctx* c, const dxmcb200_world* w)
{
    if (!c || !w || !w->density || !w->material)
        return DXMCB200_ERR_ARG;
    const uint64_t n = w->dim[0] * w->dim[1] * w->dim[2];
    if (n == 0 || n >= (1ULL << 32)) {
        c->error = "voxel count must be in [1, 2^32)";
        return DXMCB200_ERR_ARG;
    }
    CU_CHECK(c, cudaSetDevice(c->device));
    if (n != c->nVoxels) {
        hostio::poolFree(c->dAcc);
        c->dAcc = nullptr;
        c->nVoxels = 0;
        CU_CHECK(c, hostio::poolAlloc(c->device, &c->dAcc, (n + kAccPadding) * 4 * sizeof(unsigned long long)));
        c->nVoxels = n;
    }
    hostio::poolFree(c->dVoxels);
    hostio::poolFree(c->dPalette);
    cudaFree(c->dPaletteTable);
    c->dVoxels = nullptr;
    c->dPalette = nullptr;
    c->dPaletteTable = nullptr;
    // the reference's three arrays (world.hpp:48-50), staged on the device only until they are packed
    struct Staged {
        float* density = nullptr;
        uint8_t* material = nullptr;
        uint8_t* measurement = nullptr;
        ~Staged()
        {
            hostio::poolFree(density);
            hostio::poolFree(material);
            hostio::poolFree(measurement);
        }
    } staged;
    CU_CHECK(c, hostio::poolAlloc(c->device, &staged.density, n * sizeof(float)));
    CU_CHECK(c, hostio::poolAlloc(c->device, &staged.material, n));
    if (w->measurement)
        CU_CHECK(c, hostio::poolAlloc(c->device, &staged.measurement, n));
    float* const dDensity = staged.density;
    uint8_t* const dMat = staged.material;
    uint8_t* const dMeas = staged.measurement;
    CU_CHECK(c, cudaMemsetAsync(c->dAcc, 0, (n + kAccPadding) * 4 * sizeof(unsigned long long), c->stream));
    { // caller arrays are pageable: chunked copies through pinned staging on several host threads
        std::vector<hostio::Segment> up;
        up.push_back({ reinterpret_cast<char*>(const_cast<float*>(w->density)), reinterpret_cast<char*>(dDensity), n * sizeof(float) });
        up.push_back({ reinterpret_cast<char*>(const_cast<uint8_t*>(w->material)), reinterpret_cast<char*>(dMat), n });
        if (w->measurement)
            up.push_back({ reinterpret_cast<char*>(const_cast<uint8_t*>(w->measurement)), reinterpret_cast<char*>(dMeas), n });
        CU_CHECK(c, hostio::copyChunked(c->device, up, true));
    }

    // palette form when the grid holds at most 256 distinct records (4-bit indices when at most 16: a 512x512x400 grid is
    // then 52 MB and stays resident in one 63 MB L2 partition), 8-byte records otherwise
    bool palette = false;
    c->world.paletteNibbles = 0;
    if (c->allowPalette) {
        unsigned long long* dTable = nullptr;
        unsigned* dSlotIndex = nullptr; // [kPaletteSlots] + the distinct-record counter
        CU_CHECK(c, cudaMalloc(&dTable, kPaletteSlots * sizeof(unsigned long long)));
        CU_CHECK(c, cudaMalloc(&dSlotIndex, (kPaletteSlots + 1) * sizeof(unsigned)));
        CU_CHECK(c, cudaMemsetAsync(dTable, 0xff, kPaletteSlots * sizeof(unsigned long long), c->stream));
        CU_CHECK(c, cudaMemsetAsync(dSlotIndex, 0, (kPaletteSlots + 1) * sizeof(unsigned), c->stream));
        unsigned* dDistinct = dSlotIndex + kPaletteSlots;
        paletteCollectKernel<<<gridFor(c, n), 256, 0, c->stream>>>(dDensity, dMat, dMeas, n, dTable, dDistinct);
        CU_CHECK(c, cudaGetLastError());
        unsigned distinct = 0;
        CU_CHECK(c, cudaMemcpyAsync(&distinct, dDistinct, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
        if (distinct <= 256u) {
            const int nibbles = distinct <= 16u && c->allowNibbles ? 1 : 0;
            CU_CHECK(c, hostio::poolAlloc(c->device, &c->dPalette, nibbles ? (n + 1) / 2 : n));
            CU_CHECK(c, cudaMalloc(&c->dPaletteTable, 256 * sizeof(uint2)));
            CU_CHECK(c, cudaMemsetAsync(c->dPaletteTable, 0, 256 * sizeof(uint2), c->stream));
            paletteNumberKernel<<<1, 256, 0, c->stream>>>(dTable, dSlotIndex, c->dPaletteTable);
            paletteIndexKernel<<<gridFor(c, n), 256, 0, c->stream>>>(dDensity, dMat, dMeas, n, dTable, dSlotIndex, c->dPalette, nibbles);
            CU_CHECK(c, cudaGetLastError());
            CU_CHECK(c, cudaStreamSynchronize(c->stream));
            palette = true;
            c->world.paletteNibbles = static_cast<uint32_t>(nibbles);
            c->paletteCount = distinct;
            c->hPaletteBase.assign(256, make_uint2(0u, 0u));
            CU_CHECK(c, cudaMemcpy(c->hPaletteBase.data(), c->dPaletteTable, 256 * sizeof(uint2), cudaMemcpyDeviceToHost));
        }
        cudaFree(dTable);
        cudaFree(dSlotIndex);
    }
    if (!palette) {
        CU_CHECK(c, hostio::poolAlloc(c->device, &c->dVoxels, n * sizeof(uint2)));
        packVoxelsKernel<<<gridFor(c, n), 256, 0, c->stream>>>(dDensity, dMat, dMeas, c->dVoxels, n);
        CU_CHECK(c, cudaGetLastError());
    }
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 3; ++i) {
        c->world.dim[i] = static_cast<uint32_t>(w->dim[i]);
        c->world.spacing[i] = w->spacing[i];
        c->world.invSpacing[i] = 1.0f / w->spacing[i]; // IEEE single division on the host == __frcp_rn
    }
    c->world.exactInverse = 1;
    for (int i = 0; i < 3; ++i) {
        int exponent = 0;
        if (std::frexp(w->spacing[i], &exponent) != 0.5f || exponent < -60 || exponent > 60)
            c->world.exactInverse = 0;
    }
    for (int i = 0; i < 6; ++i)
        c->world.ext[i] = w->extent_safe[i];
    c->world.voxels = c->dVoxels;
    c->world.palette = c->dPalette;
    c->world.paletteTable = c->dPaletteTable;
    c->bricksValid = false;
    c->voxelsFlagged = false;

    // Optional (DXMCB200_L2PERSIST=<MB>, 1 = as much as the device allows): keep the voxel grid in L2 with a persisting carve-out
    // of that size + an access-policy window on every pipeline's stream. Measured on B200 it LOWERS throughput at every size
    // tried (profiles/README.md): the carve-out takes L2 away from the accumulator and record traffic. Off by default.
    {
        const char* env = std::getenv("DXMCB200_L2PERSIST");
        int maxPersist = 0, maxWindow = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
        cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
        cudaStreamAttrValue attr {};
        const long mb = env ? std::atol(env) : 0;
        if (mb > 0 && maxPersist > 0 && maxWindow > 0) {
            const size_t gridBytes = palette ? (c->world.paletteNibbles ? (n + 1) / 2 : n) : n * sizeof(uint2);
            const size_t persist = mb == 1 ? static_cast<size_t>(maxPersist) : std::min<size_t>(static_cast<size_t>(mb) << 20, static_cast<size_t>(maxPersist));
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
            attr.accessPolicyWindow.base_ptr = palette ? static_cast<void*>(c->dPalette) : static_cast<void*>(c->dVoxels);
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(gridBytes, static_cast<size_t>(maxWindow));
            attr.accessPolicyWindow.hitRatio = std::min(1.0f, static_cast<float>(persist) / static_cast<float>(attr.accessPolicyWindow.num_bytes));
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        } else {
            attr.accessPolicyWindow.num_bytes = 0;
        }
        for (int i = 0; i < c->nPipes; ++i)
            if (c->pipes[i].stream)
                cudaStreamSetAttribute(c->pipes[i].stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError(); // a refused hint is not an error
    }
    return DXMCB200_OK;
}

int dxmcb200_set_luts(dxmcb200_ctx* c, const dxmcb200_luts* l)
{
    if (!c || !l || !l->knots || !l->coefficients || !l->max_coefficients || !l->rita || !l->spline || !l->shells || l->n_materials == 0
        || l->n_segments == 0)
        return DXMCB200_ERR_ARG;
    if (l->n_segments >= (1u << (32 - kSegmentShift))) { // event records keep the segment in 15 bits
        c->error = "more than 32767 LUT segments";
        return DXMCB200_ERR_ARG;
    }
    CU_CHECK(c, cudaSetDevice(c->device));
    const size_t nKnots = l->n_segments;
    const size_t nCoeffIn = static_cast<size_t>(l->n_materials) * l->n_segments * 6;
    const size_t nCoeff = static_cast<size_t>(l->n_materials) * l->n_segments * kCoeffStride;
    const size_t nMax = static_cast<size_t>(l->n_segments) * 2;
    const size_t nRita = static_cast<size_t>(l->n_materials) * 4 * DXMCB200_RITA_N;
    const size_t nSpline = static_cast<size_t>(l->n_materials) * kSplineStride;
    const size_t nShell = static_cast<size_t>(l->n_materials) * DXMCB200_SHELLS * DXMCB200_SHELL_FLOATS;
    // one blob, every table 32-byte (sector) aligned
    auto pad = [](size_t n) { return (n + 7) & ~static_cast<size_t>(7); };
    const size_t total = pad(nKnots) + pad(nCoeff) + pad(nMax) + pad(nRita) + pad(nSpline) + pad(nShell);
    std::vector<float> blob(total, 0.0f);
    size_t off = 0;
    auto put = [&](const float* src, size_t n) {
        std::memcpy(blob.data() + off, src, n * sizeof(float));
        const size_t at = off;
        off += pad(n);
        return at;
    };
    // device spline records are padded to 64 floats and carry 1/step for the division-free segment index
    std::vector<float> spline(nSpline, 0.0f);
    for (uint32_t m = 0; m < l->n_materials; ++m) {
        std::memcpy(spline.data() + static_cast<size_t>(m) * kSplineStride, l->spline + static_cast<size_t>(m) * DXMCB200_SPLINE_FLOATS,
            DXMCB200_SPLINE_FLOATS * sizeof(float));
        spline[static_cast<size_t>(m) * kSplineStride + 63] = 1.0f / l->spline[static_cast<size_t>(m) * DXMCB200_SPLINE_FLOATS + 61];
    }
    // device coefficient records are padded from 6 to kCoeffStride floats: one aligned 32-byte sector each
    std::vector<float> coeff(nCoeff, 0.0f);
    for (size_t r = 0; r < nCoeffIn / 6; ++r)
        std::memcpy(coeff.data() + r * kCoeffStride, l->coefficients + r * 6, 6 * sizeof(float));
    const size_t oKnots = put(l->knots, nKnots), oCoeff = put(coeff.data(), nCoeff), oMax = put(l->max_coefficients, nMax),
                 oRita = put(l->rita, nRita), oSpline = put(spline.data(), nSpline), oShell = put(l->shells, nShell);
    cudaFree(c->dLutBlob);
    c->dLutBlob = nullptr;
    CU_CHECK(c, cudaMalloc(&c->dLutBlob, total * sizeof(float)));
    CU_CHECK(c, uploadNow(c, c->dLutBlob, blob.data(), total * sizeof(float)));
    c->lut.nMaterials = l->n_materials;
    c->lut.nSegments = l->n_segments;
    c->lut.linearIndex = l->linear_index;
    c->lut.linearStep = l->linear_step;
    c->lut.linearEnergy = l->linear_energy;
    c->lut.invLinearStep = 1.0f / l->linear_step;
    c->lut.knots = c->dLutBlob + oKnots;
    c->lut.coeff = c->dLutBlob + oCoeff;
    c->lut.maxCoeff = c->dLutBlob + oMax;
    c->lut.rita = c->dLutBlob + oRita;
    c->lut.spline = c->dLutBlob + oSpline;
    c->lut.shells = c->dLutBlob + oShell;
    c->hKnots.assign(l->knots, l->knots + nKnots);
    c->hCoeff.assign(l->coefficients, l->coefficients + nCoeffIn);
    c->hMaxCoeff.assign(l->max_coefficients, l->max_coefficients + nMax);
    c->bricksValid = false;
    return DXMCB200_OK;
}

int dxmcb200_set_beam_tables(dxmcb200_ctx* c, uint32_t nSpectra, const dxmcb200_spectrum* spectra, uint32_t nHeel, const dxmcb200_heel* heel,
    uint32_t nBowtie, const dxmcb200_bowtie* bowtie)
{
    if (!c || (nSpectra && !spectra) || (nHeel && !heel) || (nBowtie && !bowtie))
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    // size of the blob: view structs followed by their arrays
    size_t bytes = 64;
    bytes += nSpectra * sizeof(SpectrumView) + nHeel * sizeof(HeelView) + nBowtie * sizeof(BowtieView) + 48;
    for (uint32_t i = 0; i < nSpectra; ++i)
        bytes += static_cast<size_t>(spectra[i].n) * 12 + 48;
    for (uint32_t i = 0; i < nHeel; ++i)
        bytes += static_cast<size_t>(heel[i].energy_size) * heel[i].angle_size * 4 + 16;
    for (uint32_t i = 0; i < nBowtie; ++i)
        bytes += static_cast<size_t>(bowtie[i].n) * 8 + 32;
    std::vector<char> host(bytes, 0);
    cudaFree(c->dBeamBlob);
    c->dBeamBlob = nullptr;
    CU_CHECK(c, cudaMalloc(&c->dBeamBlob, bytes));
    char* dBase = static_cast<char*>(c->dBeamBlob);
    char* cursor = host.data();
    auto toDevice = [&](const void* hostPtr) { return dBase + (static_cast<const char*>(hostPtr) - host.data()); };

    auto* sv = advancePtr<SpectrumView>(cursor, nSpectra);
    auto* hv = advancePtr<HeelView>(cursor, nHeel);
    auto* bv = advancePtr<BowtieView>(cursor, nBowtie);
    for (uint32_t i = 0; i < nSpectra; ++i) {
        const auto& s = spectra[i];
        if (s.n == 0 || !s.probs || !s.alias || !s.energies)
            return DXMCB200_ERR_ARG;
        float* probs = advancePtr<float>(cursor, s.n);
        uint32_t* alias = advancePtr<uint32_t>(cursor, s.n);
        float* energies = advancePtr<float>(cursor, s.n);
        std::memcpy(probs, s.probs, s.n * 4);
        std::memcpy(alias, s.alias, s.n * 4);
        std::memcpy(energies, s.energies, s.n * 4);
        sv[i].n = s.n;
        // RandomState::randomUniform<std::size_t>(max): threshold = uint32(-max % max) in 64-bit arithmetic
        // (dxmcrandom.hpp:87), i.e. 2^64 mod n truncated to 32 bits
        const uint64_t n64 = s.n;
        sv[i].threshold = static_cast<uint32_t>((0 - n64) % n64);
        sv[i].probs = reinterpret_cast<const float*>(toDevice(probs));
        sv[i].alias = reinterpret_cast<const uint32_t*>(toDevice(alias));
        sv[i].energies = reinterpret_cast<const float*>(toDevice(energies));
    }
    for (uint32_t i = 0; i < nHeel; ++i) {
        const auto& h = heel[i];
        const size_t n = static_cast<size_t>(h.energy_size) * h.angle_size;
        if (n == 0 || !h.weights)
            return DXMCB200_ERR_ARG;
        float* w = advancePtr<float>(cursor, n);
        std::memcpy(w, h.weights, n * 4);
        hv[i] = HeelView { h.energy_start, h.energy_step, h.energy_size, h.angle_start, h.angle_step, h.angle_size,
            reinterpret_cast<const float*>(toDevice(w)) };
    }
    for (uint32_t i = 0; i < nBowtie; ++i) {
        const auto& b = bowtie[i];
        if (b.n == 0 || !b.angles || !b.weights)
            return DXMCB200_ERR_ARG;
        float* a = advancePtr<float>(cursor, b.n);
        float* w = advancePtr<float>(cursor, b.n);
        std::memcpy(a, b.angles, b.n * 4);
        std::memcpy(w, b.weights, b.n * 4);
        bv[i] = BowtieView { b.n, reinterpret_cast<const float*>(toDevice(a)), reinterpret_cast<const float*>(toDevice(w)) };
    }
    if (static_cast<size_t>(cursor - host.data()) > bytes) {
        c->error = "beam table blob overflow";
        return DXMCB200_ERR_STATE;
    }
    CU_CHECK(c, uploadNow(c, c->dBeamBlob, host.data(), bytes));
    c->beams.spectra = reinterpret_cast<const SpectrumView*>(toDevice(sv));
    c->beams.heels = reinterpret_cast<const HeelView*>(toDevice(hv));
    c->beams.bowties = reinterpret_cast<const BowtieView*>(toDevice(bv));
    return DXMCB200_OK;
}

int dxmcb200_suggest_fixed_point(uint64_t totalHistories, double maxEnergyWeight, int* energyBits, int* energySqBits)
{
    if (!energyBits || !energySqBits || maxEnergyWeight <= 0)
        return DXMCB200_ERR_ARG;
    // Russian roulette can multiply weights by 5 only while E*w < 5 keV, so one history never scores more than
    // max(maxEnergyWeight, 25) keV into one voxel in total
    const double perHistory = std::max(maxEnergyWeight, 25.0);
    const double n = static_cast<double>(std::max<uint64_t>(totalHistories, 1));
    const double boundE = n * perHistory;
    const double boundE2 = n * perHistory * perHistory;
    const int be = static_cast<int>(std::floor(62.0 - std::log2(boundE)));
    const int b2 = static_cast<int>(std::floor(63.0 - std::log2(boundE2)));
    *energyBits = std::clamp(be, 0, 40);
    *energySqBits = std::clamp(b2, 0, 40);
    return DXMCB200_OK;
}

int dxmcb200_set_fixed_point(dxmcb200_ctx* c, int energyBits, int energySqBits)
{
    if (!c || energyBits < 0 || energyBits > 40 || energySqBits < 0 || energySqBits > 40)
        return DXMCB200_ERR_ARG;
    c->energyBits = energyBits;
    c->energySqBits = energySqBits;
    return DXMCB200_OK;
}

int dxmcb200_clear(dxmcb200_ctx* c)
{
    if (!c)
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    if (c->dAcc)
        CU_CHECK(c, cudaMemsetAsync(c->dAcc, 0, (c->nVoxels + kAccPadding) * 4 * sizeof(unsigned long long), c->stream));
    CU_CHECK(c, cudaMemsetAsync(c->dCounters, 0, sizeof(Counters), c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    c->totalMs = 0;
    c->lastRunMs = 0;
    c->launches = 0;
    for (int k = 0; k < 4; ++k) {
        c->kernelMs[k] = 0;
        c->kernelLaunches[k] = 0;
    }
    return DXMCB200_OK;
}

void dxmcb200_history_stream(uint64_t seed, uint64_t exposure, uint64_t history, uint64_t out[2])
{
    uint64_t s, i;
    historyStream(seed, exposure, history, s, i);
    out[0] = s;
    out[1] = i;
}

int dxmcb200_upload_exposures(dxmcb200_ctx* c, const dxmcb200_exposure* exposures, uint64_t n)
{
    if (!c || !exposures || n == 0)
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    if (n > c->nExposuresResident) {
        cudaFree(c->dExposures);
        c->dExposures = nullptr;
        c->nExposuresResident = 0;
        CU_CHECK(c, cudaMalloc(&c->dExposures, n * sizeof(dxmcb200_exposure)));
    }
    CU_CHECK(c, uploadNow(c, c->dExposures, exposures, n * sizeof(dxmcb200_exposure)));
    c->nExposuresResident = n;
    c->hExposures.assign(exposures, exposures + n);
    return DXMCB200_OK;
}

int dxmcb200_generate_exposures(dxmcb200_ctx* c, const dxmcb200_source_params* params, const float* aecProfile, dxmcb200_exposure* out)
{
    if (!c || !params || params->exposures == 0 || params->motion > DXMCB200_SOURCE_GANTRY_TOPOGRAM || (params->aec_size && !aecProfile)
        || (params->tubes != 1 && params->tubes != 2))
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    const uint64_t n = params->exposures;
    if (n > c->nExposuresResident) {
        cudaFree(c->dExposures);
        c->dExposures = nullptr;
        c->nExposuresResident = 0;
        CU_CHECK(c, cudaMalloc(&c->dExposures, n * sizeof(dxmcb200_exposure)));
    }
    float* dProfile = nullptr;
    if (params->aec_size) {
        CU_CHECK(c, cudaMalloc(&dProfile, params->aec_size * sizeof(float)));
        cudaError_t e = cudaMemcpyAsync(dProfile, aecProfile, params->aec_size * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            cudaFree(dProfile);
            CU_CHECK(c, e);
        }
    }
    dxmc::model::SourceParams<float> block;
    std::memcpy(&block, params, sizeof(block));
    exposureKernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, c->stream>>>(block, dProfile, n, c->dExposures);
    cudaError_t e = cudaGetLastError();
    c->hExposures.resize(n);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->hExposures.data(), c->dExposures, n * sizeof(dxmcb200_exposure), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(dProfile);
    CU_CHECK(c, e);
    c->nExposuresResident = n;
    if (out)
        std::memcpy(out, c->hExposures.data(), n * sizeof(dxmcb200_exposure));
    return DXMCB200_OK;
}

int dxmcb200_exposure_table(dxmcb200_ctx* c, void** devicePtr, uint64_t* n)
{
    if (!c || !c->dExposures)
        return DXMCB200_ERR_STATE;
    if (devicePtr)
        *devicePtr = c->dExposures;
    if (n)
        *n = c->nExposuresResident;
    return DXMCB200_OK;
}

int dxmcb200_run_range(dxmcb200_ctx* c, uint64_t expBegin, uint64_t expEnd, int model, uint64_t seed, const volatile int* cancel, dxmcb200_progress_cb cb,
    void* user)
{
    if (!c || !c->dExposures || expEnd > c->nExposuresResident || expEnd < expBegin)
        return DXMCB200_ERR_STATE;
    if (expEnd == expBegin)
        return DXMCB200_OK;
    return runRange(c, c->hExposures.data(), c->dExposures, expBegin, expEnd - expBegin, 1, model, seed, cancel, cb, user);
}

int dxmcb200_run_resident(dxmcb200_ctx* c, uint64_t expBegin, uint64_t expEnd, int model, uint64_t seed)
{
    if (!c || !c->dExposures || expEnd > c->nExposuresResident)
        return DXMCB200_ERR_STATE;
    return runRange(c, c->hExposures.data(), c->dExposures, expBegin, expEnd - expBegin, 1, model, seed, nullptr, nullptr, nullptr);
}

int dxmcb200_run_strided(dxmcb200_ctx* c, uint64_t expFirst, uint64_t expStride, uint64_t expCount, int model, uint64_t seed)
{
    if (!c || !c->dExposures)
        return DXMCB200_ERR_STATE;
    if (expStride == 0)
        return DXMCB200_ERR_ARG;
    if (expCount == 0)
        return DXMCB200_OK;
    if (expFirst >= c->nExposuresResident || (c->nExposuresResident - 1 - expFirst) / expStride < expCount - 1)
        return DXMCB200_ERR_STATE;
    return runRange(c, c->hExposures.data(), c->dExposures, expFirst, expCount, expStride, model, seed, nullptr, nullptr, nullptr);
}

int dxmcb200_run_strided_monitored(dxmcb200_ctx* c, uint64_t expFirst, uint64_t expStride, uint64_t expCount, int model, uint64_t seed,
    const volatile int* cancel, dxmcb200_progress_cb cb, void* user)
{
    if (!c || !c->dExposures)
        return DXMCB200_ERR_STATE;
    if (expStride == 0)
        return DXMCB200_ERR_ARG;
    if (expCount == 0)
        return DXMCB200_OK;
    if (expFirst >= c->nExposuresResident || (c->nExposuresResident - 1 - expFirst) / expStride < expCount - 1)
        return DXMCB200_ERR_STATE;
    return runRange(c, c->hExposures.data(), c->dExposures, expFirst, expCount, expStride, model, seed, cancel, cb, user);
}

int dxmcb200_run(dxmcb200_ctx* c, const dxmcb200_exposure* exposures, uint64_t expBegin, uint64_t expEnd, int model, uint64_t seed,
    const volatile int* cancel, dxmcb200_progress_cb cb, void* user)
{
    if (!c || !exposures || expEnd < expBegin)
        return DXMCB200_ERR_ARG;
    if (expEnd == expBegin)
        return DXMCB200_OK;
    const int up = dxmcb200_upload_exposures(c, exposures, expEnd);
    if (up != DXMCB200_OK)
        return up;
    return runRange(c, exposures, c->dExposures, expBegin, expEnd - expBegin, 1, model, seed, cancel, cb, user);
}

int dxmcb200_last_run_ms(dxmcb200_ctx* c, double* ms)
{
    if (!c || !ms)
        return DXMCB200_ERR_ARG;
    *ms = c->lastRunMs;
    return DXMCB200_OK;
}

// decode voxels [first, first + n) of the accumulators into the caller's host arrays (indexed from `first`)
static int collectRange(dxmcb200_ctx* c, int mode, uint64_t totalHistories, float calibration, uint64_t first, uint64_t n, float* dose, uint32_t* nEvents,
    float* variance)
{
    if (n == 0)
        return DXMCB200_OK;
    struct Decoded { // device copies of the Result arrays, back to the pool on every exit path
        float* dose = nullptr;
        float* variance = nullptr;
        uint32_t* events = nullptr;
        ~Decoded()
        {
            hostio::poolFree(dose);
            hostio::poolFree(variance);
            hostio::poolFree(events);
        }
    } d;
    if (dose)
        CU_CHECK(c, hostio::poolAlloc(c->device, &d.dose, n * sizeof(float)));
    if (variance)
        CU_CHECK(c, hostio::poolAlloc(c->device, &d.variance, n * sizeof(float)));
    if (nEvents)
        CU_CHECK(c, hostio::poolAlloc(c->device, &d.events, n * sizeof(uint32_t)));
    const float voxelVolume = c->world.spacing[0] * c->world.spacing[1] * c->world.spacing[2] / 1000.0f;
    resultKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->dAcc, c->dVoxels, c->dPalette, c->dPaletteTable, static_cast<int>(c->world.paletteNibbles), first, n, mode,
        std::ldexp(1.0f, -c->energyBits),
        std::ldexp(1.0f, -c->energySqBits), totalHistories, calibration, voxelVolume, d.dose, d.events, d.variance);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    std::vector<hostio::Segment> down; // caller arrays are pageable: chunked copies through pinned staging
    if (dose)
        down.push_back({ reinterpret_cast<char*>(dose + first), reinterpret_cast<char*>(d.dose), n * sizeof(float) });
    if (variance)
        down.push_back({ reinterpret_cast<char*>(variance + first), reinterpret_cast<char*>(d.variance), n * sizeof(float) });
    if (nEvents)
        down.push_back({ reinterpret_cast<char*>(nEvents + first), reinterpret_cast<char*>(d.events), n * sizeof(uint32_t) });
    CU_CHECK(c, hostio::copyChunked(c->device, down, false));
    return DXMCB200_OK;
}

int dxmcb200_get_result(dxmcb200_ctx* c, int mode, uint64_t totalHistories, float calibration, float* dose, uint32_t* nEvents, float* variance)
{
    if (!c || !c->dAcc || mode < 0 || mode > 2)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    return collectRange(c, mode, totalHistories, calibration, 0, c->nVoxels, dose, nEvents, variance);
}

int dxmcb200_get_raw(dxmcb200_ctx* c, int64_t* energy, uint64_t* energySq, uint64_t* events)
{
    if (!c || !c->dAcc)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    const uint64_t n = c->nVoxels;
    struct Raw {
        long long* energy = nullptr;
        unsigned long long* energySq = nullptr;
        unsigned long long* events = nullptr;
        ~Raw()
        {
            hostio::poolFree(energy);
            hostio::poolFree(energySq);
            hostio::poolFree(events);
        }
    } d;
    if (energy)
        CU_CHECK(c, hostio::poolAlloc(c->device, &d.energy, n * 8));
    if (energySq)
        CU_CHECK(c, hostio::poolAlloc(c->device, &d.energySq, n * 8));
    if (events)
        CU_CHECK(c, hostio::poolAlloc(c->device, &d.events, n * 8));
    rawKernel<<<gridFor(c, n), 256, 0, c->stream>>>(c->dAcc, n, d.energy, d.energySq, d.events);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    std::vector<hostio::Segment> down;
    if (energy)
        down.push_back({ reinterpret_cast<char*>(energy), reinterpret_cast<char*>(d.energy), n * 8 });
    if (energySq)
        down.push_back({ reinterpret_cast<char*>(energySq), reinterpret_cast<char*>(d.energySq), n * 8 });
    if (events)
        down.push_back({ reinterpret_cast<char*>(events), reinterpret_cast<char*>(d.events), n * 8 });
    CU_CHECK(c, hostio::copyChunked(c->device, down, false));
    return DXMCB200_OK;
}

int dxmcb200_accumulators(dxmcb200_ctx* c, void** devicePtr, uint64_t* nU64)
{
    if (!c || !c->dAcc || !devicePtr || !nU64)
        return DXMCB200_ERR_STATE;
    *devicePtr = c->dAcc;
    *nU64 = c->nVoxels * 4;
    return DXMCB200_OK;
}

// ---- NCCL, loaded on first use (a single-GPU user never needs it) -------------------------------------------------------
namespace {
struct Nccl {
    // ncclUint64 = 5, ncclSum = 0
    int (*allReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*reduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*commInitAll)(void**, int, const int*) = nullptr;
    int (*commDestroy)(void*) = nullptr;
    bool ok = false;
};
const Nccl& nccl()
{
    static const Nccl api = [] {
        Nccl n;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            n.allReduce = reinterpret_cast<decltype(n.allReduce)>(dlsym(h, "ncclAllReduce"));
            n.reduceScatter = reinterpret_cast<decltype(n.reduceScatter)>(dlsym(h, "ncclReduceScatter"));
            n.commInitAll = reinterpret_cast<decltype(n.commInitAll)>(dlsym(h, "ncclCommInitAll"));
            n.commDestroy = reinterpret_cast<decltype(n.commDestroy)>(dlsym(h, "ncclCommDestroy"));
            n.ok = n.allReduce && n.reduceScatter && n.commInitAll && n.commDestroy;
        }
        return n;
    }();
    return api;
}
} // namespace

int dxmcb200_reduce(dxmcb200_ctx* c, void* comm)
{
    if (!c || !c->dAcc || !comm)
        return DXMCB200_ERR_STATE;
    if (!nccl().ok) {
        c->error = "libnccl not found";
        return DXMCB200_ERR_NCCL;
    }
    CU_CHECK(c, cudaSetDevice(c->device));
    if (nccl().allReduce(c->dAcc, c->dAcc, c->nVoxels * 4, 5, 0, comm, c->stream) != 0) {
        c->error = "ncclAllReduce failed";
        return DXMCB200_ERR_NCCL;
    }
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    return DXMCB200_OK;
}

int dxmcb200_comm_create(int n, const int* devices, void** comms)
{
    if (n < 1 || !devices || !comms)
        return DXMCB200_ERR_ARG;
    if (!nccl().ok)
        return DXMCB200_ERR_NCCL;
    return nccl().commInitAll(comms, n, devices) == 0 ? DXMCB200_OK : DXMCB200_ERR_NCCL;
}

int dxmcb200_comm_destroy(int n, void** comms)
{
    if (n < 1 || !comms)
        return DXMCB200_ERR_ARG;
    if (!nccl().ok)
        return DXMCB200_ERR_NCCL;
    for (int i = 0; i < n; ++i)
        if (comms[i])
            nccl().commDestroy(comms[i]);
    return DXMCB200_OK;
}

int dxmcb200_reduce_collect(dxmcb200_ctx* c, void* comm, int rank, int nRanks, int mode, uint64_t totalHistories, float calibration, float* dose,
    uint32_t* nEvents, float* variance)
{
    if (!c || !c->dAcc || !comm || nRanks < 1 || nRanks > static_cast<int>(kAccPadding) || rank < 0 || rank >= nRanks || mode < 0 || mode > 2)
        return DXMCB200_ERR_STATE;
    if (!nccl().ok) {
        c->error = "libnccl not found";
        return DXMCB200_ERR_NCCL;
    }
    CU_CHECK(c, cudaSetDevice(c->device));
    // equal slices of voxels (the last ones reach into the zero padding behind the grid); in place: rank r receives the sums of
    // its own slice where that slice already lies
    const uint64_t slice = (c->nVoxels + static_cast<uint64_t>(nRanks) - 1) / static_cast<uint64_t>(nRanks);
    unsigned long long* mine = c->dAcc + static_cast<uint64_t>(rank) * slice * 4;
    if (nccl().reduceScatter(c->dAcc, mine, slice * 4, 5, 0, comm, c->stream) != 0) {
        c->error = "ncclReduceScatter failed";
        return DXMCB200_ERR_NCCL;
    }
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    const uint64_t first = std::min(static_cast<uint64_t>(rank) * slice, c->nVoxels);
    const uint64_t count = std::min(slice, c->nVoxels - first);
    return collectRange(c, mode, totalHistories, calibration, first, count, dose, nEvents, variance);
}

int dxmcb200_get_stats(dxmcb200_ctx* c, dxmcb200_stats* s)
{
    if (!c || !s)
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    Counters h {};
    CU_CHECK(c, cudaMemcpy(&h, c->dCounters, sizeof(Counters), cudaMemcpyDeviceToHost));
    s->histories = h.histories;
    s->histories_in_world = h.inWorld;
    s->steps = h.steps;
    s->lookups = h.lookups;
    s->interactions = h.interactions;
    s->score_events = h.scores;
    s->kernel_launches = c->launches;
    s->kernel_ms = c->totalMs;
    s->air_walks = h.airWalks;
    s->bricks_crossed = h.bricksCrossed;
    return DXMCB200_OK;
}

int dxmcb200_get_kernel_times(dxmcb200_ctx* c, double ms[4], uint64_t launches[4])
{
    if (!c || !ms || !launches)
        return DXMCB200_ERR_ARG;
    for (int k = 0; k < 4; ++k) {
        ms[k] = c->kernelMs[k];
        launches[k] = c->kernelLaunches[k];
    }
    return DXMCB200_OK;
}

int dxmcb200_set_tracking(dxmcb200_ctx* c, int tracking, float brickMm)
{
    if (!c || tracking < 0 || tracking > 1 || (brickMm != 0.0f && !(brickMm > 0.0f)))
        return DXMCB200_ERR_ARG;
    c->tracking = tracking;
    if (brickMm > 0.0f)
        c->brickMm = brickMm;
    c->bricksValid = false;
    return DXMCB200_OK;
}

int dxmcb200_get_brick_distance(dxmcb200_ctx* c, uint8_t* distance)
{
    if (!c || !distance || (!c->dVoxels && !c->dPalette) || !c->dLutBlob)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    const int st = ensureBricks(c);
    if (st != DXMCB200_OK)
        return st;
    std::copy(c->hDistance.begin(), c->hDistance.end(), distance);
    return DXMCB200_OK;
}

int dxmcb200_get_bricks(dxmcb200_ctx* c, uint32_t shift[3], uint32_t nb[3], float* fAir, float* ratio, float* brickMax, uint8_t* air)
{
    if (!c || !shift || !nb || !fAir || (!c->dVoxels && !c->dPalette) || !c->dLutBlob)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    const int st = ensureBricks(c);
    if (st != DXMCB200_OK)
        return st;
    for (int i = 0; i < 3; ++i) {
        shift[i] = c->bricks.shift[i];
        nb[i] = c->bricks.nb[i];
    }
    *fAir = c->bricks.nWords ? c->fAir : 0.0f;
    if (ratio)
        std::copy(c->hRatio.begin(), c->hRatio.end(), ratio);
    if (brickMax)
        std::copy(c->hBrickMax.begin(), c->hBrickMax.end(), brickMax);
    if (air)
        std::copy(c->hAir.begin(), c->hAir.end(), air);
    return DXMCB200_OK;
}

int dxmcb200_get_grid_form(dxmcb200_ctx* c, int* bitsPerVoxel, uint32_t* distinctRecords)
{
    if (!c || !bitsPerVoxel || (!c->dVoxels && !c->dPalette))
        return DXMCB200_ERR_STATE;
    if (c->dLutBlob) { // the air-brick flags of tracking mode 1 may widen a palette grid
        CU_CHECK(c, cudaSetDevice(c->device));
        const int st = ensureBricks(c);
        if (st != DXMCB200_OK)
            return st;
    }
    *bitsPerVoxel = c->dPalette ? (c->world.paletteNibbles ? 4 : 8) : 64;
    if (distinctRecords)
        *distinctRecords = c->dPalette ? c->paletteCount : 0u;
    return DXMCB200_OK;
}

int dxmcb200_enable_stats(dxmcb200_ctx* c, int on)
{
    if (!c)
        return DXMCB200_ERR_ARG;
    c->collectStats = on != 0;
    return DXMCB200_OK;
}

int dxmcb200_eval_attenuation(dxmcb200_ctx* c, uint64_t n, const uint8_t* material, const float* energy, float* out3, float* outMax)
{
    if (!c || !c->dLutBlob || !material || !energy || !out3 || !outMax || n == 0)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    uint8_t* dM = nullptr;
    float *dE = nullptr, *dO = nullptr, *dX = nullptr;
    CU_CHECK(c, cudaMalloc(&dM, n));
    CU_CHECK(c, cudaMalloc(&dE, n * 4));
    CU_CHECK(c, cudaMalloc(&dO, n * 12));
    CU_CHECK(c, cudaMalloc(&dX, n * 4));
    CU_CHECK(c, uploadNow(c, dM, material, n));
    CU_CHECK(c, uploadNow(c, dE, energy, n * 4));
    evalAttenuationKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(c->lut, n, dM, dE, dO, dX);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    CU_CHECK(c, cudaMemcpy(out3, dO, n * 12, cudaMemcpyDeviceToHost));
    CU_CHECK(c, cudaMemcpy(outMax, dX, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dM);
    cudaFree(dE);
    cudaFree(dO);
    cudaFree(dX);
    return DXMCB200_OK;
}

int dxmcb200_trace_indices(dxmcb200_ctx* c, uint64_t nRays, const float* pos, const float* dir, uint32_t nSteps, const float* steps,
    int64_t* outIdx, float* outEntry)
{
    if (!c || (!c->dVoxels && !c->dPalette) || !pos || !dir || !steps || !outIdx || !outEntry || nRays == 0)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    float *dP = nullptr, *dD = nullptr, *dS = nullptr, *dEn = nullptr;
    long long* dI = nullptr;
    CU_CHECK(c, cudaMalloc(&dP, nRays * 12));
    CU_CHECK(c, cudaMalloc(&dD, nRays * 12));
    CU_CHECK(c, cudaMalloc(&dS, std::max<size_t>(nSteps, 1) * 4));
    CU_CHECK(c, cudaMalloc(&dEn, nRays * 12));
    CU_CHECK(c, cudaMalloc(&dI, nRays * (nSteps + 1) * 8));
    CU_CHECK(c, uploadNow(c, dP, pos, nRays * 12));
    CU_CHECK(c, uploadNow(c, dD, dir, nRays * 12));
    if (nSteps)
        CU_CHECK(c, uploadNow(c, dS, steps, nSteps * 4));
    traceIndicesKernel<<<static_cast<unsigned>((nRays + 127) / 128), 128, 0, c->stream>>>(c->world, nRays, dP, dD, nSteps, dS, dI, dEn);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    CU_CHECK(c, cudaMemcpy(outIdx, dI, nRays * (nSteps + 1) * 8, cudaMemcpyDeviceToHost));
    CU_CHECK(c, cudaMemcpy(outEntry, dEn, nRays * 12, cudaMemcpyDeviceToHost));
    cudaFree(dP);
    cudaFree(dD);
    cudaFree(dS);
    cudaFree(dEn);
    cudaFree(dI);
    return DXMCB200_OK;
}

int dxmcb200_tube_bremsstrahlung(float tubeVoltage, uint32_t nBins, const float* energies, const float* tungstenAttenuation, uint32_t nAngles,
    const float* angles, float* out)
{
    if (!energies || !tungstenAttenuation || !angles || !out || nBins == 0 || nAngles == 0)
        return DXMCB200_ERR_ARG;
    int devices = 0;
    if (cudaGetDeviceCount(&devices) != cudaSuccess || devices <= 0)
        return DXMCB200_ERR_NO_DEVICE;
    float *dE = nullptr, *dA = nullptr, *dT = nullptr, *dO = nullptr;
    cudaError_t e = cudaMalloc(&dE, nBins * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMalloc(&dT, nBins * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMalloc(&dA, nAngles * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMalloc(&dO, static_cast<size_t>(nBins) * nAngles * sizeof(float));
    cudaStream_t stream = nullptr;
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = spectrum::uploadTables(stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dE, energies, nBins * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dT, tungstenAttenuation, nBins * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dA, angles, nAngles * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
        spectrum::bremsstrahlungKernel<<<nBins * nAngles, spectrum::kMaxDepthSteps, 0, stream>>>(tubeVoltage, nBins, dE, dT, dA, dO);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(out, dO, static_cast<size_t>(nBins) * nAngles * sizeof(float), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(stream);
    if (stream)
        cudaStreamDestroy(stream);
    cudaFree(dE);
    cudaFree(dT);
    cudaFree(dA);
    cudaFree(dO);
    return e == cudaSuccess ? DXMCB200_OK : DXMCB200_ERR_CUDA;
}

int dxmcb200_trace_air_runs(dxmcb200_ctx* c, uint64_t nRays, const float* pos, const float* dir, float* outLength, uint32_t* outInfo, float* outEnd)
{
    if (!c || (!c->dVoxels && !c->dPalette) || !c->dLutBlob || !pos || !dir || !outLength || !outInfo || !outEnd || nRays == 0)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    const int st = ensureBricks(c);
    if (st != DXMCB200_OK)
        return st;
    if (c->bricks.nWords == 0) {
        c->error = "no air bricks (tracking mode 0, or a grid without air)";
        return DXMCB200_ERR_STATE;
    }
    float *dP = nullptr, *dD = nullptr, *dL = nullptr, *dE = nullptr;
    uint32_t* dI = nullptr;
    CU_CHECK(c, cudaMalloc(&dP, nRays * 12));
    CU_CHECK(c, cudaMalloc(&dD, nRays * 12));
    CU_CHECK(c, cudaMalloc(&dL, nRays * 4));
    CU_CHECK(c, cudaMalloc(&dI, nRays * 4));
    CU_CHECK(c, cudaMalloc(&dE, nRays * 12));
    CU_CHECK(c, uploadNow(c, dP, pos, nRays * 12));
    CU_CHECK(c, uploadNow(c, dD, dir, nRays * 12));
    traceAirRunsKernel<<<static_cast<unsigned>((nRays + 127) / 128), 128, 0, c->stream>>>(c->world, c->bricks, nRays, dP, dD, dL, dI, dE);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    CU_CHECK(c, cudaMemcpy(outLength, dL, nRays * 4, cudaMemcpyDeviceToHost));
    CU_CHECK(c, cudaMemcpy(outInfo, dI, nRays * 4, cudaMemcpyDeviceToHost));
    CU_CHECK(c, cudaMemcpy(outEnd, dE, nRays * 12, cudaMemcpyDeviceToHost));
    cudaFree(dP);
    cudaFree(dD);
    cudaFree(dL);
    cudaFree(dI);
    cudaFree(dE);
    return DXMCB200_OK;
}

int dxmcb200_sample_particles(dxmcb200_ctx* c, const dxmcb200_exposure* e, uint64_t exposureIndex, uint64_t seed, uint64_t n, float* out)
{
    if (!c || !e || !out || n == 0)
        return DXMCB200_ERR_ARG;
    if ((e->spectrum >= 0 || e->heel >= 0 || e->bowtie >= 0) && !c->dBeamBlob)
        return DXMCB200_ERR_STATE;
    CU_CHECK(c, cudaSetDevice(c->device));
    float* dO = nullptr;
    CU_CHECK(c, cudaMalloc(&dO, n * 32));
    sampleParticlesKernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(*e, c->beams, exposureIndex, seed, n, dO);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    CU_CHECK(c, cudaMemcpy(out, dO, n * 32, cudaMemcpyDeviceToHost));
    cudaFree(dO);
    return DXMCB200_OK;
}

int dxmcb200_sample_interaction(dxmcb200_ctx* c, int kind, int model, uint8_t material, float energy, uint64_t seed, uint64_t n, float* out)
{
    if (!c || !c->dLutBlob || !out || n == 0 || kind < 0 || kind > 2 || model < 0 || model > 2 || material >= c->lut.nMaterials)
        return DXMCB200_ERR_ARG;
    CU_CHECK(c, cudaSetDevice(c->device));
    float* dO = nullptr;
    CU_CHECK(c, cudaMalloc(&dO, n * 20));
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (model == 0)
        sampleInteractionKernel<0><<<grid, 256, 0, c->stream>>>(c->lut, kind, material, energy, seed, n, dO);
    else if (model == 1)
        sampleInteractionKernel<1><<<grid, 256, 0, c->stream>>>(c->lut, kind, material, energy, seed, n, dO);
    else
        sampleInteractionKernel<2><<<grid, 256, 0, c->stream>>>(c->lut, kind, material, energy, seed, n, dO);
    CU_CHECK(c, cudaGetLastError());
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    CU_CHECK(c, cudaMemcpy(out, dO, n * 20, cudaMemcpyDeviceToHost));
    cudaFree(dO);
    return DXMCB200_OK;
}

} // extern "C"
