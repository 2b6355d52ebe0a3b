// spectrum.cuh — thick-target tungsten bremsstrahlung on the device.
//
// The reference computes one Bethe-Heitler depth integral per energy bin of a tube spectrum, and again for every take-off angle of
// the heel-effect table (betheHeitlerCrossSection.hpp:368-407 called from tube.hpp:191-208 and beamfilters.hpp:465-505): 141 depths
// x 200 electron energies of pow / exp / log / sqrt per bin, 0.4 s of host time for a CT source with heel model. Here one block
// takes one (angle, bin) pair: thread k evaluates depth k's inner sum over the electron energies in the reference's order, thread
// 0 adds the depths up in the reference's order. Same single-precision operations in the same order as the host code in
// dxmc/betheHeitlerCrossSection.hpp; the results differ from the host's only by the last-ulp differences between CUDA's and the
// host libm's powf / expf / logf (tests: 2e-5 relative on the normalised spectrum). Opt-in (DXMCB200_DEVICE_SPECTRUM=1): the
// default stays the host path, whose tables are bit-identical to the reference's.
#pragma once

#include "../include/dxmc/tungsten_electron_data.hpp"

#include <cuda_runtime.h>

namespace dxmcb200::spectrum {

namespace tw = dxmc::tungsten;

struct ElectronTables { // device copy of dxmc/tungsten_electron_data.hpp (Poludniowski & Evans 2007)
    float depthF[tw::kDepths], depthM[tw::kDepths], relEnergy[tw::kEnergies];
    float twVoltage[5], twConstant[5];
    float densityF[tw::kDepths][tw::kEnergies], densityM[tw::kDepths][tw::kEnergies];
};
__constant__ ElectronTables kTables;

constexpr float kElectronMass = 510.9989461f; // constants.hpp:63
constexpr int kMaxDepthSteps = 160; // x = 0, 0.1, ... <= 14 in float accumulation: 141 values

// index of the table interval used for v: starts at the first knot >= v, pulled back so that two knots remain
template <int N>
__device__ __forceinline__ int intervalStart(const float (&knots)[N], float v)
{
    int i = 0;
    while (i < N && knots[i] < v)
        ++i;
    const int left = N - i;
    if (left < 2)
        i = N - 3 + left;
    return i;
}

__device__ __forceinline__ float bilinear(float q11, float q12, float q21, float q22, float x1, float x2, float y1, float y2, float x, float y)
{
    const float xf1 = (x2 - x) / (x2 - x1);
    const float xf2 = (x - x1) / (x2 - x1);
    const float r1 = xf1 * q11 + xf2 * q21;
    const float r2 = xf1 * q12 + xf2 * q22;
    return ((y2 - y) / (y2 - y1)) * r1 + ((y - y1) / (y2 - y1)) * r2;
}

__device__ __forceinline__ float tableDensity(const float (&depths)[tw::kDepths], const float (&table)[tw::kDepths][tw::kEnergies], float uval, float xval)
{
    const float x = fminf(fmaxf(xval, depths[0]), depths[tw::kDepths - 1]);
    const float u = fminf(fmaxf(uval, kTables.relEnergy[0]), kTables.relEnergy[tw::kEnergies - 1]);
    const int ix = intervalStart(depths, x);
    const int iu = intervalStart(kTables.relEnergy, u);
    return bilinear(table[ix][iu], table[ix][iu + 1], table[ix + 1][iu], table[ix + 1][iu + 1], depths[ix], depths[ix + 1], kTables.relEnergy[iu],
        kTables.relEnergy[iu + 1], x, u);
}

__device__ __forceinline__ float thomsonWiddingtonRange(float T0) { return 0.0119f * powf(T0, 1.513f); }

__device__ __forceinline__ float thomsonWiddingtonLaw(float x, float tubeVoltage)
{
    const int i = intervalStart(kTables.twVoltage, tubeVoltage);
    const float t1 = kTables.twVoltage[i], t2 = kTables.twVoltage[i + 1];
    const float c1 = kTables.twConstant[i], c2 = kTables.twConstant[i + 1];
    const float C = c1 + ((c2 - c1) / (t2 - t1)) * (tubeVoltage - t1);
    const float twl = (tubeVoltage * tubeVoltage - C * x) / (tubeVoltage * tubeVoltage);
    return twl < 0.0f ? 0.0f : twl;
}

__device__ __forceinline__ float numberFractionF(float x, float tubeVoltage) { return powf(thomsonWiddingtonLaw(x, tubeVoltage), 1.753f); }

__device__ __forceinline__ float numberFractionM(float x, float tubeVoltage)
{
    constexpr float K = 18.0f, Bd = 0.584f, B0 = 0.5f;
    const float grown = 1.0f - expf(-K * x / thomsonWiddingtonRange(tubeVoltage));
    const float F = Bd * grown;
    const float B = B0 + (Bd - B0) * grown;
    return numberFractionF(x, tubeVoltage) * B * (F + 1.0f) / (1.0f - B * F);
}

__device__ __forceinline__ float electronDensity(float u, float x, float tubeVoltage)
{
    const float f = thomsonWiddingtonRange(100.0f) / thomsonWiddingtonRange(tubeVoltage);
    return numberFractionF(x, tubeVoltage) * tableDensity(kTables.depthF, kTables.densityF, u, x * f)
        + numberFractionM(x, tubeVoltage) * tableDensity(kTables.depthM, kTables.densityM, u, x * f);
}

__device__ __forceinline__ float betheHeitlerCrossSection(float hv, float Ti)
{
    constexpr float phiBar = (74.0f * 74.0f) * 2.81794092E-15f * 2.81794092E-15f * 7.29735308E-03f;
    constexpr float scale = (phiBar * 2.0f) / 3.0f;
    constexpr float m = kElectronMass;
    const float Ei = m + Ti;
    const float Ef = Ei - hv;
    const float pi2 = Ei * Ei - m * m;
    const float pi = sqrtf(pi2);
    const float pf2 = Ef * Ef - m * m;
    if (pf2 <= 0.0f)
        return 0.0f;
    const float pf = sqrtf(pf2);
    const float L = 2.0f * logf((Ei * Ef + pi * pf - m * m) / (m * hv));
    const float coulomb = pi / pf;
    return scale * (4.0f * Ei * Ef * L - 7.0f * pi * pf) / (hv * pi * pi) * coulomb;
}

// out[angle][bin] = betheHeitlerSpectra(T0, energies[bin], angles[angle]); one block per (angle, bin)
__global__ void __launch_bounds__(kMaxDepthSteps) bremsstrahlungKernel(float T0, unsigned nBins, const float* __restrict__ energies,
    const float* __restrict__ tungstenAtt, const float* __restrict__ angles, float* __restrict__ out)
{
    __shared__ float depth[kMaxDepthSteps], atDepth[kMaxDepthSteps];
    __shared__ int nDepths;
    const unsigned bin = blockIdx.x % nBins, angle = blockIdx.x / nBins;
    const float hv = energies[bin];
    if (threadIdx.x == 0) { // the depths as the reference's float loop produces them: x = 0; x <= 14; x = x + 0.1
        int n = 0;
        for (float x = 0.0f; x <= 14.0f && n < kMaxDepthSteps; x = x + 0.1f)
            depth[n++] = x;
        nDepths = n;
    }
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < nDepths && hv > 0.0f) {
        const float x = depth[threadIdx.x];
        float sum = 0.0f;
        for (float u = 0.005f; u <= 1.0f; u = u + 0.005f)
            sum = sum + betheHeitlerCrossSection(hv, T0 * u) * electronDensity(u, x, T0) * 0.005f;
        atDepth[threadIdx.x] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.0f;
        if (hv > 0.0f) {
            const float att = tungstenAtt[bin], sinAngle = sinf(angles[angle]);
            for (int k = 0; k < nDepths; ++k)
                total = total + atDepth[k] * expf(-att * depth[k] * 0.001f / sinAngle) * 0.1f; // anode self-absorption, mg/cm2 -> g/cm2
        }
        out[angle * nBins + bin] = total;
    }
}

inline cudaError_t uploadTables(cudaStream_t stream)
{
    static const ElectronTables host = [] { // thread-safe one-time conversion of the double tables
        ElectronTables t {};
        for (int i = 0; i < tw::kDepths; ++i) {
            t.depthF[i] = static_cast<float>(tw::depthF[i]);
            t.depthM[i] = static_cast<float>(tw::depthM[i]);
            for (int j = 0; j < tw::kEnergies; ++j) {
                t.densityF[i][j] = static_cast<float>(tw::densityF[i][j]);
                t.densityM[i][j] = static_cast<float>(tw::densityM[i][j]);
            }
        }
        for (int j = 0; j < tw::kEnergies; ++j)
            t.relEnergy[j] = static_cast<float>(tw::relEnergy[j]);
        for (int i = 0; i < 5; ++i) {
            t.twVoltage[i] = static_cast<float>(tw::twVoltage[i]);
            t.twConstant[i] = static_cast<float>(tw::twConstant[i]);
        }
        return t;
    }();
    // per device and cheap (4.5 KB): done at every call, on the stream the kernel is launched on (ordered before it)
    return cudaMemcpyToSymbolAsync(kTables, &host, sizeof(host), 0, cudaMemcpyHostToDevice, stream);
}

} // namespace dxmcb200::spectrum
