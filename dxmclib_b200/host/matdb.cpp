// matdb.cpp — see matdb.hpp. Behaviour follows reference src/material.cpp (cited per function);
// written against the xraylib-4 C API so that a real xraylib can be swapped in with
// -DDXMCB200_USE_XRAYLIB (then link -lxrl); otherwise the in-repo xrl_lite is used.
#include "matdb.hpp"

#ifdef DXMCB200_USE_XRAYLIB
#include "xraylib.h"
#else
#include "xrl_lite.hpp"
using namespace xrl_lite;
#endif

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <execution>
#include <mutex>
#include <numeric>

namespace dxmcb200::matdb {

bool backendIsApproximate()
{
#ifdef DXMCB200_USE_XRAYLIB
    return false;
#else
    return true;
#endif
}

const char* backendName()
{
#ifdef DXMCB200_USE_XRAYLIB
    return "xraylib";
#else
    return "xrl_lite (approximate: analytic models + anchor tables, few % on mu/rho for Z <= 20 and 5-150 keV; H-Ca, I, W and a few "
           "more elements; build with a real xraylib for dosimetry: dxmclib_b200/build.py, DXMCB200_XRAYLIB)";
#endif
}

namespace {

    // A dosimetry library that silently computes with approximate cross sections is a trap: say so once per process, on the
    // first material that is looked up (DXMCB200_QUIET=1 silences it; dxmcb200_physics_backend() reports the same).
    void announceBackendOnce()
    {
        if (!backendIsApproximate())
            return;
        static std::once_flag once;
        std::call_once(once, [] {
            const char* quiet = std::getenv("DXMCB200_QUIET");
            if (quiet && quiet[0] == '1')
                return;
            std::fprintf(stderr, "[dxmcb200] physics data backend: %s\n", backendName());
        });
    }

    // mass fractions -> normalised number fractions (reference material.cpp:150-154, 347-353)
    void fillNumberFractions(Composition& c, int n, const int* Z, const double* massFractions)
    {
        c.elements.assign(Z, Z + n);
        c.numberFraction.resize(n);
        for (int i = 0; i < n; ++i)
            c.numberFraction[i] = massFractions[i] / AtomicWeight(Z[i], nullptr);
        // std::reduce, like the reference: its pairwise-by-four association decides the last bit of the sum
        const double sum = std::reduce(c.numberFraction.cbegin(), c.numberFraction.cend());
        for (auto& f : c.numberFraction)
            f /= sum;
    }

    struct ElementList {
        std::vector<int> Z;
        std::vector<double> massFraction;
    };

    // formula first, NIST name second: the order the reference uses when it needs the element list
    // of an existing material (material.cpp:259-272, 306-321 — the NIST result wins if both parse)
    bool elementList(const std::string& name, ElementList& out, bool nistOverrides)
    {
        bool found = false;
        if (compoundData* cd = CompoundParser(name.c_str(), nullptr)) {
            out.Z.assign(cd->Elements, cd->Elements + cd->nElements);
            out.massFraction.assign(cd->massFractions, cd->massFractions + cd->nElements);
            FreeCompoundData(cd);
            found = true;
            if (!nistOverrides)
                return true;
        }
        if (compoundDataNIST* cd = GetCompoundDataNISTByName(name.c_str(), nullptr)) {
            out.Z.assign(cd->Elements, cd->Elements + cd->nElements);
            out.massFraction.assign(cd->massFractions, cd->massFractions + cd->nElements);
            FreeCompoundDataNIST(cd);
            found = true;
        }
        return found;
    }

    // effective fluorescence yield of the grouped shell incl. Coster-Kronig feeding
    // (reference material.cpp:175-192)
    double groupedYield(int Z, int shell)
    {
        const auto w = [&](int s) { return FluorYield(Z, s, nullptr); };
        const auto f = [&](int t) { return CosKronTransProb(Z, t, nullptr); };
        switch (shell) {
        case 0:
            return w(K_SHELL);
        case 1:
            return w(L1_SHELL) + w(L2_SHELL) * f(FL12_TRANS)
                + (f(FL13_TRANS) + f(FL12_TRANS) * f(FL23_TRANS)) * w(L3_SHELL);
        case 2:
            return w(L2_SHELL) + w(L3_SHELL) * f(FL23_TRANS);
        case 3:
            return w(L3_SHELL);
        default:
            return 0.0;
        }
    }

    // first/last xraylib line index of the transitions filling a shell (material.cpp:172-191)
    void lineRange(int shell, int& first, int& last)
    {
        static const int ranges[4][2] = { { -1, -29 }, { -30, -58 }, { -59, -85 }, { -86, -113 } };
        if (shell >= 0 && shell < 4) {
            first = ranges[shell][0];
            last = ranges[shell][1];
        } else {
            first = 0;
            last = 0;
        }
    }

} // namespace

Composition compositionFromString(const std::string& s)
{
    announceBackendOnce();
    Composition c;
    if (compoundDataNIST* n = GetCompoundDataNISTByName(s.c_str(), nullptr)) {
        c.name = n->name;
        c.valid = true;
        c.hasDensity = true;
        c.density = n->density;
        fillNumberFractions(c, n->nElements, n->Elements, n->massFractions);
        FreeCompoundDataNIST(n);
        return c;
    }
    if (compoundData* m = CompoundParser(s.c_str(), nullptr)) {
        c.name = s;
        c.valid = true;
        c.hasDensity = false;
        fillNumberFractions(c, m->nElements, m->Elements, m->massFractions);
        FreeCompoundData(m);
    }
    return c;
}

Composition compositionFromAtomicNumber(int Z)
{
    announceBackendOnce();
    Composition c;
    if (char* sym = AtomicNumberToSymbol(Z, nullptr)) {
        c.name = sym;
        xrlFree(sym);
        c.density = ElementDensity(Z, nullptr);
        c.hasDensity = true;
        c.valid = true;
        c.elements = { Z };
        c.numberFraction = { 1.0 };
    }
    return c;
}

double photoelectric(const std::string& name, double e) { return CS_Photo_CP(name.c_str(), e, nullptr); }
double rayleigh(const std::string& name, double e) { return CS_Rayl_CP(name.c_str(), e, nullptr); }
double compton(const std::string& name, double e) { return CS_Compt_CP(name.c_str(), e, nullptr); }
double total(const std::string& name, double e) { return CS_Total_CP(name.c_str(), e, nullptr); }
double totalElement(int Z, double e) { return CS_Total(Z, e, nullptr); }
double massEnergyAbsorption(const std::string& name, double e) { return CS_Energy_CP(name.c_str(), e, nullptr); }
double atomicWeight(int Z) { return AtomicWeight(Z, nullptr); }

std::string symbol(int Z)
{
    std::string s;
    if (char* c = AtomicNumberToSymbol(Z, nullptr)) {
        s = c;
        xrlFree(c);
    }
    return s;
}

int atomicNumber(const std::string& sym) { return SymbolToAtomicNumber(sym.c_str(), nullptr); }

std::vector<std::string> nistCompoundNames()
{
    int n = 0;
    char** list = GetCompoundDataNISTList(&n, nullptr);
    std::vector<std::string> names;
    for (int i = 0; i < n; ++i)
        if (list[i]) {
            names.emplace_back(list[i]);
            xrlFree(list[i]);
        }
    xrlFree(list);
    return names;
}

// the two sums below go through the same library reductions the reference calls (with and without an execution
// policy), so that the association order of the double additions — and with it the last bit — is the same
double formFactorSquared(const Composition& c, double q)
{
    return std::transform_reduce(std::execution::par_unseq, c.elements.cbegin(), c.elements.cend(), c.numberFraction.cbegin(), 0.0, std::plus<>(),
        [=](const int Z, const double nf) {
            const double f = FF_Rayl(Z, q, nullptr);
            return nf * f * f;
        });
}

double normalizedScatterFactor(const Composition& c, double q)
{
    return std::transform_reduce(c.elements.cbegin(), c.elements.cend(), c.numberFraction.cbegin(), 0.0, std::plus<>(),
        [=](const int Z, const double nf) { return nf * SF_Compt(Z, q, nullptr) / Z; });
}

std::vector<double> bindingEnergies(const std::string& name, double minValue)
{
    std::vector<double> edges;
    ElementList el;
    if (!elementList(name, el, true))
        return edges;
    for (int Z : el.Z) {
        for (int shell = 0;; ++shell) {
            xrl_error* err = nullptr;
            const double edge = EdgeEnergy(Z, shell, &err);
            if (err)
                break;
            if (edge > minValue)
                edges.push_back(edge);
        }
    }
    std::sort(edges.begin(), edges.end(), std::greater<double>());
    return edges;
}

std::array<Shell, 12> electronConfiguration(const std::string& name)
{
    std::array<Shell, 12> result;
    ElementList el;
    if (!elementList(name, el, false))
        return result;

    std::vector<double> numberFraction(el.Z.size());
    double norm = 0;
    for (std::size_t i = 0; i < el.Z.size(); ++i) {
        numberFraction[i] = el.massFraction[i] / AtomicWeight(el.Z[i], nullptr);
        norm += numberFraction[i];
    }

    std::vector<Shell> all;
    for (std::size_t i = 0; i < el.Z.size(); ++i) {
        const int Z = el.Z[i];
        const double nf = numberFraction[i] / norm;
        xrl_error* err = nullptr;
        for (int shell = 0; !err; ++shell) {
            const double binding = EdgeEnergy(Z, shell, &err);
            if (err)
                break;
            Shell s;
            s.bindingEnergy = binding;
            s.numberElectrons = nf * ElectronConfig(Z, shell, nullptr);
            // a missing profile ends the shell scan after this entry (the reference passes the same error slot)
            s.hartreeFockOrbital_0 = ComptonProfile_Partial(Z, shell, 0.0, &err);
            s.photoIonizationProbability = 0.0;
            s.fluorescenceYield = groupedYield(Z, shell);
            s.Z = Z;
            s.shell = shell;

            // keep the three strongest lines: every candidate replaces the currently weakest slot
            std::array<double, 3> prob = { 0, 0, 0 };
            std::array<double, 3> energy = { 0, 0, 0 };
            int first, last;
            lineRange(shell, first, last);
            for (int line = first; line >= last && first != 0; --line) {
                // index of smallest |value|, ties to the lower index
                const double a = std::abs(prob[0]), b = std::abs(prob[1]), c = std::abs(prob[2]);
                const int weakest = a <= b ? (a <= c ? 0 : 2) : (b <= c ? 1 : 2);
                xrl_error* lineErr = nullptr;
                const double rate = RadRate(Z, line, &lineErr);
                if (rate > prob[weakest] && !lineErr) {
                    prob[weakest] = rate;
                    energy[weakest] = LineEnergy(Z, line, nullptr);
                }
            }
            const double probSum = prob[0] + prob[1] + prob[2];
            if (probSum > 0)
                for (auto& p : prob)
                    p /= probSum;
            s.fluorLineProbabilities = prob;
            s.fluorLineEnergies = energy;
            all.push_back(s);
        }
    }

    std::sort(all.begin(), all.end(), [](const Shell& l, const Shell& r) { return l.bindingEnergy > r.bindingEnergy; });
    for (std::size_t i = 0; i < std::min(all.size(), result.size()); ++i)
        result[i] = all[i];

    const double electrons = std::transform_reduce(result.cbegin(), result.cend(), 0.0, std::plus<>(), [](const Shell& s) { return s.numberElectrons; });
    for (auto& s : result)
        s.numberElectrons /= electrons;

    // photo-ionisation probability: occupancy-weighted sum of the partial cross section on
    // 0.5, 1.5, ... 399.5 keV; unfilled slots keep their default weight of 1 in the normalisation
    // (the sums use the same library reductions as the reference so the doubles agree to the last bit)
    std::vector<double> energy(400);
    std::iota(energy.begin(), energy.end(), 0.5);
    for (auto& s : result) {
        if (s.Z > 0) {
            const int Z = s.Z, shell = s.shell;
            const double sum = std::transform_reduce(std::execution::par_unseq, energy.cbegin(), energy.cend(), 0.0, std::plus<>(),
                [=](const double e) { return CSb_Photo_Partial(Z, shell, e, nullptr); });
            s.photoIonizationProbability = s.numberElectrons * sum;
        }
    }
    const double total = std::transform_reduce(std::execution::par_unseq, result.cbegin(), result.cend(), 0.0, std::plus<>(),
        [](const Shell& s) { return s.photoIonizationProbability; });
    if (total > 0)
        for (auto& s : result)
            s.photoIonizationProbability /= total;
    return result;
}

} // namespace dxmcb200::matdb

// C ABI (include/dxmcb200.h): which element data the host-side table builders of this library were compiled against
extern "C" const char* dxmcb200_physics_backend(int* approximate)
{
    if (approximate)
        *approximate = dxmcb200::matdb::backendIsApproximate() ? 1 : 0;
    return dxmcb200::matdb::backendName();
}
