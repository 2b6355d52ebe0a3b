// matdb — material-level physics queries on top of the element data source.
//
// This is the data seam of the hot path: the 12 `dxmc::Material` methods the reference
// implements on xraylib (reference src/material.cpp:31-377, declared in
// include/dxmc/material.hpp:62-104). Everything here is host-only and runs once per
// run; the results are flattened into device tables by the LUT builders.
//
// The functions are free functions on plain data so that both the product's
// `dxmc::Material` (dxmclib_b200/include/dxmc/material.hpp) and the oracle's shim around
// the unmodified reference `Material` class (oracle/ref_material_shim.cpp) are thin
// wrappers over the same numbers.
#pragma once

#include <array>
#include <string>
#include <vector>

namespace dxmcb200::matdb {

struct Composition {
    std::string name; // canonical name handed to the cross-section functions
    std::vector<int> elements;
    std::vector<double> numberFraction; // normalised to 1
    double density = -1.0;
    bool valid = false;
    bool hasDensity = false;
};

// the element data behind every function below: "xraylib", or the in-repo approximate xrl_lite
const char* backendName();
bool backendIsApproximate();

// NIST compound name first, chemical formula second (reference material.cpp:83-97, 340-377)
Composition compositionFromString(const std::string& nameOrFormula);
// single element (reference material.cpp:324-338)
Composition compositionFromAtomicNumber(int Z);

// mass attenuation coefficients in cm2/g, energy in keV (reference material.cpp:31-59)
double photoelectric(const std::string& name, double energy);
double rayleigh(const std::string& name, double energy);
double compton(const std::string& name, double energy);
double total(const std::string& name, double energy);
double totalElement(int Z, double energy);
double massEnergyAbsorption(const std::string& name, double energy);

double atomicWeight(int Z);
std::string symbol(int Z);
int atomicNumber(const std::string& symbol);
std::vector<std::string> nistCompoundNames();

// sum_i n_i F_i(q)^2 (reference material.cpp:133-141)
double formFactorSquared(const Composition& c, double momentumTransfer);
// sum_i n_i S_i(q)/Z_i (reference material.cpp:274-283)
double normalizedScatterFactor(const Composition& c, double momentumTransfer);

// all shell edges above minValue, descending (reference material.cpp:285-322)
std::vector<double> bindingEnergies(const std::string& name, double minValue);

struct Shell {
    double bindingEnergy = 0;
    double numberElectrons = 0;
    double hartreeFockOrbital_0 = 0;
    double photoIonizationProbability = 1;
    double fluorescenceYield = 0;
    std::array<double, 3> fluorLineProbabilities = { 1, 1, 1 };
    std::array<double, 3> fluorLineEnergies = { 0, 0, 0 };
    int Z = 0;
    int shell = 0;
};

// the 12 most tightly bound shells of the material (reference material.cpp:143-273)
std::array<Shell, 12> electronConfiguration(const std::string& name);

} // namespace dxmcb200::matdb
