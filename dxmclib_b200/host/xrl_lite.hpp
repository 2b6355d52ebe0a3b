// xrl_lite — a small, self-contained photon-interaction data source.
//
// DXMClib takes every cross section, form factor, scatter function, edge energy,
// fluorescence datum and compound definition from the third-party C library
// xraylib (reference call sites: src/material.cpp:33-59, 137, 164-199, 243, 278,
// 342-375; include/dxmc/betheHeitlerCrossSection.hpp:371). xraylib is NOT
// vendored by the reference and is absent from this build image, so this file
// provides the subset of the xraylib-4 C API the hot path's host builders need,
// under the same function names and argument conventions (trailing
// `xrl_error**`), backed by compact analytic models and short anchor tables
// (see xrl_lite.cpp and DESIGN.md "physics data"). The numbers are
// APPROXIMATE (few-% level on mu/rho for Z<=20 between 5 and 150 keV); both the
// reference oracle and the B200 product link this same source, so
// GPU-vs-reference parity does not depend on its absolute accuracy.
//
// If a real xraylib is available, compile material.cpp with
// -DDXMCB200_USE_XRAYLIB and link libxrl instead; nothing else changes.
#pragma once

#include <cstddef>

namespace xrl_lite {

struct xrl_error {
    int code;
    const char* message;
};

// shell macros (xraylib numbering)
enum : int {
    K_SHELL = 0,
    L1_SHELL = 1,
    L2_SHELL = 2,
    L3_SHELL = 3,
    M1_SHELL = 4,
    M2_SHELL = 5,
    M3_SHELL = 6,
    M4_SHELL = 7,
    M5_SHELL = 8,
    N1_SHELL = 9,
    N2_SHELL = 10,
    N3_SHELL = 11,
    N4_SHELL = 12,
    N5_SHELL = 13,
    N6_SHELL = 14,
    N7_SHELL = 15,
    O1_SHELL = 16,
    O2_SHELL = 17,
    O3_SHELL = 18,
    O4_SHELL = 19,
    O5_SHELL = 20,
    O6_SHELL = 21,
    O7_SHELL = 22,
    P1_SHELL = 23,
    P2_SHELL = 24,
    P3_SHELL = 25,
    N_SHELLS = 26
};

// Coster-Kronig transition macros
enum : int {
    FL12_TRANS = 1,
    FL13_TRANS = 2,
    FL23_TRANS = 4
};

struct compoundData {
    int nElements;
    double nAtomsAll;
    int* Elements;
    double* massFractions;
    double* nAtoms;
    double molarMass;
};

struct compoundDataNIST {
    char* name;
    int nElements;
    int* Elements;
    double* massFractions;
    double density;
};

// element data
double AtomicWeight(int Z, xrl_error** error);
double ElementDensity(int Z, xrl_error** error);
char* AtomicNumberToSymbol(int Z, xrl_error** error); // caller frees with xrlFree
int SymbolToAtomicNumber(const char* symbol, xrl_error** error);
void xrlFree(void* p);

// cross sections, cm2/g, energy in keV
double CS_Total(int Z, double E, xrl_error** error);
double CS_Photo(int Z, double E, xrl_error** error);
double CS_Rayl(int Z, double E, xrl_error** error);
double CS_Compt(int Z, double E, xrl_error** error);
double CS_Energy(int Z, double E, xrl_error** error);
double CS_Total_CP(const char* compound, double E, xrl_error** error);
double CS_Photo_CP(const char* compound, double E, xrl_error** error);
double CS_Rayl_CP(const char* compound, double E, xrl_error** error);
double CS_Compt_CP(const char* compound, double E, xrl_error** error);
double CS_Energy_CP(const char* compound, double E, xrl_error** error);

// atomic form factor and incoherent scattering function, q = E[keV]/12.398 * sin(theta/2) in 1/Angstrom
double FF_Rayl(int Z, double q, xrl_error** error);
double SF_Compt(int Z, double q, xrl_error** error);

// shell data
double EdgeEnergy(int Z, int shell, xrl_error** error);
double ElectronConfig(int Z, int shell, xrl_error** error);
double ComptonProfile_Partial(int Z, int shell, double pz, xrl_error** error);
double FluorYield(int Z, int shell, xrl_error** error);
double CosKronTransProb(int Z, int trans, xrl_error** error);
double RadRate(int Z, int line, xrl_error** error);
double LineEnergy(int Z, int line, xrl_error** error);
double CSb_Photo_Partial(int Z, int shell, double E, xrl_error** error); // barn/atom

// compounds
compoundData* CompoundParser(const char* compound, xrl_error** error);
void FreeCompoundData(compoundData* cd);
compoundDataNIST* GetCompoundDataNISTByName(const char* name, xrl_error** error);
void FreeCompoundDataNIST(compoundDataNIST* cd);
char** GetCompoundDataNISTList(int* nCompounds, xrl_error** error); // caller owns (xrlFree each + array)

// true when this element has data in xrl_lite
bool elementSupported(int Z);

} // namespace xrl_lite
