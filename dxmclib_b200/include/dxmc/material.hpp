// material.hpp — Material: composition + photon interaction data of one medium.
//
// Same public surface as the reference class (include/dxmc/material.hpp:60-139, implemented on
// xraylib in src/material.cpp). Here the class is a value wrapper around
// dxmcb200::matdb (dxmclib_b200/host/matdb.hpp), which sits on xraylib when available and on the
// in-repo xrl_lite data otherwise. Host-only: used to BUILD the device tables, never read in the
// transport kernels.
#pragma once
#include "dxmc/floating.hpp"
#include "matdb.hpp"

#include <algorithm>
#include <array>
#include <string>
#include <vector>

namespace dxmc {

template <Floating T>
struct ElectronShellConfiguration {
    T bindingEnergy = 0;
    T numberElectrons = 0;
    T hartreeFockOrbital_0 = 0;
    T photoIonizationProbability = 1;
    T fluorescenceYield = 0;
    std::array<T, 3> fluorLineProbabilities = { 1, 1, 1 };
    std::array<T, 3> fluorLineEnergies = { 0, 0, 0 };
    int Z = 0;
    int shell = 0;

    template <Floating U>
    ElectronShellConfiguration<U> cast() const
    {
        ElectronShellConfiguration<U> c;
        c.bindingEnergy = static_cast<U>(bindingEnergy);
        c.numberElectrons = static_cast<U>(numberElectrons);
        c.hartreeFockOrbital_0 = static_cast<U>(hartreeFockOrbital_0);
        c.photoIonizationProbability = static_cast<U>(photoIonizationProbability);
        c.fluorescenceYield = static_cast<U>(fluorescenceYield);
        for (std::size_t i = 0; i < 3; ++i) {
            c.fluorLineProbabilities[i] = static_cast<U>(fluorLineProbabilities[i]);
            c.fluorLineEnergies[i] = static_cast<U>(fluorLineEnergies[i]);
        }
        c.Z = Z;
        c.shell = shell;
        return c;
    }
};

class Material {
    using Db = dxmcb200::matdb::Composition;

public:
    // ---- element look-ups that need no Material
    static int getAtomicNumberFromSymbol(const std::string& symbol) { return dxmcb200::matdb::atomicNumber(symbol); }
    static std::string getSymbolFromAtomicNumber(int Z) { return dxmcb200::matdb::symbol(Z); }
    static std::string getAtomicNumberToSymbol(int Z) { return dxmcb200::matdb::symbol(Z); }
    static double getAtomicWeight(int Z) { return dxmcb200::matdb::atomicWeight(Z); }
    static double getTotalAttenuation(int atomicNumber, double energy) { return dxmcb200::matdb::totalElement(atomicNumber, energy); }
    static std::vector<std::string> getNISTCompoundNames() { return dxmcb200::matdb::nistCompoundNames(); }

    // ---- mass attenuation coefficients [cm2/g] at a photon energy [keV]
    double getTotalAttenuation(double energy) const { return dxmcb200::matdb::total(m_data.name, energy); }
    double getPhotoelectricAttenuation(double energy) const { return dxmcb200::matdb::photoelectric(m_data.name, energy); }
    double getComptonAttenuation(double energy) const { return dxmcb200::matdb::compton(m_data.name, energy); }
    double getRayleightAttenuation(double energy) const { return dxmcb200::matdb::rayleigh(m_data.name, energy); }
    double getMassEnergyAbsorbtion(double energy) const { return dxmcb200::matdb::massEnergyAbsorption(m_data.name, energy); }

    // ---- momentum-transfer functions behind the Rayleigh and Compton samplers
    double getRayleightFormFactorSquared(const double momentumTransfer) const { return dxmcb200::matdb::formFactorSquared(m_data, momentumTransfer); }
    double getComptonNormalizedScatterFactor(const double momentumTransfer) const { return dxmcb200::matdb::normalizedScatterFactor(m_data, momentumTransfer); }
    template <Floating T>
    T getRayleightFormFactorSquared(const T momentumTransfer) const { return static_cast<T>(getRayleightFormFactorSquared(static_cast<double>(momentumTransfer))); }
    template <Floating T>
    T getComptonNormalizedScatterFactor(const T momentumTransfer) const { return static_cast<T>(getComptonNormalizedScatterFactor(static_cast<double>(momentumTransfer))); }

    // ---- shell data: edges above minValue [keV]; the 12 innermost shells over all elements of the medium
    std::vector<double> getBindingEnergies(const double minValue = 1) const { return dxmcb200::matdb::bindingEnergies(m_data.name, minValue); }
    template <Floating T>
    std::vector<T> getBindingEnergies(const T minValue = 1) const
    {
        const auto edges = getBindingEnergies(static_cast<double>(minValue));
        return { edges.begin(), edges.end() };
    }
    std::array<ElectronShellConfiguration<double>, 12> getElectronConfiguration() const
    {
        std::array<ElectronShellConfiguration<double>, 12> out;
        const auto shells = dxmcb200::matdb::electronConfiguration(m_data.name);
        std::transform(shells.begin(), shells.end(), out.begin(), [](const auto& s) {
            return ElectronShellConfiguration<double> { s.bindingEnergy, s.numberElectrons, s.hartreeFockOrbital_0, s.photoIonizationProbability,
                s.fluorescenceYield, s.fluorLineProbabilities, s.fluorLineEnergies, s.Z, s.shell };
        });
        return out;
    }
    template <Floating T>
    std::array<ElectronShellConfiguration<T>, 12> getElectronConfiguration() const
    {
        std::array<ElectronShellConfiguration<T>, 12> out;
        const auto shells = getElectronConfiguration();
        std::transform(shells.begin(), shells.end(), out.begin(), [](const auto& s) { return s.template cast<T>(); });
        return out;
    }

    // ---- identity and density [g/cm3]
    const std::string& name() const { return m_data.name; }
    const std::string& prettyName() const { return m_prettyName.empty() ? m_data.name : m_prettyName; }
    bool isValid() const { return m_data.valid && m_data.hasDensity; }
    bool hasStandardDensity() const { return m_data.hasDensity; }
    double standardDensity() const { return m_data.density; }
    void setStandardDensity(double density)
    {
        if (density > 0.0) {
            m_data.density = density;
            m_data.hasDensity = true;
        }
    }

    // NIST compound name ("Water, Liquid") or chemical formula ("H2O", "C0.015N78.4O21.1Ar0.47"); or one element
    Material(const std::string& xraylibMaterialNameOrCompound = "", const std::string& prettyName = "", const double density = -1.0)
        : m_data(dxmcb200::matdb::compositionFromString(xraylibMaterialNameOrCompound))
        , m_prettyName(prettyName.empty() ? m_data.name : prettyName)
    {
        setStandardDensity(density);
    }
    Material(int atomicNumber)
        : m_data(dxmcb200::matdb::compositionFromAtomicNumber(atomicNumber))
    {
    }

private:
    Db m_data;
    std::string m_prettyName;
};
}
