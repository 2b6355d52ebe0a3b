// constants.hpp — physical constants with the reference's names and values
// (API of reference include/dxmc/constants.hpp:26-65). Each is a consteval function template so that float and
// double users get the literal rounded once to their own precision; derived constants are computed in T.
#pragma once
#include "dxmc/types.hpp"
#include <numbers>

#define DXMC_CONSTANT(NAME, ...)         \
    template <::dxmc::Floating T>        \
    consteval T NAME()                   \
    {                                    \
        return static_cast<T>(__VA_ARGS__); \
    }

namespace dxmc {
DXMC_CONSTANT(PI_VAL, std::numbers::pi_v<T>)
DXMC_CONSTANT(DEG_TO_RAD, PI_VAL<T>() / T { 180 })
DXMC_CONSTANT(RAD_TO_DEG, T { 180 } / PI_VAL<T>())
DXMC_CONSTANT(ELECTRON_REST_MASS, 510.9989461) // keV
DXMC_CONSTANT(KEV_TO_ANGSTROM, 12.398520) // hc in keV Angstrom
DXMC_CONSTANT(KEV_TO_MJ, 1.6021773e-13)
DXMC_CONSTANT(MJ_TO_KEV, T { 1 } / KEV_TO_MJ<T>())
}
#undef DXMC_CONSTANT
