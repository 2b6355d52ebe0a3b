// constants.hpp — physical constants with the reference's names and values
// (API of reference include/dxmc/constants.hpp:26-65).
#pragma once
#include "dxmc/floating.hpp"
#include <numbers>

namespace dxmc {
// clang-format off
template <Floating T> consteval T KEV_TO_ANGSTROM()    { return T { 12.398520 }; }
template <Floating T> consteval T PI_VAL()             { return std::numbers::pi_v<T>; }
template <Floating T> consteval T DEG_TO_RAD()         { return PI_VAL<T>() / T { 180 }; }
template <Floating T> consteval T RAD_TO_DEG()         { return T { 180 } / PI_VAL<T>(); }
template <Floating T> consteval T KEV_TO_MJ()          { return T { 1.6021773e-13 }; }
template <Floating T> consteval T MJ_TO_KEV()          { return T { 1 } / KEV_TO_MJ<T>(); }
template <Floating T> consteval T ELECTRON_REST_MASS() { return T { 510.9989461 }; }
// clang-format on
}
