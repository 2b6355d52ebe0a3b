// particle.hpp — photon state (API of reference include/dxmc/particle.hpp:30-47).
#pragma once
#include "dxmc/floating.hpp"
#include <array>

namespace dxmc {
template <Floating T = double>
struct Particle {
    std::array<T, 3> pos; // mm
    std::array<T, 3> dir; // treated as a unit vector
    T energy; // keV
    T weight;
};
}
