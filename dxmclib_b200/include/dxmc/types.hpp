// types.hpp — the three small vocabulary types of the dxmc API, kept in one place: the Floating concept every
// template is constrained on, the low-energy correction selector and the photon state. The reference spreads them
// over floating.hpp, lowenergycorrectionmodel.hpp and particle.hpp; those names are kept as forwarding headers.
#pragma once
#include <array>
#include <concepts>
#include <cstddef>

namespace dxmc {

template <typename T>
concept Floating = std::floating_point<T>;

// Binding-effect model of the interaction samplers (values are part of the API: they index the kernel variants).
//   NONE       free-electron Klein-Nishina Compton, Thomson Rayleigh, local photoelectric absorption
//   LIVERMORE  Compton weighted by the incoherent scatter function, form-factor Rayleigh
//   IA         impulse approximation with Doppler broadening, shell-wise photoelectric absorption with fluorescence
enum class LOWENERGYCORRECTION : int { NONE = 0, LIVERMORE = 1, IA = 2 };

// Photon state: position [mm], direction (treated as a unit vector, never renormalised), energy [keV], weight.
template <Floating T = double>
struct Particle {
    std::array<T, 3> pos, dir;
    T energy, weight;

    // move `distance` [mm] along the direction, component by component in the order the transport loop uses
    constexpr void translate(T distance) noexcept
    {
        for (std::size_t axis = 0; axis < 3; ++axis)
            pos[axis] += dir[axis] * distance;
    }
    // what Russian roulette looks at (transport.hpp:684): energy carried, in keV
    constexpr T weightedEnergy() const noexcept { return energy * weight; }
};

}
