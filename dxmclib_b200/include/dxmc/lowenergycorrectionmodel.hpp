// lowenergycorrectionmodel.hpp — binding-effect model selector
// (API of reference include/dxmc/lowenergycorrectionmodel.hpp:20-26).
#pragma once
namespace dxmc {
enum class LOWENERGYCORRECTION : int {
    NONE = 0, // free-electron Klein-Nishina, Thomson Rayleigh
    LIVERMORE = 1, // scatter-function corrected Compton, form-factor Rayleigh
    IA = 2 // impulse approximation with Doppler broadening and fluorescence
};
}
