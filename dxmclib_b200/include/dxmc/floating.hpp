// floating.hpp — the Floating concept every dxmc template is constrained on
// (API of reference include/dxmc/floating.hpp:24-25).
#pragma once
#include <concepts>

namespace dxmc {
template <typename T>
concept Floating = std::floating_point<T>;
}
