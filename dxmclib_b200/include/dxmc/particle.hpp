// particle.hpp — forwarding header: Particle<T> lives in dxmc/types.hpp.
#pragma once
#include "dxmc/types.hpp"
