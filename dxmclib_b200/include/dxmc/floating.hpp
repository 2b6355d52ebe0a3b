// floating.hpp — forwarding header: the Floating concept lives in dxmc/types.hpp.
#pragma once
#include "dxmc/types.hpp"
