// lowenergycorrectionmodel.hpp — forwarding header: LOWENERGYCORRECTION lives in dxmc/types.hpp.
#pragma once
#include "dxmc/types.hpp"
