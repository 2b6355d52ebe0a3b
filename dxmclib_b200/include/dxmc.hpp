// dxmc.hpp — umbrella header of the B200 drop-in for DXMClib's transport path
// (same include set as reference include/dxmc.hpp:19-25).
#pragma once
#include "dxmc/beamfilters.hpp"
#include "dxmc/lowenergycorrectionmodel.hpp"
#include "dxmc/material.hpp"
#include "dxmc/source.hpp"
#include "dxmc/transport.hpp"
#include "dxmc/vectormath.hpp"
#include "dxmc/world.hpp"
