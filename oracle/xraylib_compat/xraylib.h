// xraylib.h stand-in — TEST INFRASTRUCTURE.
// The reference's src/material.cpp includes "xraylib.h" (material.cpp:23); xraylib is not
// vendored by the reference and not installed in this image. This header puts the
// xraylib-4 subset implemented by dxmclib_b200/host/xrl_lite into the global namespace so
// that the reference's own material.cpp compiles UNMODIFIED from /root/reference/src.
#pragma once
#include "xrl_lite.hpp"
using namespace xrl_lite;
