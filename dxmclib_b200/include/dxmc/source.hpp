// source.hpp — x-ray sources: geometry + spectrum + dose calibration, producing Exposures.
//
// Public surface of the reference's source hierarchy (include/dxmc/source.hpp:46-1702):
//   Source                       abstract base                                       :46-207
//   PencilSource                 mono-energetic pencil beam                          :209-266
//   IsotropicSource / IsotropicCTSource   user spectrum, rectangular collimation     :268-372
//   DAPSource -> DXSource, CBCTSource     tube-based radiography / cone-beam         :374-787
//   CTBaseSource -> CTSource -> CTAxialSource, CTSpiralSource                        :789-1363
//                 CTDualSource -> CTAxialDualSource, CTSpiralDualSource              :1075-1608
//   CTTopogramSource                                                                 :1610-1700
// getExposure(i) is O(1) host code; Transport evaluates it for every exposure up front and ships
// the resulting table to the GPU. CT dose calibration (ctCalibration) runs a second Transport on
// a CTDIPhantom exactly like the reference, i.e. a second pass through the same CUDA path.
//
// The classes live in three headers by family; this one is what user code includes.
#pragma once
#include "dxmc/sourcebase.hpp" // Source, PencilSource, IsotropicSource, IsotropicCTSource
#include "dxmc/dapsource.hpp" // DAPSource, DXSource, CBCTSource
#include "dxmc/ctsource.hpp" // CTBaseSource, CTSource, CTAxial/SpiralSource, CTDualSource, CTAxial/SpiralDualSource, CTTopogramSource
