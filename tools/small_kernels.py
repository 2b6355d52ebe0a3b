"""Runs the once-per-call streaming kernels at bench size (512x512x400) so that ncu can capture them: voxel-grid packing
(palette and record form), the per-material density maxima, and the result decode (EV_PER_HISTORY and DOSE modes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dxmclib_b200 import cabi, phantoms

dim, sp = (512, 512, 400), (1.0, 1.0, 1.0)
mat, dens = phantoms.anthropomorphic(dim, sp)
half = [d * s / 2 for d, s in zip(dim, sp)]
ext = (-half[0], half[0], -half[1], half[1], -half[2], half[2])
for palette in ("1", "0"):
    os.environ["DXMCB200_PALETTE"] = palette
    ctx = cabi.Context(0)
    ctx.set_world(dim, sp, ext, dens, mat)
    print("palette", palette, "max density per material", ctx.material_max_density(10))
    ctx.set_fixed_point(20, 10)
    n = int(np.prod(dim))
    dose, ev, var = np.empty(n, np.float32), np.empty(n, np.uint32), np.empty(n, np.float32)
    import ctypes as C
    f32, u32 = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for mode in (0, 1):
        rc = ctx.l.dxmcb200_get_result(ctx.h, mode, C.c_uint64(10**10), C.c_float(1.0), dose.ctypes.data_as(f32), ev.ctypes.data_as(u32),
                                       var.ctypes.data_as(f32))
        assert rc == 0
    ctx.close()
