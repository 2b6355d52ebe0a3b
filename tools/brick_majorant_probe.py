"""CPU experiment behind DESIGN.md section 8 item 1 (regional Woodcock majorants): runs the CPU restatement's plain tracking
and the brick-majorant tracking of tools/brick_majorant_probe.cpp on a half-resolution copy of the bench workload (CT spiral
over the anthropomorphic phantom, 256x256x200 voxels of 2 mm) and reports look-ups per history, brick-face steps per history,
and how the two dose grids compare.   python tools/brick_majorant_probe.py [histories per exposure] [brick edge in voxels ...]"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
LIB = os.path.join(ROOT, "tools", "libbrickprobe.so")
subprocess.check_call(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-ffp-contract=off", f"-I{ROOT}/include",
                       os.path.join(ROOT, "tools", "brick_majorant_probe.cpp"), "-o", LIB])

from oracle import pyoracle  # noqa: E402

pyoracle.ORACLE_LIB = LIB  # the probe library contains the unmodified restatement plus brick_probe_run
import support as T  # noqa: E402
from dxmclib_b200 import cabi, phantoms  # noqa: E402
from dxmclib_b200 import scene as S  # noqa: E402

hist = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bricks = [int(x) for x in sys.argv[2:]] or [8, 16]
dim, sp = (256, 256, 200), (2.0, 2.0, 2.0)
sc = S.Scene(S.product_lib())
sc.world(dim, sp)
for name, dens in phantoms.ANTHROPOMORPHIC_MATERIALS:
    sc.add_material(name, dens)
mat, dens = phantoms.anthropomorphic(dim, sp)
sc.arrays(dens, mat)
assert sc.validate()
scan = dim[2] * sp[2]
sc.source_ct(spiral=True, voltage=120.0, al_mm=7.0, sdd=1190.0, collimation=40.0, fov=500.0, pitch=1.0, scan_length=scan,
             position=(0.0, 0.0, -scan / 2), exposure_step_deg=1.0, histories=hist, model_heel=True, ctdi_vol=10.0)
a, w = phantoms.bowtie_profile()
sc.source_bowtie(a, w)
flat = T.flatten_scene(sc)
exps = T.exposures_of(sc)
arr = (cabi.Exposure * len(exps))(*exps)
n_hist = sum(e.histories for e in exps)


def grids(o):
    d, ev, v = o.get_raw()
    return d.astype(np.float64), ev.astype(np.int64), v.astype(np.float64)


o = pyoracle.Oracle()
o.load(flat)
t0 = time.time()
o.run(exps, 0, len(exps), model=1, seed=7, per_history_streams=True)
plain, plain_ev, plain_v = grids(o)
st = o.stats()
print(f"plain Woodcock      : {st['lookups'] / n_hist:6.2f} look-ups per history, {time.time() - t0:5.1f} s, "
      f"deposited {plain.sum() / n_hist:.3f} keV per history, {plain_ev.sum() / n_hist:.3f} scoring events per history")
organ = mat.ravel()
for B in bricks:
    p = pyoracle.Oracle()
    p.load(flat)
    out = (C.c_double * 3)()
    t0 = time.time()
    rc = p.l.brick_probe_run(p.h, arr, C.c_uint64(0), C.c_uint64(len(exps)), 1, C.c_uint64(11), C.c_uint64(B), out)
    assert rc == 0
    g, g_ev, g_v = grids(p)
    # per-material (organ) totals: difference in units of the combined Monte Carlo uncertainty
    z = []
    for m in range(len(phantoms.ANTHROPOMORPHIC_MATERIALS)):
        sel = organ == m
        sigma = np.sqrt(plain_v[sel].sum() + g_v[sel].sum())
        if sigma > 0:
            z.append((plain[sel].sum() - g[sel].sum()) / sigma)
    print(f"bricks of {B:2d}^3 voxels: {out[0] / n_hist:6.2f} look-ups + {out[1] / n_hist:5.2f} face steps per history "
          f"(mean brick factor {out[2]:.3f}), {time.time() - t0:5.1f} s, deposited {g.sum() / n_hist:.3f} keV per history "
          f"({100 * (g.sum() / plain.sum() - 1):+.2f} %), events {g_ev.sum() / n_hist:.3f}, organ totals z = " + " ".join(f"{x:+.1f}" for x in z))
