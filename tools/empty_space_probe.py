"""CPU experiment behind DESIGN.md section 4b (empty-space traversal): runs the CPU restatement with plain Woodcock tracking and
with Woodcock + air-brick traversal on a half-resolution copy of the bench workload (CT spiral over the anthropomorphic phantom,
256x256x200 voxels of 2 mm) and reports voxel look-ups, air walks and bricks crossed per history, and how the dose grids compare.
    python tools/empty_space_probe.py [histories per exposure] [brick edge in mm ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle  # noqa: E402
import support as T  # noqa: E402
from dxmclib_b200 import phantoms  # noqa: E402
from dxmclib_b200 import scene as S  # noqa: E402

hist = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bricks = [float(x) for x in sys.argv[2:]] or [16.0, 32.0]
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
print(f"plain Woodcock     : {st['lookups'] / n_hist:6.2f} look-ups, {st['interactions'] / n_hist:5.2f} interactions per history, {time.time() - t0:5.1f} s, "
      f"deposited {plain.sum() / n_hist:.3f} keV per history, {plain_ev.sum() / n_hist:.3f} scoring events per history")
organ = mat.ravel()
for B in bricks:
    p = pyoracle.Oracle()
    p.load(flat)
    p.set_tracking(1, B)
    info = p.bricks()
    t0 = time.time()
    p.run(exps, 0, len(exps), model=1, seed=11, per_history_streams=True)
    g, g_ev, g_v = grids(p)
    st = p.stats()
    ws = p.walk_stats()
    z = []
    for m in range(len(phantoms.ANTHROPOMORPHIC_MATERIALS)):
        sel = organ == m
        sigma = np.sqrt(plain_v[sel].sum() + g_v[sel].sum())
        if sigma > 0:
            z.append((plain[sel].sum() - g[sel].sum()) / sigma)
    print(f"air bricks ~{B:4.0f} mm : {st['lookups'] / n_hist:6.2f} look-ups ({ws[2] / n_hist:.3f} of them in air), {ws[0] / n_hist:5.2f} walks, "
          f"{ws[1] / n_hist:5.2f} bricks crossed per history; grid {info['nb']} bricks of 2^{info['shift']} voxels, {info['air'].mean():.1%} air, "
          f"f_air {info['f_air']:.2e}; {time.time() - t0:5.1f} s, deposited {g.sum() / n_hist:.3f} keV per history "
          f"({100 * (g.sum() / plain.sum() - 1):+.2f} %), events {g_ev.sum() / n_hist:.3f}, organ totals z = " + " ".join(f"{x:+.1f}" for x in z))
