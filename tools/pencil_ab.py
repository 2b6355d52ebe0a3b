"""BASELINE config #1 (60 keV pencil beam into a 64^3 water cube, 1e7 histories: every deposit lands near one column of
voxels): histories/s of Transport::run, for the built library or for the library files given on the command line; set
DXMCB200_AGGREGATE=0/1 to force per-lane or warp-aggregated scoring."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dxmclib_b200 import scene as S

def pencil(lib, histories=1_000_000, exposures=10):
    n = 64
    sc = S.Scene(lib)
    sc.world((n, n, n), (1, 1, 1))
    sc.add_material("Water, Liquid", 1.0)
    sc.arrays(np.full(n ** 3, 1.0, np.float32), np.zeros(n ** 3, np.uint8))
    assert sc.validate()
    sc.source_pencil((0.0, 0.0, -n), (1, 0, 0, 0, 1, 0), 60.0, histories, exposures)
    return sc

for path in sys.argv[1:] or [S.PRODUCT_LIB]:
    lib = S.load(path)
    for rep in range(3):
        r = pencil(lib).transport(model=S.MODEL_LIVERMORE, output=S.OUT_EV_PER_HISTORY, seed=5)
    print(f"{os.path.basename(path):28s} pencil 64^3 water, 1e7 histories: {r.histories / r.seconds:.3e} histories/s  (total {r.dose.sum():.2f} eV/history, "
          f"events {int(r.n_events.sum())})", flush=True)
