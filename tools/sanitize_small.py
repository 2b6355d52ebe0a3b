"""A small run that goes through every wave kernel (births with air walks, stepping, air walks, interactions, forced interactions,
tile hand-over with several waves and a drain) for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ.setdefault("DXMCB200_BATCH", "16,12")  # waves of 4096 photons: many waves, re-fills from several tiles, a drain
import support as T
from dxmclib_b200 import scene as S

lib = S.product_lib()
for build in (lambda: T.air_gap_scene(lib, histories=6000, exposures=4, forced=True), lambda: T.ct_scene(lib, histories=300)):
    r = build().transport(model=1, output=S.OUT_EV_PER_HISTORY, seed=5)
    print("histories", r.histories, "events", int(r.n_events.sum()), "total", float(r.dose.sum()))
