"""Phase timings of the reference-facing call (Transport::operator() through dxs_transport) on the bench scene."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DXMCB200_TRACE"] = "1"
import bench
from dxmclib_b200 import phantoms, scene as S
hist = int(sys.argv[1]) if len(sys.argv) > 1 else 27778
t0 = time.time()
ph = phantoms.anthropomorphic(bench.DIM, bench.SPACING)
print(f"phantom {time.time()-t0:.2f}s", file=sys.stderr)
t0 = time.time()
sc = bench.build_scene(S.product_lib(), hist, phantom=ph)
print(f"build_scene {time.time()-t0:.2f}s", file=sys.stderr)
for i in range(2):
    t0 = time.time()
    r = sc.transport(model=bench.MODEL, output=S.OUT_EV_PER_HISTORY, seed=bench.SEED)
    print(f"sc.transport total {time.time()-t0:.2f}s (histories {r.histories})", file=sys.stderr)
