"""Phase timings of the reference's DEFAULT call (OUTPUTMODE::DOSE with the CT source's CTDI calibration run inside) on the bench scene."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DXMCB200_TRACE"] = "1"
os.environ.setdefault("DXMCB200_POOL_GB", "96")
import bench
from dxmclib_b200 import phantoms, scene as S
hist = int(sys.argv[1]) if len(sys.argv) > 1 else 27778
ph = phantoms.anthropomorphic(bench.DIM, bench.SPACING)
sc = bench.build_scene(S.product_lib(), hist, phantom=ph)
for i in range(3):
    t0 = time.time()
    r = sc.transport(model=bench.MODEL, output=S.OUT_DOSE, use_calibration=True, seed=bench.SEED)
    print(f"sc.transport(DOSE, calibration) total {time.time()-t0:.2f}s (histories {r.histories}, {r.units})", file=sys.stderr)
