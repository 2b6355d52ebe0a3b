import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import support as T
from dxmclib_b200 import scene as S
prod, ref = S.product_lib(), S.reference_lib()
def holes(sc, r):
    return [float(r.dose[sc.ctdi_holes(p).astype(np.int64)].astype(np.float64).mean()) for p in range(5)]
for model in (1,):
    a = T.ctdi_scene(prod, histories=60000, diameter=320); b = T.ctdi_scene(ref, histories=60000, diameter=320)
    t=time.time(); ra = a.transport(model=model, output=S.OUT_DOSE, use_calibration=False, seed=5); ta=time.time()-t
    t=time.time(); rb = b.transport(model=model, output=S.OUT_DOSE, use_calibration=False, seed=5, workers=S.WORKERS_COUNTER_STREAMS); tb=time.time()-t
    ha, hb = holes(a, ra), holes(b, rb)
    print("hist", ra.histories, "times", ta, tb)
    print("prod holes", ha); print("ref  holes", hb); print("ratio", [x/y for x,y in zip(ha,hb)])
    print("total", ra.dose.astype(np.float64).sum()/rb.dose.astype(np.float64).sum(), "events", ra.n_events.sum(), rb.n_events.sum())
for seed in (1,2,3):
    sc = T.ct_scene(prod, histories=200, aec=False, xcare=False, tilt=0.0)
    print("product calibration", sc.calibration(S.MODEL_LIVERMORE))
for k in range(2):
    sc = T.ct_scene(ref, histories=200, aec=False, xcare=False, tilt=0.0)
    print("reference calibration", sc.calibration(S.MODEL_LIVERMORE))
