"""BASELINE config #5 through ONE reference-facing call spread over the GPUs of the box (Transport::setDevices: one host thread per
device, NCCL reduce-scatter over voxel slices, every device decodes and downloads its slice): histories/s of the whole call with
host arrays in and out. usage: python tools/multidevice_call.py <n_devices> [calls]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DXMCB200_POOL_GB", "96")
import numpy as np
import bench
from dxmclib_b200 import scene as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.Workload(5, n, "weak")
sc = w.build(S.product_lib())
sc.b200_set_devices(list(range(n)))
times, checksum = [], None
for i in range(calls + 1):  # the first call carries context creation on every device
    t0 = time.perf_counter()
    r = sc.transport(model=bench.MODEL, output=S.OUT_EV_PER_HISTORY, seed=bench.SEED)
    times.append(time.perf_counter() - t0)
    checksum = float(r.dose.astype(np.float64).sum())
line = {"what": "BASELINE config #5 through one Transport call on %d devices (setDevices)" % n, "n_gpus": n, "histories_per_call": int(r.histories),
        "seconds_per_call": [round(t, 3) for t in times], "value": r.histories / float(np.mean(times[1:])), "unit": "histories/s",
        "run_seconds_last_call": r.seconds, "total_eV_per_history": checksum}
print(json.dumps(line))
