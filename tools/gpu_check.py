"""Ad-hoc GPU sanity run: product vs reference (counter streams) on the pencil-beam example."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dxmclib_b200 import scene as S


def pencil(lib, n=56, hist=100000, nexp=16, model=S.MODEL_LIVERMORE, seed=77, workers=0):
    sc = S.Scene(lib)
    sc.world((n, n, n), (1, 1, 1))
    sc.add_material("Air, Dry (near sea level)").add_material("Water, Liquid").add_element(13)
    d = [sc.material_density(i) for i in range(3)]
    dens = np.zeros((n, n, n), np.float32)
    mat = np.zeros((n, n, n), np.uint8)
    for k in range(n):
        m = 0 if k < n // 3 else (1 if k < 2 * n // 3 else 2)
        mat[k] = m
        dens[k] = d[m]
    sc.arrays(dens, mat)
    assert sc.validate()
    sc.source_pencil((0.3, 0.2, -n), (1, 0, 0, 0, 1, 0), 60.0, hist, nexp)
    t = time.time()
    r = sc.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=seed, workers=workers)
    return r, time.time() - t


for model in (S.MODEL_NONE, S.MODEL_LIVERMORE, S.MODEL_IA):
    rp, tp = pencil(S.product_lib(), model=model)
    rr, tr = pencil(S.reference_lib(), model=model, workers=-1)
    n = 56
    dp = rp.dose.reshape(n, n, n)
    dr = rr.dose.reshape(n, n, n)
    print(f"model {model}: product total {dp.sum():.3f} eV/hist ({rp.seconds:.3f}s kernel+launch, {rp.histories/rp.seconds:.3e} hist/s) "
          f"reference total {dr.sum():.3f} ({rr.seconds:.2f}s, {rr.histories/rr.seconds:.3e} hist/s) rel diff {(dp.sum()-dr.sum())/dr.sum():.2e}")
    print("   events", rp.n_events.sum(), rr.n_events.sum(), "identical event grids:", np.array_equal(rp.n_events, rr.n_events),
          "voxels differing:", int((rp.n_events != rr.n_events).sum()))
    prof_p = dp.sum(axis=(1, 2))
    prof_r = dr.sum(axis=(1, 2))
    print("   depth profile max rel diff", np.max(np.abs(prof_p - prof_r) / np.maximum(prof_r, 1e-9)))
