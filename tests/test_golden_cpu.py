"""The restatement oracle against the committed golden vectors (tests/golden/*.npz, produced from the unmodified
reference by tests/golden/make_golden.py). Runs without /root/reference and without a GPU."""
import os

import numpy as np
import pytest

import support as T
from dxmclib_b200 import scene as S
from oracle import pyoracle

G = os.path.join(T.ROOT, "tests", "golden")


def _oracle_for(sc, max_energy=None):
    o = pyoracle.Oracle()
    o.load(T.flatten_scene(sc, max_energy))
    return o


@pytest.mark.parametrize("world", ["unit", "aniso", "fine"])
def test_voxel_index_sequences_bit_exact(product, world):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_defs", os.path.join(G, "make_golden.py"))
    g = np.load(os.path.join(G, "traces.npz"))
    # world definitions are shared with the generator without importing it (it loads the reference library on import)
    worlds = {"unit": dict(dim=(64, 48, 40), spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)),
              "aniso": dict(dim=(37, 51, 29), spacing=(0.7, 1.3, 2.5), origin=(12.5, -7.25, 100.0)),
              "fine": dict(dim=(200, 10, 10), spacing=(0.1, 3.0, 3.0), origin=(-3.0, 0.5, 0.25))}
    assert spec is not None
    w = worlds[world]
    sc = S.Scene(product)
    sc.world(w["dim"], w["spacing"], w["origin"])
    sc.add_material("Water, Liquid")
    n = int(np.prod(w["dim"]))
    sc.arrays(np.ones(n, np.float32), np.zeros(n, np.uint8))
    assert sc.validate()
    o = _oracle_for(sc, 60.0)
    idx, entry = o.trace_indices(g[f"{world}_pos"], g[f"{world}_dir"], g[f"{world}_steps"])
    assert (g[f"{world}_idx"] >= 0).sum() > 500
    assert np.array_equal(idx, g[f"{world}_idx"])
    assert T.bit_equal(entry, g[f"{world}_entry"])


def test_lut_values_bit_exact(product):
    g = np.load(os.path.join(G, "lut.npz"))
    sc = T.tissue_block(product)
    o = _oracle_for(sc, 140.0)
    e = g["energy"]
    for m in range(4):
        att, mx = o.eval_attenuation(np.full(e.size, m, np.uint8), e)
        assert T.bit_equal(att, g["attenuation"][m])
        assert T.bit_equal(mx, g["max_inverse"])
        # the product's host-side evaluation of the same tables
        assert T.bit_equal(np.stack([sc.lut_attenuation(m, x) for x in e[::9]]), g["attenuation"][m][::9])
        assert T.bit_equal(np.array([sc.lut_scatter_factor(m, q) for q in g["q"]], np.float32), g["scatter_factor"][m])


SCENES = {
    "pencil": lambda lib: T.pencil_scene(lib, histories=40000, exposures=4),
    "isotropic_forced": lambda lib: T.isotropic_scene(lib, histories=30000, forced=True),
    "ct_spiral": lambda lib: T.ct_scene(lib, histories=1500),
}


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("model", [0, 1, 2])
def test_seeded_transport_matches_reference_run(product, name, model):
    """Tables built by the PRODUCT's host classes + the restatement's transport == the reference's seeded run
    (bit-exact events; totals and profiles to float summation order)."""
    g = np.load(os.path.join(G, f"transport_{name}_m{model}.npz"))
    sc = SCENES[name](product)
    o = _oracle_for(sc)
    exps = T.exposures_of(sc)
    o.run(exps, 0, len(exps), model=model, seed=T.SEED, per_history_streams=False)
    dose, ev, _ = o.get_raw()
    nx, ny, nz = sc.dim
    n = int(g["histories"])
    assert n == sum(e.histories for e in exps)
    assert int(ev.sum()) == int(g["events"])
    assert np.array_equal(ev.reshape(nz, ny, nx).sum(axis=(1, 2)), g["events_z"])
    d = dose.astype(np.float64).reshape(nz, ny, nx) * 1e3 / n
    np.testing.assert_allclose(d.sum(), float(g["total"]), rtol=1e-5)
    np.testing.assert_allclose(d.sum(axis=(1, 2)), g["profile_z"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(d.sum(axis=(0, 1)), g["profile_x"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", list(SCENES))
def test_counter_stream_transport_matches_reference_run(product, name):
    g = np.load(os.path.join(G, f"streams_{name}_m1.npz"))
    sc = SCENES[name](product)
    o = _oracle_for(sc)
    exps = T.exposures_of(sc)
    o.run(exps, 0, len(exps), model=1, seed=T.SEED, per_history_streams=True)
    dose, ev, _ = o.get_raw()
    assert T.bit_equal(ev, g["n_events"])
    np.testing.assert_allclose(dose.astype(np.float64).sum() * 1e3 / int(g["histories"]), float(g["total"]), rtol=1e-5)
