"""Pins the CPU restatement (oracle/dxmc_oracle.cpp) to the UNMODIFIED reference (oracle/_ref, built from
/root/reference): driven by the reference's own sequential PCG32 RandomState, the restatement must reproduce the
reference's dose / event / variance grids BIT FOR BIT, for every low-energy model, on scenes that cover
mono-energetic and spectrum sources, bow-tie + heel + AEC + XCare + gantry tilt CT, and forced interactions
(measurement map). Function-level entry points (LUT evaluation, voxel index traces) are pinned the same way."""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import scene as S
from oracle import pyoracle

SCENES = {
    "pencil": (lambda lib: T.pencil_scene(lib, histories=4000, exposures=3), (0, 1, 2)),
    "isotropic_spectrum": (lambda lib: T.isotropic_scene(lib, histories=4000), (0, 1, 2)),
    "isotropic_forced": (lambda lib: T.isotropic_scene(lib, histories=4000, forced=True), (0, 1, 2)),
    "isotropic_ct_mono": (lambda lib: T.isotropic_scene(lib, histories=3000, exposures=5, ct=True, mono=56.4), (1,)),
    "ct_spiral": (lambda lib: T.ct_scene(lib, histories=250), (1, 2)),
    "ct_axial": (lambda lib: T.ct_scene(lib, spiral=False, histories=250, xcare=False, tilt=0.0), (1,)),
    "ctdi_forced": (lambda lib: T.ctdi_scene(lib, histories=400), (0, 1, 2)),
}
CASES = [(name, m) for name, (_, models) in SCENES.items() for m in models]


def normalize_scoring(dose, var, n):
    """Transport::normalizeScoring in float (reference transport.hpp:780-794)."""
    h_inv = np.float32(1) / np.float32(n - 1)
    h_d_inv = np.float32(1e3) / np.float32(n)
    h_v_inv = np.float32(1e6) / np.float32(n)
    d = dose * h_d_inv
    return d, ((var * h_v_inv - d * d) * h_inv).astype(np.float32)


@pytest.mark.parametrize("name,model", CASES)
def test_restatement_is_bit_identical_to_reference(reference, name, model):
    sc = SCENES[name][0](reference)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    ref = sc.transport(model=model, output=S.OUT_EV_PER_HISTORY, seed=T.SEED)  # seeded single worker
    o = pyoracle.Oracle()
    o.load(flat)
    o.run(exps, 0, len(exps), model=model, seed=T.SEED, per_history_streams=False)
    dose, ev, var = o.get_raw()
    d, v = normalize_scoring(dose, var, ref.histories)
    assert ev.sum() > 1000
    assert T.bit_equal(ev, ref.n_events)
    assert T.bit_equal(d, ref.dose)
    assert T.bit_equal(v, ref.variance)


def test_lut_evaluation_bit_identical(reference):
    sc = T.tissue_block(reference)
    flat = T.flatten_scene(sc, max_energy=140.0)
    o = pyoracle.Oracle()
    o.load(flat)
    rng = np.random.default_rng(11)
    e = np.concatenate([rng.uniform(1.0, 140.0, 3000), [1.0, 4.0385, 4.0386, 33.1694, 33.17, 140.0]]).astype(np.float32)
    m = rng.integers(0, 4, e.size).astype(np.uint8)
    att, mx = o.eval_attenuation(m, e)
    for i in range(0, e.size, 7):
        assert T.bit_equal(att[i], sc.lut_attenuation(int(m[i]), e[i]))
        assert mx[i] == sc.lut_max_inverse(e[i])


def test_counter_stream_mode_matches_reference_counter_streams(reference):
    """Per-history stream mode (what the CUDA kernels use): same streams in the reference harness give the same events;
    dose sums agree to float accumulation order."""
    sc = T.isotropic_scene(reference, histories=3000, forced=True)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    ref = sc.transport(model=1, output=S.OUT_EV_PER_HISTORY, seed=T.SEED, workers=S.WORKERS_COUNTER_STREAMS)
    o = pyoracle.Oracle()
    o.load(flat)
    o.run(exps, 0, len(exps), model=1, seed=T.SEED, per_history_streams=True)
    dose, ev, var = o.get_raw()
    d, _ = normalize_scoring(dose, var, ref.histories)
    assert T.bit_equal(ev, ref.n_events)
    np.testing.assert_allclose(d, ref.dose, rtol=2e-5, atol=1e-6)


def test_shards_sum_to_the_whole_bit_for_bit(reference):
    """Fixed-point scoring: any partition of the exposure range sums to exactly the single-run grids."""
    sc = T.isotropic_scene(reference, histories=1500, exposures=6)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    whole = pyoracle.Oracle()
    whole.load(flat)
    whole.run(exps, 0, 6, model=1, seed=5, per_history_streams=True)
    e_all, e2_all = whole.get_fixed()
    acc_e, acc_e2 = np.zeros_like(e_all), np.zeros_like(e2_all)
    for b, e in ((0, 1), (1, 4), (4, 6)):
        part = pyoracle.Oracle()
        part.load(flat)
        part.run(exps, b, e, model=1, seed=5, per_history_streams=True)
        pe, pe2 = part.get_fixed()
        acc_e += pe
        acc_e2 += pe2
    assert e_all.any() and T.bit_equal(acc_e, e_all) and T.bit_equal(acc_e2, e2_all)
