"""Empty-space traversal (tracking mode 1, the product's default) on the GPU.

The scheme is not in the reference, so its parity has three legs:
  * the brick grid (per-material ratios, per-brick maxima, air flags, f_air) is bit-identical to the CPU restatement's;
  * the kernels follow the restatement's empty-space tracking history by history (same per-history random streams):
    event grids nearly identical, totals to 1e-4, work counters (steps, look-ups, air walks, bricks crossed) to 2e-4;
  * against plain Woodcock tracking (mode 0 = the reference's algorithm, itself pinned to the unmodified reference) the
    results agree statistically at high history counts: total energy to 0.1 %, voxels within 3 sigma.
tests/test_empty_space_cpu.py holds the CPU leg: restatement mode 1 against restatement mode 0 and the reference."""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import cabi
from oracle import pyoracle

pytestmark = pytest.mark.gpu

SCENES = {
    "air_gap": (lambda lib: T.air_gap_scene(lib, histories=30000, exposures=4), 8.0),
    "air_gap_forced": (lambda lib: T.air_gap_scene(lib, histories=30000, exposures=4, forced=True), 8.0),
    "ct_spiral": (lambda lib: T.ct_scene(lib, histories=1500), 16.0),
    "ctdi": (lambda lib: T.ctdi_scene(lib, histories=3000), 16.0),
    "pencil": (lambda lib: T.pencil_scene(lib, histories=20000, exposures=4), 8.0),
}


def _both(product, name, model=1, seed=17, palette_break=False):
    build, mm = SCENES[name]
    sc = build(product)
    flat = dict(T.flatten_scene(sc))
    if palette_break:  # more than 256 distinct records: the 8-byte record grid
        rng = np.random.default_rng(5)
        flat["density"] = (flat["density"] * rng.uniform(0.9, 1.0, flat["density"].size)).astype(np.float32)
    exps = T.exposures_of(sc)
    ctx = cabi.Context(0)
    T.load_context(ctx, flat)
    ctx.set_tracking(1, mm)
    ctx.set_fixed_point(22, 12)
    ctx.enable_stats(True)
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_tracking(1, mm)
    o.set_fixed_point(22, 12)
    return sc, flat, exps, ctx, o


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("record_grid", [False, True])
def test_brick_grid_bit_identical_to_oracle(gpu, product, name, record_grid):
    _, flat, _, ctx, o = _both(product, name, palette_break=record_grid)
    a, b = ctx.bricks(flat["luts"]["n_materials"]), o.bricks()
    assert a["shift"] == b["shift"] and a["nb"] == b["nb"]
    assert T.bit_equal(a["ratio"], b["ratio"])
    assert T.bit_equal(a["brick_max"], b["brick_max"])
    assert T.bit_equal(a["air"], b["air"]) and T.bit_equal(a["distance"], b["distance"])
    assert ((a["distance"].reshape(8, -1) > 0) == (a["air"] > 0)[None, :]).all()  # [octant][brick]
    assert np.float32(a["f_air"]).tobytes() == np.float32(b["f_air"]).tobytes()
    assert a["air"].any(), "scene without air bricks does not test the traversal"
    ctx.close()


def test_brick_grid_ignores_negative_and_nan_densities(gpu, product):
    sc = T.air_gap_scene(product)
    flat = dict(T.flatten_scene(sc))
    d = flat["density"].copy()
    d[5], d[77], d[1234] = -3.0, np.nan, -0.0
    flat["density"] = d
    ctx = cabi.Context(0)
    T.load_context(ctx, flat)
    ctx.set_tracking(1, 8.0)
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_tracking(1, 8.0)
    a, b = ctx.bricks(flat["luts"]["n_materials"]), o.bricks()
    assert T.bit_equal(a["brick_max"], b["brick_max"]) and T.bit_equal(a["air"], b["air"])
    assert np.isfinite(a["brick_max"]).all() and (a["brick_max"] >= 0).all()
    ctx.close()


@pytest.mark.parametrize("name", list(SCENES))
@pytest.mark.parametrize("model", [0, 1, 2])
def test_follows_oracle_history_by_history(gpu, product, name, model):
    if model != 1 and name not in ("air_gap", "ct_spiral"):
        pytest.skip("models 0 and 2 on two scenes")
    sc, flat, exps, ctx, o = _both(product, name, model)
    ctx.run(exps, 0, len(exps), model=model, seed=17)
    e, e2, ev = ctx.get_raw()
    s = ctx.stats()
    o.run(exps, 0, len(exps), model=model, seed=17, per_history_streams=True)
    oe, oe2 = o.get_fixed()
    _, oev, _ = o.get_raw()
    t = o.stats()
    w = o.walk_stats()
    assert ev.sum() > 5000 and w[0] > 1000
    assert (ev.astype(np.int64) != oev.astype(np.int64)).mean() < 5e-3
    np.testing.assert_allclose(e.sum() / 2.0 ** 22, oe.sum() / 2.0 ** 22, rtol=2e-4)
    assert s["histories"] == t["histories"]
    for k in ("histories_in_world", "steps", "lookups", "interactions", "score_events"):
        assert abs(s[k] - t[k]) <= 2e-4 * t[k] + 5, (k, s[k], t[k])
    assert abs(s["air_walks"] - w[0]) <= 2e-4 * w[0] + 5, (s["air_walks"], w[0])
    assert abs(s["bricks_crossed"] - w[1]) <= 2e-4 * w[1] + 5, (s["bricks_crossed"], w[1])
    ctx.close()


@pytest.mark.parametrize("name", ["air_gap", "ct_spiral", "ctdi"])
def test_air_run_traversal_bit_exact_and_true(gpu, product, name):
    """The ray / brick-grid traversal of the air walk (the path's Siddon / Amanatides-Woo style traversal) for fixed rays:
    run length, cubes crossed and the exit flag are the bits of the CPU restatement's, and the run is TRUE against dense sampling
    of the brick flags along the ray: every point of the run lies in an air brick (or outside the grid), and a run that ended
    neither at the grid's edge nor at the cube cap ends on the face of a non-air brick."""
    _, flat, _, ctx, o = _both(product, name)
    pos32, d32 = T.air_run_rays(flat, 20000, seed=11)
    got = ctx.trace_air_runs(pos32, d32)
    want = o.trace_air_runs(pos32, d32)
    for a, b in zip(got, want):
        assert T.bit_equal(np.ascontiguousarray(a), np.ascontiguousarray(b))
    _, entry = ctx.trace_indices(pos32, d32, np.zeros(0, np.float32))
    T.assert_air_runs_true(flat, ctx.bricks(flat["luts"]["n_materials"]), entry, d32, got)
    ctx.close()


def test_record_grid_follows_oracle(gpu, product):
    sc, flat, exps, ctx, o = _both(product, "air_gap", palette_break=True)
    ctx.run(exps, 0, len(exps), model=1, seed=23)
    e, _, ev = ctx.get_raw()
    o.run(exps, 0, len(exps), model=1, seed=23, per_history_streams=True)
    oe, _ = o.get_fixed()
    _, oev, _ = o.get_raw()
    assert (ev.astype(np.int64) != oev.astype(np.int64)).mean() < 5e-3
    np.testing.assert_allclose(e.sum() / 2.0 ** 22, oe.sum() / 2.0 ** 22, rtol=2e-4)
    ctx.close()


# the scene has air + three denser materials: 3 * tissue_records + air_records plain records, air_records flagged twins
@pytest.mark.parametrize("tissue_records,air_records,bits", [(4, 1, 4), (4, 3, 8), (83, 7, 64)])
def test_air_brick_flags_widen_the_palette_and_change_nothing(gpu, product, tissue_records, air_records, bits, monkeypatch):
    """Tracking mode 1 writes "this voxel lies in an air brick" into the voxel records; a palette grid gets flagged twins of the
    entries that occur in air bricks and is widened (4-bit -> byte -> 8-byte records) when they do not fit. Whatever form results,
    the accumulators are the bits of the plain 8-byte record grid."""
    sc = T.air_gap_scene(product, histories=20000, exposures=4)
    flat = dict(T.flatten_scene(sc))
    d, m = flat["density"].copy(), flat["material"]
    air = d < 0.01
    assert air.any() and (~air).any()
    idx = np.arange(d.size)
    dense = np.flatnonzero(~air)
    d[dense] = (d[dense] * (1.0 - 1e-3 * (idx[dense] % tissue_records))).astype(np.float32)
    thin = np.flatnonzero(air)
    d[thin] = (d[thin] * (1.0 - 1e-2 * (idx[thin] % air_records))).astype(np.float32)
    flat["density"] = d
    exps = T.exposures_of(sc)
    grids = {}
    for palette in ("1", "0"):
        monkeypatch.setenv("DXMCB200_PALETTE", palette)
        ctx = cabi.Context(0)
        T.load_context(ctx, flat)
        ctx.set_tracking(1, 8.0)
        ctx.set_fixed_point(22, 12)
        ctx.run(exps, 0, len(exps), model=1, seed=31)
        grids[palette] = (ctx.get_raw(), ctx.grid_form())
        ctx.close()
    (a, form_a), (b, form_b) = grids["1"], grids["0"]
    assert form_b[0] == 64
    assert form_a[0] == bits, f"{form_a[1]} distinct records ended up as {form_a[0]} bits per voxel"
    for x, y in zip(a, b):
        assert T.bit_equal(x, y)
    assert a[2].sum() > 5000


@pytest.mark.parametrize("name,histories", [("air_gap", 60_000_000), ("ct_spiral", 2_000_000)])
def test_statistically_equivalent_to_plain_woodcock(gpu, product, name, histories):
    """Mode 1 against mode 0 (the reference's algorithm) with independent seeds: total deposited energy within 0.1 % (north_star:
    0.5 %), per-voxel dose within 3 sigma of the combined uncertainty in voxels with under 2 % relative error (a 0.3 % tail and
    5 sigma worst case allowed over tens of thousands of voxels, as in test_live_reference_three_sigma)."""
    build, mm = SCENES[name]
    sc = T.air_gap_scene(product, histories=histories, exposures=4) if name == "air_gap" else T.ct_scene(product, histories=histories)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    n = sum(x.histories for x in exps)
    grids = []
    for tracking, seed in ((0, 101), (1, 202)):
        ctx = cabi.Context(0)
        T.load_context(ctx, flat)
        ctx.set_tracking(tracking, mm)
        bits = cabi.suggest_fixed_point(n, 150.0)
        ctx.set_fixed_point(*bits)
        ctx.run(exps, 0, len(exps), model=1, seed=seed)
        e, e2, ev = ctx.get_raw()
        grids.append((e.astype(np.float64) / 2.0 ** bits[0], e2.astype(np.float64) / 2.0 ** bits[1], ev))
        ctx.close()
    (a, a2, aev), (b, b2, bev) = grids
    sigma_total = np.sqrt(a2.sum() + b2.sum())
    assert abs(a.sum() - b.sum()) < max(4.0 * sigma_total, 1e-3 * a.sum()), (a.sum(), b.sum(), sigma_total)
    assert abs(a.sum() - b.sum()) / a.sum() < 1e-3
    assert abs(int(aev.sum()) - int(bev.sum())) < 2e-3 * int(aev.sum())
    outside, tested, worst = T.compare_dose(a, a2, b, b2)
    assert tested >= 100, "scene too sparse for the per-voxel criterion"
    assert outside * tested <= T.allowed_outliers(tested) and worst < 5.0, f"{outside:.4%} of {tested} voxels beyond 3 sigma (worst {worst:.2f})"


def test_schedule_invariance_with_air_walks(gpu, product, monkeypatch):
    """Bit-identical accumulators for any wave size, re-fill batch, pipeline count and grid layout, with the air walk on."""
    sc = T.air_gap_scene(product, histories=30000, exposures=4, forced=True)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)

    def run(batch, palette, pipes):
        monkeypatch.setenv("DXMCB200_BATCH", batch)
        monkeypatch.setenv("DXMCB200_PALETTE", palette)
        monkeypatch.setenv("DXMCB200_PIPES", pipes)
        ctx = cabi.Context(0)
        T.load_context(ctx, flat)
        ctx.set_tracking(1, 8.0)
        ctx.set_fixed_point(20, 10)
        ctx.run(exps, 0, len(exps), model=1, seed=99)
        raw = ctx.get_raw()
        ctx.close()
        return raw

    base = run("8,26", "1", "2")
    assert base[2].sum() > 30000
    for setting in [("1,10", "1", "2"), ("32,12", "0", "1"), ("8,14", "8", "3"), ("4,11", "1", "1")]:
        other = run(*setting)
        for x, y in zip(base, other):
            assert T.bit_equal(x, y), f"grids differ for DXMCB200_BATCH/PALETTE/PIPES = {setting}"
