"""Empty-space traversal, CPU leg: the restatement's mode 1 (Woodcock + air-brick traversal, the
scheme the CUDA kernels run by default) against its own mode 0, which is pinned bit for bit to the unmodified reference
(tests/test_oracle_pinned.py). The two consume different random numbers, so the comparison is statistical: the criteria of
BASELINE.json north_star (total energy 0.5 %, voxels within 3 sigma where the relative error is under 2 %)."""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import scene as S
from oracle import pyoracle


def _run(flat, exps, tracking, mm, seed, model=1):
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_tracking(tracking, mm)
    o.run(exps, 0, len(exps), model=model, seed=seed, per_history_streams=True)
    d, ev, v = o.get_raw()
    return d.astype(np.float64), v.astype(np.float64), ev.astype(np.int64), o.stats(), (o.walk_stats() if tracking else [0, 0, 0]), o.bricks()


@pytest.mark.parametrize("name,build,mm,model", [
    ("air_gap", lambda lib: T.air_gap_scene(lib, histories=1500000, exposures=4), 8.0, 1),
    ("air_gap_forced_ia", lambda lib: T.air_gap_scene(lib, histories=1000000, exposures=4, forced=True), 8.0, 2),
    ("ct_spiral", lambda lib: T.ct_scene(lib, histories=60000), 16.0, 1),
])
def test_mode1_statistically_equivalent_to_reference_tracking(product, name, build, mm, model):
    sc = build(product)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    a, a2, aev, sa, _, _ = _run(flat, exps, 0, mm, 5, model)
    b, b2, bev, sb, walks, bricks = _run(flat, exps, 1, mm, 6, model)
    assert bricks["air"].mean() > 0.3 and walks[0] > sa["histories"] * 0.5
    assert sb["lookups"] < 0.8 * sa["lookups"], "the traversal should remove most look-ups in air"
    sigma = np.sqrt(a2.sum() + b2.sum())  # of the difference of the two totals
    assert sigma / a.sum() < 2.5e-3, "sample too small to resolve the north_star bound of 0.5 %"
    assert abs(a.sum() - b.sum()) < 3.5 * sigma, (a.sum(), b.sum(), sigma)
    assert abs(a.sum() - b.sum()) / a.sum() < 5e-3 + 2.0 * sigma / a.sum()  # north_star bound, widened by the sample's own noise
    assert abs(int(aev.sum()) - int(bev.sum())) < 1e-2 * int(aev.sum())
    # the CPU affords about a million histories: the per-voxel criterion is applied to blocks of 4 x 4 x 4 voxels
    nx, ny, nz = (int(x) for x in flat["dim"])

    def blocks(g):
        g = g.reshape(nz, ny, nx)[: nz // 4 * 4, : ny // 4 * 4, : nx // 4 * 4]
        return g.reshape(nz // 4, 4, ny // 4, 4, nx // 4, 4).sum(axis=(1, 3, 5)).ravel()

    outside, tested, worst = T.compare_dose(blocks(a), blocks(a2), blocks(b), blocks(b2), rel_err_limit=0.05)
    assert tested >= 50
    assert outside * tested <= T.allowed_outliers(tested) and worst < 5.0, f"{outside:.4%} of {tested} voxels beyond 3 sigma (worst {worst:.2f})"


def test_mode1_without_air_bricks_is_mode0(product):
    """A grid with no air brick (homogeneous block): mode 1 must consume the same random numbers and give the same bits as mode 0."""
    sc = T.isotropic_scene(product, histories=4000, exposures=2)
    flat = T.flatten_scene(sc)
    exps = T.exposures_of(sc)
    a = _run(flat, exps, 0, 16.0, 9)
    b = _run(flat, exps, 1, 16.0, 9)
    assert not b[5]["air"].any()
    assert T.bit_equal(a[0], b[0]) and T.bit_equal(a[2], b[2])


def test_brick_layout_rule(product):
    """Brick edges are powers of two (voxels) closest to the requested size; the grid never exceeds 16384 bricks."""
    sc = S.Scene(product)
    sc.world((200, 120, 90), (0.5, 1.0, 2.5))
    sc.add_material(T.AIR, 0.001205)
    sc.arrays(np.full(200 * 120 * 90, 0.001205, np.float32), np.zeros(200 * 120 * 90, np.uint8))
    assert sc.validate()
    flat = T.flatten_scene(sc, max_energy=100.0)
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_tracking(1, 8.0)
    b = o.bricks()
    assert b["shift"] == [4, 3, 2] and b["nb"] == [13, 15, 23]
    o.set_tracking(1, 1.0)  # 400 x 120 x 90 bricks of about 1 mm would be too many: axes with the shortest edge grow first
    b = o.bricks()
    assert int(np.prod(b["nb"])) <= 16384
    # a homogeneous world has no brick that is thin compared with the majorant (which is the same material): no air bricks
    assert not b["air"].any() and b["f_air"] == 0


@pytest.mark.parametrize("name,build,mm", [
    ("air_gap", lambda lib: T.air_gap_scene(lib), 8.0),
    ("ct_spiral", lambda lib: T.ct_scene(lib, histories=100), 16.0),
])
def test_air_run_traversal_is_true_against_dense_sampling(product, name, build, mm):
    """The restatement's ray / brick-grid traversal (airRunLength: the path's Siddon / Amanatides-Woo style traversal, crossing
    whole all-air cubes) against ground truth: dense sampling of the brick flags along 20000 fixed rays. The kernels' traversal is
    compared with this one bit for bit in tests/test_gpu_empty_space.py."""
    flat = T.flatten_scene(build(product))
    o = pyoracle.Oracle()
    o.load(flat)
    o.set_tracking(1, mm)
    pos32, d32 = T.air_run_rays(flat, 20000, seed=3)
    result = o.trace_air_runs(pos32, d32)
    _, entry = o.trace_indices(pos32, d32, np.zeros(0, np.float32))
    T.assert_air_runs_true(flat, o.bricks(), entry, d32, result)
