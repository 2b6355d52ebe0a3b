"""Host-side builders of the drop-in C++ classes against the UNMODIFIED reference, bit for bit: material queries,
attenuation LUT / majorant / RITA / spline / shell tables, exposures of every source type, tube spectrum, alias, heel
and bow-tie tables, CTDI phantom. (reference files cited in the headers under
dxmclib_b200/include/dxmc/.)"""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import phantoms
from dxmclib_b200 import scene as S

MATERIALS = ["Water, Liquid", "Air, Dry (near sea level)", "Polymethyl Methacralate (Lucite, Perspex)", "Bone, Cortical (ICRP)",
             "Tissue, Soft (ICRP)", T.SOFT, T.THYROID, "H2O", "Ca5(PO4)3F"]


def _material_scene(lib):
    sc = S.Scene(lib)
    sc.world((4, 4, 4), (1, 1, 1))
    for m in MATERIALS:
        sc.add_material(m, 1.1)
    for z in (13, 29, 74, 82):
        sc.add_element(z)
    n_mat = len(MATERIALS) + 4
    mat = (np.arange(64) % n_mat).astype(np.uint8)
    sc.arrays(np.linspace(0.5, 2.0, 64).astype(np.float32), mat)
    assert sc.validate()
    return sc, n_mat


def test_material_queries_bit_exact(product, reference):
    a, n = _material_scene(product)
    b, _ = _material_scene(reference)
    for i in range(n):
        for e in (1.0, 4.04, 10.0, 33.17, 60.0, 88.1, 150.0):
            assert T.bit_equal(a.material_attenuation(i, e), b.material_attenuation(i, e))
        for q in (0.0, 0.3, 2.5, 11.0):
            assert a.material_form_factor_sq(i, q) == b.material_form_factor_sq(i, q)
            assert a.material_scatter_factor(i, q) == b.material_scatter_factor(i, q)
        assert T.bit_equal(a.material_binding_energies(i, 1.0), b.material_binding_energies(i, 1.0))
        assert T.bit_equal(a.material_shells(i), b.material_shells(i))
        assert a.material_density(i) == b.material_density(i)


@pytest.mark.parametrize("max_energy", [30.0, 60.0, 150.0])
def test_lut_tables_bit_exact(product, reference, max_energy):
    a, _ = _material_scene(product)
    b, _ = _material_scene(reference)
    a.lut_generate(max_energy)
    b.lut_generate(max_energy)
    for what in range(6):
        ta, tb = a.lut_table(what), b.lut_table(what)
        assert ta.size > 0 and T.bit_equal(ta, tb), f"LUT table {what} differs at max energy {max_energy}"
    rng = np.random.default_rng(3)
    for e in rng.uniform(1.0, max_energy, 200).astype(np.float32):
        for m in (0, 3, 6, 11):
            assert T.bit_equal(a.lut_attenuation(m, e), b.lut_attenuation(m, e))
        assert a.lut_max_inverse(e) == b.lut_max_inverse(e)
    for q in rng.uniform(0, 12, 50).astype(np.float32):
        assert a.lut_scatter_factor(4, q) == b.lut_scatter_factor(4, q)
    assert T.bit_equal(a.lut_sample_form_factor(0, 9.0, (5, 7), 2000), b.lut_sample_form_factor(0, 9.0, (5, 7), 2000))


def _same_exposures(a, b, idx):
    assert a.total_exposures() == b.total_exposures()
    for i in idx:
        ea, eb = a.exposure(i), b.exposure(i)
        for k in ea:
            assert np.array_equal(np.asarray(ea[k]), np.asarray(eb[k])), f"exposure {i} field {k}: {ea[k]} vs {eb[k]}"


def test_pencil_and_isotropic_exposures(product, reference):
    _same_exposures(T.pencil_scene(product), T.pencil_scene(reference), range(4))
    _same_exposures(T.isotropic_scene(product), T.isotropic_scene(reference), range(3))
    a, b = T.isotropic_scene(product, ct=True, exposures=7), T.isotropic_scene(reference, ct=True, exposures=7)
    _same_exposures(a, b, range(7))
    for what in range(3):
        assert T.bit_equal(a.source_table(what), b.source_table(what))


@pytest.mark.parametrize("spiral", [True, False])
def test_ct_exposures_and_beam_tables_bit_exact(product, reference, spiral):
    a, b = T.ct_scene(product, spiral=spiral), T.ct_scene(reference, spiral=spiral)
    n = a.total_exposures()
    _same_exposures(a, b, list(range(0, n, max(1, n // 9))) + [n - 1])
    ea, wa = a.spectrum()
    eb, wb = b.spectrum()
    assert T.bit_equal(ea, eb) and T.bit_equal(wa, wb), "tube spectrum differs"
    for what in range(7):
        ta, tb = a.source_table(what), b.source_table(what)
        assert ta.size > 0 and T.bit_equal(ta, tb), f"beam table {what} differs"
    assert a.max_energy() == b.max_energy()


@pytest.mark.parametrize("build", [lambda lib: T.ct_dual_scene(lib, True), lambda lib: T.ct_dual_scene(lib, False), T.topogram_scene, T.cbct_scene],
                         ids=["dual_spiral", "dual_axial", "topogram", "cbct"])
def test_dual_source_topogram_and_cone_beam_exposures_bit_exact(product, reference, build):
    """Every remaining source type of the reference (source.hpp:679-787 cone beam, :1369-1600 dual source, :1618-1700 topogram):
    all exposures, evaluated from the product's parameter block on the host, equal the reference's getExposure bit for bit."""
    a, b = build(product), build(reference)
    _same_exposures(a, b, range(a.total_exposures()))


def test_dx_source(product, reference):
    def build(lib):
        sc = T.tissue_block(lib)
        sc.source_dx(voltage=90.0, al_mm=3.0, sdd=1000.0, field_size=(120.0, 90.0), source_angles_deg=(20.0, -10.0),
                     tube_rotation_deg=15.0, dap=2.0, histories=1000, exposures=5, position=(3.0, -2.0, 10.0))
        return sc

    a, b = build(product), build(reference)
    _same_exposures(a, b, range(5))
    for what in range(5):
        assert T.bit_equal(a.source_table(what), b.source_table(what))
    assert a.calibration() == b.calibration()


def test_pencil_calibration(product, reference):
    assert T.pencil_scene(product).calibration() == T.pencil_scene(reference).calibration()


@pytest.mark.parametrize("diameter", [160, 320])
def test_ctdi_phantom_identical(product, reference, diameter):
    a, b = S.Scene(product).ctdi_phantom(diameter), S.Scene(reference).ctdi_phantom(diameter)
    assert a.dim == b.dim
    assert a.validate() and b.validate()
    for x, y in zip(a.get_arrays(), b.get_arrays()):
        assert T.bit_equal(x, y)
    for pos in range(5):
        assert np.array_equal(a.ctdi_holes(pos), b.ctdi_holes(pos))
    da, sa, ea = a.dimensions()
    db, sb, eb = b.dimensions()
    assert da == db and T.bit_equal(sa, sb) and T.bit_equal(ea, eb)


def test_world_validation_rules(product, reference):
    """isValid() rejects what the reference rejects (world.hpp:174-227)."""
    for lib in (product, reference):
        sc = S.Scene(lib)
        sc.world((4, 4, 4), (1, 1, 1))
        sc.add_material("Water, Liquid")
        assert not sc.validate()  # no arrays
        sc.arrays(np.ones(64, np.float32), np.ones(64, np.uint8))
        assert not sc.validate()  # material index 1 without a second material
        sc.arrays(np.ones(64, np.float32), np.zeros(64, np.uint8))
        assert sc.validate()
        sc2 = S.Scene(lib)
        sc2.world((4, 4, 4), (1, 1, 1), cosines=(1, 0, 0, 0.5, 0.5, 0))  # not orthogonal
        sc2.add_material("Water, Liquid")
        sc2.arrays(np.ones(64, np.float32), np.zeros(64, np.uint8))
        assert not sc2.validate()
        with pytest.raises(S.SceneError):
            S.Scene(lib).world((2, 2, 2), (1, 1, 1)).add_material("NotAMaterial")
