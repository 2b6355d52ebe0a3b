"""On-device exposure generation: exposureKernel evaluates exposure i of a source from its parameter block for
all i at once (dxmcb200_generate_exposures). The table it makes is compared with what the UNMODIFIED reference returns from
Source::getExposure(i) + Exposure::alignToDirectionCosines for every source type the reference has. The host evaluation of the
same block is bit-identical to the reference (tests/test_host_tables_parity.py); the device evaluates sin / cos / atan in
double and rounds, where the host calls the float libm functions, so the device table may differ in the last units:
positions within 4e-7 of the source-isocentre distance, unit vectors within 4e-7, angles and weights within 1e-6 relative."""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import cabi
from dxmclib_b200 import scene as S

pytestmark = pytest.mark.gpu

SCENES = {
    "pencil": lambda lib: T.pencil_scene(lib),
    "isotropic_ct": lambda lib: T.isotropic_scene(lib, ct=True, exposures=19),
    "dx": lambda lib: T.dx_slab_scene(lib, histories=100, exposures=3),
    "ct_spiral_aec_xcare_tilt": lambda lib: T.ct_scene(lib, spiral=True, histories=10),
    "ct_axial": lambda lib: T.ct_scene(lib, spiral=False, histories=10, tilt=0.0),
    "dual_spiral": lambda lib: T.ct_dual_scene(lib, True),
    "dual_axial": lambda lib: T.ct_dual_scene(lib, False),
    "topogram": T.topogram_scene,
    "cbct": T.cbct_scene,
    "rotated_world": lambda lib: _rotated(lib),
}


def _rotated(lib):
    c, s = np.cos(0.3), np.sin(0.3)
    sc = T.tissue_block(lib, cosines=(c, s, 0, -s, c, 0))
    sc.source_ct(spiral=True, voltage=100.0, al_mm=5.0, collimation=20.0, scan_length=60.0, position=(5, 6, -20), exposure_step_deg=20.0,
                 histories=10, gantry_tilt_deg=3.0)
    return sc


@pytest.mark.parametrize("name", list(SCENES))
def test_device_exposure_table_matches_reference(gpu, product, reference, name):
    a, b = SCENES[name](product), SCENES[name](reference)
    n = a.total_exposures()
    assert n == b.total_exposures()
    a.b200_prepare(device=0, model=1, seed=3)
    ctx = cabi.Context(handle=a.b200_context())
    table = ctx.resident_exposures(n)  # the table the kernels read
    scale = max(600.0, float(np.abs(table["position"]).max()))
    for i in range(n):
        want = b.exposure(i)
        got = table[i]
        np.testing.assert_allclose(got["position"], want["position"], rtol=0, atol=4e-7 * scale)
        np.testing.assert_allclose(got["cosines"], want["cosines"], rtol=0, atol=4e-7)
        np.testing.assert_allclose(got["beam_direction"], want["beam_direction"], rtol=0, atol=4e-7)
        np.testing.assert_allclose(got["collimation"], want["collimation"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got["weight"], want["weight"], rtol=2e-6)
        assert got["mono_energy"] == want["mono_energy"] and int(got["histories"]) == want["histories"]
        assert (got["spectrum"] >= 0) == bool(want["has_spectrum"]) and (got["heel"] >= 0) == bool(want["has_heel"])
        assert (got["bowtie"] >= 0) == bool(want["has_bowtie"])
    a.b200_release()


def test_transport_from_device_generated_exposures_matches_reference(gpu, product, reference):
    """End to end through Transport::operator(): a dual-source spiral scan, exposures made on the device, against the reference."""
    a = T.ct_dual_scene(product, True, histories=200000).transport(model=1, output=S.OUT_EV_PER_HISTORY, seed=T.SEED)
    b = T.ct_dual_scene(reference, True, histories=200000).transport(model=1, output=S.OUT_EV_PER_HISTORY, seed=T.SEED, workers=S.WORKERS_COUNTER_STREAMS)
    assert a.histories == b.histories
    ta, tb = float(a.dose.astype(np.float64).sum()), float(b.dose.astype(np.float64).sum())
    assert abs(ta - tb) / tb < 5e-3
