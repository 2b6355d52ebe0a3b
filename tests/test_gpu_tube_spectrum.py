"""the tube model's bremsstrahlung depth integrals on the device (dxmcb200_tube_bremsstrahlung, opt-in with
DXMCB200_DEVICE_SPECTRUM=1) against the host evaluation, which is bit-identical to the reference's
(tests/test_host_tables_parity.py): the source's normalised spectrum and its heel-effect table agree to 2e-5 relative, and the
device takes a fraction of the host's time."""
import time

import numpy as np
import pytest

import support as T
from dxmclib_b200 import scene as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("builder", ["ct", "dx"])
def test_device_spectrum_matches_host(gpu, product, builder, monkeypatch):
    def build():
        if builder == "ct":
            sc = T.ct_scene(product, histories=10)
        else:
            sc = T.dx_slab_scene(product, histories=10, exposures=1)
        t0 = time.perf_counter()
        heel = sc.source_table(4)  # Source::validate(): builds the tube spectrum, its alias table and the heel-effect table
        assert heel.size > 0
        return sc, time.perf_counter() - t0

    monkeypatch.setenv("DXMCB200_DEVICE_SPECTRUM", "0")
    host, t_host = build()
    monkeypatch.setenv("DXMCB200_DEVICE_SPECTRUM", "1")
    build()  # context creation and module load out of the timed call
    dev, t_dev = build()
    eh, wh = host.spectrum()
    ed, wd = dev.spectrum()
    assert T.bit_equal(eh, ed) and wh.size > 50
    np.testing.assert_allclose(wd, wh, rtol=2e-5, atol=1e-9)
    assert abs(float(wd.sum()) - 1.0) < 1e-5
    for what in (3, 4):  # heel table: header, weights
        a, b = host.source_table(what), dev.source_table(what)
        assert a.size == b.size and a.size > 0
        np.testing.assert_allclose(b, a, rtol=5e-5)
    print(f"{builder}: source set-up {t_host * 1e3:.0f} ms on the host, {t_dev * 1e3:.0f} ms with the device spectrum")
    assert t_dev < t_host
