"""Transport::operator() with a ProgressBar, the way the reference's callers use it (validation/validation.cpp:269-290,
SURVEY 8b "Threading"): the call blocks one thread while another polls ProgressBar::getETA / computeDoseProgressImage
and may call setCancel. Semantics taken from the reference (transport.hpp:160-186, progressbar.hpp:59-109): progress
reaches 100 %, the live dose buffer is visible as a non-empty preview image while the run is in flight, a cancelled run
returns all zeros with numberOfHistories == 0. The same monitor code (include/dxmcb200_scene_monitor.hpp) drives both
implementations."""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import scene as S


def _scene(lib, histories, exposures):
    return T.isotropic_scene(lib, histories=histories, exposures=exposures)


@pytest.mark.skipif(not T.have_reference(), reason="oracle/_ref not built")
def test_reference_progress_and_cancel_semantics():
    lib = S.reference_lib()
    res, rep = _scene(lib, 40000, 40).transport_monitored(model=S.MODEL_LIVERMORE)
    assert res.histories == 40 * 40000 and res.dose.sum() > 0
    assert rep["percent_final"] == 100.0 and not rep["cancelled"]
    assert rep["images_nonzero"] >= 1 and (rep["image_width"], rep["image_height"]) == (24, 28)  # MIP along y: nx x nz
    res, rep = _scene(lib, 40000, 40).transport_monitored(model=S.MODEL_LIVERMORE, cancel_at_percent=20.0)
    assert rep["cancelled"] and rep["percent_final"] < 100.0
    assert res.histories == 0 and not res.dose.any() and not res.n_events.any() and not res.variance.any()


@pytest.mark.gpu
def test_product_progress_preview_does_not_disturb_the_result(gpu, product, monkeypatch):
    monkeypatch.setenv("DXMCB200_BATCH", "8,14")  # 16 K-photon waves: hundreds of progress call-backs per run
    plain = _scene(product, 200000, 64).transport(model=S.MODEL_LIVERMORE, seed=77)
    res, rep = _scene(product, 200000, 64).transport_monitored(model=S.MODEL_LIVERMORE, seed=77)
    assert rep["percent_final"] == 100.0 and not rep["cancelled"]
    assert rep["images_nonzero"] >= 1 and (rep["image_width"], rep["image_height"]) == (24, 28)
    # the live-dose refresh reads the accumulators while waves are in flight; it must not change them
    assert res.histories == plain.histories == 64 * 200000
    assert T.bit_equal(res.dose, plain.dose) and T.bit_equal(res.n_events, plain.n_events) and T.bit_equal(res.variance, plain.variance)


@pytest.mark.gpu
def test_product_cancel_returns_zero_result(gpu, product, monkeypatch):
    monkeypatch.setenv("DXMCB200_BATCH", "8,14")
    res, rep = _scene(product, 200000, 64).transport_monitored(model=S.MODEL_LIVERMORE, seed=77, cancel_at_percent=20.0)
    assert rep["cancelled"] and rep["percent_final"] < 100.0
    assert res.histories == 0 and not res.dose.any() and not res.n_events.any() and not res.variance.any()
    # the context is reusable after a cancelled run
    again = _scene(product, 20000, 4).transport(model=S.MODEL_LIVERMORE, seed=77)
    assert again.histories == 80000 and again.dose.sum() > 0
