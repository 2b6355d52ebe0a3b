"""Several GPUs behind the one Transport::operator() call (Transport::setDevices / DXMCB200_DEVICES): one host thread per
device, exposures interleaved, ONE NCCL reduce-scatter of the fixed-point grids over voxel slices, every device decodes and
downloads its own slice. The Result must be bit-identical to the single-GPU Result for the same seed. Needs two GPUs
(skipped on a one-GPU box; the gloo tests in tests/test_multirank_gloo.py cover the partition logic on the CPU)."""
import numpy as np
import pytest

import support as T
from dxmclib_b200 import cabi
from dxmclib_b200 import scene as S

pytestmark = pytest.mark.gpu


def _two_gpus():
    return cabi.device_count() >= 2


@pytest.mark.skipif(not T.have_gpu() or not _two_gpus(), reason="needs two CUDA devices")
@pytest.mark.parametrize("output,calibrate", [(S.OUT_EV_PER_HISTORY, False), (S.OUT_DOSE, False)])
def test_two_devices_bit_identical_to_one(gpu, product, output, calibrate):
    one = T.ct_scene(product, histories=40000).transport(model=1, output=output, use_calibration=calibrate, seed=T.SEED)
    sc = T.ct_scene(product, histories=40000)
    sc.b200_set_devices([0, 1])
    two = sc.transport(model=1, output=output, use_calibration=calibrate, seed=T.SEED)
    assert one.histories == two.histories and one.units == two.units
    assert T.bit_equal(one.n_events, two.n_events)
    assert T.bit_equal(one.dose, two.dose) and T.bit_equal(one.variance, two.variance)
    assert one.n_events.sum() > 100000


@pytest.mark.skipif(not T.have_gpu() or not _two_gpus(), reason="needs two CUDA devices")
def test_two_devices_uneven_split_and_small_grid(gpu, product):
    """An odd number of exposures and a grid whose voxel count is odd: slices of unequal size, padding behind the grid."""
    def scene():
        sc = T.pencil_scene(product, n=31, histories=30000, exposures=5)
        return sc
    one = scene().transport(model=1, output=S.OUT_EV_PER_HISTORY, seed=7)
    sc = scene()
    sc.b200_set_devices([1, 0])
    two = sc.transport(model=1, output=S.OUT_EV_PER_HISTORY, seed=7)
    assert T.bit_equal(one.n_events, two.n_events) and T.bit_equal(one.dose, two.dose)
