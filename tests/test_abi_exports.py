"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

import support as T
from dxmclib_b200 import cabi
from dxmclib_b200 import scene as S


def declared_functions(header):
    text = open(os.path.join(T.ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dxmcb200_[a-z0-9_]+|dxs_[a-z0-9_]+)\s*\(", text)))


def test_device_abi_symbols_exported():
    lib = ctypes.CDLL(S.PRODUCT_LIB)
    names = [n for n in declared_functions("dxmcb200.h") if n not in ("dxmcb200_progress_cb",)]
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"libdxmcb200.so does not export {n}"
    assert sorted(cabi.CABI_SYMBOLS) == names


def test_scene_abi_symbols_exported_by_product_and_reference():
    names = declared_functions("dxmcb200_scene.h")
    assert sorted(S.SCENE_SYMBOLS) == names
    libs = [ctypes.CDLL(S.PRODUCT_LIB)]
    if T.have_reference():
        libs.append(ctypes.CDLL(S.REFERENCE_LIB))
    for lib in libs:
        for n in names:
            assert hasattr(lib, n), f"{n} missing"


def test_backend_names(product):
    assert S.Scene(product).backend == "dxmc-b200"
    if T.have_reference():
        assert S.Scene(S.reference_lib()).backend == "dxmclib-reference"


@pytest.mark.skipif(T.have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(product):
    """Without a CUDA device the product fails loudly instead of computing on the host."""
    with pytest.raises(cabi.CabiError):
        cabi.Context(0)
    sc = T.pencil_scene(product, histories=10, exposures=1)
    with pytest.raises(S.SceneError) as e:
        sc.transport()
    assert "-4" in str(e.value) or "CUDA" in str(e.value)


def test_history_stream_matches_oracle():
    from oracle import pyoracle

    for key in [(0, 0, 0), (1, 2, 3), (0xD1C02026, 3599, 2777777), (2**63 + 5, 2**40, 2**33)]:
        assert cabi.history_stream(*key) == pyoracle.history_stream(*key)
        assert cabi.history_stream(*key)[1] & 1 == 1
