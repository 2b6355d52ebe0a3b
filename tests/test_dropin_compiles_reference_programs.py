"""Source compatibility of the drop-in headers, checked with the reference's OWN programs: its example, its TG-195
validation program and the unit-test programs that stay on the public API are compiled and linked UNCHANGED (from /root/reference, nothing is copied) against
dxmclib_b200/include and libdxmcb200.so. Not covered: testtransport.cpp, which subclasses Transport to call the per-thread
interaction samplers (computeInteractions, comptonScatter, ...: protected internals of the CPU hot path that are CUDA
kernels here, tested through the C ABI instead), and testdxmclib.cpp / validatedxmclib.cpp, which include headers
(dxmc/transport.h, dxmc/attenuationlut.h) the reference itself no longer has."""
import os
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

import pytest

import support as T

REF = os.environ.get("DXMC_REFERENCE", "/root/reference")
PROGRAMS = ["examples/pencilbeam/pencilbeam.cpp", "validation/validation.cpp", "tests/testattenuationlut.cpp", "tests/testbeamfilters.cpp", "tests/testexposure.cpp",
            "tests/testinterpolation.cpp", "tests/testmaterial.cpp", "tests/testrandom.cpp", "tests/testsource.cpp", "tests/testtube.cpp",
            "tests/testvectormath.cpp", "tests/testworld.cpp"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "dxmc")), reason="the reference tree is not present here")
def test_reference_programs_compile_and_link_against_the_dropin_headers():
    lib_dir = os.path.join(T.ROOT, "dxmclib_b200")
    assert os.path.exists(os.path.join(lib_dir, "libdxmcb200.so"))
    with tempfile.TemporaryDirectory() as tmp:
        def build(rel):
            out = os.path.join(tmp, os.path.basename(rel)[:-4])
            cmd = ["g++", "-O0", "-std=c++20", "-pthread", "-Wno-narrowing", f"-I{T.ROOT}/include", f"-I{lib_dir}/include", f"-I{lib_dir}/host",
                   f"-I{T.ROOT}/oracle/xraylib_compat", os.path.join(REF, rel), "-o", out, f"-L{lib_dir}", "-ldxmcb200"]
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            return rel, p.returncode, p.stderr[-1500:]

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            results = list(pool.map(build, PROGRAMS))
    failed = [(rel, err) for rel, rc, err in results if rc != 0]
    assert not failed, failed
