"""The xraylib code path of the host-side physics data layer (dxmclib_b200/host/matdb.cpp with -DDXMCB200_USE_XRAYLIB) is
type-checked against a declarations-only stand-in of xraylib 4's header on every run, so that it cannot rot while no real
xraylib is installed; and the library reports which backend it was built on."""
import os
import shutil
import subprocess

import support as T
from dxmclib_b200 import cabi


def test_matdb_compiles_against_the_xraylib_api():
    cxx = shutil.which("g++")
    assert cxx
    host = os.path.join(T.ROOT, "dxmclib_b200", "host")
    cmd = [cxx, "-std=c++20", "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration", "-DDXMCB200_USE_XRAYLIB",
           f"-I{os.path.join(T.ROOT, 'tests', 'stubs', 'xraylib')}", f"-I{host}", os.path.join(host, "matdb.cpp")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout


def test_backend_is_reported():
    name, approximate = cabi.physics_backend()
    assert name
    assert approximate == name.startswith("xrl_lite")
