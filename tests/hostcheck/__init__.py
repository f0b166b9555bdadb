"""
Builds and loads the CPU harness (hostcheck.cpp) around the csrc/ arithmetic headers -- test
infrastructure only.  ``load()`` returns a ctypes handle or raises pytest.skip if g++ is missing.
"""

import ctypes
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE.parent.parent / "tapqir_b200" / "csrc"
SO = HERE / "_hostcheck.so"
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [HERE / "hostcheck.cpp"] + sorted(CSRC.glob("*.cuh"))
    if not SO.exists() or any(s.stat().st_mtime > SO.stat().st_mtime for s in srcs):
        gxx = shutil.which("g++")
        if gxx is None:
            import pytest

            pytest.skip("g++ not available")
        cmd = [gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", str(HERE / "hostcheck.cpp"), "-I", str(CSRC),
               "-I", "/usr/local/cuda/include", "-o", str(SO)]
        subprocess.run(cmd, check=True)
    _lib = ctypes.CDLL(str(SO))
    return _lib
