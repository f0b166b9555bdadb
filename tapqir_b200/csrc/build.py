"""
Builds tapqir_b200/lib/libtapqir_b200.so (the C-ABI library of include/tapqir_b200.h) with nvcc for
sm_100a only.  Run directly (``python tapqir_b200/csrc/build.py``) or through
``__graft_entry__.build()``.  The .so is git-ignored but travels to the GPU box with the tree.
"""

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_DIR = HERE.parent / "lib"
LIB = LIB_DIR / "libtapqir_b200.so"
SOURCES = sorted(HERE.glob("*.cu"))
HEADERS = sorted(HERE.glob("*.cuh")) + [HERE.parent.parent / "include" / "tapqir_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; the sm_100a library cannot be built")
    return exe


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force=False, verbose=False, variant=None, defines=()):
    """``variant`` / ``defines``: an experimental build with extra -D flags into lib/libtapqir_b200.<variant>.so (selected
    at run time with TQ_LIB=<path>; used for A/B timings of kernel variants, never shipped as the default)."""
    lib = LIB if variant is None else LIB_DIR / f"libtapqir_b200.{variant}.so"
    if variant is None and not force and not needs_build():
        return LIB
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / ("obj" if variant is None else f"obj_{variant}")
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = obj_dir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src.name}")
        objs.append(str(obj))
    (LIB_DIR / ("ptxas.log" if variant is None else f"ptxas.{variant}.log")).write_text("\n".join(log))
    cmd = [nvcc, "-shared", "-o", str(lib), *objs, "-lcudart"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed")
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    variant = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), None)
    defines = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant, defines=defines))
