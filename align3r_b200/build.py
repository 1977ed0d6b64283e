"""Builds liba3r_ga.so in-tree for sm_100a:  python -m align3r_b200.build  (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["a3r_common.cu", "ga_kernels.cu", "ga_handle.cu", "mst_kernels.cu", "motion_kernels.cu"]
HEADERS = ["a3r_common.cuh", "ga_math.cuh", "ptx_helpers.cuh", "ga_scan_v4.cuh", "ga_scan_v5.cuh", "ga_small.cuh"]


def lib_path():
    return os.environ.get("A3R_LIB") or os.path.join(HERE, "lib", "liba3r_ga.so")


def needs_build():
    out = lib_path()
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(HERE, "csrc", f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "a3r_ga.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    out = lib_path()
    if not force and not needs_build():
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"),
           "-o", out] + [os.path.join(HERE, "csrc", f) for f in SOURCES]
    for macro in ("A3R_V4_WARPS", "A3R_V4_SLOTS", "A3R_V4_STAGES", "A3R_V4_EVICT_FIRST", "A3R_V4_TILE", "A3R_V4_NOFLUSH"):      # tuning knobs of the scan kernel (csrc/ga_scan_v4.cuh)
        if os.environ.get(macro):
            cmd.append(f"-D{macro}={os.environ[macro]}")
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
