"""The C-ABI shared library loads and exports every symbol include/a3r_ga.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from align3r_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "a3r_ga.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(a3r_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    path = build.build()
    assert os.path.exists(path) and path == _lib.LIB_PATH


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(build.build())
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/a3r_ga.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_versions_and_constants_agree_with_the_header():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "a3r_ga.h")).read()
    get = lambda name: int(re.search(rf"#define {name}\s+\(?(-?\d+)", hdr).group(1))
    assert lib.a3r_abi_version() == get("A3R_ABI_VERSION") == _lib.ABI_VERSION
    assert get("A3R_SLOT_ACC") == _lib.SLOT_ACC and get("A3R_FRAME_ACC") == _lib.FRAME_ACC and get("A3R_SCAL") == _lib.SCAL
    assert lib.a3r_tile_pixels() > 0 and lib.a3r_max_slots() >= 20      # swin-5 frames carry 20 slots
    for bit, name in ((_lib.F_SHARED_FOCAL, "A3R_F_SHARED_FOCAL"), (_lib.F_HAS_FLOW, "A3R_F_HAS_FLOW"),
                      (_lib.F_TRAIN_ADAPTORS, "A3R_F_TRAIN_ADAPTORS")):
        assert bit == 1 << int(re.search(rf"#define {name}\s+\(1u << (\d+)\)", hdr).group(1))


def test_struct_layout_matches_the_compiler():
    """sizeof(A3rProblem) as gcc lays it out == ctypes' layout."""
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sz.c")
        open(c, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "a3r_ga.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                           'sizeof(A3rProblem), offsetof(A3rProblem, pix_stride), offsetof(A3rProblem, edge_ij),'
                           'offsetof(A3rProblem, loss_hist));return 0;}\n')
        exe = os.path.join(d, "sz")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        size, o1, o2, o3 = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    P = _lib.A3rProblem
    assert (ctypes.sizeof(P), P.pix_stride.offset, P.edge_ij.offset, P.loss_hist.offset) == (size, o1, o2, o3)


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.a3r_ga_prepare(None, 0, None)
    assert rc == -1 and b"NULL" in lib.a3r_last_error()
    with pytest.raises(RuntimeError, match="a3r_ga_prepare"):
        _lib.check(rc, "a3r_ga_prepare")
    assert lib.a3r_occ_masks(None, 128, None, None, 1, 1, 8, 3.0, None, None) == -1 and b"a3r_occ_masks" in lib.a3r_last_error()
    assert lib.a3r_occ_masks(None, 128, None, None, 1, 8, 8, 3.0, None, None) == -1 and b"NULL" in lib.a3r_last_error()
    assert lib.a3r_occ_masks(None, 128, None, None, 0, 8, 8, 3.0, None, None) == 0
    assert lib.a3r_occ_masks(None, 32, None, None, 1, 8, 8, 3.0, None, None) == -1          # the image does not fit the planes
