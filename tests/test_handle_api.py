"""Handle layer of the C ABI (include/a3r_ga.h: a3r_ga_create / destroy / pack_inputs / reset_adam / step / forward / grads,
a3r_mst_edge_scores - the export set SURVEY.md section 8(b) names).  Argument handling without a GPU; on the GPU box the
layer is compared with the stateless entry points the Python class drives (tests/handle_api_worker.py, own process)."""
import ctypes as C
import os
import subprocess
import sys

import pytest

from align3r_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_argument_errors_of_the_handle_layer():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.a3r_ga_create(None, C.byref(h)) == -1 and b"a3r_ga_create" in lib.a3r_last_error()
    empty = _lib.A3rProblem()
    assert lib.a3r_ga_create(C.byref(empty), C.byref(h)) == -1 and b"tile_pixels" in lib.a3r_last_error() and not h.value
    empty.tile_pixels = lib.a3r_tile_pixels()
    assert lib.a3r_ga_create(C.byref(empty), C.byref(h)) == -1 and b"bad sizes" in lib.a3r_last_error()
    empty.n_frames, empty.n_edges, empty.n_ctas, empty.pix_stride = 3, 4, 16, 4 * empty.tile_pixels
    assert lib.a3r_ga_create(C.byref(empty), C.byref(h)) == -1 and b"NULL" in lib.a3r_last_error() and not h.value
    assert lib.a3r_ga_destroy(None) == 0
    assert not lib.a3r_ga_problem(None)
    assert lib.a3r_ga_grads(None, None) == -1
    assert lib.a3r_ga_step(None, 0, 1, 0.01, 0, None) == -1 and b"handle is NULL" in lib.a3r_last_error()
    assert lib.a3r_ga_forward(None, 0, None, None) == -1
    assert lib.a3r_ga_reset_adam(None, None) == -1
    assert lib.a3r_ga_pack_inputs(None, None, None, None, 0, 1, None, None, 0, None, None, None) == -1
    assert lib.a3r_mst_edge_scores(None, None, None, -1, None, None) == -1
    assert lib.a3r_mst_edge_scores(None, None, None, 0, None, None) == 0
    assert lib.a3r_mst_edge_scores(None, None, None, 5, None, None) == -1
    assert C.sizeof(_lib.A3rGrads) == 7 * C.sizeof(C.c_void_p)


@pytest.mark.gpu
def test_handle_layer_against_the_stateless_entry_points(cuda_device):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "handle_api_worker.py")], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "handle layer: ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
