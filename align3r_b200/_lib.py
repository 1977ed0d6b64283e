"""ctypes binding of liba3r_ga.so (C ABI: include/a3r_ga.h).  There is no fallback: if the shared
library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("A3R_LIB") or os.path.join(HERE, "lib", "liba3r_ga.so")   # A3R_LIB: tuning builds
ABI_VERSION = 7

# A3rProblem.flags
F_SHARED_FOCAL, F_NORM_PW_SCALE, F_DIST_L2 = 1 << 0, 1 << 1, 1 << 2
F_TRAIN_PW_POSES, F_TRAIN_DEPTH, F_TRAIN_IM_POSES, F_TRAIN_FOCALS = 1 << 3, 1 << 4, 1 << 5, 1 << 6
F_TRAIN_PP, F_TRAIN_ADAPTORS, F_HAS_FLOW, F_ROW_TILES = 1 << 7, 1 << 8, 1 << 9, 1 << 10
# mode bits
M_FLOW_ACTIVE, M_APPLY_ADAM = 1 << 0, 1 << 1
SLOT_ACC, FRAME_ACC, PRIOR_ACC, GLOB_ACC, SCAL = 28, 12, 4, 8, 16
PAIR_ROW, CAM_ROW = 32, 20
S_LOSS, S_ALIGN, S_TEMPORAL, S_FLOW, S_FLOW_DROPPED, S_SC_I, S_SC_J, S_FLAG_STICKY = range(8)
S_PRIOR = 10
CONF_MODES = {"log": 0, "sqrt": 1, "m1": 2, "id": 3, "none": 3}

_i32, _u32, _i64, _f32, _vp = C.c_int32, C.c_uint32, C.c_int64, C.c_float, C.c_void_p

_POINTER_FIELDS = [
    "edge_ij", "frame_hw",
    "pw_poses", "pw_poses_m", "pw_poses_v", "pw_adaptors", "pw_adaptors_m", "pw_adaptors_v",
    "im_poses", "im_poses_m", "im_poses_v", "im_focals", "im_focals_m", "im_focals_v",
    "im_pp", "im_pp_m", "im_pp_v",
    "g_pw_poses", "g_pw_adaptors", "g_im_poses", "g_im_focals", "g_im_pp",
    "frame_par", "edge_par", "slot_par",
    "lframe_id", "wf_lframe", "wf_slot_begin", "wf_tile_begin", "wf_accum", "lslot_gslot", "cta_tile_begin",
    "predw", "flow", "static_mask",
    "logdepth", "logdepth_m", "logdepth_v", "g_align", "g_flow_i", "g_flow_j",
    "acc_slot", "acc_frame", "frame_grad", "edge_scratch", "scal", "loss_hist",
    "logdepth_init", "prior_acc",
    "acc_glob", "acc_alt", "pair_scratch", "cam_cache", "inc_begin", "inc_items", "sync_ctr",
]


class A3rProblem(C.Structure):
    _fields_ = (
        [(n, _i32) for n in ("n_frames", "n_edges", "n_local_frames", "n_local_slots", "n_work_frames",
                             "n_tiles", "n_ctas", "tile_pixels")]
        + [("flags", _u32), ("pix_stride", _i64)]
        + [(n, _f32) for n in ("focal_break", "pw_break", "base_scale", "flow_loss_weight", "flow_loss_thre",
                               "pxl_thre", "temporal_smoothing_weight", "translation_weight", "inv_area_i",
                               "inv_area_j", "adam_beta1", "adam_beta2", "adam_eps", "depth_regularize_weight")]
        + [(n, _vp) for n in _POINTER_FIELDS]
    )


MAX_PEERS = 8


class A3rPeers(C.Structure):
    _fields_ = [("world", _i32), ("rank", _i32), ("n_acc", _i64), ("base", _vp * MAX_PEERS), ("win", _i32 * (MAX_PEERS + 1))]


class A3rAdamStep(C.Structure):
    _fields_ = [("step_size", _f32), ("bc2_sqrt", _f32), ("iter", _i32), ("_pad", _i32)]


class A3rGrads(C.Structure):
    _fields_ = [(n, _vp) for n in ("pw_poses", "pw_adaptors", "im_poses", "im_focals", "im_pp", "logdepth", "scal")]


_SIGNATURES = {
    "a3r_abi_version": (C.c_int, []),
    "a3r_last_error": (C.c_char_p, []),
    "a3r_tile_pixels": (C.c_int, []),
    "a3r_max_slots": (C.c_int, []),
    "a3r_max_ctas": (C.c_int, [C.c_int]),
    "a3r_scan_align_warps": (C.c_int, []),
    "a3r_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "a3r_pack_pred": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "a3r_pack_flow": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, _vp]),
    "a3r_ga_prepare": (C.c_int, [C.POINTER(A3rProblem), _u32, _vp]),
    "a3r_ga_accumulate": (C.c_int, [C.POINTER(A3rProblem), _u32, _vp]),
    "a3r_ga_accumulate_folded": (C.c_int, [C.POINTER(A3rProblem), _u32, C.POINTER(A3rAdamStep), _u32, _vp]),
    "a3r_ga_update": (C.c_int, [C.POINTER(A3rProblem), _u32, C.POINTER(A3rAdamStep), _vp]),
    "a3r_ga_run": (C.c_int, [C.POINTER(A3rProblem), _i32, C.POINTER(_u32), C.POINTER(A3rAdamStep), _vp]),
    "a3r_ga_unproject": (C.c_int, [C.POINTER(A3rProblem), _vp, _vp]),
    # handle layer (SURVEY.md section 8(b) export set)
    "a3r_ga_create": (C.c_int, [C.POINTER(A3rProblem), C.POINTER(_vp)]),
    "a3r_ga_destroy": (C.c_int, [_vp]),
    "a3r_ga_problem": (C.POINTER(A3rProblem), [_vp]),
    "a3r_ga_pack_inputs": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "a3r_ga_reset_adam": (C.c_int, [_vp, _vp]),
    "a3r_ga_step": (C.c_int, [_vp, _i32, _i32, _f32, _u32, _vp]),
    "a3r_ga_forward": (C.c_int, [_vp, _u32, _vp, _vp]),
    "a3r_ga_grads": (C.c_int, [_vp, C.POINTER(A3rGrads)]),
    "a3r_mst_edge_scores": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "a3r_peer_region_bytes": (C.c_size_t, [_i64]),
    "a3r_peer_alloc": (C.c_int, [_i64, C.POINTER(_vp), C.c_char_p]),
    "a3r_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "a3r_peer_close": (C.c_int, [_vp]),
    "a3r_peer_free": (C.c_int, [_vp]),
    "a3r_ga_run_peers": (C.c_int, [C.POINTER(A3rProblem), C.POINTER(A3rPeers), _u32, _i32, C.POINTER(_u32), C.POINTER(A3rAdamStep), _vp]),
    "a3r_clean_pointcloud": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _f32, _f32, _vp, _vp]),
    "a3r_clean_pointcloud_range": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _f32, _f32, _vp, _vp]),
    "a3r_procrustes_batched": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "a3r_weiszfeld_focal_batched": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "a3r_init_depth": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp]),
    "a3r_pnp_inliers": (C.c_int, [_vp, _vp, _i32, _i64, C.POINTER(C.c_double), _vp, _vp, _vp]),
    "a3r_similarity_points": (C.c_int, [_vp, _i64, C.POINTER(C.c_float), _vp, _vp]),
    "a3r_pnp_gather": (C.c_int, [_vp, _vp, C.POINTER(C.c_int32), _i32, _vp, _vp]),
    "a3r_occ_masks": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp]),
    "a3r_motion_minmax": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "a3r_debug_k2_trace": (C.c_int, [C.c_int, _vp]),
    "a3r_debug_timeline": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "a3r_motion_masks": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp, _vp, _vp]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load liba3r_ga.so (built by ``python -m align3r_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"align3r_b200: CUDA extension {LIB_PATH} is missing; run `python -m align3r_b200.build` "
                           "(there is no CPU or PyTorch fallback for the alignment kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    got = lib.a3r_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"align3r_b200: {LIB_PATH} has ABI {got}, python side expects {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().a3r_last_error().decode(errors="replace")
        raise RuntimeError(f"align3r_b200: {what} failed (code {rc}): {msg}")


def ptr(t):
    """Device (or host) address of a torch tensor, or NULL for None."""
    return None if t is None else C.c_void_p(t.data_ptr())
