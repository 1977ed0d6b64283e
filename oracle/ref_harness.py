"""TEST INFRASTRUCTURE - runs the UNMODIFIED reference global aligner from /root/reference on CPU.

The reference tree is ``/root/reference`` in the build container and its verbatim git-ignored copy ``oracle/_ref``
(made by ``oracle/build_ref.py``) on the GPU box.  Used by
``oracle/make_golden.py`` to generate the committed fixtures in ``tests/golden/`` and by the
``not gpu`` tests to cross-check ``oracle/ga_oracle.py`` live when the reference tree is present.

Five third-party modules the reference imports at module scope are absent from the image and are
stubbed (they are never *called* on the alignment path): matplotlib(.pyplot), seaborn, evo.*,
sam2.build_sam.  ``roma`` is restated in oracle/roma_shim.py.  RAFT weights are absent, so
``PointCloudOptimizer.get_flow`` is overridden to return injected flow tensors.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_root() -> str:
    """/root/reference in the build container; on the GPU box the verbatim, git-ignored copy oracle/_ref made by
    oracle/build_ref.py (it travels with the snapshot)."""
    env = os.environ.get("A3R_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "dust3r", "cloud_opt_flow")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dust3r", "cloud_opt_flow"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_IMPORTED = {}


def import_reference():
    """Returns (global_aligner, GlobalAlignerMode, PointCloudOptimizer, base_opt_module)."""
    if _IMPORTED:
        return _IMPORTED["v"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import roma_shim
    sys.modules["roma"] = roma_shim

    def _noop(*a, **k):
        return None

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            plt = _stub("matplotlib.pyplot", grid=_noop, ion=_noop, imsave=_noop, figure=_noop, savefig=_noop,
                        title=_noop, xlabel=_noop, ylabel=_noop)
            mpl = _stub("matplotlib", pyplot=plt)
            mpl.scale = _stub("matplotlib.scale", scale_factory=_noop)       # dust3r/cloud_opt/__init__.py:10 (unused import)
    if "seaborn" not in sys.modules:
        try:
            import seaborn  # noqa: F401
        except Exception:
            _stub("seaborn", heatmap=_noop)

    class _Dummy:
        translation_part = rotation_angle_deg = frames = None

    evo_names = dict(PoseRelation=_Dummy, Unit=_Dummy, PosePath3D=_Dummy, PoseTrajectory3D=_Dummy)
    core = _stub("evo.core", **evo_names)
    for sub in ("sync", "metrics", "trajectory"):
        setattr(core, sub, _stub(f"evo.core.{sub}", **evo_names))
    tools = _stub("evo.tools")
    tools.file_interface = _stub("evo.tools.file_interface")
    tools.plot = _stub("evo.tools.plot")
    _stub("evo", core=core, tools=tools, main_ape=_stub("evo.main_ape"), main_rpe=_stub("evo.main_rpe"))
    if "wandb" not in sys.modules:      # dust3r/cloud_opt/__init__.py:11 imports torch through wandb (an IDE auto-import)
        import torch as _torch
        wb = _stub("wandb")
        wb.wandb_torch = _stub("wandb.wandb_torch", torch=_torch)
    sam2 = _stub("sam2")
    sam2.build_sam = _stub("sam2.build_sam", build_sam2_video_predictor=_noop)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from dust3r.cloud_opt_flow import global_aligner, GlobalAlignerMode
        from dust3r.cloud_opt_flow.optimizer import PointCloudOptimizer
        import dust3r.cloud_opt_flow.base_opt as base_opt
    _IMPORTED["v"] = (global_aligner, GlobalAlignerMode, PointCloudOptimizer, base_opt)
    return _IMPORTED["v"]


@contextlib.contextmanager
def quiet():
    """The reference prints the flow loss every iteration (optimizer.py:539) and tqdm bars."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def build_reference_scene(scene, **kw):
    """Construct the reference PointCloudOptimizer on CPU for an align3r_b200.synth.SyntheticScene."""
    import copy
    import torch
    global_aligner, Mode, PCO, _ = import_reference()
    flow_on = kw.get("flow_loss_weight", 0) > 0
    if flow_on:
        E, _, H, W = scene.flow_ij.shape
        ones = torch.ones(E, 1, H, W, dtype=torch.bool)
        fij, fji = scene.flow_ij.clone(), scene.flow_ji.clone()
        PCO.get_flow = lambda self, sintel_ckpt=False: (fij, fji, ones, ones.clone())
    out = copy.deepcopy(scene.output)
    with quiet():
        ref = global_aligner(out, "cpu", mode=Mode.PointCloudOptimizer, **kw)
    return ref


def to_double(ref):
    """fp64 arbiter mode of the reference."""
    import torch
    ref.double()
    if hasattr(ref, "flow_ij"):
        ref.flow_ij = ref.flow_ij.double()
        ref.flow_ji = ref.flow_ji.double()
    H, W = ref.imshape
    ref.depth_wrapper.generate_grid(H, W, "cpu")
    ref.depth_wrapper.coord = ref.depth_wrapper.coord.double()
    return ref
