"""Global alignment wrapper - drop-in for reference dust3r/cloud_opt_flow/__init__.py:14-33."""
from enum import Enum

from .optimizer import PointCloudOptimizer, global_alignment_loop

GlobalAlignerMode = Enum("GlobalAlignerMode", {name: name for name in ("PointCloudOptimizer", "ModularPointCloudOptimizer", "PairViewer")})
GlobalAlignerMode.__doc__ = "Members and values as in the reference (cloud_opt_flow/__init__.py:14-17)."


def _pair_viewer():
    from .pair_viewer import PairViewer
    return PairViewer


_BUILDERS = {
    # the mode every caller of the flow-aware aligner uses for more than two images (tool/demo.py:150, tool/pose_test.py:169)
    GlobalAlignerMode.PointCloudOptimizer: lambda: PointCloudOptimizer,
    # what the tools select for two images (tool/demo.py:151-153)
    GlobalAlignerMode.PairViewer: _pair_viewer,
}


def global_aligner(dust3r_output, device, mode=GlobalAlignerMode.PointCloudOptimizer, **optim_kw):
    """Build the scene optimiser from the pairwise inference dict (reference __init__.py:20-33): the dict's four entries
    become the optimiser's positional arguments, everything else is passed through.  ``ModularPointCloudOptimizer`` (no
    caller in the reference's tools) raises NotImplementedError, as the reference does for unknown modes."""
    if mode not in _BUILDERS:
        if mode == GlobalAlignerMode.ModularPointCloudOptimizer:
            raise NotImplementedError(f'{mode} is outside the B200 hot path (SURVEY.md section 8(b)); '
                                      'use GlobalAlignerMode.PointCloudOptimizer')
        raise NotImplementedError(f'Unknown mode {mode}')
    cls = _BUILDERS[mode]()
    args = [dust3r_output[k] for k in ('view1', 'view2', 'pred1', 'pred2')]
    return cls(*args, device=device, **optim_kw).to(device)
