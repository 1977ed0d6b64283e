"""Global alignment wrapper - drop-in for reference dust3r/cloud_opt_flow/__init__.py:14-33."""
from enum import Enum

from .optimizer import PointCloudOptimizer, global_alignment_loop


class GlobalAlignerMode(Enum):
    PointCloudOptimizer = "PointCloudOptimizer"
    ModularPointCloudOptimizer = "ModularPointCloudOptimizer"
    PairViewer = "PairViewer"


def global_aligner(dust3r_output, device, mode=GlobalAlignerMode.PointCloudOptimizer, **optim_kw):
    """Build the scene optimiser from the pairwise inference dict (reference __init__.py:20-33).

    ``GlobalAlignerMode.PointCloudOptimizer`` is the mode every caller of the flow-aware aligner uses for more than two
    images (tool/demo.py:150, tool/pose_test.py:169); ``PairViewer`` is what they select for two (tool/demo.py:151-153).
    ``ModularPointCloudOptimizer`` (no caller) raises NotImplementedError, as the reference does for unknown modes.
    """
    view1, view2, pred1, pred2 = [dust3r_output[k] for k in 'view1 view2 pred1 pred2'.split()]
    if mode == GlobalAlignerMode.PointCloudOptimizer:
        net = PointCloudOptimizer(view1, view2, pred1, pred2, device=device, **optim_kw).to(device)
    elif mode == GlobalAlignerMode.PairViewer:
        from .pair_viewer import PairViewer
        net = PairViewer(view1, view2, pred1, pred2, device=device, **optim_kw).to(device)
    elif mode == GlobalAlignerMode.ModularPointCloudOptimizer:
        raise NotImplementedError(f'{mode} is outside the B200 hot path (SURVEY.md section 8(b)); '
                                  'use GlobalAlignerMode.PointCloudOptimizer')
    else:
        raise NotImplementedError(f'Unknown mode {mode}')
    return net
