"""Global alignment wrapper - drop-in for reference dust3r/cloud_opt_flow/__init__.py:14-33."""
from enum import Enum

from .optimizer import PointCloudOptimizer, global_alignment_loop


class GlobalAlignerMode(Enum):
    PointCloudOptimizer = "PointCloudOptimizer"
    ModularPointCloudOptimizer = "ModularPointCloudOptimizer"
    PairViewer = "PairViewer"


def global_aligner(dust3r_output, device, mode=GlobalAlignerMode.PointCloudOptimizer, **optim_kw):
    """Build the scene optimiser from the pairwise inference dict (reference __init__.py:20-33).

    Only ``GlobalAlignerMode.PointCloudOptimizer`` is built here: it is the mode every caller of the
    flow-aware aligner uses (tool/demo.py:150, tool/pose_test.py:169).  The other two members are kept
    so that existing call sites import cleanly; selecting them raises NotImplementedError, as the
    reference does for unknown modes.
    """
    view1, view2, pred1, pred2 = [dust3r_output[k] for k in 'view1 view2 pred1 pred2'.split()]
    if mode == GlobalAlignerMode.PointCloudOptimizer:
        net = PointCloudOptimizer(view1, view2, pred1, pred2, device=device, **optim_kw).to(device)
    elif mode in (GlobalAlignerMode.ModularPointCloudOptimizer, GlobalAlignerMode.PairViewer):
        raise NotImplementedError(f'{mode} is outside the B200 hot path (SURVEY.md section 8(b)); '
                                  'use GlobalAlignerMode.PointCloudOptimizer')
    else:
        raise NotImplementedError(f'Unknown mode {mode}')
    return net
