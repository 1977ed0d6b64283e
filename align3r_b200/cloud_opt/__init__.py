"""The plain (flow-free) global aligner - drop-in for reference dust3r/cloud_opt/__init__.py:19-39, the variant
tool/depth_test.py runs for the video-depth protocol (depth_test.py:29, 642, 662).

It is the same optimisation path with the flow / temporal / depth-prior terms off (reference cloud_opt/optimizer.py:223-241
against cloud_opt_flow/optimizer.py:500-557), a cosine schedule ending at ``lr_min = 1e-6`` (cloud_opt/base_opt.py:424 against
1e-3) and, typically, non-symmetrised complete clip graphs: the same kernels, flags off.  The reference's signature carries two
extra positional arguments, ``if_use_mono`` and ``mono_depths`` (depth = mono_depth * exp(scalemap) + shift,
cloud_opt/optimizer.py:29-33,178-184); every caller passes ``if_use_mono=False`` (depth_test.py:572).  The mono-depth
parameterisation is not built.
"""
from ..cloud_opt_flow import GlobalAlignerMode
from ..cloud_opt_flow.optimizer import PointCloudOptimizer as _FlowPointCloudOptimizer
from ..cloud_opt_flow.optimizer import global_alignment_loop as _loop


class PointCloudOptimizer(_FlowPointCloudOptimizer):
    """reference dust3r/cloud_opt/optimizer.py:18-241 (``if_use_mono=False``)."""

    def __init__(self, view1, view2, pred1, pred2, if_use_mono=False, mono_depths=None, **kw):
        if if_use_mono:
            raise NotImplementedError("if_use_mono=True (mono-depth scale/shift parameterisation, reference "
                                      "cloud_opt/optimizer.py:29-33,178-184) is not built; every caller passes False")
        for k in ("flow_loss_weight", "temporal_smoothing_weight", "depth_regularize_weight"):
            if kw.get(k):
                raise TypeError(f"{k} belongs to cloud_opt_flow.PointCloudOptimizer")
        self.if_use_mono, self.mono_depths = False, mono_depths
        super().__init__(view1, view2, pred1, pred2, flow_loss_weight=0.0, temporal_smoothing_weight=0,
                         depth_regularize_weight=0.0, **kw)

    def compute_global_alignment(self, init=None, init_priors=None, niter_PnP=10, **kw):
        kw.setdefault("lr_min", 1e-6)                       # cloud_opt/base_opt.py:424
        return super().compute_global_alignment(init=init, init_priors=init_priors, niter_PnP=niter_PnP, **kw)

    def forward(self, epoch=9999):                          # the reference's forward() takes no epoch
        return super().forward(epoch)


def global_alignment_loop(net, lr=0.01, niter=300, schedule='cosine', lr_min=1e-6):
    return _loop(net, lr=lr, niter=niter, schedule=schedule, lr_min=lr_min)


def global_aligner(dust3r_output, if_use_mono, mono_depths, device, mode=GlobalAlignerMode.PointCloudOptimizer, **optim_kw):
    view1, view2, pred1, pred2 = [dust3r_output[k] for k in 'view1 view2 pred1 pred2'.split()]
    if mode == GlobalAlignerMode.PointCloudOptimizer:
        return PointCloudOptimizer(view1, view2, pred1, pred2, if_use_mono, mono_depths, device=device, **optim_kw).to(device)
    if mode == GlobalAlignerMode.PairViewer:                # reference cloud_opt/pair_viewer.py = cloud_opt_flow/pair_viewer.py
        from ..cloud_opt_flow.pair_viewer import PairViewer
        return PairViewer(view1, view2, pred1, pred2, device=device, **optim_kw).to(device)
    if mode == GlobalAlignerMode.ModularPointCloudOptimizer:
        raise NotImplementedError(f'{mode} is outside the B200 hot path (SURVEY.md section 8(b))')
    raise NotImplementedError(f'Unknown mode {mode}')
