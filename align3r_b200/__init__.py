"""align3r_b200 - B200-native global alignment (the dust3r/cloud_opt_flow PointCloudOptimizer loop).

    from align3r_b200.cloud_opt_flow import global_aligner, GlobalAlignerMode
    scene = global_aligner(output, device, mode=GlobalAlignerMode.PointCloudOptimizer, **kw)
    loss = scene.compute_global_alignment(init='mst', niter=300, schedule='linear', lr=0.01)
"""
__version__ = "0.1.0"
