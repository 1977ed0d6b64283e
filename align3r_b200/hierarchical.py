"""Hierarchical (key frames, then clips) alignment driver - the caller-side loop of the reference tools around the
alignment path (tool/demo.py:173-251 `get_reconstructed_scene_hierachical`, tool/depth_test.py:620-676), with the pairwise
network left to the caller:

    key frames = every ``clip_size``-th image, aligned on their complete (non-symmetrised) graph;
    every clip = ``clip_size`` consecutive images starting at a key frame, aligned on its complete graph with the MST
    initialisation started from the key frame's pose and focal (``init_priors``, reference init_im_poses.py:173-215);
    the clip trajectories are re-based on the key poses (``get_tum_poses(init_keypose)``, base_opt.py:305-330) and concatenated.

``inference(pairs)`` is the caller's pairwise network (reference dust3r/inference.py): it maps a list of (img_i, img_j)
dict pairs to the ``{'view1','view2','pred1','pred2'}`` dict the aligner takes; it may also return
``(output, (flow_ij, flow_ji))`` when the flow term is used (RAFT is outside this package).
"""
from __future__ import annotations

import os

import numpy as np


def choose_clip_size(n_imgs, clip_size=10):
    """tool/demo.py:199-201, depth_test.py:628-629: shrink the clip until no clip is left with a single image."""
    while n_imgs % clip_size == 1 or n_imgs % clip_size == 0 or clip_size > n_imgs:
        clip_size -= 1
    return clip_size


def my_make_pairs(imgs, clip_size, variant='depth_test'):
    """Returns (coarse_init_pairs, keyframes_id, all_clips_pairs, all_clips_id).  The reference has two drivers with two pair
    builders of this name:
      ``variant='depth_test'``  tool/depth_test.py:395-435 (flow-free aligner): complete, NON-symmetrised graphs for the key
                                frames and for every clip;
      ``variant='demo'``        tool/demo.py:357-397 (flow aligner): the key-frame pairs are symmetrised, a clip pairs image i
                                with i+1, i+3, i+5, ... (stride 2) and mirrors those pairs - the symmetric graphs
                                ``use_self_mask=True`` (demo.py's default) asserts on (optimizer.py:157).
    The key frames and the images of a clip are re-indexed from 0 (``'idx'``) on copies: the caller's dicts are left alone
    (the reference renumbers them in place)."""
    if variant not in ('depth_test', 'demo'):
        raise ValueError(f'bad {variant=}')
    demo = variant == 'demo'
    keyframes_id = list(range(0, len(imgs), clip_size))
    key_imgs = [dict(imgs[i], idx=k) for k, i in enumerate(keyframes_id)]
    coarse_init_pairs = [(key_imgs[a], key_imgs[b]) for a in range(len(key_imgs) - 1) for b in range(a + 1, len(key_imgs))]
    if demo:
        coarse_init_pairs += [(b, a) for a, b in coarse_init_pairs]
    all_clips_pairs, all_clips_id = [], []
    for start in keyframes_id:
        clip = imgs[start:start + clip_size]
        all_clips_id.append([im['idx'] for im in clip])
        local = [dict(im, idx=k) for k, im in enumerate(clip)]
        pairs = [(dict(local[a]), dict(local[b])) for a in range(len(local) - 1) for b in range(a + 1, len(local), 2 if demo else 1)]
        if demo:
            pairs += [(b, a) for a, b in pairs]
        all_clips_pairs.append(pairs)
    return coarse_init_pairs, keyframes_id, all_clips_pairs, all_clips_id


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


def hierarchical_alignment(imgs, inference, device='cuda', clip_size=10, niter=300, schedule='cosine', lr=0.05,
                           flow=False, conf_to=None, save_folder=None, pairs=None, **aligner_kw):
    """Key-frame alignment, then one alignment per clip started from its key frame.

    ``flow=False``: the flow-free aligner of tool/depth_test.py (``align3r_b200.cloud_opt``, lr 0.05, cosine) on that
    tool's pair graphs (complete, not symmetrised);
    ``flow=True``: ``align3r_b200.cloud_opt_flow`` as in tool/demo.py (pass lr=0.01 and its keyword arguments) on demo.py's
    pair graphs (symmetrised; clips pair every image with every second later one) - ``pairs=`` overrides the choice.
    ``conf_to``: depth_test.py:638-639,659-660 sets every confidence above 1 to 10 before the alignment (``conf_to=10``).
    Returns a dict: ``traj`` = [tum poses (M,7), timestamps (M,)], ``depths`` / ``confs`` lists over all frames of all
    clips, ``focals``, ``key_poses``, ``losses`` (key alignment first).  With ``save_folder`` the per-clip results are also
    written in the reference's formats (trajectory, intrinsics, depth / confidence maps)."""
    from .cloud_opt_flow import GlobalAlignerMode
    if flow:
        from .cloud_opt_flow import global_aligner as _ga
        from .cloud_opt_flow.writeout import save_trajectory_tum_format

        def build(out, fl):
            return _ga(out, device, mode=GlobalAlignerMode.PointCloudOptimizer, flow=fl, **aligner_kw)
    else:
        from .cloud_opt import global_aligner as _ga
        from .cloud_opt_flow.writeout import save_trajectory_tum_format

        def build(out, fl):
            return _ga(out, False, [], device, mode=GlobalAlignerMode.PointCloudOptimizer, **aligner_kw)

    def run_inference(pairs):
        res = inference(pairs)
        out, fl = res if isinstance(res, tuple) else (res, None)
        if conf_to is not None:
            for p in ('pred1', 'pred2'):
                c = out[p]['conf']
                c[c > 1] = conf_to
        return out, fl

    clip_size = choose_clip_size(len(imgs), clip_size)
    coarse_init_pairs, keyframes_id, all_clips_pairs, all_clips_id = my_make_pairs(
        imgs, clip_size, variant=pairs or ('demo' if flow else 'depth_test'))
    key_out, key_flow = run_inference(coarse_init_pairs)
    if len(imgs) > 2:                                        # tool/demo.py:204-216, depth_test.py:641-647
        scene = build(key_out, key_flow)
        losses = [scene.compute_global_alignment(init='mst', niter=niter, schedule=schedule, lr=lr)]
    else:
        scene = _ga(key_out, device, mode=GlobalAlignerMode.PairViewer, **aligner_kw) if flow else \
            _ga(key_out, False, [], device, mode=GlobalAlignerMode.PairViewer, **aligner_kw)
        losses = [float('nan')]
    key_poses = _np(scene.get_im_poses()).tolist()
    key_depths = [_np(d) for d in scene.get_depthmaps()]
    key_focals = _np(scene.get_focals()).tolist()

    traj = [np.zeros((0, 7)), np.zeros((0,))]
    depths, confs, focals = [], [], []
    offset = 0
    if save_folder is not None:
        os.makedirs(save_folder, exist_ok=True)
    for key_pose, key_depth, key_focal, clip_pairs in zip(key_poses, key_depths, key_focals, all_clips_pairs):
        if not clip_pairs:
            continue
        out, fl = run_inference(clip_pairs)
        clip = build(out, fl)
        losses.append(clip.compute_global_alignment(init='mst', init_priors=[key_pose, key_depth, key_focal], niter=niter,
                                                    schedule=schedule, lr=lr))
        t = clip.get_tum_poses(key_pose)
        traj[0] = np.concatenate([traj[0], t[0]], axis=0)
        traj[1] = np.concatenate([traj[1], t[1] + offset], axis=0)
        depths += [_np(d) for d in clip.get_depthmaps()]
        confs += [_np(c) for c in clip.get_conf()]
        focals += _np(clip.get_focals()).reshape(-1).tolist() if not getattr(clip, 'shared_focal', False) \
            else [float(_np(clip.get_focals()).reshape(-1)[0])] * len(t[1])
        if save_folder is not None:
            save_trajectory_tum_format(traj, f'{save_folder}/pred_traj.txt')
            clip.save_intrinsics(f'{save_folder}/pred_intrinsics.txt')
            clip.save_depth_maps(save_folder, offset)
            clip.save_conf_maps(save_folder, offset)
            clip.save_init_conf_maps(save_folder, offset)
            if flow and getattr(clip, 'dynamic_masks', None) is not None:
                clip.save_dynamic_masks(save_folder, offset)                 # tool/demo.py:240
        offset += t[0].shape[0]
        del clip
    return dict(traj=traj, depths=depths, confs=confs, focals=focals, key_poses=key_poses, key_focals=key_focals,
                keyframes_id=keyframes_id, all_clips_id=all_clips_id, clip_size=clip_size, losses=losses, key_scene=scene)
