"""Result write-out of the aligner: the on-disk formats the reference's tools and viewers consume
(reference dust3r/cloud_opt_flow/base_opt.py:332-427, dust3r/utils/vo_eval.py:308-316).

    pred_traj.txt        one line per frame "t x y z qw qx qy qz"  (TUM trajectory)
    pred_focal.txt       one focal per row, %.6f
    pred_intrinsics.txt  nine numbers per row (K row-major), %.6f, APPENDED (clips of one video share the file)
    conf_<k>.npy / init_conf_<k>.npy        transformed confidence maps
    frame_<k:04d>.npy (+ .png JET colour map, + _depth_maps.gif)   depth maps
    frame_<k:04d>_rgb.png (+ _rgb.gif), dynamic_mask_<k>.png
    depthmaps_<i>_iter_<n>.png           progress pictures of the loop (depth_map_save_dir, verbose runs)

Host code only (numpy / cv2 / PIL); bound to PointCloudOptimizer as methods of the same names.
"""
from __future__ import annotations

import os

import numpy as np


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def save_trajectory_tum_format(traj, path, verbose=True):
    """reference utils/vo_eval.py:308-316 for traj = [poses (M,7) x y z qw qx qy qz, timestamps (M,)]."""
    poses, tt = traj
    with open(path, "w") as f:
        for t, p in zip(tt, np.asarray(poses, dtype=np.float64)):
            f.write(f"{t} " + " ".join(map(str, p[:3])) + " " + " ".join(map(str, p[3:7])) + "\n")
    if verbose:
        print(f"Saved trajectory to {path}")


def save_tum_poses(self, path):
    traj = self.get_tum_poses()
    save_trajectory_tum_format(traj, path, self.verbose)
    return traj[0]


def save_focals(self, path):
    focals = self.get_focals()
    np.savetxt(path, _np(focals), fmt='%.6f')
    return focals


def save_intrinsics(self, path):
    K = self.get_intrinsics()
    with open(path, 'a') as f:
        np.savetxt(f, _np(K.reshape(-1, 9)), fmt='%.6f')
    return K


def save_conf_maps(self, path, start):
    conf = self.get_conf()
    for i, c in enumerate(conf):
        np.save(f'{path}/conf_{start + i}.npy', _np(c))
    return conf


def save_init_conf_maps(self, path, start):
    conf = self.get_init_conf()
    for i, c in enumerate(conf):
        np.save(f'{path}/init_conf_{start + i}.npy', _np(c))
    return conf


def _append_gif(gif_path, files):
    from PIL import Image, ImageSequence
    frames = [Image.open(p) for p in files]
    if os.path.exists(gif_path):
        old = Image.open(gif_path)
        frames = [fr.copy() for fr in ImageSequence.Iterator(old)] + frames
    frames[0].save(gif_path, save_all=True, append_images=frames[1:], duration=100, loop=0)


def save_rgb_imgs(self, path, start):
    import cv2
    files = []
    for i, img in enumerate(self.imgs):
        name = f'{path}/frame_{start + i:04d}_rgb.png'
        cv2.imwrite(name, np.ascontiguousarray(np.asarray(img)[..., ::-1]) * 255)      # RGB -> BGR
        files.append(name)
    _append_gif(os.path.join(path, '_rgb.gif'), files)
    return self.imgs


def save_dynamic_masks(self, path, start):
    import cv2
    masks = self.dynamic_masks if getattr(self, 'sam2_dynamic_masks', None) is None else self.sam2_dynamic_masks
    for i, m in enumerate(masks):
        cv2.imwrite(f'{path}/dynamic_mask_{start + i}.png', (_np(m) * 255).astype(np.uint8))
    return masks


def save_depth_maps(self, path, start, depth_maps_init=None):
    import cv2
    refine = depth_maps_init is not None
    depth_maps = depth_maps_init if refine else [_np(d) for d in self.get_depthmaps()]
    stem = 'frame_refine_' if refine else 'frame_'
    files = []
    for i, d in enumerate(depth_maps):
        name = f'{path}/{stem}{start + i:04d}'
        cv2.imwrite(name + '.png', cv2.applyColorMap((np.asarray(d) * 255).astype(np.uint8), cv2.COLORMAP_JET))
        np.save(name + '.npy', d)
        files.append(name + '.png')
    _append_gif(os.path.join(path, '_depth_maps_refine.gif' if refine else '_depth_maps.gif'), files)
    return depth_maps


def save_depthmap_pictures(depth_maps, save_dir, iteration):
    """reference base_opt.py:531-540: ``depthmaps_<i>_iter_<n>.png`` of every frame, min-max normalised and jet-coloured
    (matplotlib's ``imsave(cmap='jet')`` where matplotlib exists, OpenCV's JET colour map otherwise)."""
    os.makedirs(save_dir, exist_ok=True)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        plt = None
    files = []
    for i, d in enumerate(depth_maps):
        d = _np(d).astype(np.float64)
        path = os.path.join(save_dir, f'depthmaps_{i}_iter_{iteration}.png')
        if plt is not None:
            plt.imsave(path, d, cmap='jet')
        else:
            import cv2
            lo, hi = float(d.min()), float(d.max())
            norm = (d - lo) / (hi - lo) if hi > lo else np.zeros_like(d)
            cv2.imwrite(path, cv2.applyColorMap(np.round(norm * 255).astype(np.uint8), cv2.COLORMAP_JET))
        files.append(path)
    return files


METHODS = (save_tum_poses, save_focals, save_intrinsics, save_conf_maps, save_init_conf_maps, save_rgb_imgs,
           save_dynamic_masks, save_depth_maps)
