"""TEST INFRASTRUCTURE - generates tests/golden/small_hierarchical.pt by running the REAL reference (CPU).

    python oracle/make_golden_hierarchical.py          (build container only: needs /root/reference)

The hierarchical clip driver of tool/depth_test.py:620-676 on a synthetic 11-frame video: key frames 0, 4, 8 aligned on their
complete graph by the reference's flow-free aligner (dust3r.cloud_opt, lr 0.05, cosine, confidences above 1 set to 10), then
every clip aligned with init_priors = [key pose, key depth, key focal].  `my_make_pairs` is taken from the reference's own
source (extracted with ast and executed here - tool/depth_test.py cannot be imported: it loads the networks at import), the
driver loop below follows depth_test.py:631-676 line by line with the pairwise network replaced by the synthetic generator.
"""
import ast
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
from align3r_b200.synth import make_scene                 # noqa: E402

N, H, W, CLIP, NITER, SEED = 11, 32, 48, 4, 30, 4


def reference_function(path, name):
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    ns = {}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    return ns[name]


def synthetic_inference(pairs):
    """The pairwise network of the driver, replaced: pointmaps of the synthetic video for the pairs' frames."""
    edges = [(a["frame"], b["frame"]) for a, b in pairs]
    sc = make_scene(N, H, W, "complete", seed=SEED, with_flow=False, edges=edges)
    out = copy.deepcopy(sc.output)
    out["view1"]["idx"] = [a["idx"] for a, _ in pairs]
    out["view2"]["idx"] = [b["idx"] for _, b in pairs]
    return out


def main():
    rh.import_reference()
    from dust3r.cloud_opt import global_aligner, GlobalAlignerMode
    from dust3r.utils.device import to_numpy
    my_make_pairs = reference_function(os.path.join(rh.REFERENCE_ROOT, "tool", "depth_test.py"), "my_make_pairs")
    imgs = [dict(idx=k, frame=k) for k in range(N)]
    clip_size = CLIP
    while len(imgs) % clip_size == 1 or len(imgs) % clip_size == 0 or clip_size > len(imgs):
        clip_size -= 1
    coarse_init_pairs, keyframes_id, all_clips_pairs, all_clips_id = my_make_pairs(imgs, clip_size)
    pairs_idx = dict(key=[(a["idx"], b["idx"], a["frame"], b["frame"]) for a, b in coarse_init_pairs],
                     clips=[[(a["idx"], b["idx"], a["frame"], b["frame"]) for a, b in c] for c in all_clips_pairs])
    key_output = synthetic_inference(coarse_init_pairs)
    key_output['pred1']['conf'][key_output['pred1']['conf'] > 1] = 10
    key_output['pred2']['conf'][key_output['pred2']['conf'] > 1] = 10
    mode = GlobalAlignerMode.PointCloudOptimizer
    with rh.quiet():
        scene = global_aligner(key_output, False, [], device='cpu', mode=mode, verbose=False, min_conf_thr=3)
        loss = scene.compute_global_alignment(init='mst', niter=NITER, schedule='cosine', lr=0.05)
    losses = [float(loss)]
    init_keyposes = to_numpy(scene.get_im_poses()).tolist()
    init_keydepths = to_numpy(scene.get_depthmaps())
    init_keyfocals = to_numpy(scene.get_focals()).tolist()
    depths, confs, poses, focals = [], [], [], []
    for init_keypose, init_keydepth, init_keyfocal, clip_pair in zip(init_keyposes, init_keydepths, init_keyfocals, all_clips_pairs):
        clipwise_output = synthetic_inference(clip_pair)
        clipwise_output['pred1']['conf'][clipwise_output['pred1']['conf'] > 1] = 10
        clipwise_output['pred2']['conf'][clipwise_output['pred2']['conf'] > 1] = 10
        with rh.quiet():
            scene_clip_i = global_aligner(clipwise_output, False, [], device='cpu', mode=mode, verbose=False, min_conf_thr=3)
            init_priors = [init_keypose, init_keydepth, init_keyfocal]
            loss = scene_clip_i.compute_global_alignment(init='mst', init_priors=init_priors, niter=NITER, schedule='cosine', lr=0.05)
        losses.append(float(loss))
        depths += [torch.from_numpy(np.array(d)) for d in to_numpy(scene_clip_i.get_depthmaps())]
        confs += [torch.from_numpy(np.array(c)) for c in to_numpy(scene_clip_i.get_conf())]
        poses.append(scene_clip_i.get_im_poses().detach().clone())
        focals.append(scene_clip_i.get_focals().detach().clone())
    out = dict(N=N, H=H, W=W, clip=CLIP, niter=NITER, seed=SEED, clip_size=clip_size, keyframes_id=keyframes_id, all_clips_id=all_clips_id,
               pairs=pairs_idx, losses=losses, key_poses=torch.tensor(init_keyposes), key_focals=torch.tensor(init_keyfocals),
               depths=torch.stack(depths), confs=torch.stack(confs), clip_poses=poses, clip_focals=focals)
    path = os.path.join(ROOT, "tests", "golden", "small_hierarchical.pt")
    torch.save(out, path)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB", "losses", losses, "clip size", clip_size, keyframes_id, all_clips_id)


if __name__ == "__main__":
    main()
