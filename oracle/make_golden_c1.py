"""TEST INFRASTRUCTURE - generates tests/golden/c1_reference_run.pt by running the REAL reference (CPU).

    python oracle/make_golden_c1.py          (build container only: needs /root/reference; ~5 minutes)

BASELINE.json configs[0] at FULL size: PointCloudOptimizer global alignment on the synthetic 10-frame 288x512 sequence,
swin-3 pair graph, init='mst', niter=300 (canonical hyper-parameters: shared focal, flow 0.01 + temporal 0.01, linear
schedule, lr 0.01), run by the unmodified reference.  The fixture keeps what a user reads out after the run - final loss,
camera poses, focals, the pair poses, a stride-8 sample of the log-depth maps and the Sintel-protocol depth metrics /
ATE / RPE against the synthetic ground truth - small enough to commit; the inputs are regenerated from the recipe.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
import ga_oracle as go                                     # noqa: E402
from align3r_b200.synth import make_scene, CANONICAL_KW, CONFIGS   # noqa: E402
from make_golden import checksum                           # noqa: E402

NITER = 300
STRIDE = 8


def main():
    rh.import_reference()
    n, H, W, graph = CONFIGS["c1"]
    recipe = dict(n_frames=n, H=H, W=W, scene_graph=graph, seed=0, with_flow=True)
    scene = make_scene(**recipe)
    kw = dict(CANONICAL_KW, num_total_iter=NITER)
    torch.manual_seed(0)
    ref = rh.build_reference_scene(scene, **kw)
    t0 = time.time()
    with rh.quiet():
        final = ref.compute_global_alignment(init="mst", niter=NITER, schedule="linear", lr=0.01)
    dt = time.time() - t0
    depth = torch.stack([d.detach() for d in ref.get_depthmaps()])                 # (n, H, W)
    poses = ref.get_im_poses().detach().clone()
    gt_depth = np.stack([scene.gt.depth(k).numpy() for k in range(n)])
    out = dict(recipe=recipe, kw=kw, niter=NITER, checksum=checksum(scene), edges=scene.edges, seed=0,
               loss=float(final), flow_dropped=bool(ref.flow_loss_flag),
               im_poses_4x4=poses, focals=ref.get_focals().detach().clone(),
               pw_poses_4x4=ref.get_pw_poses().detach().clone(),
               logdepth_sample=depth[:, ::STRIDE, ::STRIDE].log().clone(), stride=STRIDE,
               logdepth_mean=float(depth.log().double().mean()),
               depth_metrics=go.depth_metrics(depth.numpy(), gt_depth),
               pose_metrics=go.pose_metrics(poses.numpy(), scene.gt.cam2world),
               reference_seconds=dt, reference_threads=torch.get_num_threads())
    path = os.path.join(ROOT, "tests", "golden", "c1_reference_run.pt")
    torch.save(out, path)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB  final loss {final}  dropped {out['flow_dropped']}  "
          f"{NITER / dt:.2f} it/s on {torch.get_num_threads()} threads")
    print(out["depth_metrics"], out["pose_metrics"])


if __name__ == "__main__":
    main()
