"""Worker of tests/test_gpu_multi.py (launched with torchrun, one rank per GPU, NCCL)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from align3r_b200.cloud_opt_flow import GlobalAlignerMode, global_aligner
from align3r_b200.plan import frame_costs, partition_frames
from tests.helpers import golden_case


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    name = sys.argv[1]
    fx, scene, prob = golden_case(name)
    kw = dict(fx["kw"])
    edges = scene.edges
    H, W = prob.H, prob.W
    wins = partition_frames(frame_costs(np.asarray(edges, dtype=np.int32), np.asarray([(H, W)] * prob.n_imgs, dtype=np.int32)), world)
    lo, hi = wins[rank]
    rows = [e for e, (i, j) in enumerate(edges) if lo <= i < hi or lo <= j < hi]
    out = scene.output
    sub = dict(view1={"idx": out["view1"]["idx"], "dynamic_mask": out["view1"]["dynamic_mask"][rows]},
               view2={"idx": out["view2"]["idx"], "dynamic_mask": out["view2"]["dynamic_mask"][rows]},
               pred1={k: v[rows] for k, v in out["pred1"].items()}, pred2={k: v[rows] for k, v in out["pred2"].items()})
    net = global_aligner(sub, dev, mode=GlobalAlignerMode.PointCloudOptimizer, flow=(scene.flow_ij[rows], scene.flow_ji[rows]),
                         shard=dict(world=world, rank=rank, group=None), edge_rows=rows, **kw)
    if len(sys.argv) > 3 and sys.argv[3] == "readout":
        return readout(net, scene, prob, rank, world, lo, hi)
    if len(sys.argv) > 3 and sys.argv[3] == "known_poses":
        return known_poses(net, fx, prob, rank, world, lo, hi)
    loss = net.compute_global_alignment(init="mst", niter=fx["niter"], schedule="linear", lr=0.01)
    # replicated parameters must be bit-identical on every rank
    small = torch.cat([net.pw_poses.detach().flatten(), net.im_poses.detach().flatten(), net.im_focals.detach().flatten()])
    ref = small.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(ref, small)
    depth_local = net.im_depthmaps.detach()[:, :prob.P].contiguous()
    gathered = [None] * world
    masks = torch.stack([net.dynamic_masks[n] for n in range(lo, hi)]).cpu() if kw.get("use_self_mask") else None
    dist.all_gather_object(gathered, (lo, hi, depth_local.cpu(), same, loss, bool(net.flow_loss_flag), masks))
    if rank == 0:
        torch.save(dict(parts=gathered, im_poses=net.get_im_poses().detach().cpu(), focals=net.get_focals().detach().cpu(),
                        exchange="peer" if net._peers is not None else "nccl"), sys.argv[2])
    dist.barrier()
    dist.destroy_process_group()


def known_poses(net, fx, prob, rank, world, lo, hi):
    """init='known_poses' on frame windows (every rank is handed all cameras here, as a caller holding the trajectory would)."""
    net.camera_poses = [p for p in fx["known_c2w"]]
    net.compute_global_alignment(init="known_poses", niter=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, net.im_depthmaps.detach()[:, :prob.P].cpu(), net.get_pw_poses().detach().cpu(),
                                      net.get_im_poses().detach().cpu()))
    if rank == 0:
        torch.save(dict(parts=gathered), sys.argv[2])
    dist.barrier()
    dist.destroy_process_group()


def readout(net, scene, prob, rank, world, lo, hi):
    """The post-step and the lazy pre-step read-outs on frame windows, at parameters both sides build from the same recipe:
    clean_pointcloud (every frame meets every other frame's depth map) and the flow validity maps (a pair's two flow planes
    may live on different GPUs)."""
    from align3r_b200.synth import load_params, near_solution_params
    load_params(net, *near_solution_params(scene.gt, scene.edges, seed=0))
    net.clean_pointcloud()
    conf = torch.stack([c.detach().cpu() for c in net.im_conf])
    vi, vj = net.flow_valid_mask_i, net.flow_valid_mask_j
    rows_i, rows_j = net.flow_valid_rows(0), net.flow_valid_rows(1)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, conf, rows_i, vi[rows_i].cpu(), rows_j, vj[rows_j].cpu()))
    net.close()
    from align3r_b200.cloud_opt_flow.optimizer import free_peer_regions
    free_peer_regions()
    if rank == 0:
        torch.save(dict(parts=gathered), sys.argv[2])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
