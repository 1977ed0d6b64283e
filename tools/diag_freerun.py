"""Diagnostic (not a test): where does the free-running difference to the reference come from?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import golden_case, go, rel_err
from tests.test_gpu_parity import build, set_params
from align3r_b200.cloud_opt_flow import init_im_poses as init_fun
from align3r_b200.cloud_opt_flow.optimizer import global_alignment_loop

for name in ["small_flow", "small_perframe_focal", "small_gate_trips"]:
    fx, scene, prob = golden_case(name)
    ref = fx["free_run"]
    # (a) full pipeline
    net = build(scene, fx["kw"])
    init_fun.init_minimum_spanning_tree(net)
    ip = fx["init_params"]
    Pm, Pr = net.get_im_poses().detach().cpu(), go.poses_to_matrices(ip["im_poses"])
    print(name, "init pose diff per frame", (Pm - Pr).abs().flatten(1).max(1).values.tolist())
    print("   init depth diff per frame", (net.im_depthmaps.detach().cpu()[:, :prob.P] - ip["im_depthmaps"]).abs().max(1).values.tolist())
    la = global_alignment_loop(net, lr=0.01, niter=fx["niter"], schedule="linear")
    # (b) loop only, from the reference's init
    net2 = build(scene, fx["kw"])
    set_params(net2, ip)
    lb = global_alignment_loop(net2, lr=0.01, niter=fx["niter"], schedule="linear")
    print("   final loss: ref %.8f  ours(full) %.8f (%.1e)  ours(from ref init) %.8f (%.1e)" % (
        ref["loss"], la, abs(la - ref["loss"]) / ref["loss"], lb, abs(lb - ref["loss"]) / ref["loss"]))
    d = (net2.im_depthmaps.detach().cpu()[:, :prob.P] - ref["params"]["im_depthmaps"]).abs()
    print("   from ref init: depth median %.2e max %.2e; poses %.2e" % (float(d.median()), float(d.max()),
          float((net2.get_im_poses().detach().cpu() - ref["im_poses_4x4"]).abs().max())))
    hist = net2._bufs.loss_hist[:fx["niter"]].cpu()
    # oracle trajectory from the same init for comparison
    _, losses, _ = go.run_alignment(prob, ip, niter=fx["niter"], schedule="linear", lr=0.01)
    print("   loss traj rel diff vs oracle at it 0,5,10,20,last:", [abs(float(hist[i]) - losses[i]) / losses[i] for i in (0, 5, 10, min(20, fx["niter"] - 1), fx["niter"] - 1)])
