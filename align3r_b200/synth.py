"""Seeded synthetic pairwise-prediction scenes (the inputs of the global alignment).

The global aligner consumes the dict produced by the reference's
``dust3r.inference.inference`` (reference: dust3r/inference.py:55-72): ``view1``/``view2``
(frame indices, optional images and dynamic masks) and ``pred1``/``pred2`` (pointmaps
and confidences of every directed pair).  There is no network here, so this module
fabricates that dict from a known ground-truth scene:

  * depth      d_n(u,v) = 2 + 0.5 sin(u/60 + 0.2 n) cos(v/45) + 0.05 n
  * intrinsics focal 0.9 W, principal point (W/2, H/2)
  * cameras    yaw 0.02 n rad, translation (0.05 n, 0, 0.02 n)   (cam-to-world)
  * pair (i,j) pred1 = s_ij X_i + N(0, sigma),  pred2 = s_ij (T_i^-1 T_j X_j) + N(0, sigma)
               with a per-pair scale s_ij ~ U(1, 1.1) (DUSt3R's scale ambiguity)
  * conf       1 + 3 U(0,1)
  * flow       exact ego-flow of the ground-truth scene + N(0, 0.5 px); inside two
               random "dynamic" rectangles per frame an extra U(-10, 10) px is added.

Every directed edge draws from its own generator seeded by (seed, i, j), so any rank can
build any subset of edges independently and get bit-identical tensors.
Everything here is host-side numpy/torch-CPU; it is input fabrication, not part of the
alignment path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch


# ---------------------------------------------------------------------------------------
# pair graphs (edge semantics of reference dust3r/image_pairs.py:11-75, restated)
# ---------------------------------------------------------------------------------------
def make_edge_list(n_frames: int, scene_graph: str = "swin-5-noncyclic", symmetrize: bool = True):
    """Directed edge list for the reference's scene-graph strings.

    Supports ``complete``, ``swin-w``, ``swinstride-w``, ``swin2stride-w``, ``logwin-w``,
    ``oneref-k`` with the ``-noncyclic`` suffix (reference image_pairs.py:13-65).  The
    undirected pairs are emitted in sorted order (the reference iterates a Python set, whose
    order is an implementation detail); symmetrisation appends the reversed pairs, exactly as
    image_pairs.py:66-67 does, so edge ``e`` and ``e + E/2`` are mirror images.
    """
    n = n_frames
    pairs = []
    if scene_graph == "complete":
        pairs = [(i, j) for i in range(n) for j in range(i)]
    elif scene_graph.startswith("swin"):
        cyclic = not scene_graph.endswith("noncyclic")
        try:
            win = int(scene_graph.split("-")[1])
        except Exception:
            win = 3
        stride = 2 if scene_graph.startswith("swinstride") else 3 if scene_graph.startswith("swin2stride") else 1
        ids = set()
        for i in range(n):
            for off in range(1, stride * win + 1, stride):
                j = i + off
                if cyclic:
                    j %= n
                if j >= n or j == i:
                    continue
                ids.add((i, j) if i < j else (j, i))
        pairs = sorted(ids)
    elif scene_graph.startswith("logwin"):
        cyclic = not scene_graph.endswith("noncyclic")
        try:
            win = int(scene_graph.split("-")[1])
        except Exception:
            win = 3
        ids = set()
        for i in range(n):
            for off in [2 ** k for k in range(win)]:
                for j in (i - off, i + off):
                    if cyclic:
                        j %= n
                    if j < 0 or j >= n or j == i:
                        continue
                    ids.add((i, j) if i < j else (j, i))
        pairs = sorted(ids)
    elif scene_graph.startswith("oneref"):
        ref = int(scene_graph.split("-")[1]) if "-" in scene_graph else 0
        pairs = [(ref, j) for j in range(n) if j != ref]
    else:
        raise ValueError(f"bad scene_graph {scene_graph!r}")
    if symmetrize:
        pairs = pairs + [(j, i) for i, j in pairs]
    return pairs


# ---------------------------------------------------------------------------------------
# ground truth
# ---------------------------------------------------------------------------------------
@dataclass
class GroundTruth:
    n_frames: int
    H: int
    W: int
    focal: float
    pp: tuple
    cam2world: np.ndarray          # (N,4,4) float64
    motion: float = 1.0

    def depth(self, n: int) -> torch.Tensor:
        v, u = torch.meshgrid(torch.arange(self.H, dtype=torch.float64),
                              torch.arange(self.W, dtype=torch.float64), indexing="ij")
        return 2 + 0.5 * torch.sin(u / 60 + 0.2 * n) * torch.cos(v / 45) + 0.05 * n

    def cam_points(self, n: int) -> torch.Tensor:
        """(H,W,3) float64 points of frame n in its own camera frame."""
        d = self.depth(n)
        v, u = torch.meshgrid(torch.arange(self.H, dtype=torch.float64),
                              torch.arange(self.W, dtype=torch.float64), indexing="ij")
        x = d * (u - self.pp[0]) / self.focal
        y = d * (v - self.pp[1]) / self.focal
        return torch.stack((x, y, d), dim=-1)

    def intrinsics(self) -> np.ndarray:
        return np.array([[self.focal, 0, self.pp[0]], [0, self.focal, self.pp[1]], [0, 0, 1.0]])


def make_ground_truth(n_frames: int, H: int, W: int, motion: float = 1.0) -> GroundTruth:
    poses = np.zeros((n_frames, 4, 4))
    for n in range(n_frames):
        a = 0.02 * n * motion
        c, s = math.cos(a), math.sin(a)
        poses[n] = np.array([[c, 0, s, 0.05 * n * motion],
                             [0, 1, 0, 0.0],
                             [-s, 0, c, 0.02 * n * motion],
                             [0, 0, 0, 1.0]])
    return GroundTruth(n_frames, H, W, 0.9 * W, (W / 2, H / 2), poses, motion)


def ego_flow(gt: GroundTruth, src: int, tgt: int) -> torch.Tensor:
    """(2,H,W) float64 optical flow induced by camera motion src->tgt on the GT scene.

    Same projective arithmetic as the flow term (reference goem_opt.py:195-236), in fp64.
    """
    K = torch.from_numpy(gt.intrinsics())
    Kinv = torch.linalg.inv(K)
    Ts, Tt = torch.from_numpy(gt.cam2world[src]), torch.from_numpy(gt.cam2world[tgt])
    relR = Tt[:3, :3].T @ Ts[:3, :3]
    relt = Tt[:3, :3].T @ (Ts[:3, 3] - Tt[:3, 3])
    A = K @ relR @ Kinv
    b = K @ relt
    v, u = torch.meshgrid(torch.arange(gt.H, dtype=torch.float64),
                          torch.arange(gt.W, dtype=torch.float64), indexing="ij")
    disp = 1.0 / (gt.depth(src) + 1e-6)
    q = [A[r, 0] * u + A[r, 1] * v + A[r, 2] + disp * b[r] for r in range(3)]
    den = q[2] + 1e-6
    return torch.stack((q[0] / den - u, q[1] / den - v), dim=0)


def frame_dynamic_mask(seed: int, n: int, H: int, W: int) -> torch.Tensor:
    """(H,W) bool, two random rectangles covering ~10 % of the frame."""
    g = torch.Generator().manual_seed((seed * 7919 + 17) * 1000003 + n)
    m = torch.zeros(H, W, dtype=torch.bool)
    rh, rw = max(1, int(0.24 * H)), max(1, int(0.205 * W))
    for _ in range(2):
        y0 = int(torch.randint(0, max(1, H - rh), (1,), generator=g))
        x0 = int(torch.randint(0, max(1, W - rw), (1,), generator=g))
        m[y0:y0 + rh, x0:x0 + rw] = True
    return m


# ---------------------------------------------------------------------------------------
# the dust3r_output-shaped scene
# ---------------------------------------------------------------------------------------
@dataclass
class SyntheticScene:
    edges: list
    view1: dict
    view2: dict
    pred1: dict
    pred2: dict
    flow_ij: torch.Tensor | None
    flow_ji: torch.Tensor | None
    dynamic_masks: list
    gt: GroundTruth
    meta: dict = field(default_factory=dict)

    @property
    def output(self) -> dict:
        return dict(view1=self.view1, view2=self.view2, pred1=self.pred1, pred2=self.pred2)


def _edge_generator(seed: int, i: int, j: int, n: int) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + i) * 1000003 + j + 7 * n)


def make_scene(n_frames: int, H: int, W: int, scene_graph: str = "swin-5-noncyclic", seed: int = 0,
               with_flow: bool = True, with_img: bool = False, edges=None,
               pixel_noise: float = 0.01, frame_noise: float = 0.0, flow_noise: float = 0.5,
               motion: float = 1.0, pin_memory: bool = False) -> SyntheticScene:
    """Build the pairwise-prediction dict for a synthetic video.

    ``edges`` restricts generation to a subset of the graph's directed edges (used by the
    frame-window sharded multi-GPU path: each rank generates only the edges it touches).
    ``frame_noise`` > 0 makes the noise per-frame coherent (X_n (1+eps_n)), the
    "Depth-Pro-conditioned" variant of SURVEY.md section 8(d).
    """
    all_edges = make_edge_list(n_frames, scene_graph)
    edges = list(all_edges) if edges is None else [tuple(e) for e in edges]
    E = len(edges)
    gt = make_ground_truth(n_frames, H, W, motion)

    def alloc(*shape, dtype=torch.float32):
        t = torch.empty(*shape, dtype=dtype)
        return t.pin_memory() if pin_memory and torch.cuda.is_available() else t

    pts1, pts2 = alloc(E, H, W, 3), alloc(E, H, W, 3)
    conf1, conf2 = alloc(E, H, W), alloc(E, H, W)
    flow_ij = alloc(E, 2, H, W) if with_flow else None
    flow_ji = alloc(E, 2, H, W) if with_flow else None
    masks = [frame_dynamic_mask(seed, n, H, W) for n in range(n_frames)]
    feps = None
    if frame_noise > 0:
        g = torch.Generator().manual_seed(seed * 31 + 5)
        feps = 1 + frame_noise * torch.randn(n_frames, generator=g, dtype=torch.float64)

    cam_cache: dict = {}

    def cam_pts(n):
        if n not in cam_cache:
            if len(cam_cache) > 16:
                cam_cache.pop(next(iter(cam_cache)))
            p = gt.cam_points(n)
            cam_cache[n] = p * feps[n] if feps is not None else p
        return cam_cache[n]

    flow_cache: dict = {}
    for e, (i, j) in enumerate(edges):
        g = _edge_generator(seed, i, j, n_frames)
        s_ij = 1.0 + 0.1 * float(torch.rand(1, generator=g))
        Ti, Tj = torch.from_numpy(gt.cam2world[i]), torch.from_numpy(gt.cam2world[j])
        rel = torch.linalg.inv(Ti) @ Tj                          # cam j -> cam i
        Xi = cam_pts(i)
        Xj_in_i = cam_pts(j) @ rel[:3, :3].T + rel[:3, 3]
        pts1[e] = (s_ij * Xi).float() + pixel_noise * torch.randn(H, W, 3, generator=g)
        pts2[e] = (s_ij * Xj_in_i).float() + pixel_noise * torch.randn(H, W, 3, generator=g)
        conf1[e] = 1 + 3 * torch.rand(H, W, generator=g)
        conf2[e] = 1 + 3 * torch.rand(H, W, generator=g)
        if with_flow:
            for (s, t, dst) in ((i, j, flow_ij), (j, i, flow_ji)):
                if (s, t) not in flow_cache:
                    if len(flow_cache) > 64:
                        flow_cache.pop(next(iter(flow_cache)))
                    flow_cache[(s, t)] = ego_flow(gt, s, t).float()
                f = flow_cache[(s, t)] + flow_noise * torch.randn(2, H, W, generator=g)
                f = f + masks[s][None] * (20 * torch.rand(2, H, W, generator=g) - 10)
                dst[e] = f

    idx1 = [i for i, _ in edges]
    idx2 = [j for _, j in edges]
    view1 = {"idx": idx1}
    view2 = {"idx": idx2}
    if with_flow:
        mstack = torch.stack(masks)
        view1["dynamic_mask"] = mstack[idx1]
        view2["dynamic_mask"] = mstack[idx2]
    if with_img:
        gi = torch.Generator().manual_seed(seed * 13 + 3)
        frames = 2 * torch.rand(n_frames, 3, H, W, generator=gi) - 1
        view1["img"] = frames[idx1]
        view2["img"] = frames[idx2]
    pred1 = {"pts3d": pts1, "conf": conf1}
    pred2 = {"pts3d_in_other_view": pts2, "conf": conf2}
    meta = dict(n_frames=n_frames, H=H, W=W, scene_graph=scene_graph, seed=seed, n_edges_total=len(all_edges))
    return SyntheticScene(edges, view1, view2, pred1, pred2, flow_ij, flow_ji, masks, gt, meta)


# canonical hyper-parameters used by every caller of the reference
CANONICAL_KW = dict(shared_focal=True, flow_loss_weight=0.01, flow_loss_fn="smooth_l1",
                    depth_regularize_weight=0.0, num_total_iter=300, temporal_smoothing_weight=0.01,
                    translation_weight=1.0, flow_loss_start_epoch=0.1, flow_loss_thre=25,
                    use_self_mask=False, sam2_mask_refine=False, pxl_thre=50, verbose=False)

# BASELINE.json configs: name -> (n_frames, H, W, scene_graph)
CONFIGS = {
    "c1": (10, 288, 512, "swin-3-noncyclic"),
    "c2": (50, 288, 512, "swin-5-noncyclic"),
    "c3": (90, 288, 512, "swin-5-noncyclic"),
    "c4": (110, 384, 512, "swin-5-noncyclic"),
    "c5": (1000, 288, 512, "swin-5-noncyclic"),
}
