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
               motion: float = 1.0, pin_memory: bool = False, workers: int = 1) -> SyntheticScene:
    """Build the pairwise-prediction dict for a synthetic video.

    ``edges`` restricts generation to a subset of the graph's directed edges (used by the
    frame-window sharded multi-GPU path: each rank generates only the edges it touches).
    ``frame_noise`` > 0 makes the noise per-frame coherent (X_n (1+eps_n)), the
    "Depth-Pro-conditioned" variant of SURVEY.md section 8(d).
    ``workers`` > 1 draws the edges on a thread pool (every edge has its own generator: same tensors, bit for bit).
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

    import threading
    lock = threading.Lock()
    cam_cache: dict = {}

    def cam_pts(n):
        with lock:
            p = cam_cache.get(n)
        if p is None:
            p = gt.cam_points(n)
            p = p * feps[n] if feps is not None else p
            with lock:
                if len(cam_cache) > 32:
                    cam_cache.pop(next(iter(cam_cache)))
                cam_cache[n] = p
        return p

    flow_cache: dict = {}

    def flow_of(s, t):
        with lock:
            f = flow_cache.get((s, t))
        if f is None:
            f = ego_flow(gt, s, t).float()
            with lock:
                if len(flow_cache) > 96:
                    flow_cache.pop(next(iter(flow_cache)))
                flow_cache[(s, t)] = f
        return f

    def draw(e):
        i, j = edges[e]
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
                f = flow_of(s, t) + flow_noise * torch.randn(2, H, W, generator=g)
                f = f + masks[s][None] * (20 * torch.rand(2, H, W, generator=g) - 10)
                dst[e] = f

    if workers > 1 and E > 1:
        from concurrent.futures import ThreadPoolExecutor
        nthr = torch.get_num_threads()
        torch.set_num_threads(1)                 # the pool is the parallelism; per-edge ops are small
        try:
            with ThreadPoolExecutor(max_workers=int(workers)) as pool:
                list(pool.map(draw, range(E)))
        finally:
            torch.set_num_threads(nthr)
    else:
        for e in range(E):
            draw(e)

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


# ---------------------------------------------------------------------------------------
# long videos: the same recipe drawn on the GPU, row by row, when the aligner packs them
# ---------------------------------------------------------------------------------------
class DeviceRows:
    """(E, *row_shape) float32 rows that are drawn on a CUDA device when somebody indexes them.

    The 1000-frame workload (BASELINE configs[4]) has 70 GB of pairwise predictions: drawing them with the host
    generator takes minutes per rank and a pinned host copy of that size.  This sequence hands the aligner's packing
    loop device tensors instead: ``rows[k]``, ``rows[a:b]`` and ``rows[[k0, k1, ...]]`` return freshly drawn rows; every
    row has its own Philox stream seeded by (seed, i, j, kind), so any rank draws the same values for the same edge.
    Input fabrication, not part of the alignment path.
    """

    def __init__(self, n_rows, row_shape, draw, device):
        self.shape = (int(n_rows),) + tuple(row_shape)
        self._draw, self.device, self.dtype = draw, torch.device(device), torch.float32

    def __len__(self):
        return self.shape[0]

    def numel(self):
        return int(np.prod(self.shape))

    def element_size(self):
        return 4

    def __getitem__(self, k):
        if isinstance(k, slice):
            k = list(range(*k.indices(self.shape[0])))
        if isinstance(k, (list, tuple)):
            return torch.stack([self._draw(int(r)) for r in k]) if len(k) else torch.empty((0,) + self.shape[1:], device=self.device)
        return self._draw(int(k))


def make_scene_device(n_frames: int, H: int, W: int, scene_graph: str = "swin-5-noncyclic", seed: int = 0, device="cuda",
                      edges=None, pixel_noise: float = 0.01, flow_noise: float = 0.5, motion: float = 1.0) -> SyntheticScene:
    """``make_scene``'s recipe (same ground truth, same noise model, same per-pair scales) with every per-edge tensor a
    ``DeviceRows`` sequence drawn on ``device`` on demand.  The values differ from the host generator's (another random
    stream); ranks agree with each other.  The dynamic masks are host tensors as in ``make_scene``."""
    device = torch.device(device)
    all_edges = make_edge_list(n_frames, scene_graph)
    edges = list(all_edges) if edges is None else [tuple(e) for e in edges]
    E = len(edges)
    gt = make_ground_truth(n_frames, H, W, motion)
    masks = [frame_dynamic_mask(seed, n, H, W) for n in range(n_frames)]
    cams = torch.from_numpy(gt.cam2world).to(device)
    K = torch.from_numpy(gt.intrinsics()).to(device)
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float64, device=device), torch.arange(W, dtype=torch.float64, device=device), indexing="ij")
    cache: dict = {}

    def cached(key, fn, limit=64):
        if key not in cache:
            if len(cache) > limit:
                cache.pop(next(iter(cache)))
            cache[key] = fn()
        return cache[key]

    def depth(n):
        return 2 + 0.5 * torch.sin(u / 60 + 0.2 * n) * torch.cos(v / 45) + 0.05 * n

    def cam_pts(n):
        def fn():
            d = depth(n)
            return torch.stack((d * (u - gt.pp[0]) / gt.focal, d * (v - gt.pp[1]) / gt.focal, d), dim=-1)
        return cached(("X", n), fn)

    def flow_gt(s, t):
        def fn():
            relR = cams[t, :3, :3].T @ cams[s, :3, :3]
            relt = cams[t, :3, :3].T @ (cams[s, :3, 3] - cams[t, :3, 3])
            A = K @ relR @ torch.linalg.inv(K)
            b = K @ relt
            disp = 1.0 / (depth(s) + 1e-6)
            q = [A[r, 0] * u + A[r, 1] * v + A[r, 2] + disp * b[r] for r in range(3)]
            den = q[2] + 1e-6
            return torch.stack((q[0] / den - u, q[1] / den - v), dim=0).float()
        return cached(("F", s, t), fn)

    def mask_dev(n):
        return cached(("M", n), lambda: masks[n].to(device))

    def gen(i, j, kind):
        g = torch.Generator(device=device)
        g.manual_seed((((seed * 1000003 + i) * 1000003 + j) * 16 + kind) % (1 << 62))
        return g

    def scale(i, j):
        return 1.0 + 0.1 * float(torch.rand(1, generator=_edge_generator(seed, i, j, n_frames)))

    def pts1(e):
        i, j = edges[e]
        return (scale(i, j) * cam_pts(i)).float() + pixel_noise * torch.randn(H, W, 3, generator=gen(i, j, 0), device=device)

    def pts2(e):
        i, j = edges[e]
        rel = torch.linalg.inv(cams[i]) @ cams[j]
        x = cam_pts(j) @ rel[:3, :3].T + rel[:3, 3]
        return (scale(i, j) * x).float() + pixel_noise * torch.randn(H, W, 3, generator=gen(i, j, 1), device=device)

    def conf(kind):
        def fn(e):
            i, j = edges[e]
            return 1 + 3 * torch.rand(H, W, generator=gen(i, j, kind), device=device)
        return fn

    def flow(kind):
        def fn(e):
            i, j = edges[e]
            s, t = (i, j) if kind == 4 else (j, i)
            g = gen(i, j, kind)
            f = flow_gt(s, t) + flow_noise * torch.randn(2, H, W, generator=g, device=device)
            return f + mask_dev(s)[None] * (20 * torch.rand(2, H, W, generator=g, device=device) - 10)
        return fn

    idx1 = [i for i, _ in edges]
    idx2 = [j for _, j in edges]
    mstack = torch.stack(masks)
    view1 = {"idx": idx1, "dynamic_mask": mstack[idx1]}
    view2 = {"idx": idx2, "dynamic_mask": mstack[idx2]}
    pred1 = {"pts3d": DeviceRows(E, (H, W, 3), pts1, device), "conf": DeviceRows(E, (H, W), conf(2), device)}
    pred2 = {"pts3d_in_other_view": DeviceRows(E, (H, W, 3), pts2, device), "conf": DeviceRows(E, (H, W), conf(3), device)}
    meta = dict(n_frames=n_frames, H=H, W=W, scene_graph=scene_graph, seed=seed, n_edges_total=len(all_edges), drawn_on=str(device))
    return SyntheticScene(edges, view1, view2, pred1, pred2, DeviceRows(E, (2, H, W), flow(4), device),
                          DeviceRows(E, (2, H, W), flow(5), device), masks, gt, meta)


# ---------------------------------------------------------------------------------------
# parameters by recipe: a point near the solution, reproducible anywhere without running an initialisation
# ---------------------------------------------------------------------------------------
def near_solution_params(gt: GroundTruth, edges, seed: int = 0, base_scale: float = 0.5, focal_break: float = 20.0,
                         shared_focal: bool = True, wobble: float = 0.02):
    """Parameter values of the aligner (the reference's parameterisation, base_opt.py:202-247, optimizer.py:392-436)
    derived in closed form from the ground truth plus a small deterministic wobble, so that a teacher-forced evaluation
    can be repeated on another machine from the recipe alone (a 1000-frame depth stack does not fit a fixture).
    The scene is rescaled by c = base_scale * exp(mean log s_ij) so that the pair scales need no normalisation shift.
    Returns (small, logdepth): small = dict(pw_poses (E,8), im_poses (N,7), im_focals (1|N,1)) fp32 CPU tensors,
    logdepth(n, device='cpu') -> (H*W,) fp32 log-depth of frame n."""
    from scipy.spatial.transform import Rotation
    N, E = gt.n_frames, len(edges)
    s_ij = np.array([1.0 + 0.1 * float(torch.rand(1, generator=_edge_generator(seed, i, j, N))) for i, j in edges])
    c = base_scale * float(np.exp(np.mean(np.log(s_ij))))
    rng = np.random.RandomState(seed * 7 + 11)

    def slog1p(x):
        return np.sign(x) * np.log1p(np.abs(x))

    im = np.zeros((N, 7))
    for n in range(N):
        rv = Rotation.from_matrix(gt.cam2world[n, :3, :3]) * Rotation.from_rotvec(wobble * 0.1 * rng.randn(3))
        im[n, :4] = rv.as_quat() * (1.0 + 0.3 * rng.rand())              # raw quaternions are not unit in the reference either
        im[n, 4:] = slog1p(c * gt.cam2world[n, :3, 3] + wobble * 0.05 * rng.randn(3))
    pw = np.zeros((E, 8))
    for e, (i, j) in enumerate(edges):
        s = c / s_ij[e] * np.exp(wobble * rng.randn())
        rv = Rotation.from_matrix(gt.cam2world[i, :3, :3]) * Rotation.from_rotvec(wobble * 0.1 * rng.randn(3))
        pw[e, :4] = rv.as_quat() * (1.0 + 0.3 * rng.rand())
        pw[e, 4:7] = slog1p((c * gt.cam2world[i, :3, 3] + wobble * 0.05 * rng.randn(3)) / s)      # y = s (R x + T / s)
        pw[e, 7] = np.log(s)
    nf = 1 if shared_focal else N
    foc = focal_break * np.log(gt.focal * (1.0 + wobble * 0.1 * rng.randn(nf, 1)))
    small = dict(pw_poses=torch.from_numpy(pw).float(), im_poses=torch.from_numpy(im).float(), im_focals=torch.from_numpy(foc).float())

    def logdepth(n, device="cpu"):
        v, u = torch.meshgrid(torch.arange(gt.H, dtype=torch.float64, device=device),
                              torch.arange(gt.W, dtype=torch.float64, device=device), indexing="ij")
        d = 2 + 0.5 * torch.sin(u / 60 + 0.2 * n) * torch.cos(v / 45) + 0.05 * n
        wob = wobble * torch.sin(0.05 * u + 0.07 * v + 1.3 * n)
        return (torch.log(c * d) + wob).float().reshape(-1)

    return small, logdepth


def load_params(net, small, logdepth):
    """Writes near_solution_params' values into an aligner (ours; the frames this device owns)."""
    with torch.no_grad():
        net.pw_poses.data.copy_(small["pw_poses"].to(net.device))
        net.im_poses.data.copy_(small["im_poses"].to(net.device))
        net.im_focals.data.copy_(small["im_focals"].reshape(net.im_focals.shape).to(net.device))
        for k, n in enumerate(net._plan.lframe_id.tolist()):
            ld = logdepth(n, net.device)
            net.im_depthmaps.data[k, :ld.numel()] = ld


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
