"""Host-side planning of the alignment scan: slots, work frames, tiles, CTA ranges, GPU shards.

Pure numpy (no CUDA): everything here is unit-tested on CPU.

Vocabulary (include/a3r_ga.h): every directed edge e = (i, j) of the pair graph has two *slots*
(edge-sides): global slot 2e lives on frame i's pixel grid (pred1 pointmap, reference
optimizer.py:87,511), global slot 2e+1 on frame j's grid (pred2 pointmap, optimizer.py:88,512).
The scan is frame-major: a CTA walks *tiles* (runs of ``tile_pixels`` pixels) of one frame and, for
each pixel, loops over the slots living on that frame.  A frame with more than ``max_slots`` slots is
split into several *work frames*.  For multi-GPU runs frames are cut into contiguous windows; a GPU
owns the per-pixel planes of its frames and all slots living on them, so no per-pixel data crosses
GPUs (partition A).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ShardPlan:
    n_frames: int
    n_edges: int
    edges: np.ndarray              # (E,2) int32
    frame_hw: np.ndarray           # (N,2) int32  (H, W)
    pix_stride: int                # padded pixels per plane
    tile_pixels: int
    max_slots: int
    n_ctas: int
    # local shard
    lframe_id: np.ndarray          # (nl,) int32 global frame ids, ascending
    lframe_slot_begin: np.ndarray  # (nl+1,) int32
    lslot_gslot: np.ndarray        # (nls,) int32 global slot ids, frame-major
    wf_lframe: np.ndarray          # (nwf,) int32
    wf_slot_begin: np.ndarray      # (nwf+1,) int32
    wf_tile_begin: np.ndarray      # (nwf+1,) int32
    wf_accum: np.ndarray           # (nwf,) int32
    cta_tile_begin: np.ndarray     # (n_ctas+1,) int32
    total_area_i: float            # sum over edges of area(frame i)  (reference optimizer.py:91)
    total_area_j: float

    @property
    def n_local_frames(self):
        return len(self.lframe_id)

    @property
    def n_local_slots(self):
        return len(self.lslot_gslot)

    @property
    def n_work_frames(self):
        return len(self.wf_lframe)

    @property
    def n_tiles(self):
        return int(self.wf_tile_begin[-1])

    def local_slot_of(self):
        """dict global slot -> local slot index."""
        return {int(g): k for k, g in enumerate(self.lslot_gslot)}

    def local_frame_of(self):
        return {int(n): k for k, n in enumerate(self.lframe_id)}

    def pixel_slots(self):
        """Number of (pixel, slot) pairs this shard scans per iteration."""
        areas = self.frame_hw[:, 0].astype(np.int64) * self.frame_hw[:, 1]
        deg = np.diff(self.lframe_slot_begin).astype(np.int64)
        return int((deg * areas[self.lframe_id]).sum())


def slots_per_frame(edges: np.ndarray, n_frames: int) -> np.ndarray:
    """Number of slots living on every frame: out-degree + in-degree of the directed graph."""
    deg = np.zeros(n_frames, dtype=np.int64)
    np.add.at(deg, edges[:, 0], 1)
    np.add.at(deg, edges[:, 1], 1)
    return deg


def check_edges(edges) -> int:
    """reference base_opt.py:182-185: frame indices must be 0..N-1 without holes."""
    idx = sorted({int(v) for e in edges for v in e})
    assert idx == list(range(len(idx))), 'bad pair indices: missing values '
    return len(idx)


def partition_frames(frame_cost: np.ndarray, world: int):
    """Cut frames into ``world`` contiguous windows of (nearly) equal cost.  Returns [(lo, hi)]."""
    n = len(frame_cost)
    if world <= 1:
        return [(0, n)]
    pref = np.concatenate(([0.0], np.cumsum(frame_cost, dtype=np.float64)))
    cuts = [0]
    for r in range(1, world):
        target = pref[-1] * r / world
        k = int(np.searchsorted(pref, target))
        # nearest boundary, keep windows non-empty and monotone when there are enough frames
        if k > 0 and abs(pref[k - 1] - target) <= abs(pref[min(k, n)] - target):
            k -= 1
        k = max(k, cuts[-1] + (1 if n >= world else 0))
        k = min(k, n - (world - r) if n >= world else n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def frame_costs(edges: np.ndarray, frame_hw: np.ndarray) -> np.ndarray:
    """Scan cost of a frame ~ (slots + per-pixel overhead) * pixels."""
    areas = frame_hw[:, 0].astype(np.float64) * frame_hw[:, 1]
    return (slots_per_frame(edges, len(frame_hw)) + 2.0) * areas


def build_plan(edges, frame_hw, tile_pixels: int, max_slots: int, n_ctas: int, frame_range=None, align_warps: int = 0) -> ShardPlan:
    """``align_warps`` > 0 (class-specialised scan, include/a3r_ga.h a3r_scan_align_warps): cta_tile_begin holds three
    partitions of the tile list behind each other - n_ctas ranges (every warp an align warp: flow term off), then
    align_warps/16 * n_ctas ranges (the align warps of a flow-on launch), then the rest (its flow warps)."""
    edges = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
    frame_hw = np.asarray(frame_hw, dtype=np.int32).reshape(-1, 2)
    N, E = len(frame_hw), len(edges)
    assert check_edges(edges.tolist()) == N, "frame_hw does not match the edge list"
    areas = frame_hw[:, 0].astype(np.int64) * frame_hw[:, 1]
    assert int(areas.max()) < (1 << 24), "frames of 2^24 pixels or more are not supported (the scan derives (u, v) from a float quotient)"
    # whole tiles per plane: the TMA-fed scan copies full tiles, padding is zero (weight 0, not static)
    unit = int(np.lcm(4, int(tile_pixels)))
    pix_stride = int((areas.max() + unit - 1) // unit * unit)
    lo, hi = (0, N) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
    assert 0 <= lo <= hi <= N

    # slots living on each frame, ordered by (edge, side)
    per_frame = [[] for _ in range(N)]
    for e, (i, j) in enumerate(edges.tolist()):
        per_frame[i].append(2 * e)
        per_frame[j].append(2 * e + 1)

    lframe_id = np.arange(lo, hi, dtype=np.int32)
    slot_begin, lslots = [0], []
    wf_lframe, wf_slot_begin, wf_tile_begin, wf_accum = [], [0], [0], []
    for k, n in enumerate(lframe_id.tolist()):
        sl = sorted(per_frame[n], key=lambda g: (g & 1, g))     # side-i slots first (the v4 scan swaps its flow-gradient class once per tile)
        base = len(lslots)
        lslots.extend(sl)
        slot_begin.append(len(lslots))
        ntile = int((areas[n] + tile_pixels - 1) // tile_pixels)
        groups = max(1, (len(sl) + max_slots - 1) // max_slots)
        for g in range(groups):
            a, b = g * max_slots, min(len(sl), (g + 1) * max_slots)
            wf_lframe.append(k)
            wf_slot_begin.append(base + b)
            wf_tile_begin.append(wf_tile_begin[-1] + ntile)
            wf_accum.append(1 if groups > 1 else 0)
            assert wf_slot_begin[-2] == base + a
    wf_tile_begin = np.asarray(wf_tile_begin, dtype=np.int64)
    n_tiles = int(wf_tile_begin[-1])

    # balanced contiguous tile ranges: cost of a tile ~ slots of its work frame + per-pixel overhead
    n_ctas = max(1, int(n_ctas))

    def partition(parts):
        if n_tiles > 0:
            deg = np.diff(np.asarray(wf_slot_begin)).astype(np.float64) + 2.0
            ntile_wf = np.diff(wf_tile_begin)
            cost = np.repeat(deg, ntile_wf)
            pref = np.concatenate(([0.0], np.cumsum(cost)))
            targets = pref[-1] * np.arange(parts + 1) / parts
            c = np.searchsorted(pref, targets, side="left")
            c[0], c[-1] = 0, n_tiles
            return np.maximum.accumulate(c)
        return np.zeros(parts + 1, dtype=np.int64)

    if align_warps > 0:
        assert n_ctas % 16 == 0 and 0 < align_warps < 16
        S = n_ctas // 16
        cta = np.concatenate([partition(n_ctas), partition(align_warps * S), partition((16 - align_warps) * S)])
    else:
        cta = partition(n_ctas)

    return ShardPlan(
        n_frames=N, n_edges=E, edges=edges, frame_hw=frame_hw, pix_stride=pix_stride,
        tile_pixels=int(tile_pixels), max_slots=int(max_slots), n_ctas=n_ctas,
        lframe_id=lframe_id, lframe_slot_begin=np.asarray(slot_begin, dtype=np.int32),
        lslot_gslot=np.asarray(lslots, dtype=np.int32),
        wf_lframe=np.asarray(wf_lframe, dtype=np.int32), wf_slot_begin=np.asarray(wf_slot_begin, dtype=np.int32),
        wf_tile_begin=wf_tile_begin.astype(np.int32), wf_accum=np.asarray(wf_accum, dtype=np.int32),
        cta_tile_begin=cta.astype(np.int32),
        total_area_i=float(areas[edges[:, 0]].sum()), total_area_j=float(areas[edges[:, 1]].sum()))
