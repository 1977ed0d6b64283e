"""Scene-graph builder with the reference's interface (dust3r/image_pairs.py:11-118): which image pairs the pairwise
network is run on, and the sequence-distance edge filters.  Host-side graph logic only - the callers either side of the
alignment path use it (`tool/demo.py:139-147`, `pose_test.py:31`); the edge list it produces is what
``PointCloudOptimizer`` receives as ``view1['idx'] / view2['idx']``.

The undirected pairs of the window graphs are collected in a Python ``set`` of index tuples and emitted in its iteration
order, like the reference does, so that with the same interpreter the edge ORDER is the reference's as well (the optimiser
itself only needs edge ``e`` and ``e + E/2`` to be mirror images after symmetrisation, which holds for any order).
``align3r_b200.synth.make_edge_list`` is the index-only variant with sorted pairs used by the synthetic scenes.
"""
from __future__ import annotations

import numpy as np
import torch


def _window_size(scene_graph, default=3):
    try:
        return int(scene_graph.split('-')[1])
    except Exception:
        return default


def pair_indices(n, scene_graph='complete', symmetrize=True):
    """Index pairs (i, j) of ``make_pairs`` for ``n`` images (reference image_pairs.py:11-67)."""
    ids = []
    if scene_graph == 'complete':
        ids = [(i, j) for i in range(n) for j in range(i)]
    elif scene_graph.startswith('swin'):
        cyclic = not scene_graph.endswith('noncyclic')
        win = _window_size(scene_graph)
        stride = 2 if scene_graph.startswith('swinstride') else 3 if scene_graph.startswith('swin2stride') else 1
        seen = set()
        for i in range(n):
            for off in range(1, stride * win + 1, stride):
                j = (i + off) % n if cyclic else i + off
                if j < n:
                    seen.add((i, j) if i < j else (j, i))
        ids = list(seen)
    elif scene_graph.startswith('logwin'):
        cyclic = not scene_graph.endswith('noncyclic')
        offsets = [2 ** k for k in range(_window_size(scene_graph))]
        seen = set()
        for i in range(n):
            for j in [i - off for off in offsets] + [i + off for off in offsets]:
                if cyclic:
                    j %= n
                if 0 <= j < n and j != i:
                    seen.add((i, j) if i < j else (j, i))
        ids = list(seen)
    elif scene_graph.startswith('oneref'):
        ref = int(scene_graph.split('-')[1]) if '-' in scene_graph else 0
        ids = [(ref, j) for j in range(n) if j != ref]
    if symmetrize:
        ids = ids + [(j, i) for i, j in ids]
    return ids


def make_pairs(imgs, scene_graph='complete', prefilter=None, symmetrize=True):
    """reference image_pairs.py:11-75: list of (img_i, img_j) tuples of the image dicts in ``imgs``."""
    pairs = [(imgs[i], imgs[j]) for i, j in pair_indices(len(imgs), scene_graph, symmetrize)]
    if isinstance(prefilter, str) and prefilter.startswith('seq'):
        pairs = filter_pairs_seq(pairs, int(prefilter[3:]))
    if isinstance(prefilter, str) and prefilter.startswith('cyc'):
        pairs = filter_pairs_seq(pairs, int(prefilter[3:]), cyclic=True)
    return pairs


def sel(x, kept):
    """reference image_pairs.py:78-84: keep rows ``kept`` of a (nested) container."""
    if isinstance(x, dict):
        return {k: sel(v, kept) for k, v in x.items()}
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x[kept]
    if isinstance(x, (tuple, list)):
        return type(x)([x[k] for k in kept])


def _filter_edges_seq(edges, seq_dis_thr, cyclic=False):
    n = max(max(e) for e in edges) + 1
    kept = []
    for e, (i, j) in enumerate(edges):
        dis = abs(i - j)
        if cyclic:
            dis = min(dis, abs(i + n - j), abs(i - n - j))
        if dis <= seq_dis_thr:
            kept.append(e)
    return kept


def filter_pairs_seq(pairs, seq_dis_thr, cyclic=False):
    """reference image_pairs.py:102-105."""
    kept = _filter_edges_seq([(a['idx'], b['idx']) for a, b in pairs], seq_dis_thr, cyclic=cyclic)
    return [pairs[k] for k in kept]


def filter_edges_seq(view1, view2, pred1, pred2, seq_dis_thr, cyclic=False):
    """reference image_pairs.py:108-112 (same console line)."""
    edges = [(int(i), int(j)) for i, j in zip(view1['idx'], view2['idx'])]
    kept = _filter_edges_seq(edges, seq_dis_thr, cyclic=cyclic)
    print(f'>> Filtering edges more than {seq_dis_thr} frames apart: kept {len(kept)}/{len(edges)} edges')
    return sel(view1, kept), sel(view2, kept), sel(pred1, kept), sel(pred2, kept)
