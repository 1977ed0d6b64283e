"""Host-side planning: slots, work frames, tiles, CTA ranges, frame-window shards (pure numpy)."""
import numpy as np
import pytest

from align3r_b200.plan import build_plan, partition_frames, frame_costs, slots_per_frame, check_edges
from align3r_b200.synth import make_edge_list


def _plan(n, graph, H=288, W=512, ctas=296, rng=None, max_slots=32):
    edges = make_edge_list(n, graph)
    return edges, build_plan(edges, [(H, W)] * n, 1024, max_slots, ctas, rng)


@pytest.mark.parametrize("n,graph,E", [(10, "swin-3-noncyclic", 48), (50, "swin-5-noncyclic", 470),
                                       (90, "swin-5-noncyclic", 870), (1000, "swin-5-noncyclic", 9970)])
def test_edge_counts_match_survey_table(n, graph, E):
    assert len(make_edge_list(n, graph)) == E


def test_slots_are_frame_major_and_complete():
    edges, pl = _plan(50, "swin-5-noncyclic")
    assert pl.n_local_slots == 2 * len(edges)
    assert sorted(pl.lslot_gslot.tolist()) == list(range(2 * len(edges)))
    for k, n in enumerate(pl.lframe_id.tolist()):
        for s in range(pl.lframe_slot_begin[k], pl.lframe_slot_begin[k + 1]):
            gs = int(pl.lslot_gslot[s])
            assert edges[gs >> 1][gs & 1] == n          # the slot lives on this frame's grid
    deg = slots_per_frame(np.asarray(edges), 50)
    assert deg.max() == 20 and deg.min() == 10 and int(deg.sum()) == 2 * len(edges)
    assert pl.pixel_slots() == 2 * len(edges) * 288 * 512


def test_cta_ranges_cover_tiles_once_and_are_balanced():
    edges, pl = _plan(50, "swin-5-noncyclic")
    c = pl.cta_tile_begin
    assert c[0] == 0 and c[-1] == pl.n_tiles and np.all(np.diff(c) >= 0)
    deg = np.diff(pl.wf_slot_begin) + 2.0
    cost = np.repeat(deg, np.diff(pl.wf_tile_begin))
    per_cta = np.array([cost[c[k]:c[k + 1]].sum() for k in range(pl.n_ctas)])
    assert per_cta.max() <= per_cta.mean() + cost.max() + 1e-9


def test_high_degree_frames_are_split_into_work_frames():
    edges, pl = _plan(20, "complete", H=32, W=32, ctas=8, max_slots=8)
    assert pl.n_work_frames > pl.n_local_frames
    assert np.all(np.diff(pl.wf_slot_begin) <= 8)
    split = {int(f) for f in pl.wf_lframe[pl.wf_accum == 1]}
    counts = np.bincount(pl.wf_lframe)
    assert all(counts[f] > 1 for f in split) and all(counts[f] == 1 for f in range(20) if f not in split)
    # every slot appears in exactly one work frame
    assert pl.wf_slot_begin[-1] == pl.n_local_slots


def test_ragged_and_tiny_inputs():
    pl = build_plan([(0, 1), (1, 0)], [(3, 5), (4, 4)], 8, 32, 4)
    assert pl.pix_stride == 16 and pl.n_tiles == 4          # planes are padded to whole tiles
    assert pl.total_area_i == 15 + 16 and pl.total_area_j == 16 + 15
    with pytest.raises(AssertionError):
        check_edges([(0, 2)])                      # reference base_opt.py:184: indices must be contiguous


@pytest.mark.parametrize("world", [2, 4, 8])
def test_frame_windows_partition_and_balance(world):
    n = 1000
    edges = np.asarray(make_edge_list(n, "swin-5-noncyclic"))
    hw = np.asarray([(288, 512)] * n)
    cost = frame_costs(edges, hw)
    wins = partition_frames(cost, world)
    assert wins[0][0] == 0 and wins[-1][1] == n
    assert all(wins[r][1] == wins[r + 1][0] for r in range(world - 1))
    shares = np.array([cost[a:b].sum() for a, b in wins])
    assert shares.max() / shares.mean() < 1.02
    seen = []
    for a, b in wins:
        pl = build_plan(edges, hw, 1024, 32, 16, (a, b))
        seen += pl.lslot_gslot.tolist()
        assert pl.lframe_id.tolist() == list(range(a, b))
    assert sorted(seen) == list(range(2 * len(edges)))          # every slot is owned by exactly one window


def test_more_ranks_than_frames():
    wins = partition_frames(np.ones(3), 8)
    assert wins[0][0] == 0 and wins[-1][1] == 3
    assert sum(b - a for a, b in wins) == 3


def test_incidence_lists_cover_every_row_once_per_camera():
    """The update kernel gathers, per frame, the rows of pair_scratch that carry a gradient contribution to its camera
    (align3r_b200/problem.py::incidence_lists): every flow slot appears once as source (on the frame it lives on) and once as
    target (on the pair's other frame), every temporal pair once on each of its two frames, in ascending row order."""
    from align3r_b200.problem import incidence_lists
    from align3r_b200.synth import make_edge_list
    for n, graph in ((7, "swin-3-noncyclic"), (5, "complete"), (9, "swinstride-2-noncyclic")):
        edges = make_edge_list(n, graph)
        E = len(edges)
        begin, items = incidence_lists(edges, n)
        assert begin[0] == 0 and begin[-1] == len(items) == 4 * E + 2 * (n - 1)
        seen = set()
        for f in range(n):
            codes = items[begin[f]:begin[f + 1]].tolist()
            assert codes == sorted(codes)
            for c in codes:
                row, part = c >> 1, c & 1
                if row < 2 * E:
                    e, side = row >> 1, row & 1
                    src, tgt = edges[e][side], edges[e][1 - side]
                    assert f == (tgt if part else src)
                else:
                    k = row - 2 * E                  # temporal pair (source k+1, target k)
                    assert f == (k if part else k + 1)
                assert c not in seen
                seen.add(c)
        assert len(seen) == 4 * E + 2 * (n - 1)


def test_recipe_parameters_are_reproducible_and_near_the_solution():
    """synth.near_solution_params: the closed-form parameter recipe the full-size fixtures rely on gives the same values on
    every call and a loss of the order of the MST initialisation's (checked through the oracle)."""
    import torch
    from align3r_b200.synth import CANONICAL_KW, make_scene, near_solution_params
    from oracle import ga_oracle as go
    sc = make_scene(7, 24, 32, "swin-2-noncyclic", seed=4)
    a, lda = near_solution_params(sc.gt, sc.edges, seed=0)
    b, ldb = near_solution_params(sc.gt, sc.edges, seed=0)
    assert all(torch.equal(a[k], b[k]) for k in a) and torch.equal(lda(3), ldb(3))
    prob = go.problem_from_output(sc.output, sc.flow_ij, sc.flow_ji, **dict(CANONICAL_KW, num_total_iter=10))
    params, _ = go.mst_init(prob)
    l_mst, _ = go.total_loss(prob, params, 9999)
    p = dict(params, pw_poses=a["pw_poses"], im_poses=a["im_poses"], im_focals=a["im_focals"].reshape(params["im_focals"].shape),
             im_depthmaps=torch.stack([lda(n) for n in range(7)]))
    l_rec, info = go.total_loss(prob, p, 9999)
    assert float(l_rec) < 10 * float(l_mst) and not info["flow_dropped"]


def test_host_pool_is_shared_between_ranks(monkeypatch):
    from align3r_b200._timing import host_workers
    import os
    monkeypatch.delenv("A3R_HOST_THREADS", raising=False)
    cores = os.cpu_count() or 1
    assert host_workers(1000) == cores and host_workers(3) == min(3, cores)
    assert host_workers(1000, dict(world=8, rank=0)) == max(1, cores // 8)
    # a rank that knows the jobs of all ranks takes its share of the cores: 3 of 12 jobs on 16 cores are 3 threads, not 16 // 8
    monkeypatch.setattr(os, "cpu_count", lambda: 16)
    assert host_workers(3, dict(world=8, rank=0), total_jobs=12) == 3
    assert host_workers(1, dict(world=8, rank=0), total_jobs=12) == 1
    assert host_workers(12, None, total_jobs=12) == 12
    assert host_workers(40, dict(world=8, rank=0), total_jobs=320) == 2
    monkeypatch.setenv("A3R_HOST_THREADS", "2")
    assert host_workers(1000, dict(world=8, rank=0)) == 2
