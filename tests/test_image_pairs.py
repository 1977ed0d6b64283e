"""Scene-graph builder (align3r_b200/image_pairs.py) against the reference's make_pairs (image_pairs.py:11-118): live when
/root/reference is present (build container), and against the closed-form pair counts everywhere."""
import contextlib
import io
import os
import sys

import pytest
import torch

from align3r_b200 import image_pairs as ip
from align3r_b200.synth import make_edge_list

GRAPHS = ["complete", "swin-3", "swin-5-noncyclic", "swin-2-noncyclic", "swinstride-3-noncyclic", "swinstride-2", "swin2stride-2-noncyclic",
          "logwin-3", "logwin-4-noncyclic", "oneref", "oneref-2", "swin"]


def imgs(n):
    return [dict(idx=k, instance=str(k)) for k in range(n)]


@pytest.mark.parametrize("graph", GRAPHS)
@pytest.mark.parametrize("n", [3, 7, 12])
def test_pair_graphs(graph, n):
    pairs = ip.make_pairs(imgs(n), graph, prefilter=None, symmetrize=True)
    edges = [(a["idx"], b["idx"]) for a, b in pairs]
    E = len(edges)
    assert E % 2 == 0
    assert all(edges[k + E // 2] == (edges[k][1], edges[k][0]) for k in range(E // 2)), "edge k + E/2 mirrors edge k"
    # a cyclic window that wraps onto its own frame makes the reference emit self pairs (image_pairs.py:31-37); kept as is
    wraps = graph.startswith("swin") and not graph.endswith("noncyclic") and any(i == j for i, j in edges)
    if not wraps:
        assert len(set(edges)) == E and all(i != j for i, j in edges)
        # same graph as the sorted index-only variant of the synthetic scenes
        assert set(edges) == set(make_edge_list(n, graph))
    half = ip.make_pairs(imgs(n), graph, symmetrize=False)
    assert len(half) == E // 2
    if graph == "complete":
        assert E == n * (n - 1)
    if graph == "swin-5-noncyclic":
        w = min(5, n - 1)
        assert E == 2 * (w * n - w * (w + 1) // 2)


def test_sequence_filters():
    pairs = ip.make_pairs(imgs(9), "complete", prefilter="seq3")
    assert all(abs(a["idx"] - b["idx"]) <= 3 for a, b in pairs) and len(pairs) == 2 * (3 * 9 - 6)
    cyc = ip.make_pairs(imgs(9), "complete", prefilter="cyc2")
    assert len(cyc) == 2 * 2 * 9
    v1 = dict(idx=[0, 0, 1, 5], img=torch.arange(4.0))
    v2 = dict(idx=[1, 4, 2, 1], img=torch.arange(4.0) + 10)
    p1, p2 = dict(conf=torch.arange(4.0)), dict(conf=[0, 1, 2, 3])
    with contextlib.redirect_stdout(io.StringIO()) as out:
        a, b, c, d = ip.filter_edges_seq(v1, v2, p1, p2, 2)
    assert a["idx"] == [0, 1] and b["idx"] == [1, 2] and c["conf"].tolist() == [0.0, 2.0] and d["conf"] == [0, 2]
    assert "kept 2/4 edges" in out.getvalue()


@pytest.mark.skipif(not os.path.isdir("/root/reference/dust3r"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("graph", GRAPHS)
def test_same_pairs_in_the_same_order_as_the_reference(graph):
    sys.path.insert(0, "/root/reference")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from dust3r.image_pairs import make_pairs as ref_make_pairs
            for n in (2, 5, 11):
                for sym in (True, False):
                    for pre in (None, "seq2", "cyc1"):
                        try:
                            ref = ref_make_pairs(imgs(n), graph, prefilter=pre, symmetrize=sym)
                        except IndexError:          # oneref-k with k >= n
                            with pytest.raises(IndexError):
                                ip.make_pairs(imgs(n), graph, prefilter=pre, symmetrize=sym)
                            continue
                        ours = ip.make_pairs(imgs(n), graph, prefilter=pre, symmetrize=sym)
                        assert [(a["idx"], b["idx"]) for a, b in ours] == [(a["idx"], b["idx"]) for a, b in ref], (graph, n, sym, pre)
    finally:
        sys.path.remove("/root/reference")
