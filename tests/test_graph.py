"""Graph tables of the product package: bit-exact with the reference (golden fixture; live reference when mounted)."""
import os

import numpy as np
import pytest

import dsgcn_b200
from dsgcn_b200.graph import Graph, k_adjacency
from oracle import ref_loader as rl

G = os.path.join(os.path.dirname(__file__), "golden")


def test_graph_vs_golden():
    t = np.load(os.path.join(G, "graph_tables.npz"))
    for key in t.files:
        parts = key.split("|")
        if parts[1] == "node_type":
            assert Graph(layout=parts[0]).node_type == t[key].tolist()
        elif parts[1] == "edge_type":
            et = Graph(layout=parts[0]).edge_type
            assert et.dtype == np.float64 and np.array_equal(et, t[key])
        elif parts[1] == "random":
            np.random.seed(7)
            assert np.array_equal(Graph(layout=parts[0], mode="random", num_filter=3, init_off=.04, init_std=.02).A, t[key])
        else:
            assert np.array_equal(Graph(layout=parts[0], mode=parts[1], max_hop=int(parts[2])).A, t[key]), key


def test_graph_errors_like_reference():
    with pytest.raises(AssertionError):
        Graph(layout="nope")
    with pytest.raises(AssertionError):
        Graph(layout="coco", mode="does_not_exist")
    with pytest.raises(AssertionError):
        Graph(layout="coco", mode="spatial", nx_node=2)
    assert not hasattr(Graph(layout="openpose"), "node_type")       # reference: openpose has no semantic tables


@pytest.mark.skipif(not rl.available(), reason="/root/reference not mounted")
def test_graph_vs_live_reference():
    ns = rl.load()
    for layout in ("nturgb+d", "coco", "openpose"):
        for mode, kw in (("spatial", {}), ("stgcn_spatial", {"max_hop": 3}), ("binary_adj", {}), ("random", {"num_filter": 8, "nx_node": 2})):
            np.random.seed(3); a = ns.Graph(layout=layout, mode=mode, **kw)
            np.random.seed(3); b = Graph(layout=layout, mode=mode, **kw)
            assert np.array_equal(a.A, b.A) and np.array_equal(a.hop_dis, b.hop_dis)
            assert a.inward == b.inward and a.neighbor == b.neighbor and a.center == b.center and a.num_node == b.num_node
    A = ns.Graph(layout="coco", mode="binary_adj").A[0]
    for k in range(4):
        assert np.array_equal(ns.graph_module.k_adjacency(A, k, with_self=True), k_adjacency(A, k, with_self=True))
