"""Skeleton graph tables — same constructor, attributes and values as pyskl/utils/graph.py:58-187.

`Graph(layout, mode, max_hop, nx_node, num_filter, init_std, init_off)` exposes `.A [K,V,V] float64`,
`.node_type` (list[int]), `.edge_type` ([V,V] float64 holding integers 0..14), `.num_node`, `.inward`,
`.outward`, `.neighbor`, `.self_link`, `.center`, `.hop_dis`.  Integer tables are bit-exact with the
reference (tests/test_graph.py); 'random' mode consumes numpy's global RNG like the reference does.
"""
import numpy as np
import torch

_NTU = ((1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9), (11, 10), (12, 11), (13, 1),
        (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19), (22, 8), (23, 8), (24, 12), (25, 12))

LAYOUTS = {
    "openpose": dict(num_node=18, center=1, node_type=None,
                     inward=[(4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2), (5, 1),
                             (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14)]),
    "nturgb+d": dict(num_node=25, center=20, inward=[(i - 1, j - 1) for i, j in _NTU],
                     node_type=[0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 0, 1, 1, 2, 2]),
    "coco": dict(num_node=17, center=0, node_type=[0, 0, 0, 0, 0, 1, 2, 1, 2, 1, 2, 3, 4, 3, 4, 3, 4],
                 inward=[(15, 13), (13, 11), (16, 14), (14, 12), (11, 5), (12, 6), (9, 7), (7, 5), (10, 8), (8, 6), (5, 0),
                         (6, 0), (1, 0), (3, 1), (2, 0), (4, 2)]),
}


def k_adjacency(A, k, with_self=False, self_factor=1):
    """graph.py:5-16"""
    if isinstance(A, torch.Tensor):
        A = A.data.cpu().numpy()
    eye = np.eye(len(A), dtype=A.dtype)
    if k == 0:
        return eye
    reach = lambda p: np.minimum(np.linalg.matrix_power(A + eye, p), 1)
    Ak = reach(k) - reach(k - 1)
    if with_self:
        Ak += self_factor * eye
    return Ak


def edge2mat(link, num_node):
    A = np.zeros((num_node, num_node))
    for i, j in link:
        A[j, i] = 1
    return A


def normalize_digraph(A, dim=0):
    deg = np.sum(A, dim)
    inv = np.zeros((A.shape[1], A.shape[1]))
    nz = deg > 0
    inv[nz, nz] = deg[nz] ** (-1)
    return np.dot(A, inv)


def get_hop_distance(num_node, edge, max_hop=1):
    A = np.eye(num_node)
    for i, j in edge:
        A[i, j] = A[j, i] = 1
    hop = np.zeros((num_node, num_node)) + np.inf
    reach = np.stack([np.linalg.matrix_power(A, d) for d in range(max_hop + 1)]) > 0
    for d in range(max_hop, -1, -1):
        hop[reach[d]] = d
    return hop


def semantic_tables(node_type):
    """edge_type[u,w] = rank of s_u*s_w among the distinct products, s = (t+1)*(-1)^(t+1)   (graph.py:119-126)"""
    t = np.asarray(node_type).reshape(-1, 1) + 1
    s = t * np.power(-1, t)
    prod = np.dot(s, s.T)
    uniq = np.unique(prod)
    return np.searchsorted(uniq, prod).astype(np.float64), uniq


class Graph:
    def __init__(self, layout="coco", mode="spatial", max_hop=1, nx_node=1, num_filter=3, init_std=0.02, init_off=0.04):
        self.max_hop, self.layout, self.mode = max_hop, layout, mode
        self.num_filter, self.init_std, self.init_off, self.nx_node = num_filter, init_std, init_off, nx_node
        assert nx_node == 1 or mode == "random", "nx_node can be > 1 only if mode is 'random'"
        assert layout in LAYOUTS
        spec = LAYOUTS[layout]
        self.num_node, self.center = spec["num_node"], spec["center"]
        self.inward = list(spec["inward"])
        if spec["node_type"] is not None:
            self.node_type = list(spec["node_type"])
            self.edge_type, self.edge_type_num = semantic_tables(self.node_type)
        self.self_link = [(i, i) for i in range(self.num_node)]
        self.outward = [(j, i) for i, j in self.inward]
        self.neighbor = self.inward + self.outward
        self.hop_dis = get_hop_distance(self.num_node, self.inward, max_hop)
        assert hasattr(self, mode), f"Do Not Exist This Mode: {mode}"
        self.A = getattr(self, mode)()

    def __str__(self):
        return self.A

    def stgcn_spatial(self):
        adj = (self.hop_dis <= self.max_hop).astype(np.float64)
        nadj = normalize_digraph(adj)
        hc = self.hop_dis[:, self.center]
        far_or_equal = hc[:, None] >= hc[None, :]
        A = []
        for hop in range(self.max_hop + 1):
            sel = self.hop_dis == hop
            A.append(np.where(sel & far_or_equal, nadj, 0.0))
            if hop > 0:
                A.append(np.where(sel & ~far_or_equal, nadj, 0.0))
        return np.stack(A)

    def spatial(self):
        iden = edge2mat(self.self_link, self.num_node)
        return np.stack((iden, normalize_digraph(edge2mat(self.inward, self.num_node)),
                         normalize_digraph(edge2mat(self.outward, self.num_node))))

    def binary_adj(self):
        return edge2mat(self.inward + self.outward, self.num_node)[None]

    def random(self):
        n = self.num_node * self.nx_node
        return np.random.randn(self.num_filter, n, n) * self.init_std + self.init_off
