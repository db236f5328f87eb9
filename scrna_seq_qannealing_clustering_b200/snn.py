"""Synthetic planted-partition SNN graphs (no PBMC download): Gaussian-mixture "PCA" embedding -> kNN (incl. self)
-> Jaccard shared-nearest-neighbour weights -> prune -> symmetric degree trim.

The recipe restates the Seurat pipeline of the reference's notebooks (R/pbmc3k/Pbmc3k_general_data_preparation.Rmd:47-75,
R/benchmarks/Benchmark.Rmd:150-166) and was validated against the shipped fixtures R/benchmarks/graph_*.gexf
(exact match on noisy_circles, aniso, no_structure).  Output format is the hot path's input
contract (create_graphs.py:5-8): undirected weighted graph, node ids '0'..'n-1'.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

Graph = Tuple[int, np.ndarray, np.ndarray, np.ndarray]  # (n, eu, ev, w) with eu < ev, sorted by (eu, ev)


def gaussian_mixture_embedding(n: int, dim: int = 15, centres: int = 8, sep: float = 4.0, seed: int = 0):
    """x_i = centre[label_i] + N(0, I); centres ~ N(0, sep^2 I); labels uniform.  Returns (X [n][dim], labels [n])."""
    rng = np.random.default_rng(seed)
    mu = rng.normal(0.0, sep, size=(centres, dim))
    labels = rng.integers(0, centres, size=n)
    X = mu[labels] + rng.normal(0.0, 1.0, size=(n, dim))
    return X, labels


def knn_including_self(X: np.ndarray, k: int, block: int = 2048) -> np.ndarray:
    """Exact k nearest neighbours of every row, the row itself included (Seurat FindNeighbors k.param).  [n][k] indices."""
    n = X.shape[0]
    k = min(k, n)
    sq = np.einsum("ij,ij->i", X, X)
    out = np.empty((n, k), dtype=np.int64)
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        d = sq[lo:hi, None] + sq[None, :] - 2.0 * (X[lo:hi] @ X.T)
        d[np.arange(hi - lo), np.arange(lo, hi)] = -1.0  # self first
        part = np.argpartition(d, k - 1, axis=1)[:, :k]
        rows = np.arange(hi - lo)[:, None]
        order = np.argsort(d[rows, part], axis=1, kind="stable")
        out[lo:hi] = part[rows, order]
    return out


def snn_graph(X: np.ndarray, k: int = 5, prune: float = 1.0 / 15.0, max_degree: Optional[int] = 15) -> Graph:
    """Jaccard SNN: w_ij = s/(2k - s), s = |kNN(i) & kNN(j)|; drop w < prune; symmetric trim to ``max_degree``."""
    import scipy.sparse as sp

    n = X.shape[0]
    nn = knn_including_self(X, k)
    kk = nn.shape[1]
    A = sp.csr_matrix((np.ones(n * kk, dtype=np.int32), (np.repeat(np.arange(n), kk), nn.ravel())), shape=(n, n))
    S = (A @ A.T).tocoo()
    keep = S.row < S.col
    eu, ev, s = S.row[keep].astype(np.int64), S.col[keep].astype(np.int64), S.data[keep].astype(np.float64)
    w = s / (2.0 * kk - s)
    keep = w >= prune
    eu, ev, w = eu[keep], ev[keep], w[keep]
    order = np.lexsort((ev, eu))
    eu, ev, w = eu[order], ev[order], w[order]
    if max_degree is not None:
        eu, ev, w = symmetric_degree_trim(n, eu, ev, w, max_degree)
    return n, eu, ev, w


def symmetric_degree_trim(n: int, eu, ev, w, max_degree: int):
    """Sequential, in place: for i = 0..n-1 keep the ``max_degree`` heaviest entries of column i (ties -> lower index),
    zero the rest in column i AND row i (Pbmc3k_general_data_preparation.Rmd:69-75)."""
    adj = [dict() for _ in range(n)]
    for u, v, x in zip(eu.tolist(), ev.tolist(), w.tolist()):
        adj[u][v] = x
        adj[v][u] = x
    for i in range(n):
        if len(adj[i]) > max_degree:
            ranked = sorted(adj[i].items(), key=lambda t: (-t[1], t[0]))
            for j, _ in ranked[max_degree:]:
                del adj[i][j]
                del adj[j][i]
    out_u, out_v, out_w = [], [], []
    for u in range(n):
        for v in sorted(adj[u]):
            if u < v:
                out_u.append(u)
                out_v.append(v)
                out_w.append(adj[u][v])
    return np.asarray(out_u, dtype=np.int64), np.asarray(out_v, dtype=np.int64), np.asarray(out_w, dtype=np.float64)


def synthetic_snn(n: int, k: int = 5, dim: int = 15, centres: int = 8, sep: float = 4.0, max_degree: Optional[int] = 15,
                  prune: float = 1.0 / 15.0, seed: int = 0):
    """(graph, planted labels) for ``n`` cells: the benchmark input of BASELINE.json's configs."""
    X, labels = gaussian_mixture_embedding(n, dim, centres, sep, seed)
    return snn_graph(X, k, prune, max_degree), labels


def subsample_problems(n_cells: int, num_problems: int, cells_per_problem: int, k: int = 10, dim: int = 30,
                       max_degree: int = 15, centres: int = 8, seed: int = 0):
    """Config 4: disjoint random subsets of a large embedding, each with its OWN SNN graph (QA_subsampling inputs)."""
    X, _ = gaussian_mixture_embedding(n_cells, dim, centres, 4.0, seed)
    rng = np.random.default_rng(seed + 1)
    perm = rng.permutation(n_cells)
    graphs = []
    for p in range(num_problems):
        idx = np.sort(perm[p * cells_per_problem:(p + 1) * cells_per_problem])
        graphs.append(snn_graph(X[idx], k, 1.0 / 15.0, max_degree))
    return graphs


def to_networkx(graph: Graph):
    """networkx.Graph with string node ids '0'..'n-1' and ``weight`` edge attributes, like ``nx.read_gexf`` returns."""
    import networkx as nx

    n, eu, ev, w = graph
    G = nx.Graph()
    G.add_nodes_from(str(i) for i in range(n))
    G.add_weighted_edges_from((str(u), str(v), float(x)) for u, v, x in zip(eu.tolist(), ev.tolist(), w.tolist()))
    return G


def gaussian_affinity(X: np.ndarray, k: int = 10) -> np.ndarray:
    """Dense affinity A_ij = exp(-|x_i - x_j|^2 / (2 sigma^2)), sigma = median k-NN distance (BASELINE.json config 5)."""
    sq = np.einsum("ij,ij->i", X, X)
    d2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * (X @ X.T), 0.0)
    kth = np.sqrt(np.partition(d2, min(k, len(X) - 1), axis=1)[:, min(k, len(X) - 1)])
    sigma = float(np.median(kth))
    A = np.exp(-d2 / (2.0 * sigma * sigma))
    np.fill_diagonal(A, 0.0)
    return A
