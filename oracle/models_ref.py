"""Line-by-line restatement of the reference's model-construction loops (pure Python dicts).

TEST INFRASTRUCTURE (see oracle/cpu_sa_ref.cpp header for who may import oracle/).  Each function follows the cited
lines of /root/reference and returns what the reference hands to its sampler; tests/golden/ holds the same objects
captured from the reference's OWN functions (tools/make_golden.py), which pins these restatements.
"""
from __future__ import annotations

from collections import defaultdict
from itertools import combinations


def qubo_clustering_bqm(G, gamma_factor, k=8):
    """BQM_clustering.py:29-47 (k = 8 at :33).  Returns (Q, gamma)."""
    edges_weights = G.size(weight="weight")
    nodes_len = len(G.nodes)
    gamma = gamma_factor * edges_weights / nodes_len
    Q = defaultdict(int)
    for u, v in G.edges:
        Q[(u, u)] += k * G.get_edge_data(u, v)["weight"]
        Q[(v, v)] += k * G.get_edge_data(u, v)["weight"]
        Q[(u, v)] += k * -2 * G.get_edge_data(u, v)["weight"]
    for i in G.nodes:
        Q[(i, i)] += gamma * (1 - len(G.nodes))
    for i, j in combinations(G.nodes, 2):
        Q[(i, j)] += 2 * gamma
    return Q, gamma


def qubo_clustering_bqm_2(G, gamma_factor, k):
    """BQM_clustering.py:210-236.  Returns (Q, gamma)."""
    nodes_len = len(G.nodes)
    weights_sum = G.size(weight="weight")
    gamma = (weights_sum / nodes_len) * gamma_factor
    Q = defaultdict(int)
    for u, v in G.edges:
        Q[(u, u)] += k * G.get_edge_data(u, v)["weight"]
        Q[(v, v)] += k * G.get_edge_data(u, v)["weight"]
        Q[(u, v)] += k * -2 * G.get_edge_data(u, v)["weight"]
    for i in G.nodes:
        Q[(i, i)] += gamma
    return Q, gamma


def qubo_clustering_bqm_3(G, gamma_factor, size_limit, k=8):
    """BQM_clustering.py:357-380.  Returns (Q, constraint terms, lb, ub, lagrange_multiplier)."""
    edges_weights = G.size(weight="weight")
    nodes_len = len(G.nodes)
    gamma = gamma_factor * edges_weights / nodes_len
    Q = defaultdict(int)
    for u, v in G.edges:
        Q[(u, u)] += k * G.get_edge_data(u, v)["weight"]
        Q[(v, v)] += k * G.get_edge_data(u, v)["weight"]
        Q[(u, v)] += k * -2 * G.get_edge_data(u, v)["weight"]
    x = [str(n) for n in G.nodes()]
    c1 = [(x[int(n)], 1) for n in G.nodes()]
    return Q, c1, size_limit, len(G.nodes) / 6, gamma


def qubo_graph_subsampling(G, gamma):
    """QA_subsampling.py:26-35 (P = 1)."""
    P = 1
    Q = defaultdict(int)
    for u, v in G.edges:
        Q[(u, u)] += -P * (1 - G.get_edge_data(u, v)["weight"])
        Q[(v, v)] += -P * (1 - G.get_edge_data(u, v)["weight"])
        Q[(u, v)] += P * (1 - G.get_edge_data(u, v)["weight"])
    for i in G.nodes:
        Q[(i, i)] += gamma
    return Q


def dqm_clustering(G, num_of_clusters, gamma):
    """DQM_clustering.py:25-43 with dimod's *set* (overwrite) semantics.  Returns (linear, quadratic):
    linear[node] = list of K biases; quadratic[(u, v)] = {(c, c): bias} keyed by the first orientation seen."""
    nodes = G.nodes
    edges = G.edges
    clusters = [i for i in range(0, num_of_clusters)]
    linear, quadratic = {}, {}
    for node in nodes:
        linear[node] = [0.0] * num_of_clusters
    for node in nodes:
        linear[node] = [gamma * (1 - len(G.nodes) / num_of_clusters) for cluster in clusters]
    for i, j in combinations(nodes, 2):
        quadratic.setdefault((i, j) if (j, i) not in quadratic else (j, i), {}).update(
            {(cluster, cluster): 2 * gamma for cluster in clusters})
    for u, v in edges:
        key = (u, v) if (v, u) not in quadratic else (v, u)
        quadratic.setdefault(key, {}).update({(cluster, cluster): -2 * G.get_edge_data(u, v)["weight"] for cluster in clusters})
        linear[u] = [G.get_edge_data(u, v)["weight"] for cluster in clusters]
        linear[v] = [G.get_edge_data(u, v)["weight"] for cluster in clusters]
    return linear, quadratic


def cqm_clustering(G, num_of_clusters, min_size=20):
    """CQM_clustering.py:26-48 as plain coefficient dicts.  Returns (objective_linear, objective_quadratic, discretes,
    size_constraints) with variable labels 'v_{i},{k}'."""
    nodes = G.nodes
    edges = G.edges
    clusters = range(num_of_clusters)
    lin, quad = defaultdict(float), defaultdict(float)
    discretes = {f"one-hot-node-{i}": [f"v_{i},{k}" for k in clusters] for i in nodes}
    for i, j in edges:
        for p in clusters:
            lin[f"v_{i},{p}"] += 1.0
            lin[f"v_{j},{p}"] += 1.0
            quad[(f"v_{i},{p}", f"v_{j},{p}")] += -2 * G.get_edge_data(i, j)["weight"]
    size = {f"cluster_size{j}": ([f"v_{i},{j}" for i in nodes], min_size) for j in clusters}
    return dict(lin), dict(quad), discretes, size
