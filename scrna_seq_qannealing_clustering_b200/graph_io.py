"""Graph input / labelled-graph output in the reference's on-disk formats.

Input side (Python_Functions/create_graphs.py:5-18): the SNN graph arrives as a ``.gexf`` written by the R notebooks
(``nx.read_gexf``: node ids are strings '0'..'n-1', edge attribute ``weight``) or as a ``.csv`` edge list whose columns 1-3
are (u, v, weight).  Output side (plot_and_save.py:34, 44, 63, 83, 102, 126): the labelled graph is written back with
``nx.write_gexf`` -- labels live in node attributes (``label<iteration>`` from the recursive bipartition, ``label1`` for the
DQM / CQM / pruning results, ``z_cluster`` for ``clustering_cqm_2``).  The matplotlib drawings of the reference are out of
scope; only the data products are reproduced.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Mapping, Optional, Union

import networkx as nx


def create_graph(path, layout: bool = True, seed: Optional[int] = None):
    """``create_graphs.create_graph``: read a ``.gexf`` graph; returns ``(G, pos)`` with a spring layout like the reference
    (``layout=False`` skips it -- it is only used for plotting and is slow on large graphs; ``pos`` is then ``None``)."""
    G = nx.read_gexf(path)
    pos = nx.spring_layout(G, seed=seed) if layout else None
    return G, pos


def create_graph_csv(dirs: Union[Mapping, str], layout: bool = True, seed: Optional[int] = None):
    """``create_graphs.create_graph_csv``: columns 1-3 of the csv (header row skipped) are weighted edges ``(u, v, w)``.
    ``dirs`` is the reference's directory mapping (key ``"graph_in_csv"``) or a path."""
    import pandas as pd
    path = dirs["graph_in_csv"] if isinstance(dirs, Mapping) else dirs
    data = pd.read_csv(path, header=0, usecols=[1, 2, 3])
    G = nx.Graph()
    G.add_weighted_edges_from(list(data.to_records(index=False)))
    pos = nx.spring_layout(G, seed=seed) if layout else None
    return G, pos


def _path(dirs, key):
    return dirs[key] if isinstance(dirs, Mapping) else dirs


def save_graph_out_bqm(G, dirs, key: str = "graph_out_bqm"):
    """``plot_and_save_graph_out_bqm`` without the drawing: the ``label<i>`` node attributes set by the bipartition
    (``clustering_bqm`` / ``_2`` / ``_3``) are written with the graph.  Returns (cut edges, uncut edges) as the reference
    computes them from the LAST label of every node."""
    last = {n: list(d.values())[-1] for n, d in G.nodes(data=True) if d}
    cut = [(u, v) for u, v in G.edges if last.get(u) != last.get(v)]
    uncut = [(u, v) for u, v in G.edges if last.get(u) == last.get(v)]
    nx.write_gexf(G, _path(dirs, key))
    return cut, uncut


def save_graph_out_dqm(G, dirs, sampleset, key: str = "graph_out_dqm"):
    """``plot_and_save_graph_out_dqm``: ``label1`` = the case of every cell in ``sampleset.first.sample``."""
    lut = dict(sampleset.first.sample)
    nx.set_node_attributes(G, lut, name="label1")
    nx.write_gexf(G, _path(dirs, key))
    return lut


def cqm_labels(G, sample, num_of_clusters: int, by_subindex: bool = False):
    """Cluster of every node from a CQM sample over the binaries ``'v_{i},{p}'`` (plot_and_save.py:46-63; with
    ``by_subindex`` the index is the node's ``subindex`` attribute as in ``clustering_cqm_2``, :65-83).  Nodes without a set
    bit get no label (``defaultdict(int)`` in the reference reads them as 0 only when accessed); with several bits set the
    last one wins, as in the reference's loop."""
    labels = defaultdict(int)
    for node in G.nodes:
        idx = int(G.nodes[node]["subindex"]) if by_subindex else int(node)
        for p in range(num_of_clusters):
            if sample[f"v_{idx},{p}"] == 1:
                labels[idx if by_subindex else node] = p
                if by_subindex:
                    G.nodes[node]["z_cluster"] = p
    return labels


def save_graph_out_cqm(G, dirs, sampleset, num_of_clusters: int, key: str = "graph_out_cqm", by_subindex: bool = False):
    """``plot_and_save_graph_out_cqm`` / ``_cqm_2``: decode ``sampleset.first.sample``, store ``label1``, write the gexf."""
    labels = cqm_labels(G, sampleset.first.sample, num_of_clusters, by_subindex)
    nx.set_node_attributes(G, labels, name="label1")
    nx.write_gexf(G, _path(dirs, key))
    return dict(labels)


def save_graph_out_mvc(G, dirs, key: str = "graph_out_pru1"):
    """``plot_and_save_graph_out_mvc``: the pruning result (``label1`` in {0, 1} set by ``graph_subsampling``); returns the
    (included, excluded) edge lists of the reference."""
    included = [(u, v) for u, v in G.edges if G.nodes[u]["label1"] == 1 or G.nodes[v]["label1"] == 1]
    inc = set(included)
    excluded = [(u, v) for u, v in G.edges if (u, v) not in inc]
    nx.write_gexf(G, _path(dirs, key))
    return included, excluded


def save_graphs_out_cqm_multi(G, sampleset, num_of_clusters: int, number_of_samples: int, prefix: str = "./graphs_multi_samples/sample_number"):
    """``plot_and_save_graph_out_cqm_multi``: one gexf per sample for the first ``number_of_samples - 1`` samples."""
    written = []
    for i, sample in enumerate(list(sampleset.samples())[: number_of_samples - 1]):
        labels = cqm_labels(G, sample, num_of_clusters)
        nx.set_node_attributes(G, labels, name="label1")
        path = f"{prefix}{i}.gexf"
        nx.write_gexf(G, path)
        written.append(path)
    return written
