"""Graph I/O in the reference's on-disk formats (create_graphs.py:5-18; the gexf export of plot_and_save.py): round trips of
the fixture graphs through .gexf and .csv, and the label decoding the reference's writers perform."""
from pathlib import Path
from types import SimpleNamespace

import networkx as nx
import numpy as np
import pytest

from scrna_seq_qannealing_clustering_b200 import graph_io

GOLD = Path(__file__).parent / "golden"


def fixture_graph(name):
    g = np.load(GOLD / "graphs.npz")
    labels = [str(x) for x in g[f"{name}_labels"]]
    G = nx.Graph()
    G.add_nodes_from(labels)
    for u, v, w in zip(g[f"{name}_eu"], g[f"{name}_ev"], g[f"{name}_w"]):
        G.add_edge(labels[u], labels[v], weight=float(w))
    return G


@pytest.mark.parametrize("name", ["noisy_circles", "varied"])
def test_gexf_round_trip_keeps_nodes_order_and_weights(tmp_path, name):
    G = fixture_graph(name)
    path = tmp_path / "g.gexf"
    nx.write_gexf(G, path)
    H, pos = graph_io.create_graph(path, layout=False)
    assert pos is None
    assert list(H.nodes) == list(G.nodes)                       # node ids stay the strings '0'..'n-1', in file order
    assert [(u, v) for u, v in H.edges] == [(u, v) for u, v in G.edges]
    assert all(H[u][v]["weight"] == G[u][v]["weight"] for u, v in G.edges)
    H2, pos2 = graph_io.create_graph(path, layout=True, seed=1)
    assert set(pos2) == set(H2.nodes)


def test_csv_edge_list_uses_columns_1_to_3(tmp_path):
    import pandas as pd
    G = fixture_graph("noisy_moons")
    rows = [(i, int(u), int(v), d["weight"]) for i, (u, v, d) in enumerate(G.edges(data=True))]
    path = tmp_path / "edges.csv"
    pd.DataFrame(rows, columns=["", "from", "to", "weight"]).to_csv(path, index=False)
    H, _ = graph_io.create_graph_csv({"graph_in_csv": str(path)}, layout=False)
    assert H.number_of_edges() == G.number_of_edges()
    for u, v, d in G.edges(data=True):
        assert H[int(u)][int(v)]["weight"] == pytest.approx(d["weight"])
    H2, _ = graph_io.create_graph_csv(str(path), layout=False)
    assert H2.number_of_edges() == H.number_of_edges()


def test_bipartition_labels_are_exported_and_cut_edges_follow_the_last_label(tmp_path):
    G = fixture_graph("noisy_circles")
    comps = sorted(nx.connected_components(G), key=len)
    for n in G.nodes:
        G.nodes[n]["label0"] = 7 if n in comps[0] else 150
    half = set(list(comps[0])[: len(comps[0]) // 2])
    for n in comps[0]:
        G.nodes[n]["label1"] = 3 if n in half else 130
    cut, uncut = graph_io.save_graph_out_bqm(G, {"graph_out_bqm": str(tmp_path / "out.gexf")})
    assert len(cut) + len(uncut) == G.number_of_edges()
    assert all((u in half) != (v in half) or (u in comps[0]) != (v in comps[0]) for u, v in cut)
    H = nx.read_gexf(tmp_path / "out.gexf")
    assert all(H.nodes[n]["label0"] == G.nodes[n]["label0"] for n in G.nodes)
    assert all(H.nodes[n].get("label1") == G.nodes[n].get("label1") for n in G.nodes)


def test_dqm_and_cqm_samples_become_label1(tmp_path):
    G = nx.path_graph(6)
    G = nx.relabel_nodes(G, {i: str(i) for i in range(6)})
    K = 3
    want = {str(i): i % K for i in range(6)}
    dqm = SimpleNamespace(first=SimpleNamespace(sample=dict(want)))
    graph_io.save_graph_out_dqm(G, str(tmp_path / "dqm.gexf"), dqm)
    assert {n: d["label1"] for n, d in nx.read_gexf(tmp_path / "dqm.gexf").nodes(data=True)} == want
    sample = {f"v_{i},{p}": int(want[str(i)] == p) for i in range(6) for p in range(K)}
    cqm = SimpleNamespace(first=SimpleNamespace(sample=sample), samples=lambda: [sample, sample, sample])
    labels = graph_io.save_graph_out_cqm(G, {"graph_out_cqm": str(tmp_path / "cqm.gexf")}, cqm, K)
    assert labels == want
    assert {n: d["label1"] for n, d in nx.read_gexf(tmp_path / "cqm.gexf").nodes(data=True)} == want
    # clustering_cqm_2 keys the binaries on the per-component 'subindex' attribute
    for i, n in enumerate(reversed(list(G.nodes))):
        G.nodes[n]["subindex"] = i
    sample2 = {f"v_{G.nodes[n]['subindex']},{p}": int(want[n] == p) for n in G.nodes for p in range(K)}
    cqm2 = SimpleNamespace(first=SimpleNamespace(sample=sample2))
    graph_io.save_graph_out_cqm(G, str(tmp_path / "cqm2.gexf"), cqm2, K, by_subindex=True)
    assert all(G.nodes[n]["z_cluster"] == want[n] for n in G.nodes)
    written = graph_io.save_graphs_out_cqm_multi(G, cqm, K, 3, prefix=str(tmp_path / "multi"))
    assert len(written) == 2 and all(Path(p).exists() for p in written)


def test_pruning_result_edges(tmp_path):
    G = nx.cycle_graph(5)
    for n in G.nodes:
        G.nodes[n]["label1"] = int(n in (0, 2))
    inc, exc = graph_io.save_graph_out_mvc(G, str(tmp_path / "pru.gexf"))
    assert set(inc) == {(0, 1), (0, 4), (1, 2), (2, 3)} and set(exc) == {(3, 4)}
