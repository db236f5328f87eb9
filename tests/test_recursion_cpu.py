"""Host logic of the level-batched recursion (clustering.recursive_bipartition_batched) on CPU: the device entry points it
drives (qa_graph_split, qa_build_cut_balance, qa_model_concat, qa_sa_sample_model_batch) are replaced by numpy / oracle-backed
stand-ins with the same semantics, so the level bookkeeping, the per-problem seeds / initial states / beta schedules and the
termination rules are checked without a GPU (tests/test_gpu_recursion.py runs the same driver through the C ABI)."""
import networkx as nx
import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import clustering, models, schedule, snn


def host_split(graph, part_of, P):
    """Restatement of qa_graph_split: children keep the parent's node order and edge order, nodes relabelled by rank."""
    n, eu, ev, w = graph
    out = []
    for p in range(P):
        nodes = np.flatnonzero(part_of == p)
        local = np.full(n, -1, dtype=np.int64)
        local[nodes] = np.arange(len(nodes))
        keep = (part_of[eu] == p) & (part_of[ev] == p)
        out.append((len(nodes), local[eu[keep]], local[ev[keep]], np.asarray(w)[keep]))
    return out


class _Split:
    def __init__(self, graphs):
        self.graphs = graphs

    def device_graph(self, p):
        return self.graphs[p]

    graph = device_graph

    def close(self):
        pass


class _Model:
    def __init__(self, m):
        self.m = m
        self.num_variables = m.num_variables

    def get_ising(self):
        return self.m.h, self.m.starts, self.m.ends, self.m.weights

    def close(self):
        pass


class OracleBatchContext:
    """Stand-in for the `Context` methods the driver uses; every batched launch is counted."""

    def __init__(self):
        self.calls = 0

    def split_graph(self, graph, part_of, P):
        return _Split(host_split(graph, np.asarray(part_of), P))

    def build_cut_balance(self, graph, gamma_factor, k):
        m = models.cut_balance_model(graph, gamma_factor, k=k, structured=True)
        return _Model(m), m.offset, m.meta["gamma"]

    def build_cut_linear(self, graph, gamma_factor, k):
        m = models.cut_linear_model(graph, gamma_factor, k)
        return _Model(m), m.offset, m.meta["gamma"]

    def concat_models(self, gms):
        return type("Batch", (), {"parts": list(gms), "close": lambda self: None})()

    def sample_model_batch(self, bm, rpp, states, betas, spb, seeds):
        self.calls += 1
        e = np.empty(len(bm.parts) * rpp)
        off = 0
        for p, gm in enumerate(bm.parts):
            m, n = gm.m, gm.num_variables
            st = states[off:off + rpp * n].reshape(rpp, n).copy()
            b = betas[p] if np.ndim(betas) == 2 else betas
            ee, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, st, np.ascontiguousarray(b), spb, seeds[p * rpp:(p + 1) * rpp],
                                        groups=m.groups.astuple() if m.groups is not None else None)
            states[off:off + rpp * n] = st.ravel()
            e[p * rpp:(p + 1) * rpp] = ee
            off += rpp * n
        return e, None, rpp


def test_four_blobs_are_recovered_with_one_launch_per_level():
    X, truth = snn.gaussian_mixture_embedding(240, dim=8, centres=4, sep=9.0, seed=1)
    G = snn.to_networkx(snn.snn_graph(X, k=10))
    ctx = OracleBatchContext()
    labels, levels, energies = clustering.recursive_bipartition_batched(G, gamma_factor=0.05, k=8.0, size_limit=30, iter_limit=2,
                                                                        num_reads=96, num_sweeps=300, seed=3, context=ctx)
    # min_size rule, literally: a level that stops writes nothing, so the blobs (halves <= size_limit) are the leaves
    assert ctx.calls == 3 and [len(lv) for lv in levels] == [1, 2, 4]
    assert set(labels) == set(G.nodes) and len(set(labels.values())) == 4
    node_truth = {str(i): int(t) for i, t in enumerate(truth)}
    for leaf in set(labels.values()):
        assert len({node_truth[n] for n, l in labels.items() if l == leaf}) == 1      # every leaf is one planted blob
    # level 0 is exactly one neal-style call on the whole graph with the documented seeds / initial states / default range
    # (graph given as arrays: networkx's G.size() sums with Python 3.12's compensated sum(), the device builders and the array
    # path of models.py with plain left-to-right adds -- one ulp apart in gamma)
    lab, eu, ev, w = models.graph_arrays(G)
    m = models.cut_balance_model((lab, eu, ev, w), 0.05, k=8.0, structured=True)
    br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
    betas, spb = schedule.make_beta_schedule(br, 300, 1, "geometric")
    st = schedule.random_spin_states(96, m.num_variables, 3)
    e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, st, betas, spb, schedule.per_read_seeds(3, 96), groups=m.groups.astuple())
    assert energies[0][0] == float(e.min() + m.offset)


def test_size_limit_and_iteration_limit_stop_the_recursion():
    X, _ = snn.gaussian_mixture_embedding(120, dim=6, centres=2, sep=8.0, seed=2)
    G = snn.to_networkx(snn.snn_graph(X, k=8))
    ctx = OracleBatchContext()
    labels, levels, _ = clustering.recursive_bipartition_batched(G, 0.05, size_limit=1000, iter_limit=5, num_reads=32, num_sweeps=100,
                                                                 seed=1, context=ctx)
    assert ctx.calls == 1 and len(set(labels.values())) == 1            # min_size writes nothing when it stops: one cluster
    ctx = OracleBatchContext()
    labels, levels, _ = clustering.recursive_bipartition_batched(G, 0.05, size_limit=5, iter_limit=0, num_reads=32, num_sweeps=100,
                                                                 seed=1, context=ctx)
    assert ctx.calls == 1                                               # iteration limit reached at the first level
    ctx = OracleBatchContext()
    labels, levels, _ = clustering.recursive_bipartition_batched(G, 0.05, terminate_on="once", num_reads=32, num_sweeps=100, seed=1,
                                                                 context=ctx)
    assert ctx.calls == 1 and len(set(labels.values())) == 2            # `once`: one split, both halves labelled


class _RecordingSampler:
    """Records (node set, S0, S1) of every sampler call of the call-by-call recursion; anneals with the oracle."""

    def __init__(self):
        self.calls = []

    def sample(self, model, num_reads=None, num_sweeps=None, seed=None, sorted=False, **kw):
        from scrna_seq_qannealing_clustering_b200.sampleset import SampleSet
        br = schedule.default_ising_beta_range(model.h, model.starts, model.ends, model.weights, None)
        betas, spb = schedule.make_beta_schedule(br, num_sweeps, 1, "geometric")
        st = schedule.random_spin_states(num_reads, model.num_variables, seed)
        e, _ = oracle.sample_ising(model.h, model.starts, model.ends, model.weights, st, betas, spb,
                                   schedule.per_read_seeds(seed, num_reads), groups=model.groups.astuple())
        ss = SampleSet.from_samples((((st + 1) // 2).astype(np.int8), model.labels), energy=e + model.offset, vartype="BINARY").sorted()
        first = ss.first.sample
        self.calls.append((frozenset(model.labels), frozenset(v for v in model.labels if not first[v])))
        return ss


@pytest.mark.parametrize("rule,kw", [("min_size", {"size_limit": 30, "iter_limit": 3}), ("conf", {"iter_limit": 2}),
                                     ("iter_limit", {"iter_limit": 2}), ("once", {})])
def test_batched_driver_builds_the_tree_of_the_call_by_call_recursion(rule, kw):
    """Every termination rule: the level-batched driver anneals exactly the sub-graphs the reference-shaped recursion
    (clustering_bqm calling itself on G.subgraph(S0 / S1)) anneals, and splits them identically."""
    X, _ = snn.gaussian_mixture_embedding(200, dim=8, centres=4, sep=9.0, seed=4)
    G = snn.to_networkx(snn.snn_graph(X, k=10))
    rec = _RecordingSampler()
    clustering.clustering_bqm(G.copy(), 0, {"name": "t"}, "rec", 0.05, terminate_on=rule, sampler=rec, device_build=False,
                              num_reads=24, num_sweeps=60, seed=9, **kw)
    ctx = OracleBatchContext()
    _, levels, _ = clustering.recursive_bipartition_batched(G, 0.05, terminate_on=rule, num_reads=24, num_sweeps=60, seed=9, context=ctx,
                                                            **kw)
    assert {c[0] for c in rec.calls} == {frozenset(part) for lv in levels for part in lv}
    assert ctx.calls == len(levels)


def test_batched_driver_runs_the_cut_linear_model_too():
    """clustering_bqm_2's sparse model (k * cut + gamma * sum x) through the same driver, `min_size` rule of bqm2_rule."""
    X, _ = snn.gaussian_mixture_embedding(160, dim=8, centres=4, sep=9.0, seed=6)
    G = snn.to_networkx(snn.snn_graph(X, k=10))
    ctx = OracleBatchContext()
    labels, levels, _ = clustering.recursive_bipartition_batched(G, 0.01, k=1.0, model="cut_linear", terminate_on="min_size", size_limit=30,
                                                                 num_reads=16, num_sweeps=40, seed=2, context=ctx)
    assert set(labels) == set(G.nodes) and ctx.calls == len(levels)


def test_split_keeps_the_order_networkx_subgraphs_have():
    """qa_graph_split's contract (restated by host_split): a child's edge list equals graph_arrays(G.subgraph(part))."""
    gz = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "graphs.npz")
    for name in ("blobs", "noisy_moons"):
        labels = [str(x) for x in gz[f"{name}_labels"]]
        G = nx.Graph()
        G.add_nodes_from(labels)
        for u, v, w in zip(gz[f"{name}_eu"], gz[f"{name}_ev"], gz[f"{name}_w"]):
            G.add_edge(labels[u], labels[v], weight=float(w))
        lab, eu, ev, w = models.graph_arrays(G)
        rng = np.random.default_rng(3)
        part_of = rng.integers(-1, 3, size=len(lab)).astype(np.int32)
        for p, child in enumerate(host_split((len(lab), eu, ev, w), part_of, 3)):
            sub = G.subgraph([lab[i] for i in np.flatnonzero(part_of == p)])
            sl, su, sv, sw = models.graph_arrays(sub)
            assert sl == [lab[i] for i in np.flatnonzero(part_of == p)]
            assert np.array_equal(su, child[1]) and np.array_equal(sv, child[2]) and np.array_equal(sw, child[3])


# ---- the reference's termination rules, restated literally (the default `conf` path must match) ----------
class StubSampler:
    """Returns a crafted energy-sorted SampleSet: the first half of the nodes in S1, energies as given per call."""

    def __init__(self, energies_per_call):
        self.energies = list(energies_per_call)
        self.calls = []

    def sample(self, model, **kw):
        from scrna_seq_qannealing_clustering_b200.sampleset import SampleSet
        labels = list(model.labels)
        n = len(labels)
        self.calls.append(labels)
        e = self.energies.pop(0) if self.energies else [-1.0, -1.0, -1.0, -1.0]
        rows = np.zeros((len(e), n), dtype=np.int8)
        rows[:, : n // 2] = 1
        return SampleSet.from_samples((rows, labels), energy=np.asarray(e, dtype=float), vartype="BINARY")


def _path_graph(n):
    import networkx as nx
    G = nx.Graph()
    G.add_nodes_from(str(i) for i in range(n))
    for i in range(n - 1):
        G.add_edge(str(i), str(i + 1), weight=0.5)
    return G


def test_bqm_rule_matches_the_reference_branches():
    r = clustering.bqm_rule
    # conf, > 3 energies: ratio e0/e3 > 1.5, both halves > 5, iteration < iter_limit -> recurse; label ends as one colour
    assert r("conf", 10, 10, [-30.0, -29.0, -28.0, -15.0], 0, 40, 2) == (True, "single")
    assert r("conf", 10, 10, [-30.0, -29.0, -28.0, -25.0], 0, 40, 2) == (False, "single")      # ratio 1.2
    assert r("conf", 10, 5, [-30.0, -29.0, -28.0, -15.0], 0, 40, 2) == (False, "single")       # min(len) > 5 fails
    assert r("conf", 10, 10, [-30.0, -29.0, -28.0, -15.0], 2, 40, 2) == (False, "single")      # iteration limit
    assert r("conf", 10, 10, [-30.0, -29.0, -28.0, 0.05], 0, 40, 2) == (False, "single")       # |e3| <= 0.1: early return
    assert r("conf", 10, 10, [3.0, 2.0, 1.5, 1.0], 0, 40, 2) == (True, "single")               # positive energies: ratio 3
    # conf, <= 3 energies: size rule alone
    assert r("conf", 10, 10, [-3.0, -2.0], 0, 40, 2) == (True, "split")
    assert r("conf", 10, 4, [-3.0, -2.0], 0, 40, 2) == (False, "single")
    # min_size writes nothing when it stops; iter_limit ignores the sizes
    assert r("min_size", 41, 41, [], 0, 40, 2) == (True, "split") and r("min_size", 41, 40, [], 0, 40, 2) == (False, None)
    assert r("iter_limit", 1, 0, [], 1, 40, 2) == (True, "split") and r("iter_limit", 50, 50, [], 2, 40, 2) == (False, None)
    assert r("once", 50, 50, [], 0, 40, 2) == (False, "split")
    r2 = clustering.bqm2_rule
    assert r2("min_size", 41, 41, [], 40) == (True, "color") and r2("min_size", 41, 40, [], 40) == (False, "color")
    assert r2("conf", 10, 10, [-30.0, -29.0, -28.0, -15.0], 40) == (True, "split")
    assert r2("conf", 10, 10, [-30.0, -29.0, -28.0, -25.0], 40) == (False, None)              # difference <= 10


def test_conf_recursion_tree_and_labels_follow_the_reference():
    G = _path_graph(48)
    # level 0: recurse (ratio 2); the two children: one stops on the ratio, one on |e3| <= 0.1
    stub = StubSampler([[-30.0, -29.0, -28.0, -15.0], [-8.0, -7.5, -7.2, -7.0], [-8.0, -7.5, -7.2, 0.0]])
    clustering.clustering_bqm(G, 0, {"name": "t"}, "stub", 0.05, terminate_on="conf", size_limit=40, iter_limit=2, sampler=stub)
    assert [len(c) for c in stub.calls] == [48, 24, 24]
    assert len({G.nodes[n]["label0"] for n in G.nodes}) == 1            # overwritten with one colour after the recursion
    assert all("label1" in G.nodes[n] for n in G.nodes) and all("label2" not in G.nodes[n] for n in G.nodes)
    # the default size_limit plays no role in `conf`: halves of 24 > 5 recursed although size_limit = 40


def test_min_size_and_bqm2_recursion():
    G = _path_graph(40)
    stub = StubSampler([])
    clustering.clustering_bqm(G, 0, {"name": "t"}, "stub", 0.05, terminate_on="min_size", size_limit=9, iter_limit=5, sampler=stub)
    assert [len(c) for c in stub.calls] == [40, 20, 10, 10, 20, 10, 10]       # depth-first: halves of 5 are not > 9... 10 > 9 recursed
    G = _path_graph(40)
    stub = StubSampler([])
    clustering.clustering_bqm_2(G, 0, {"name": "t"}, "stub", 0.01, color=0, terminate_on="min_size", size_limit=9, k=1, sampler=stub)
    assert [len(c) for c in stub.calls] == [40, 20, 10, 10, 20, 10, 10]
    assert {G.nodes[n]["label0"] for n in G.nodes} == {100, -100}             # 100 - color / color - 100
    assert {G.nodes[n]["label1"] for n in G.nodes} == {80, -80}
    import pytest
    with pytest.raises(NotImplementedError):
        clustering.clustering_bqm_2(G, 0, {"name": "t"}, "stub", 0.01, terminate_on="iter_limit", sampler=stub)
