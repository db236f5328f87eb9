"""Recursion driver on the device (csrc/recursion.cu; reference: clustering_bqm calling itself on
G.subgraph(S0) / G.subgraph(S1), Python_Functions/BQM_clustering.py:113-203): qa_graph_split, qa_model_concat,
qa_sa_sample_model_batch, and the level-batched driver against the call-by-call recursion through the dimod-shaped sampler."""
import json
from pathlib import Path

import networkx as nx
import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import _lib, clustering, models, schedule, snn
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel
from scrna_seq_qannealing_clustering_b200.sampler import B200SimulatedAnnealingSampler
from test_recursion_cpu import host_split

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def ctx(built):
    c = Context(0)
    yield c
    c.close()


def load_graph(name):
    g = np.load(GOLD / "graphs.npz")
    labels = [str(x) for x in g[f"{name}_labels"]]
    G = nx.Graph()
    G.add_nodes_from(labels)
    for u, v, w in zip(g[f"{name}_eu"], g[f"{name}_ev"], g[f"{name}_w"]):
        G.add_edge(labels[u], labels[v], weight=float(w))
    return G


def test_graph_split_equals_networkx_subgraphs(ctx):
    G = load_graph("blobs")
    lab, eu, ev, w = models.graph_arrays(G)
    rng = np.random.default_rng(0)
    part_of = rng.integers(-1, 5, size=len(lab)).astype(np.int32)
    part_of[part_of == 3] = 0                      # an empty part in the middle
    root = (len(lab), eu.astype(np.int32), ev.astype(np.int32), w)
    dg = ctx.split_graph(root, part_of, 5)
    try:
        want = host_split((len(lab), eu, ev, w), part_of, 5)
        for p in range(5):
            n, gu, gv, gw = dg.graph(p)
            assert n == want[p][0]
            assert np.array_equal(gu, want[p][1]) and np.array_equal(gv, want[p][2]) and np.array_equal(gw, want[p][3])
            assert np.array_equal(dg.nodes(p), np.flatnonzero(part_of == p))
            sub = G.subgraph([lab[i] for i in np.flatnonzero(part_of == p)])
            sl, su, sv, sw = models.graph_arrays(sub)
            assert np.array_equal(su, gu) and np.array_equal(sv, gv) and np.array_equal(sw, gw)
        # a device graph can be split again (device pointers in, device pointers out)
        inner = np.arange(dg.num_nodes(0)) % 2
        dg2 = ctx.split_graph(dg.device_graph(0), inner.astype(np.int32), 2)
        want2 = host_split(want[0], inner, 2)
        for p in range(2):
            got = dg2.graph(p)
            assert got[0] == want2[p][0] and np.array_equal(got[1], want2[p][1]) and np.array_equal(got[3], want2[p][3])
        dg2.close()
    finally:
        dg.close()


def test_concatenated_structured_models_anneal_like_separate_calls(ctx):
    """qa_model_concat + qa_sa_sample_model_batch with one beta schedule per problem == one qa_sa_sample_model call per problem
    (rank-1 groups included): states bytewise, energies bitwise, against the oracle as well."""
    graphs = [snn.synthetic_snn(n, k=5, seed=s)[0] for n, s in ((150, 1), (97, 2), (260, 3))]
    R, sweeps = 40, 50
    host = [models.cut_balance_model(g, 0.05, k=8.0, structured=True) for g in graphs]
    gms = []
    for g in graphs:
        gm, off, gam = ctx.build_cut_balance((g[0], g[1].astype(np.int32), g[2].astype(np.int32), g[3]), 0.05, 8.0)
        gms.append(gm)
    betas = np.stack([schedule.make_beta_schedule(schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None), sweeps, 1,
                                                  "geometric")[0] for m in host])
    seeds = np.concatenate([schedule.per_read_seeds(11 + p, R) for p in range(3)])
    inits = [schedule.random_spin_states(R, m.num_variables, 5 + p) for p, m in enumerate(host)]
    states = np.concatenate([s.ravel() for s in inits]).copy()
    bm = ctx.concat_models(gms)
    try:
        e, st, done = ctx.sample_model_batch(bm, R, states, betas, 1, seeds)
    finally:
        bm.close()
    assert done == R and ctx.last_kernel == _lib.QA_KERNEL_WARP_PER_READ
    off = 0
    for p, m in enumerate(host):
        n = m.num_variables
        ref = inits[p].copy()
        ref_e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, ref, betas[p].copy(), 1, seeds[p * R:(p + 1) * R],
                                       groups=m.groups.astuple())
        got = states[off:off + R * n].reshape(R, n)
        off += R * n
        assert np.array_equal(got, ref), p
        assert np.array_equal(e[p * R:(p + 1) * R].view(np.uint64), ref_e.view(np.uint64)), p
        single = inits[p].copy()
        se, _, _ = gms[p].sample(single, betas[p].copy(), 1, seeds[p * R:(p + 1) * R])
        assert np.array_equal(single, ref) and np.array_equal(se.view(np.uint64), ref_e.view(np.uint64))
    for gm in gms:
        gm.close()


class _Recording(B200SimulatedAnnealingSampler):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.calls = []

    def sample(self, model, **kw):
        ss = super().sample(model, **kw)
        first = ss.first.sample
        self.calls.append((frozenset(model.labels), frozenset(v for v in model.labels if not first[v]), float(ss.first.energy)))
        return ss


@pytest.mark.parametrize("rule,reads,kw,min_levels", [("min_size", 48, {"size_limit": 30, "iter_limit": 3}, 2),
                                                      ("conf", 48, {"iter_limit": 2}, 1),      # ratio branch: e0 / e3 decides
                                                      ("conf", 3, {"iter_limit": 2}, 3),       # <= 3 energies: size branch recurses
                                                      ("iter_limit", 16, {"iter_limit": 2}, 3), ("once", 16, {}, 1)])
def test_batched_recursion_reproduces_clustering_bqm_on_graph_blobs(ctx, rule, reads, kw, min_levels):
    """the level-batched device recursion anneals the same sub-graphs, finds the same halves and the same best
    energies as `clustering_bqm` calling the sampler once per sub-graph (both build on the device, both use per-call seeding),
    for every termination rule of the reference."""
    G = load_graph("blobs")
    rec = _Recording(context=ctx)
    clustering.clustering_bqm(G.copy(), 0, {"name": "blobs"}, "b200", 0.05, terminate_on=rule, sampler=rec, num_reads=reads,
                              num_sweeps=120, seed=21, **kw)
    labels, levels, energies = clustering.recursive_bipartition_batched(G, 0.05, terminate_on=rule, num_reads=reads, num_sweeps=120,
                                                                        seed=21, context=ctx, write_labels=True, **kw)
    want = {c[0]: c for c in rec.calls}
    got = {frozenset(part): e for lv, es in zip(levels, energies) for part, e in zip(lv, es)}
    assert set(want) == set(got)
    for nodes, e in got.items():
        assert e == want[nodes][2]                      # best energy of every sub-graph, bit for bit
    assert len(levels) >= min_levels and set(labels) == set(G.nodes)
    assert all("label0" in G.nodes[v] for v in G.nodes) or rule in ("min_size", "iter_limit")
