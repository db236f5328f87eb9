"""SNN graph construction on the device (csrc/snn.cu, qa_snn_build) against the host specification
snn.py and against the reference's own fixture graphs (R/benchmarks/graph_*.gexf, regenerated from the sklearn datasets of
Benchmark.Rmd:33-55): identical edge sets, identical fp64 weights."""
import time

import numpy as np
import pytest

from scrna_seq_qannealing_clustering_b200 import models, snn
from scrna_seq_qannealing_clustering_b200.engine import Context
from test_snn_cpu import as_dict, benchmark_datasets, fixture_edges

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = Context(0)
    yield c
    c.close()


def assert_same_graph(got, want):
    assert got[0] == want[0]
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert np.array_equal(got[3].view(np.uint64), np.asarray(want[3], dtype=np.float64).view(np.uint64))


@pytest.mark.parametrize("name", ["noisy_circles", "aniso", "no_structure"])
def test_device_snn_reproduces_the_reference_fixture(ctx, name):
    X = np.asarray(benchmark_datasets()[name], dtype=np.float64)
    g = ctx.build_snn(X, k=10, prune=0.0, max_degree=None)
    try:
        assert as_dict(g.graph(0)) == fixture_edges(name)
    finally:
        g.close()


@pytest.mark.parametrize("n,k,dim,trim", [(512, 5, 15, 15), (2048, 5, 15, 15), (1000, 10, 30, 15), (700, 8, 10, 9), (300, 20, 5, None),
                                          (7, 10, 3, 15)])
def test_device_snn_equals_host_snn(ctx, n, k, dim, trim):
    X, _ = snn.gaussian_mixture_embedding(n, dim=dim, centres=6, seed=n)
    want = snn.snn_graph(X, k, 1.0 / 15.0, trim)
    g = ctx.build_snn(X, k=k, prune=1.0 / 15.0, max_degree=trim)
    try:
        assert_same_graph(g.graph(0), want)
    finally:
        g.close()


def test_config3_graph_16384_cells_under_half_a_second(ctx):
    X, _ = snn.gaussian_mixture_embedding(16384, dim=15, centres=8, seed=0)
    ctx.build_snn(X[:512], k=5).close()          # warm-up (module load, scratch)
    ctx.synchronize()
    t0 = time.perf_counter()
    g = ctx.build_snn(X, k=5, prune=1.0 / 15.0, max_degree=15)
    dt = time.perf_counter() - t0
    try:
        assert_same_graph(g.graph(0), snn.snn_graph(X, 5, 1.0 / 15.0, 15))
    finally:
        g.close()
    assert dt < 0.5, dt


def test_batched_point_sets_build_their_own_graphs(ctx):
    """Config 4's inputs: disjoint subsets of one embedding, each with its own SNN graph, one call."""
    X, _ = snn.gaussian_mixture_embedding(6000, dim=30, centres=8, seed=0)
    rng = np.random.default_rng(1)
    perm = rng.permutation(6000)
    sizes = [1000, 1000, 777, 1000, 223, 1000, 1000]
    off = np.cumsum([0] + sizes)
    idx = [np.sort(perm[off[p]:off[p + 1]]) for p in range(len(sizes))]
    g = ctx.build_snn(np.concatenate([X[i] for i in idx]), k=10, prune=1.0 / 15.0, max_degree=15, offsets=off)
    try:
        total = 0
        for p, i in enumerate(idx):
            want = snn.snn_graph(X[i], 10, 1.0 / 15.0, 15)
            assert_same_graph(g.graph(p), want)
            total += len(want[1])
        assert g.num_edges() == total
    finally:
        g.close()


def test_device_graph_feeds_the_model_builders_without_leaving_the_gpu(ctx):
    X, _ = snn.gaussian_mixture_embedding(1024, dim=15, centres=8, seed=3)
    host = snn.snn_graph(X, 5, 1.0 / 15.0, 15)
    ref = models.cqm_model(host, 8, min_size=20)
    g = ctx.build_snn(X, k=5)
    try:
        gm, off = ctx.build_cqm_penalty(g.device_graph(0), 8, 20, ref.meta["onehot_penalty"], ref.meta["size_penalty"])
        h, s, e, w = gm.get_ising()
        gm.close()
    finally:
        g.close()
    assert np.array_equal(h.view(np.uint64), ref.h.view(np.uint64))
    assert np.array_equal(s, ref.starts) and np.array_equal(e, ref.ends)
    assert np.array_equal(w.view(np.uint64), ref.weights.view(np.uint64))
    assert off == ref.offset
