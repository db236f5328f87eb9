"""BASELINE.json's configurations at their own model sizes (reads / sweeps reduced so that the CPU oracle finishes in seconds),
through the automatic kernel selection: states bytewise, energies bitwise against the oracle.

config 1: 2-way cut+balance partition of a 512-cell SNN graph, 1000 reads          (clustering_bqm, BQM_clustering.py:25)
config 2: 4-way DQM on a 2048-cell SNN graph = 8192 binary variables, 10 000 reads (clustering_dqm, DQM_clustering.py:24)
config 3: 8-way CQM on 16 384 cells -- full size is checked by bench.py's cpu_baseline parity_check; here 1024 cells
config 4: QA_subsampling batch of independent 1000-cell sub-graph QUBOs, 100 reads each, one launch
config 5: dense Gaussian-affinity 4-way model (no tensor-core path yet: the dense rows run on the eager lockstep kernel)
"""
import networkx as nx
import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = Context(0)   # QA_KERNEL_AUTO
    yield c
    c.close()


def _parity(ctx, model, R, sweeps, seed, beta_range):
    n = model.num_variables
    groups = model.groups.astuple() if model.groups is not None else None
    betas, spb = schedule.make_beta_schedule(beta_range, sweeps, 1, "geometric")
    seeds = schedule.per_read_seeds(seed, R)
    init = schedule.random_spin_states(R, n, seed)
    ref = init.copy()
    ref_e, ref_st = oracle.sample_ising(model.h, model.starts, model.ends, model.weights, ref, betas, spb, seeds, groups=groups)
    gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
    if groups is not None:
        gm.set_groups(*groups)
    states = init.copy()
    e, st, done = gm.sample(states, betas, spb, seeds)
    gm.close()
    assert done == R
    bad = np.nonzero((states != ref).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {R} reads differ, first {bad[:8]}"
    assert np.array_equal(e.view(np.uint64), ref_e.view(np.uint64))
    for key in ("attempts", "candidates", "draws", "accepted", "nbr_updates"):
        assert getattr(st, key) == ref_st[key], key
    assert st.near_ties == 0
    return ctx.last_kernel


def test_config1_two_way_partition_512_cells(ctx):
    g = snn.synthetic_snn(512, k=5, seed=0)[0]
    m = models.cut_balance_model(g, 0.05, k=8.0, structured=True)
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    assert _parity(ctx, m, 1000, 100, 1234, (hot, 10.0)) == _lib.QA_KERNEL_WARP_PER_READ


def test_config2_four_way_dqm_2048_cells(ctx):
    g = snn.synthetic_snn(2048, k=5, seed=0)[0]
    m = models.dqm_model(g, 4, 0.005, semantics="intended")
    assert m.num_variables == 8192
    br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
    assert _parity(ctx, m, 10000, 24, 7, br) == _lib.QA_KERNEL_REPLAY


def test_config3_eight_way_cqm_1024_cells(ctx):
    g = snn.synthetic_snn(1024, k=5, seed=0)[0]
    m = models.cqm_model(g, 8, min_size=20)
    br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
    assert _parity(ctx, m, 6400, 30, 11, br) == _lib.QA_KERNEL_REPLAY


def test_config4_subsampling_batch_of_1000_cell_problems(ctx):
    P, cells, rpp = 64, 1000, 100
    graphs = snn.subsample_problems(P * cells, P, cells, k=10, dim=30, seed=0)
    ms = [models.subsampling_model(g, 7.0) for g in graphs]
    br = schedule.default_ising_beta_range(ms[0].h, ms[0].starts, ms[0].ends, ms[0].weights)
    betas, spb = schedule.make_beta_schedule(br, 40, 1, "geometric")
    voff = np.cumsum([0] + [m.num_variables for m in ms])
    coff = np.cumsum([0] + [m.num_couplers for m in ms])
    seeds = schedule.per_read_seeds(5, rpp * P)
    inits = [schedule.random_spin_states(rpp, m.num_variables, 300 + i) for i, m in enumerate(ms)]
    states = np.concatenate([s.ravel() for s in inits]).copy()
    e, st, done = ctx.sample_ising_batch(voff, coff, np.concatenate([m.h for m in ms]), np.concatenate([m.starts for m in ms]),
                                         np.concatenate([m.ends for m in ms]), np.concatenate([m.weights for m in ms]), rpp, states,
                                         betas, spb, seeds)
    assert done == rpp and ctx.last_kernel == _lib.QA_KERNEL_REPLAY
    off = 0
    for i, m in enumerate(ms):
        ref = inits[i].copy()
        ref_e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, ref, betas, spb, seeds[i * rpp:(i + 1) * rpp])
        got = states[off:off + rpp * m.num_variables].reshape(rpp, m.num_variables)
        off += rpp * m.num_variables
        assert np.array_equal(got, ref), i
        assert np.array_equal(e[i * rpp:(i + 1) * rpp].view(np.uint64), ref_e.view(np.uint64)), i


def test_config5_dense_gaussian_affinity_four_way(ctx):
    """Dense affinity model: every same-case pair of cells is coupled (rows of degree ~n), so the slab format does not apply
    and the eager lockstep kernel runs it, bit-exact.  (The fp64 tensor-core path of config 5 is not built yet.)"""
    X, _ = snn.gaussian_mixture_embedding(160, dim=15, centres=4, sep=4.0, seed=2)
    A = snn.gaussian_affinity(X, k=10)
    G = nx.Graph()
    G.add_nodes_from(str(i) for i in range(len(X)))
    for i in range(len(X)):
        for j in range(i):
            G.add_edge(str(j), str(i), weight=float(A[i, j]))
    m = models.dqm_model(G, 4, 0.05, semantics="intended")
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    k = _parity(ctx, m, 6400, 20, 3, (hot, 20.0))
    assert k == _lib.QA_KERNEL_WARP_PER_READ or k == _lib.QA_KERNEL_LOCKSTEP_PUSH
