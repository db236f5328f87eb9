"""Parity at BASELINE.json's FULL model sizes ("harden parity at full size"): the models are as large as the
configs say, only reads / sweeps are cut so that the CPU oracle finishes in seconds.  Same bar as everywhere: final states
bytewise, energies bitwise, event counters equal.

    config 3  8-way CQM on 16 384 cells (131 184 variables) on the replay kernel -- the benched launch shape's model
    config 1  2-way cut+balance on 512 cells, MATERIALISED K_512 (130 816 couplers), 1000 reads
    config 5  dense Gaussian-affinity 4-way model on 1024 cells (4096 variables, 2.1 M couplers)
    more than one wave of CTAs on the replay kernel; an interrupt callback on the replay kernel
"""
import networkx as nx
import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel

from test_gpu_replay import _replay_ctx

pytestmark = pytest.mark.gpu


def _parity(ctx, model, R, sweeps, seed, beta_range, nthreads=0, interrupt=None):
    n = model.num_variables
    groups = model.groups.astuple() if model.groups is not None else None
    betas, spb = schedule.make_beta_schedule(beta_range, sweeps, 1, "geometric")
    seeds = schedule.per_read_seeds(seed, R)
    init = schedule.random_spin_states(R, n, seed)
    ref = init.copy()
    ref_e, ref_st = oracle.sample_ising(model.h, model.starts, model.ends, model.weights, ref, betas, spb, seeds, groups=groups,
                                        nthreads=nthreads)
    gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
    if groups is not None:
        gm.set_groups(*groups)
    states = init.copy()
    e, st, done = gm.sample(states, betas, spb, seeds, interrupt_function=interrupt)
    gm.close()
    if interrupt is None:
        assert done == R
        for key in ("attempts", "candidates", "draws", "accepted", "nbr_updates"):
            assert getattr(st, key) == ref_st[key], key
    bad = np.nonzero((states[:done] != ref[:done]).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {done} reads differ, first {bad[:8]}"
    assert np.array_equal(e[:done].view(np.uint64), ref_e[:done].view(np.uint64))
    assert st.near_ties == 0
    return done, init, states


def test_config3_full_size_model_on_the_replay_kernel(built):
    g = snn.synthetic_snn(16384, k=5, dim=15, centres=8, seed=0)[0]
    m = models.cqm_model(g, 8, min_size=20)
    assert m.num_variables == 16384 * 8 + 8 * 14
    br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
    for permille, sweeps in ((0, 5), (1000, 4)):      # all sweeps replayed / one replay sweep, catch-up, then pushes
        ctx = _replay_ctx(permille=permille)
        try:
            _parity(ctx, m, 64, sweeps, 1234, br, nthreads=8)
            assert ctx.last_kernel == _lib.QA_KERNEL_REPLAY
        finally:
            ctx.close()


def test_config1_materialised_k512_1000_reads(built):
    g = snn.synthetic_snn(512, k=5, seed=0)[0]
    m = models.cut_balance_model(g, 0.05, k=8.0, structured=False)
    assert m.num_couplers == 512 * 511 // 2
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    with Context(0) as ctx:
        _parity(ctx, m, 1000, 30, 1234, (hot, 10.0), nthreads=8)


def test_config5_dense_affinity_1024_cells(built):
    X, _ = snn.gaussian_mixture_embedding(1024, dim=15, centres=4, sep=4.0, seed=2)
    A = snn.gaussian_affinity(X, k=10)
    iu, ju = np.tril_indices(len(X), -1)
    G = nx.Graph()
    G.add_nodes_from(str(i) for i in range(len(X)))
    G.add_weighted_edges_from((str(j), str(i), float(A[i, j])) for i, j in zip(iu, ju))
    m = models.dqm_model(G, 4, 0.05, semantics="intended")
    assert m.num_variables == 4096
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    with Context(0) as ctx:
        _parity(ctx, m, 128, 10, 3, (hot, 20.0), nthreads=8)


def test_more_than_one_wave_of_ctas_on_the_replay_kernel(built):
    """One-warp CTAs: 148 SMs hold at most a few hundred of them, 20 000 reads are 625 -- later CTAs reuse scratch slots."""
    g = snn.synthetic_snn(256, k=5, seed=3)[0]
    m = models.subsampling_model(g, 7.0)
    ctx = _replay_ctx(warps=1)
    try:
        _parity(ctx, m, 20000, 20, 9, (0.05, 6.0), nthreads=8)
        assert ctx.last_kernel == _lib.QA_KERNEL_REPLAY
    finally:
        ctx.close()


def test_interrupt_function_stops_the_replay_kernel_between_read_groups(built):
    """neal polls interrupt_function between reads; the replay kernel polls a host-mapped flag whenever a CTA pulls its next
    group of reads.  The reads handed out before the stop are complete and equal to the oracle's; the rest keep their
    initial states."""
    g = snn.synthetic_snn(256, k=5, seed=3)[0]
    m = models.subsampling_model(g, 7.0)
    ctx = _replay_ctx(warps=1)
    try:
        calls = []

        def stop_at_once():
            calls.append(1)
            return True

        R = 60000
        done, init, states = _parity(ctx, m, R, 40, 11, (0.05, 6.0), nthreads=8, interrupt=stop_at_once)
        assert ctx.last_kernel == _lib.QA_KERNEL_REPLAY and calls
        assert 0 < done <= R and done % 32 == 0 or done == R
        assert np.array_equal(states[done:], init[done:])        # untouched
        # a callback that never fires: the whole job completes
        done2, _, _ = _parity(ctx, m, 2000, 10, 12, (0.05, 6.0), nthreads=8, interrupt=lambda: False)
        assert done2 == 2000
    finally:
        ctx.close()
