"""Device model builders (qa_build_*) against the host builders (models.py), which are themselves pinned bit for bit
against the reference's own functions (tests/test_golden_models.py).  Bar: h, coupler indices and weights bit-identical;
offsets to 1e-12 relative; and the built model anneals to the same states as the host-built one."""
from pathlib import Path

import numpy as np
import pytest

from scrna_seq_qannealing_clustering_b200 import models, schedule, snn
from scrna_seq_qannealing_clustering_b200.engine import IsingModel

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def fixture_graph(name):
    g = np.load(GOLD / "graphs.npz")
    return (256, g[f"{name}_eu"].astype(np.int64), g[f"{name}_ev"].astype(np.int64), g[f"{name}_w"])


GRAPHS = {"noisy_moons": lambda: fixture_graph("noisy_moons"), "blobs": lambda: fixture_graph("blobs"),
          "synthetic_700": lambda: snn.synthetic_snn(700, k=5, seed=4)[0]}


def assert_same_vectors(gm, model, offset):
    h, s, e, w = gm.get_ising()
    assert np.array_equal(s, model.starts) and np.array_equal(e, model.ends)
    assert np.array_equal(w.view(np.uint64), model.weights.view(np.uint64))
    assert np.array_equal(h.view(np.uint64), model.h.view(np.uint64))
    assert offset == pytest.approx(model.offset, rel=1e-12, abs=1e-9)


def assert_same_anneal(ctx, gm, model, seed=3):
    betas, spb = schedule.make_beta_schedule((0.02, 6.0), 40, 1, "geometric")
    seeds = schedule.per_read_seeds(seed, 64)
    init = schedule.random_spin_states(64, model.num_variables, seed)
    a = init.copy()
    ea, _, _ = gm.sample(a, betas, spb, seeds)
    ref = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
    if model.groups is not None:
        ref.set_groups(*model.groups.astuple())
    b = init.copy()
    eb, _, _ = ref.sample(b, betas, spb, seeds)
    ref.close()
    assert np.array_equal(a, b) and np.array_equal(ea.view(np.uint64), eb.view(np.uint64))


@pytest.mark.parametrize("name", list(GRAPHS))
def test_cut_balance(gpu_ctx, name):
    g = GRAPHS[name]()
    model = models.cut_balance_model(g, 0.05, structured=True)
    gm, off, gamma = gpu_ctx.build_cut_balance(g, 0.05, 8.0)
    assert gamma == model.meta["gamma"]
    assert_same_vectors(gm, model, off)
    assert_same_anneal(gpu_ctx, gm, model)
    gm.close()


@pytest.mark.parametrize("name", list(GRAPHS))
def test_cut_linear(gpu_ctx, name):
    """clustering_bqm_2's sparse model (BQM_clustering.py:210-236), main.py:154's arguments k = 1, gamma_factor = 0.01."""
    g = GRAPHS[name]()
    model = models.cut_linear_model(g, 0.01, 1.0)
    gm, off, gamma = gpu_ctx.build_cut_linear(g, 0.01, 1.0)
    assert gamma == model.meta["gamma"]
    assert_same_vectors(gm, model, off)
    assert_same_anneal(gpu_ctx, gm, model)
    gm.close()


@pytest.mark.parametrize("name", list(GRAPHS))
def test_subsampling(gpu_ctx, name):
    g = GRAPHS[name]()
    model = models.subsampling_model(g, 7.0)
    gm, off = gpu_ctx.build_subsampling(g, 7.0)
    assert_same_vectors(gm, model, off)
    gm.close()


@pytest.mark.parametrize("name", list(GRAPHS))
@pytest.mark.parametrize("semantics", ["as_written", "intended"])
def test_dqm(gpu_ctx, name, semantics):
    g = GRAPHS[name]()
    model = models.dqm_model(g, 4, 0.005, penalty=33.5, semantics=semantics)
    gm, off = gpu_ctx.build_dqm_onehot(g, 4, 0.005, 33.5, semantics)
    assert_same_vectors(gm, model, off)
    assert_same_anneal(gpu_ctx, gm, model)
    gm.close()


@pytest.mark.parametrize("name", list(GRAPHS))
def test_cqm(gpu_ctx, name):
    g = GRAPHS[name]()
    model = models.cqm_model(g, 5, min_size=20, onehot_penalty=41.0, size_penalty=1.5)
    gm, off = gpu_ctx.build_cqm_penalty(g, 5, 20, 41.0, 1.5)
    assert gm.num_variables == model.num_variables
    assert_same_vectors(gm, model, off)
    assert_same_anneal(gpu_ctx, gm, model)
    gm.close()


def test_builder_rejects_bad_edges(gpu_ctx):
    from scrna_seq_qannealing_clustering_b200 import _lib
    with pytest.raises(_lib.QAnnealError) as ei:
        gpu_ctx.build_subsampling((4, np.array([0, 9]), np.array([1, 2]), np.array([0.5, 0.5])), 7.0)
    assert ei.value.code == -2
