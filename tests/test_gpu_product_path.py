"""The device work on the PRODUCT path: models built by the qa_build_* kernels when a clustering
function is handed a graph, the state matrix resident on the GPU from creation (counter-based generator) to the top-k export,
duplicate aggregation and ranking on the device."""
import numpy as np
import pytest

import scrna_seq_qannealing_clustering_b200 as qa
from scrna_seq_qannealing_clustering_b200 import clustering, models, schedule, snn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sampler(built):
    s = qa.B200SimulatedAnnealingSampler(device=0)
    yield s
    s.close()


def test_counter_generator_is_the_same_function_on_host_and_device(gpu_ctx):
    for (R, n, seed, first) in ((5, 1, 0, 0), (33, 64, 7, 0), (70, 131, 2 ** 31 + 5, 12345), (3, 1000, 9, 2 ** 33)):
        dev = gpu_ctx.random_states(seed, first, R, n)
        got = dev.download()
        dev.close()
        want = schedule.counter_spin_states(R, n, seed, first_read=first)
        assert got.dtype == np.int8 and np.array_equal(got, want)
        assert set(np.unique(got)) <= {-1, 1}
    # a function of the GLOBAL read index: shards concatenate to the unsharded matrix
    whole = schedule.counter_spin_states(20, 77, 3)
    assert np.array_equal(np.vstack([schedule.counter_spin_states(8, 77, 3), schedule.counter_spin_states(12, 77, 3, first_read=8)]), whole)
    assert abs(float(schedule.counter_spin_states(64, 4096, 1).mean())) < 0.01


def test_best_k_keeps_the_state_matrix_on_the_device(sampler):
    g = snn.synthetic_snn(256, k=5, seed=3)[0]
    model = models.cqm_model(g, 4, min_size=20)
    kw = dict(num_reads=300, num_sweeps=120, beta_range=(0.02, 6.0), seed=5, initial_states_generator="counter")
    full = sampler.sample(model, **kw)
    best = sampler.sample(model, return_samples="best_k", num_best=7, **kw)
    order = np.argsort(full.record.energy, kind="stable")
    assert len(best) == 7
    assert np.array_equal(best.record.energy, full.record.energy[order[:7]])          # same reads, same energies, bitwise
    assert np.array_equal(best.record.sample, full.record.sample[order[:7]])
    assert np.array_equal(best.info["best_read_index"], order[:7])
    assert np.array_equal(best.info["energies"], full.record.energy)                  # all R energies still come back
    assert best.first.energy == full.first.energy and best.info["b200"]["return_samples"] == "best_k"
    # given initial states take the same route (uploaded once, ranked on the device)
    init = schedule.random_spin_states(64, model.num_variables, 11)
    a = sampler.sample(model, num_reads=64, num_sweeps=60, beta_range=(0.02, 6.0), seed=5, initial_states=(init, model.labels))
    b = sampler.sample(model, num_reads=64, num_sweeps=60, beta_range=(0.02, 6.0), seed=5, initial_states=(init, model.labels),
                       return_samples="best_k", num_best=3)
    assert np.array_equal(b.record.sample, a.record.sample[np.argsort(a.record.energy, kind="stable")[:3]])


def test_duplicate_aggregation_on_the_device(sampler, gpu_ctx):
    rng = np.random.default_rng(0)
    base = (rng.integers(0, 2, size=(9, 300), dtype=np.int8) * 2 - 1).astype(np.int8)
    pick = rng.integers(0, 9, size=500)
    states = base[pick]
    first, count = gpu_ctx.aggregate_reads(states)
    _, idx, cnt = np.unique(pick, return_index=True, return_counts=True)
    order = np.argsort(idx)
    assert np.array_equal(first, idx[order]) and np.array_equal(count, cnt[order])     # order of first occurrence, like dimod
    # through the sampler: a cold anneal of a tiny model returns few distinct samples
    g = snn.synthetic_snn(40, k=4, seed=1)[0]
    m = models.subsampling_model(g, 7.0)
    kw = dict(num_reads=400, num_sweeps=80, beta_range=(0.05, 20.0), seed=3)
    host = sampler.sample(m, aggregate=True, sorted=True, **kw)
    dev = sampler.sample(m, aggregate=True, return_samples="best_k", num_best=10 ** 6, **kw)
    assert len(dev) == len(host) and int(dev.record.num_occurrences.sum()) == 400
    assert np.array_equal(np.sort(dev.record.energy), np.sort(host.record.energy))
    key = lambda ss: sorted((row.tobytes(), int(c)) for row, c in zip(ss.record.sample, ss.record.num_occurrences))  # noqa: E731
    assert key(dev) == key(host)


@pytest.mark.parametrize("fn", ["cqm", "dqm", "bqm", "subsampling"])
def test_clustering_functions_build_their_models_on_the_device(sampler, fn):
    """Same SampleSet whether the model vectors come from the host builders (models.py) or from the qa_build_* kernels."""
    G = snn.to_networkx(snn.synthetic_snn(200, k=5, seed=2)[0])
    kw = dict(num_reads=96, num_sweeps=80, beta_range=(0.02, 6.0), seed=4, sampler=sampler)
    if fn == "cqm":
        a = clustering.clustering_cqm(G, 4, device_build=False, **kw)
        b = clustering.clustering_cqm(G, 4, device_build=True, **kw)
        assert np.array_equal(a.record.is_feasible, b.record.is_feasible)
    elif fn == "dqm":
        a = clustering.clustering_dqm(G, 3, 0.005, semantics="intended", device_build=False, **kw)
        b = clustering.clustering_dqm(G, 3, 0.005, semantics="intended", device_build=True, **kw)
    elif fn == "bqm":
        a = clustering.clustering_bqm(G.copy(), 0, {"name": "t"}, "b200", 0.05, device_build=False, **kw)
        b = clustering.clustering_bqm(G.copy(), 0, {"name": "t"}, "b200", 0.05, device_build=True, **kw)
    else:
        a = clustering.graph_subsampling(G.copy(), 7, device_build=False, **kw)
        b = clustering.graph_subsampling(G.copy(), 7, device_build=True, **kw)
    assert list(a.variables) == list(b.variables)
    assert np.array_equal(a.record.sample, b.record.sample)
    assert np.allclose(a.record.energy, b.record.energy, rtol=1e-12, atol=1e-9)
