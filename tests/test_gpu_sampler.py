"""The dimod-facing sampler on the GPU: the call sites of the reference, end to end (Q dict / BQM / DQM / CQM in,
SampleSet out), checked against the oracle and the known answers."""
import json
import warnings
from pathlib import Path

import networkx as nx
import numpy as np
import pytest

import scrna_seq_qannealing_clustering_b200 as qa
from oracle import models_ref, oracle
from scrna_seq_qannealing_clustering_b200 import models, schedule, snn

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def load_graph(name):
    g = np.load(GOLD / "graphs.npz")
    labels = [str(x) for x in g[f"{name}_labels"]]
    G = nx.Graph()
    G.add_nodes_from(labels)
    for u, v, w in zip(g[f"{name}_eu"], g[f"{name}_ev"], g[f"{name}_w"]):
        G.add_edge(labels[u], labels[v], weight=float(w))
    return G


@pytest.fixture(scope="module")
def sampler(built):
    s = qa.B200SimulatedAnnealingSampler(device=0)
    yield s
    s.close()


def test_sample_qubo_like_the_reference_call_site(sampler):
    """BQM_clustering.py:57 -- sampler.sample_qubo(Q, label=...) on the reference's own Q (materialised, 32 896 entries)."""
    G = load_graph("noisy_circles")
    Q, gamma = models_ref.qubo_clustering_bqm(G, 0.05)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        response = sampler.sample_qubo(Q, label="noisy_circles_hybrid", chain_strength=20, num_reads=64, num_sweeps=300,
                                       beta_range=(0.01, 10.0), seed=7)
    assert sum("Ignoring unknown kwarg" in str(x.message) for x in w) == 2
    known = json.loads((GOLD / "known_answers.json").read_text())["noisy_circles"]
    assert response.first.energy == pytest.approx(known["lower_bound"], rel=1e-10)   # provable ground state reached
    lut = response.first.sample
    S0 = [node for node in G.nodes if not lut[node]]
    S1 = [node for node in G.nodes if lut[node]]
    assert sorted(map(len, (S0, S1))) == [128, 128]
    comps = list(nx.connected_components(G))
    assert set(S0) in [set(c) for c in comps]
    rows = list(response.data(fields=["sample", "energy", "num_occurrences"]))
    assert len(rows) == 64 and rows[0].energy <= rows[-1].energy and rows[0].num_occurrences == 1
    assert response.vartype.name == "BINARY" and set(np.unique(response.record.sample)) <= {0, 1}
    # every returned energy equals the QUBO energy of its sample (dimod bqm.energies) to 1e-12 relative
    bqm = qa.BinaryQuadraticModel.from_qubo(Q)
    e = bqm.energies((response.record.sample, response.variables))
    scale = sum(abs(v) for v in Q.values())
    assert np.allclose(e, response.record.energy, rtol=1e-12, atol=1e-12 * scale)
    assert set(response.info) >= {"beta_range", "beta_schedule_type", "timing"}


def test_sample_matches_oracle_bit_for_bit_through_the_public_api(sampler):
    g = snn.synthetic_snn(300, k=5, seed=9)[0]
    model = models.subsampling_model(g, 0.8)
    ss = sampler.sample(model, num_reads=50, num_sweeps=80, beta_range=(0.05, 8.0), seed=11)
    betas, spb = schedule.make_beta_schedule((0.05, 8.0), 80, 1, "geometric")
    ref = schedule.random_spin_states(50, model.num_variables, 11)
    ref_e, _ = oracle.sample_ising(model.h, model.starts, model.ends, model.weights, ref, betas, spb, schedule.per_read_seeds(11, 50))
    assert np.array_equal(ss.record.sample, (ref + 1) // 2)
    assert np.array_equal(ss.record.energy, ref_e + model.offset)
    # neal's own seeding (one RNG stream over all reads) is available as seed_mode='stream'
    st = sampler.sample(model, num_reads=6, num_sweeps=40, beta_range=(0.05, 8.0), seed=5, seed_mode="stream")
    betas, spb = schedule.make_beta_schedule((0.05, 8.0), 40, 1, "geometric")
    ref = schedule.random_spin_states(6, model.num_variables, 5)
    oracle.sample_ising(model.h, model.starts, model.ends, model.weights, ref, betas, spb, np.array([5], dtype=np.uint64), seed_mode=1)
    assert np.array_equal(st.record.sample, (ref + 1) // 2)


def test_spin_bqm_initial_states_and_interrupt(sampler):
    bqm = qa.BinaryQuadraticModel.from_ising({"a": 0.5, "b": -0.25, "c": 0.0}, {("a", "b"): -1.0, ("b", "c"): 0.75})
    init = (np.array([[1, -1, 1], [-1, -1, -1]], dtype=np.int8), ["a", "b", "c"])
    ss = sampler.sample(bqm, initial_states=init, num_sweeps=0, beta_schedule_type="custom", beta_schedule=[])
    assert ss.vartype.name == "SPIN" and ss.record.sample.tolist() == [[1, -1, 1], [-1, -1, -1]]
    assert ss.record.energy.tolist() == pytest.approx(bqm.energies(init).tolist())
    calls = []
    ss = sampler.sample(bqm, num_reads=10, num_sweeps=20, interrupt_function=lambda: calls.append(1) or False, seed=1)
    assert len(ss) == 10 and ss.first.energy == pytest.approx(bqm.energies(ss.record.sample).min())
    with pytest.raises(ValueError):
        sampler.sample(bqm, num_reads=2, beta_schedule_type="bogus")


def test_clustering_functions(sampler):
    """The reference's clustering functions with the sampler injected (clustering.py)."""
    G = load_graph("blobs")                                    # three components: 86 + 85 + 85
    dirs = {"name": "blobs"}
    r = qa.clustering_bqm(G, 1, dirs, "hybrid", 0.05, terminate_on="once", sampler=sampler, num_reads=64, num_sweeps=300,
                          beta_range=(0.01, 10.0), seed=3)
    labels = nx.get_node_attributes(G, "label1")
    assert len(labels) == 256 and len(set(labels.values())) == 2
    known = json.loads((GOLD / "known_answers.json").read_text())["blobs"]
    assert r.first.energy <= known["E_largest_component"] + 1e-6     # at least as good as the known split
    assert np.all(np.diff(r.record.energy) >= 0)                      # sorted record, like a QPU answer

    G2 = load_graph("noisy_moons")
    r2 = qa.graph_subsampling(G2, 0.5, "hybrid", sampler=sampler, num_sweeps=200, beta_range=(0.05, 10.0), seed=4)
    assert len(r2) == 100 and set(nx.get_node_attributes(G2, "label1").values()) <= {0, 1}

    r3 = qa.clustering_bqm_3(G2, 1, dirs, "hybrid", 0.05, sampler=sampler, num_reads=20, num_sweeps=200, beta_range=(0.01, 10.0), seed=5)
    assert any(str(v).startswith("slack_c1_constraint_") for v in r3.variables)

    ss = qa.clustering_dqm(G2, 3, 0.005, sampler=sampler, num_reads=40, num_sweeps=300, beta_range=(0.02, 10.0), seed=6)
    assert ss.variables == list(G2.nodes) and set(np.unique(ss.record.sample)) <= {0, 1, 2}
    assert ss.record.is_feasible.mean() > 0.9
    best = ss.first
    assert len(set(best.sample.values())) >= 2

    cq = qa.clustering_cqm(G2, 3, sampler=sampler, num_reads=40, num_sweeps=400, beta_range=(0.02, 10.0), seed=7)
    assert cq.variables[0] == "v_0,0" and cq.record.cluster_sizes.shape == (40, 3)
    feas = cq.record.is_feasible
    assert feas.any()
    sizes = cq.record.cluster_sizes[feas]
    assert (sizes >= 20).all() and (sizes.sum(axis=1) == 256).all()


def test_clustering_cqm_defaults_give_feasible_reads(sampler):
    """The product's defaults (one-hot penalty A, size penalty B = A / n, slack bits initialised consistently with the cell bits)
    must leave SA-reachable feasible states: with round 1's B = 1 the binary slack froze and no read ended one-hot."""
    G = snn.to_networkx(snn.synthetic_snn(1024, k=5, seed=4)[0])
    cq = qa.clustering_cqm(G, 4, sampler=sampler, num_reads=64, num_sweeps=300, seed=7)
    assert cq.info["slack_init"] == "consistent"
    assert cq.info["size_penalty"] == pytest.approx(cq.info["onehot_penalty"] / 1024)
    assert cq.record.is_feasible.mean() >= 0.9
    sizes = cq.record.cluster_sizes[cq.record.is_feasible]
    assert (sizes >= 20).all() and (sizes.sum(axis=1) == 1024).all()
    hard = qa.clustering_cqm(G, 4, sampler=sampler, num_reads=64, num_sweeps=300, seed=7, size_penalty=1.0)
    assert hard.record.is_feasible.mean() <= cq.record.is_feasible.mean()


def test_generic_dqm_and_cqm_objects(sampler):
    d = qa.DiscreteQuadraticModel()
    for v in range(6):
        d.add_variable(3, label=v)
    for v in range(6):
        d.set_linear(v, [0.1 * v, 0.0, 0.2])
    for u in range(5):
        d.set_quadratic(u, u + 1, {(c, c): -1.0 for c in range(3)})
    ss = sampler.sample_dqm(d, num_reads=30, num_sweeps=200, beta_range=(0.1, 10.0), seed=2)
    assert ss.record.is_feasible.all()
    assert ss.first.energy == pytest.approx(d.energies([[1] * 6])[0])      # all in case 1: -5.0

    x = [qa.Binary(f"x{i}") for i in range(6)]
    c = qa.ConstrainedQuadraticModel()
    c.set_objective(sum(xi * (i + 1) for i, xi in enumerate(x)))
    c.add_constraint(sum(x) >= 3, label="atleast3")
    c.add_discrete(["x4", "x5"], label="d")
    ss = sampler.sample_cqm(c, num_reads=40, num_sweeps=300, beta_range=(0.05, 10.0), seed=3, onehot_penalty=20.0,
                            constraint_penalty=20.0)
    best = next(r for r in ss.data(fields=["sample", "energy", "is_feasible"]) if r.is_feasible)
    chosen = [v for v in ("x0", "x1", "x2", "x3", "x4", "x5") if best.sample[v]]
    assert chosen == ["x0", "x1", "x4"]                                     # cheapest feasible: 1 + 2 + 5


def test_throughput_mode_through_the_sampler(sampler):
    G = load_graph("noisy_circles")
    model = models.cut_balance_model(G, 0.05)
    ss = sampler.sample(model, num_reads=256, num_sweeps=200, beta_range=(0.01, 10.0), seed=8, mode="throughput")
    known = json.loads((GOLD / "known_answers.json").read_text())["noisy_circles"]
    assert ss.first.energy == pytest.approx(known["lower_bound"], rel=1e-10)
    assert np.allclose(ss.record.energy, model.energies(2 * ss.record.sample.astype(np.int64) - 1), rtol=1e-12, atol=1e-8)


# ---- SampleSet post-processing on the device (qa_sort_reads / qa_gather_samples / qa_decode_onehot) --------------------
def test_device_sort_topk_and_onehot_decode(built):
    from scrna_seq_qannealing_clustering_b200.engine import Context
    rng = np.random.default_rng(5)
    R, cells, K, slack = 3000, 37, 4, 11
    energies = np.round(rng.normal(size=R), 1)          # many ties: the order must be stable
    energies[7] = -np.inf
    energies[11] = np.inf
    onehot = rng.integers(0, K, size=(R, cells))
    states = -np.ones((R, cells * K + slack), dtype=np.int8)
    for c in range(cells):
        states[np.arange(R), c * K + onehot[:, c]] = 1
    states[5, 0:K] = 1                                   # read 5: cell 0 has four bits set
    states[6, K:2 * K] = -1                              # read 6: cell 1 has none
    states[:, cells * K:] = rng.choice(np.array([-1, 1], dtype=np.int8), size=(R, slack))
    with Context(0) as ctx:
        order = ctx.sort_reads(energies)
        assert np.array_equal(order, np.argsort(energies, kind="stable"))
        best = ctx.gather_samples(states, order[:16])
        assert np.array_equal(best, states[order[:16]])
        labels, viol = ctx.decode_onehot(states, cells, K, on_value=1, min_size=700)
        if torch_cuda():
            import torch
            d_states = torch.from_numpy(states).cuda()
            d_e = torch.from_numpy(energies).cuda()
            assert np.array_equal(ctx.sort_reads(d_e), order)
            assert np.array_equal(ctx.gather_samples(d_states, order[:16]), best)
            l2, v2 = ctx.decode_onehot(d_states, cells, K, on_value=1, min_size=700)
            assert np.array_equal(l2, labels) and np.array_equal(v2, viol)
    want = onehot.astype(np.int32).copy()
    want[5, 0] = -1
    want[6, 1] = -1
    assert np.array_equal(labels, want)
    not_onehot = (want < 0).sum(axis=1)
    sizes = np.stack([(want == k).sum(axis=1) for k in range(K)], axis=1)
    assert np.array_equal(viol[:, 0], not_onehot) and np.array_equal(viol[:, 1], (sizes < 700).sum(axis=1))


def torch_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_level_batched_recursive_bipartition(built):
    """All sub-graphs of a recursion level are extracted, built and annealed on the device in one batched launch; level 0's best
    energy equals the oracle's on the same inputs, and four well-separated blobs are recovered as the four leaves."""
    from scrna_seq_qannealing_clustering_b200 import clustering
    from scrna_seq_qannealing_clustering_b200.engine import Context
    X, truth = snn.gaussian_mixture_embedding(240, dim=8, centres=4, sep=9.0, seed=1)
    G = snn.to_networkx(snn.snn_graph(X, k=10))
    with Context(0) as ctx:
        labels, levels, energies = clustering.recursive_bipartition_batched(G, gamma_factor=0.05, k=8.0, size_limit=30, iter_limit=2,
                                                                            num_reads=96, num_sweeps=300, seed=3, context=ctx)
    assert [len(lv) for lv in levels] == [1, 2, 4]      # one launch per level
    assert set(labels) == set(G.nodes) and len(set(labels.values())) == 4
    node_truth = {str(i): int(t) for i, t in enumerate(truth)}
    for leaf in set(labels.values()):
        assert len({node_truth[n] for n, l in labels.items() if l == leaf}) == 1      # every leaf is one planted blob
    # level 0 reproduced with the oracle: same structured model, same seeds, same initial states -> same best energy
    lab, eu, ev, w = models.graph_arrays(G)     # as arrays: plain left-to-right sums like the device builder (G.size() compensates)
    m = models.cut_balance_model((lab, eu, ev, w), 0.05, k=8.0, structured=True)
    br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
    betas, spb = schedule.make_beta_schedule(br, 300, 1, "geometric")
    st = schedule.random_spin_states(96, m.num_variables, 3)
    e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, st, betas, spb, schedule.per_read_seeds(3, 96), groups=m.groups.astuple())
    assert energies[0][0] == float(e.min() + m.offset)
