"""Model construction pinned against the reference's OWN functions.

tests/golden/q_*.npz hold the Q dicts / DQM / CQM coefficients that Python_Functions/{BQM,DQM,CQM}_clustering.py and
QA_subsampling.py built on the six fixture graphs of R/benchmarks (captured by tools/make_golden.py with the D-Wave
packages stubbed).  Checked here, all on CPU:
  (1) oracle/models_ref.py (line-by-line restatement) reproduces every captured coefficient bit for bit, in order;
  (2) the vectorised builders of the product (models.py) give bit-identical Ising vectors to the dimod-style lowering of
      the captured Q, and structured (rank-1) forms agree in energy to 1e-12 relative;
  (3) the known-answer energies of SURVEY.md section 4.
"""
import json
from pathlib import Path

import networkx as nx
import numpy as np
import pytest

from oracle import models_ref
from scrna_seq_qannealing_clustering_b200 import models
from scrna_seq_qannealing_clustering_b200.bqm import BinaryQuadraticModel

GOLD = Path(__file__).parent / "golden"
NAMES = ["noisy_circles", "noisy_moons", "varied", "aniso", "blobs", "no_structure"]


def load_graph(name):
    g = np.load(GOLD / "graphs.npz")
    labels = [str(x) for x in g[f"{name}_labels"]]
    G = nx.Graph()
    G.add_nodes_from(labels)
    for u, v, w in zip(g[f"{name}_eu"], g[f"{name}_ev"], g[f"{name}_w"]):
        G.add_edge(labels[u], labels[v], weight=float(w))
    return G, labels


def q_from_npz(z, prefix, labels):
    return {(labels[i], labels[j]): q for i, j, q in zip(z[f"{prefix}_i"], z[f"{prefix}_j"], z[f"{prefix}_q"].tolist())}


def assert_same_dict(got, want):
    assert list(got.keys()) == list(want.keys())
    a = np.array(list(got.values()), dtype=np.float64)
    b = np.array(list(want.values()), dtype=np.float64)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def bqm_in_node_order(Q, labels):
    bqm = BinaryQuadraticModel({}, {}, 0.0, "BINARY")
    for v in labels:
        bqm.add_variable(v, 0.0)
    for (u, v), q in Q.items():
        if u == v:
            bqm.add_linear(u, q)
        else:
            bqm.add_quadratic(u, v, q)
    return bqm


def assert_same_lowering(model, ref):
    assert np.array_equal(model.starts, ref.starts) and np.array_equal(model.ends, ref.ends)
    assert np.array_equal(model.weights.view(np.uint64), ref.weights.view(np.uint64))
    assert np.array_equal(model.h.view(np.uint64), ref.h.view(np.uint64))
    assert model.offset == pytest.approx(ref.offset, rel=1e-12, abs=1e-9)


@pytest.mark.parametrize("name", NAMES)
def test_restatement_reproduces_reference_q(name):
    G, labels = load_graph(name)
    z = np.load(GOLD / f"q_{name}.npz")
    assert_same_dict(dict(models_ref.qubo_clustering_bqm(G, 0.05)[0]), q_from_npz(z, "bqm", labels))
    assert_same_dict(dict(models_ref.qubo_clustering_bqm_2(G, 0.01, 1)[0]), q_from_npz(z, "bqm2", labels))
    Q3, c1, lb, ub, lam = models_ref.qubo_clustering_bqm_3(G, 0.05, 40)
    assert_same_dict(dict(Q3), q_from_npz(z, "bqm3", labels))
    assert [lb, ub, lam] == z["bqm3_constraint"].tolist()
    assert [labels.index(v) for v, _ in c1] == z["bqm3_terms"].tolist()
    assert_same_dict(dict(models_ref.qubo_graph_subsampling(G, 7)), q_from_npz(z, "sub", labels))


@pytest.mark.parametrize("name", NAMES)
def test_restatement_reproduces_reference_dqm(name):
    G, labels = load_graph(name)
    z = np.load(GOLD / f"q_{name}.npz")
    lin, quad = models_ref.dqm_clustering(G, 3, 0.005)
    assert np.array_equal(np.array([lin[v] for v in labels]).view(np.uint64), z["dqm_linear"].view(np.uint64))
    keys = list(quad.keys())
    assert [labels.index(a) for a, _ in keys] == z["dqm_u"].tolist()
    assert [labels.index(b) for _, b in keys] == z["dqm_v"].tolist()
    diag = np.array([[quad[k][(c, c)] for c in range(3)] for k in keys])
    assert np.array_equal(diag.view(np.uint64), z["dqm_diag"].view(np.uint64))
    assert int(z["dqm_offdiag_count"][0]) == 0  # only same-case interactions exist (DQM_clustering.py:37,41)


def test_restatement_reproduces_reference_cqm():
    G, labels = load_graph("noisy_circles")
    z = np.load(GOLD / "cqm_noisy_circles.npz")
    lin, quad, disc, size = models_ref.cqm_clustering(G, 3)
    pos = {f"v_{v},{k}": i * 3 + k for i, v in enumerate(labels) for k in range(3)}
    got = np.zeros(len(pos))
    for v, b in lin.items():
        got[pos[v]] = b
    assert np.array_equal(got, z["lin"])
    assert [pos[a] for a, _ in quad] == z["qu"].tolist() and [pos[b] for _, b in quad] == z["qv"].tolist()
    assert np.array_equal(np.array(list(quad.values())).view(np.uint64), z["qq"].view(np.uint64))
    assert [[pos[v] for v in grp] for grp in disc.values()] == z["discrete"].tolist()
    assert z["size_rhs"].tolist() == [20.0] * 3 and z["size_sense_ge"].all()
    assert [[pos[v] for v in vs] for vs, _ in size.values()] == z["size_vars"].tolist()
    assert float(z["offset"][0]) == 0.0


@pytest.mark.parametrize("name", NAMES)
def test_builders_match_dimod_style_lowering_of_reference_q(name):
    G, labels = load_graph(name)
    z = np.load(GOLD / f"q_{name}.npz")
    for prefix, model in (("bqm", models.cut_balance_model(G, 0.05, structured=False)),
                          ("bqm2", models.cut_linear_model(G, 0.01, 1)),
                          ("sub", models.subsampling_model(G, 7))):
        ref = models.lowered_from_bqm(bqm_in_node_order(q_from_npz(z, prefix, labels), labels))
        assert model.labels == labels
        assert_same_lowering(model, ref)


@pytest.mark.parametrize("name", ["noisy_circles", "blobs"])
def test_structured_forms_agree_with_materialised(name):
    G, labels = load_graph(name)
    rng = np.random.default_rng(0)
    dense = models.cut_balance_model(G, 0.05, structured=False)
    rank1 = models.cut_balance_model(G, 0.05, structured=True)
    s = rng.integers(0, 2, size=(64, len(labels))) * 2 - 1
    assert np.allclose(dense.energies(s), rank1.energies(s), rtol=1e-12, atol=1e-9)
    assert np.allclose(rank1.materialise().energies(s), rank1.energies(s), rtol=1e-12, atol=1e-9)
    assert rank1.num_couplers == G.number_of_edges() and dense.num_couplers == len(labels) * (len(labels) - 1) // 2


@pytest.mark.parametrize("name", ["noisy_moons", "aniso"])
@pytest.mark.parametrize("structured", [True, False])
def test_dqm_builder_matches_reference_dqm_energy(name, structured):
    G, labels = load_graph(name)
    z = np.load(GOLD / f"q_{name}.npz")
    n, K = len(labels), 3
    model = models.dqm_model(G, K, 0.005, penalty=50.0, semantics="as_written", structured=structured)
    rng = np.random.default_rng(1)
    cases = rng.integers(0, K, size=(16, n))
    # reference DQM energy from the captured coefficients
    lin, u, v, diag = z["dqm_linear"], z["dqm_u"], z["dqm_v"], z["dqm_diag"]
    e_ref = lin[np.arange(n)[None, :], cases].sum(axis=1)
    same = cases[:, u] == cases[:, v]
    e_ref = e_ref + (same * diag[np.arange(len(u))[None, :], cases[:, u]]).sum(axis=1)
    bits = np.zeros((16, n * K), dtype=np.int8)
    bits[np.repeat(np.arange(16), n), (np.arange(n)[None, :] * K + cases).ravel()] = 1
    e = model.energies(2 * bits.astype(np.int64) - 1)
    # one-hot satisfied: the penalty contributes nothing.  Tolerance 1e-12 RELATIVE TO THE SUMMED MAGNITUDES: the spin
    # form carries an offset of ~penalty*n = 1.3e4 that cancels against the couplers, so absolute 1e-12 * 1e5 = 1e-7
    scale = abs(model.offset) + np.abs(model.h).sum() + np.abs(model.weights).sum()
    assert np.allclose(e, e_ref, rtol=1e-12, atol=1e-12 * scale)
    # a violated one-hot costs exactly `penalty` per extra / missing bit
    bits2 = bits.copy()
    bits2[:, 0:K] = 0
    e2 = model.energies(2 * bits2.astype(np.int64) - 1)
    assert np.all(e2 - e > 50.0 - 5.0)


def test_cqm_builder_matches_reference_objective_and_constraints():
    G, labels = load_graph("noisy_circles")
    z = np.load(GOLD / "cqm_noisy_circles.npz")
    n, K = len(labels), 3
    A, B = 40.0, 2.0
    model = models.cqm_model(G, K, min_size=20, onehot_penalty=A, size_penalty=B)
    assert model.labels[: n * K] == [f"v_{v},{k}" for v in labels for k in range(K)]
    rng = np.random.default_rng(2)
    coeffs = np.array(model.meta["slack_coefficients"])
    nb = len(coeffs)
    assert coeffs.sum() == n - 20 and model.num_variables == n * K + K * nb
    x = (rng.random((32, n * K)) < 0.4).astype(np.int64)
    sl = (rng.random((32, K * nb)) < 0.5).astype(np.int64)
    obj = x @ z["lin"] + (x[:, z["qu"]] * x[:, z["qv"]]) @ z["qq"]
    onehot = A * ((x[:, z["discrete"]].sum(axis=2) - 1) ** 2).sum(axis=1)
    size = np.zeros(32)
    for j in range(K):
        N = x[:, z["size_vars"][j]].sum(axis=1)
        S = sl[:, j * nb:(j + 1) * nb] @ coeffs
        size += B * (N - 20 - S) ** 2
    e = model.energies(2 * np.concatenate([x, sl], axis=1) - 1)
    assert np.allclose(e, obj + onehot + size, rtol=1e-12, atol=1e-8)
    # materialised form = same energies
    small = models.cqm_model((12, *_ring(12)), 2, min_size=3, onehot_penalty=5.0, size_penalty=1.5)
    s = rng.integers(0, 2, size=(50, small.num_variables)) * 2 - 1
    assert np.allclose(small.materialise().energies(s), small.energies(s), rtol=1e-12, atol=1e-10)


def _ring(n):
    eu = np.arange(n)
    ev = (np.arange(n) + 1) % n
    return np.minimum(eu, ev), np.maximum(eu, ev), np.linspace(0.2, 1.0, n)


def test_bqm3_inequality_lowering_matches_dimod_semantics():
    G, labels = load_graph("noisy_moons")
    z = np.load(GOLD / "q_noisy_moons.npz")
    lb, ub, lam = z["bqm3_constraint"].tolist()
    bqm = models.cut_inequality_bqm(G, 0.05, 40)
    slack = [v for v in bqm.variables if str(v).startswith("slack_c1_constraint_")]
    # ub = 256/6 = 42.67, lb = 40 -> slack range int(2.67) = 2 -> coefficients [1, 1]; penalty keeps the fractional ub
    assert len(slack) == 2 and bqm.num_variables == len(labels) + 2
    rng = np.random.default_rng(3)
    Q = q_from_npz(z, "bqm3", labels)
    base = bqm_in_node_order(Q, labels)
    for _ in range(5):
        x = (rng.random(len(labels)) < 0.16).astype(int)
        s = rng.integers(0, 2, size=2)
        sample = dict(zip(labels, x.tolist()))
        e0 = base.energy(sample)
        sample.update(dict(zip(slack, s.tolist())))
        want = e0 + lam * (x.sum() + s.sum() - ub) ** 2
        assert bqm.energy(sample) == pytest.approx(want, rel=1e-10)


def test_known_answer_energies():
    known = json.loads((GOLD / "known_answers.json").read_text())
    # SURVEY.md section 4 (independently computed there)
    assert known["noisy_circles"]["E_parity"] == 638.2080612198255
    assert known["noisy_circles"]["E_first_half"] == 792.3621025689133
    assert known["noisy_circles"]["E_largest_component"] == -2951.8108596597763
    assert known["blobs"]["E_largest_component"] == -2630.105678136055
    assert known["noisy_moons"]["E_parity"] == 820.7776167417034
    for name in NAMES:
        G, labels = load_graph(name)
        k = known[name]
        n = k["n"]
        for structured in (True, False):
            m = models.cut_balance_model(G, 0.05, structured=structured)
            assert m.meta["gamma"] == k["gamma"] and m.meta["W"] == k["W"]
            ids = np.array([int(v) for v in labels])
            x_par = ids % 2
            x_half = (ids < n // 2).astype(int)
            e = m.energies(np.stack([2 * x_par - 1, 2 * x_half - 1]))
            assert e[0] == pytest.approx(k["E_parity"], rel=1e-12)
            assert e[1] == pytest.approx(k["E_first_half"], rel=1e-12)
            if "E_largest_component" in k:
                big = max(nx.connected_components(G), key=len)
                x = np.array([int(v in big) for v in labels])
                assert m.energies(2 * x - 1)[0] == pytest.approx(k["E_largest_component"], rel=1e-12)
    # noisy_circles: component split has cut = 0 and perfect balance -> provably optimal E* = -gamma n^2 / 4
    k = known["noisy_circles"]
    assert k["E_largest_component"] == pytest.approx(k["lower_bound"], rel=1e-14)


def test_dense_kway_builder_equals_the_dqm_builder_on_a_complete_graph():
    """models.dense_kway_model (BASELINE config 5) emits, without a networkx round trip, exactly the couplers of
    dqm_model(..., semantics='intended', structured=False) on the complete graph of the affinity matrix -- same order, same bits;
    the linear part differs only by the summation order of the weighted degrees."""
    from scrna_seq_qannealing_clustering_b200 import snn
    X, _ = snn.gaussian_mixture_embedding(36, dim=6, centres=3, seed=2)
    A = snn.gaussian_affinity(X, k=5)
    G = nx.Graph()
    G.add_nodes_from(str(i) for i in range(len(X)))
    for i in range(len(X)):
        for j in range(i):
            G.add_edge(str(j), str(i), weight=float(A[i, j]))
    for K in (1, 2, 4):
        a = models.dqm_model(G, K, 0.05, penalty=30.0, semantics="intended", structured=False)
        b = models.dense_kway_model(A, K, 0.05, penalty=30.0)
        assert np.array_equal(a.starts, b.starts) and np.array_equal(a.ends, b.ends)
        assert np.array_equal(a.weights.view(np.uint64), b.weights.view(np.uint64))
        assert np.allclose(a.h, b.h, rtol=0, atol=1e-12) and a.offset == pytest.approx(b.offset, abs=1e-9)
        assert b.meta["num_cases"] == K and b.meta["dense"] is True
