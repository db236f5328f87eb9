"""Host logic (no GPU): dimod-compatible containers, neal argument handling, the C-ABI surface, loud failure without CUDA."""
import ctypes
import re
import warnings
from pathlib import Path

import numpy as np
import pytest

import scrna_seq_qannealing_clustering_b200 as qa
from scrna_seq_qannealing_clustering_b200 import _lib, cqm, schedule
from scrna_seq_qannealing_clustering_b200.bqm import BinaryQuadraticModel
from scrna_seq_qannealing_clustering_b200.sampleset import SampleSet

ROOT = Path(__file__).resolve().parents[1]


# ---- C ABI ---------------------------------------------------------------------------------------------------------
def declared_functions():
    text = (ROOT / "include" / "qanneal.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qa_[a-z0-9_]+)\s*\(", text)) - {"qa_interrupt_fn"})


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/qanneal.h but not exported"
    assert sorted(_lib.SIGNATURES) == names  # the ctypes table covers exactly the header
    assert _lib.load().qa_version() == 200      # round-2 library (split translation units)


def test_stats_struct_matches_header():
    text = (ROOT / "include" / "qanneal.h").read_text()
    body = text[text.index("typedef struct qa_stats {"):text.index("} qa_stats;")]
    fields = re.findall(r"^\s*(?:uint64_t|uint32_t|double)\s+(\w+);", body, flags=re.M)
    assert fields == [f for f, _ in _lib.QAStats._fields_]


def test_no_cpu_fallback(built):
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.QAnnealError) as ei:
        qa.B200SimulatedAnnealingSampler().sample_qubo({(0, 0): -1.0, (0, 1): 2.0}, num_reads=2, num_sweeps=5)
    assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_the_oracle():
    for path in (ROOT / "scrna_seq_qannealing_clustering_b200").rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", path.read_text(), flags=re.M), path


# ---- BQM -----------------------------------------------------------------------------------------------------------
def test_from_qubo_variable_order_and_vectors():
    Q = {("b", "b"): 1.0, ("a", "b"): -2.0, ("c", "a"): 0.5, ("a", "a"): 3.0, ("b", "a"): 0.25}
    bqm = BinaryQuadraticModel.from_qubo(Q)
    assert bqm.variables == ["b", "a", "c"]  # first appearance
    assert bqm.get_quadratic("a", "b") == -1.75 and bqm.num_interactions == 2
    ldata, (irow, icol, qdata), off = bqm.to_numpy_vectors()
    assert ldata.tolist() == [1.0, 3.0, 0.0]
    assert list(zip(irow.tolist(), icol.tolist(), qdata.tolist())) == [(1, 0, -1.75), (2, 1, 0.5)]  # row = larger index
    sp = bqm.spin
    x = np.array([[0, 1, 1], [1, 1, 0], [0, 0, 0]])
    assert np.allclose(bqm.energies(x), sp.energies(2 * x - 1))
    assert np.allclose(sp.binary.energies(x), bqm.energies(x))


def test_equality_and_inequality_constraints():
    bqm = BinaryQuadraticModel({}, {}, 0.0, "BINARY")
    terms = [("x0", 1), ("x1", 1), ("x2", 1), ("x3", 1)]
    slack = bqm.add_linear_inequality_constraint(terms, lagrange_multiplier=2.0, label="c", lb=1, ub=3)
    assert [c for _, c in slack] == [1, 1] and [v for v, _ in slack] == ["slack_c_0", "slack_c_1"]
    for bits in range(16):
        x = [(bits >> i) & 1 for i in range(4)]
        best = min(bqm.energy({**dict(zip(["x0", "x1", "x2", "x3"], x)), "slack_c_0": a, "slack_c_1": b})
                   for a in (0, 1) for b in (0, 1))
        s = sum(x)
        viol = 0 if 1 <= s <= 3 else min((s - 1) ** 2, (s - 3) ** 2)
        assert best == pytest.approx(2.0 * viol)
    with pytest.warns(UserWarning):
        assert BinaryQuadraticModel({}, {}, 0.0, "BINARY").add_linear_inequality_constraint(terms, 1.0, "free", lb=0, ub=4) == []
    with pytest.raises(ValueError):
        BinaryQuadraticModel({}, {}, 0.0, "BINARY").add_linear_inequality_constraint(terms, 1.0, "bad", lb=5, ub=6)


# ---- SampleSet (the access patterns of the reference) ----------------------------------------------
def test_sampleset_access_patterns():
    samples = np.array([[1, 0, 1], [0, 0, 1], [1, 1, 1], [0, 0, 1]], dtype=np.int8)
    ss = SampleSet.from_samples((samples, ["a", "b", "c"]), energy=[3.0, -1.0, 2.0, -1.0], vartype="BINARY", info={"k": 1})
    rows = list(ss.data(fields=["sample", "energy", "num_occurrences"]))
    assert [r.energy for r in rows] == [-1.0, -1.0, 2.0, 3.0] and rows[0].sample == {"a": 0, "b": 0, "c": 1}
    first = ss.first
    assert first.energy == -1.0 and first.sample["c"] == 1 and first.num_occurrences == 1
    lut = first.sample
    assert [v for v in ["a", "b", "c"] if not lut[v]] == ["a", "b"]           # BQM_clustering.py:105-109
    assert ss.record.energy.tolist() == [3.0, -1.0, 2.0, -1.0]                # read order, like neal
    assert ss.sorted().record.energy.tolist() == [-1.0, -1.0, 2.0, 3.0]      # QPU-like order for the `conf` rule
    agg = ss.aggregate()
    assert len(agg) == 3 and sorted(agg.record.num_occurrences.tolist()) == [1, 1, 2]
    top2 = ss.samples()[:2]                                                   # plot_and_save.py:106
    assert len(top2) == 2 and dict(top2[0]) == {"a": 0, "b": 0, "c": 1}
    assert ss.change_vartype("SPIN", inplace=False).record.sample.min() == -1
    assert len(ss.lowest()) == 2 and len(ss.truncate(1)) == 1


# ---- neal argument handling -----------------------------------------------------------------------------------------
def test_beta_schedules():
    b, spb = schedule.make_beta_schedule((0.1, 10.0), 1000, 1, "geometric")
    assert len(b) == 1000 and np.array_equal(b, np.geomspace(0.1, 10.0, 1000))
    b, spb = schedule.make_beta_schedule((0.1, 10.0), 100, 5, "linear")
    assert len(b) == 20 and spb == 5 and np.array_equal(b, np.linspace(0.1, 10.0, 20))
    b, _ = schedule.make_beta_schedule((0.1, 10.0), 1, 1, "geometric")
    assert b.tolist() == [10.0]
    b, _ = schedule.make_beta_schedule(None, None, 2, "custom", beta_schedule=[1.0, 2.0, 3.0])
    assert b.tolist() == [1.0, 2.0, 3.0]
    with pytest.raises(ValueError):
        schedule.make_beta_schedule((0.1, 10.0), 10, 3, "geometric")
    with pytest.raises(ValueError):
        schedule.make_beta_schedule((0.1, 10.0), 10, 1, "exponential")
    with pytest.raises(ValueError):
        schedule.make_beta_schedule(None, 7, 2, "custom", beta_schedule=[1.0, 2.0, 3.0])


def test_default_beta_range_matches_neal_formula():
    h = np.array([0.5, 0.0, -2.0])
    irow, icol, q = np.array([1, 2]), np.array([0, 0]), np.array([1.5, -0.25])
    hot, cold = schedule.default_ising_beta_range(h, irow, icol, q)
    assert hot == pytest.approx(np.log(2) / (0.5 + 1.5 + 0.25)) and cold == pytest.approx(np.log(100) / 0.25)
    assert schedule.default_ising_beta_range(np.zeros(2), np.zeros(0, int), np.zeros(0, int), np.zeros(0)) == (0.1, 1.0)


def test_seeds_and_initial_states():
    s = schedule.per_read_seeds(1234, 1000)
    assert s.dtype == np.uint64 and s.max() < 2 ** 32 and len(np.unique(s)) > 990
    assert np.array_equal(schedule.per_read_seeds(1234, 10, first_read=990), s[990:])  # sharding-invariant
    a = schedule.random_spin_states(4, 16, 7)
    assert a.dtype == np.int8 and set(np.unique(a)) <= {-1, 1}
    rs = np.random.RandomState(7)
    assert np.array_equal(a, rs.choice(np.asarray([1, -1], dtype=np.int8), size=(4, 16)))
    with pytest.raises(ValueError):
        schedule.resolve_seed(-1)
    with pytest.raises(TypeError):
        schedule.resolve_seed(1.5)
    st = schedule.parse_initial_states(3, ["a", "b", "c"], False, (np.array([[1, 0, 1]]), ["c", "b", "a"]), "tile", 3, 0)
    assert st.tolist() == [[1, -1, 1]] * 3
    with pytest.raises(ValueError):
        schedule.parse_initial_states(3, ["a", "b", "c"], False, None, "none", 2, 0)


# ---- CQM / DQM containers --------------------------------------------------------------------------------------------
def test_cqm_expression_algebra_and_lowering():
    x = [cqm.Binary(f"x{i}") for i in range(4)]
    model = cqm.ConstrainedQuadraticModel()
    model.add_discrete(["x0", "x1"], label="d0")
    model.set_objective(sum([x[0] + x[2] - 2 * 0.5 * x[0] * x[2], x[1] + x[3] - 2 * 0.25 * x[1] * x[3]]))
    model.add_constraint(sum(x) >= 2, label="atleast2")
    low = model.to_lowered(onehot_penalty=10.0, constraint_penalty=3.0)
    labels = low.labels
    for bits in range(16):
        xv = {f"x{i}": (bits >> i) & 1 for i in range(4)}
        obj = xv["x0"] + xv["x2"] - xv["x0"] * xv["x2"] + xv["x1"] + xv["x3"] - 0.5 * xv["x1"] * xv["x3"]
        pen = 10.0 * (xv["x0"] + xv["x1"] - 1) ** 2
        slack_labels = [v for v in labels if str(v).startswith("slack_")]
        best = np.inf
        for sb in range(1 << len(slack_labels)):
            full = dict(xv)
            full.update({v: (sb >> i) & 1 for i, v in enumerate(slack_labels)})
            s = np.array([[2 * full[v] - 1 for v in labels]])
            best = min(best, float(low.energies(s)[0]))
        total = sum(xv.values())
        assert best == pytest.approx(obj + pen + 3.0 * max(0, 2 - total) ** 2)
    feas = model.check_feasible(np.array([[1, 0, 1, 0], [1, 1, 1, 1], [1, 0, 0, 0]]), ["x0", "x1", "x2", "x3"])
    assert feas.tolist() == [True, False, False]


def test_dqm_container_set_semantics():
    d = qa.DiscreteQuadraticModel()
    for v in "abc":
        d.add_variable(2, label=v)
    d.set_linear("a", [1.0, 2.0])
    d.set_quadratic("a", "b", {(0, 0): 5.0, (1, 1): 5.0})
    d.set_quadratic("a", "b", {(0, 0): -1.0})          # overwrites that entry only
    d.set_quadratic("c", "b", {(1, 0): 2.0})
    assert d.get_quadratic("a", "b") == {(0, 0): -1.0, (1, 1): 5.0} and d.get_quadratic("b", "c") == {(0, 1): 2.0}
    assert d.energies([[0, 0, 1], [1, 1, 0]]).tolist() == [1.0 - 1.0 + 2.0, 2.0 + 5.0]
    low = d.to_lowered(penalty=7.0)
    s = -np.ones((1, 6))
    s[0, [0, 2, 5]] = 1  # a=0, b=0, c=1 one-hot
    assert low.energies(s)[0] == pytest.approx(2.0)


def test_sampler_surface():
    s = qa.B200SimulatedAnnealingSampler()
    assert set(["beta_range", "num_reads", "num_sweeps", "num_sweeps_per_beta", "beta_schedule_type", "seed",
                "interrupt_function", "beta_schedule", "initial_states", "initial_states_generator"]) <= set(s.parameters)
    assert s.properties["beta_schedule_options"] == ("linear", "geometric", "custom")
    assert qa.SimulatedAnnealingSampler is qa.B200SimulatedAnnealingSampler
    # empty model: empty SampleSet without touching the GPU (neal returns an empty SampleSet too)
    ss = s.sample(BinaryQuadraticModel({}, {}, 0.0, "SPIN"), num_reads=3)
    assert len(ss) == 0
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        s.sample(BinaryQuadraticModel({}, {}, 0.0, "SPIN"), label="x", chain_strength=4)
    assert len(w) == 2


def test_dwave_samplers_default_beta_range_rule():
    """The newer default range (dwave-samplers >= 1.0) on a hand-computable model: h = (1, 0, -2), J01 = 0.5, J12 = -0.25."""
    h = np.array([1.0, 0.0, -2.0])
    irow = np.array([1, 2], dtype=np.int32)
    icol = np.array([0, 1], dtype=np.int32)
    q = np.array([0.5, -0.25])
    hot, cold = schedule.default_ising_beta_range_samplers(h, irow, icol, q)
    # fields: v0 1.5, v1 0.75, v2 2.25 -> hot = ln2 / 4.5 ; smallest non-zero bias per variable: 0.5, 0.25, 0.25 -> two minima
    assert hot == pytest.approx(np.log(2) / 4.5)
    assert cold == pytest.approx(np.log(2 / 0.01) / 0.5)
    _, cold1 = schedule.default_ising_beta_range_samplers(h, irow, icol, q, scale_T_with_N=False)
    assert cold1 == pytest.approx(np.log(1 / 0.01) / 0.5)
    assert schedule.default_ising_beta_range_samplers(np.zeros(3), np.zeros(0, int), np.zeros(0, int), np.zeros(0)) == (1.0, 1.0)


def test_consistent_slack_makes_the_size_penalty_start_at_zero():
    """models.consistent_slack (used by sample_cqm for generated initial states): slack bits encode max(N_j - min_size, 0)
    exactly, for dimod's binary encoding with its odd last coefficient, and clamp at the ends of the range."""
    import numpy as np
    from scrna_seq_qannealing_clustering_b200 import models, schedule, snn
    g = snn.synthetic_snn(300, k=5, seed=2)[0]
    m = models.cqm_model(g, 3, min_size=20)
    assert m.meta["size_penalty"] == m.meta["onehot_penalty"] / 300      # default B = A / n
    K, nx = 3, m.meta["num_cell_variables"]
    co = np.array(m.meta["slack_coefficients"])
    st = schedule.random_spin_states(40, m.num_variables, 1)
    st[0, :nx] = -1                                   # empty clusters: below min_size -> slack 0
    st[1, :nx] = 1                                    # every bit set: N_j = n -> slack n - min_size (the top of the range)
    models.consistent_slack(st, m.meta)
    x = (st + 1) // 2
    N = x[:, :nx].reshape(40, -1, K).sum(axis=1)
    slack = x[:, nx:].reshape(40, K, len(co)) @ co
    assert np.array_equal(slack, np.clip(N - 20, 0, co.sum()))
    assert (slack[0] == 0).all() and (slack[1] == 300 - 20).all()
    assert set(np.unique(st)) <= {-1, 1}
