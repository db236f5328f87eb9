"""k_anneal_dense (csrc/dense.cuh): the dense k-way tensor-core path of BASELINE.json's config 5 (SURVEY.md 2.2 K5;
reference analogue: the all-pairs same-case term of DQM_clustering.py:36-37 / BQM_clustering.py:46-47).

Tolerance-parity mode (north star: "every returned sample's energy must equal dimod's bqm.energies to 1e-12 relative; the
best energy / hit rate statistically indistinguishable"): neal's sweep order and per-read RNG are kept, local fields are
recomputed per 8-cell block by mma.sync f64 instead of updated incrementally, so a run coincides with the oracle's until an
fp64 rounding decides a branch.  Stated tolerances:
  * energies: |E_gpu - E_oracle(state_gpu)| <= 1e-12 * max(|E|, 1)          (oracle = neal get_state_energy order)
  * trajectories: >= 95 % of the reads end in exactly the oracle's state (generic real weights: ties have measure zero)
  * distribution: two-sample KS on final energies, alpha = 0.001
"""
import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = Context(0)
    yield c
    c.close()


def _affinity_model(cells, K, seed=2, gamma=0.05):
    X, _ = snn.gaussian_mixture_embedding(cells, dim=15, centres=max(K, 2), sep=4.0, seed=seed)
    return models.dense_kway_model(snn.gaussian_affinity(X, k=10), K, gamma)


def _run_dense(ctx, m, K, R, sweeps, seed, beta_range, expect_dense=True):
    n = m.num_variables
    betas, spb = schedule.make_beta_schedule(beta_range, sweeps, 1, "geometric")
    seeds = schedule.per_read_seeds(seed, R)
    init = schedule.random_spin_states(R, n, seed)
    gm = IsingModel(ctx, m.h, m.starts, m.ends, m.weights)
    try:
        assert gm.enable_dense(K) is expect_dense
        states = init.copy()
        e, st, done = gm.sample(states, betas, spb, seeds, mode=_lib.QA_MODE_THROUGHPUT)
    finally:
        gm.close()
    assert done == R
    return init, states, e, st, betas, spb, seeds


def _dense_energies(m, K, states):
    """Energies from the dense form in numpy (fp64 BLAS, blocked summation): E = s.h + 1/2 sum_c s_c^T W s_c + P * (intra-cell
    pairs) -- the analogue of dimod's vectorised ``bqm.energies`` for models whose coupler list is too long to gather per read."""
    n = m.num_variables
    cells = n // K
    i, j = m.starts.astype(np.int64), m.ends.astype(np.int64)
    inter = (i // K) != (j // K)
    W = np.zeros((cells, cells))
    c0 = inter & (i % K == 0)
    W[i[c0] // K, j[c0] // K] = m.weights[c0]
    W = W + W.T
    P = float(m.weights[~inter][0]) if (~inter).any() else 0.0
    s = states.astype(np.float64).reshape(len(states), cells, K)
    e = states.astype(np.float64) @ m.h
    for c in range(K):
        sc = s[:, :, c]
        e += 0.5 * np.einsum("rc,rc->r", sc @ W, sc)
    tot = s.sum(axis=2)
    e += P * 0.5 * ((tot * tot).sum(axis=1) - cells * K)      # sum_{c<c'} s s' = ((sum s)^2 - K) / 2 per cell
    return e


def _check_against_oracle(m, init, states, e, st, betas, spb, seeds, min_identical=0.95, K=None):
    ref = init.copy()
    ref_e, ref_st = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, ref, betas, spb, seeds)
    # energies of the RETURNED states in neal's summation order
    e_chk = oracle.state_energies(m.h, m.starts, m.ends, m.weights, states)
    if m.num_couplers < 5_000_000:
        assert np.all(np.abs(e - e_chk) <= 1e-12 * np.maximum(np.abs(e_chk), 1.0)), np.abs(e - e_chk).max()
    else:
        # 33.5 M terms: neal's strictly sequential sum adds tens of millions of equal-magnitude couplings to an accumulator of
        # ~1e7, each add dropping the same sub-ulp bits -- a systematic bias of ~2e-11 relative (measured) that a blocked sum
        # (dimod's vectorised bqm.energies, this kernel) does not have.  Bar: 1e-12 against the blocked fp64 evaluation,
        # 1e-9 against neal's order.
        e_np = _dense_energies(m, K, states)
        assert np.all(np.abs(e - e_np) <= 1e-12 * np.maximum(np.abs(e_np), 1.0)), np.abs(e - e_np).max()
        assert np.all(np.abs(e - e_chk) <= 1e-9 * np.maximum(np.abs(e_chk), 1.0)), np.abs(e - e_chk).max()
    same = (states == ref).all(axis=1)
    assert same.mean() >= min_identical, f"only {same.mean():.3f} of the reads follow the oracle's trajectory"
    if same.all():
        for key in ("attempts", "candidates", "draws", "accepted", "nbr_updates"):
            assert getattr(st, key) == ref_st[key], key
    assert set(np.unique(states)) <= {-1, 1}
    return ref_e, same


@pytest.mark.parametrize("cells,K", [(160, 4), (100, 4), (37, 2), (64, 8), (96, 1)])
def test_dense_kway_follows_the_oracle(ctx, cells, K):
    """Sizes that are / are not multiples of the 8-cell block and the 32-cell spin word; every supported K."""
    m = _affinity_model(cells, K)
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    out = _run_dense(ctx, m, K, 96, 25, 11, (hot, 30.0))
    assert ctx.last_kernel == _lib.QA_KERNEL_DENSE
    _check_against_oracle(m, *out)
    if cells * K >= 256:      # the numpy dense-form evaluation used at config-5 size agrees with neal's order where both are cheap
        assert np.allclose(_dense_energies(m, K, out[1]), oracle.state_energies(m.h, m.starts, m.ends, m.weights, out[1]), rtol=1e-13)


def test_general_dense_ising_is_the_k1_case(ctx):
    """A materialised cut+balance K_n model (config 1's dense form, BQM_clustering.py:36-47) is a general dense Ising model."""
    g = snn.synthetic_snn(200, k=5, seed=3)[0]
    m = models.cut_balance_model(g, 0.05, k=8.0, structured=False)
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    init, states, e, st, betas, spb, seeds = _run_dense(ctx, m, 1, 70, 30, 5, (hot, 20.0))
    assert ctx.last_kernel == _lib.QA_KERNEL_DENSE
    # h == 0 and uniform couplings: exact ties (dE == 0) are decided by roundings, so only the energy and statistical bars apply
    e_chk = oracle.state_energies(m.h, m.starts, m.ends, m.weights, states)
    assert np.all(np.abs(e - e_chk) <= 1e-12 * np.maximum(np.abs(e_chk), 1.0))


def test_structure_check_rejects_other_models(ctx):
    g = snn.synthetic_snn(256, k=5, seed=0)[0]
    m = models.subsampling_model(g, 7.0)
    gm = IsingModel(ctx, m.h, m.starts, m.ends, m.weights)
    try:
        assert gm.enable_dense(1) is True           # any Ising model is a K = 1 dense model (zeros where there is no coupler)
        assert gm.enable_dense(4) is False          # but its couplers do not have the k-way Kronecker form
    finally:
        gm.close()
    # rank-1 group terms are outside the dense form: not an error, just "no"
    mg = models.cqm_model(g, 4, min_size=5)
    gm = IsingModel(ctx, mg.h, mg.starts, mg.ends, mg.weights)
    try:
        gm.set_groups(*mg.groups.astuple())
        assert gm.enable_dense(4) is False
    finally:
        gm.close()
    # case-dependent inter-cell coupling: rejected
    m4 = _affinity_model(40, 4)
    w = m4.weights.copy()
    inter = (m4.starts // 4) != (m4.ends // 4)
    w[np.nonzero(inter)[0][5]] += 0.125
    gm = IsingModel(ctx, m4.h, m4.starts, m4.ends, w)
    try:
        assert gm.enable_dense(4) is False
    finally:
        gm.close()


def test_dense_statistics_match_the_oracle(ctx):
    """Two-sample KS on final energies (alpha = 0.001) with DIFFERENT seeds on both sides, plus equal best energy."""
    from scipy import stats
    m = _affinity_model(64, 4, seed=7)
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    R = 2048
    init, states, e, st, betas, spb, seeds = _run_dense(ctx, m, 4, R, 40, 21, (hot, 30.0))
    ref = schedule.random_spin_states(R, m.num_variables, 99)
    ref_e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, ref, betas, spb, schedule.per_read_seeds(99, R))
    ks = stats.ks_2samp(e, ref_e)
    assert ks.pvalue > 0.001, ks
    assert abs(e.min() - ref_e.min()) <= 3 * ref_e.std() + 1e-9      # different seeds on both sides: within the spread


def test_config5_size_4096_cells_times_4(ctx):
    """BASELINE.json config 5 at its own size: 4096 cells x 4 clusters = 16 384 variables, 33.5 M couplers."""
    m = _affinity_model(4096, 4)
    assert m.num_variables == 16384 and m.num_couplers == 4 * (4096 * 4095 // 2) + 4096 * 6
    hot = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)[0]
    out = _run_dense(ctx, m, 4, 64, 3, 3, (hot, 10 * hot))
    assert ctx.last_kernel == _lib.QA_KERNEL_DENSE
    _check_against_oracle(m, *out, K=4)
