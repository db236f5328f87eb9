"""The replay kernel (QA_KERNEL_REPLAY, csrc/replay.cuh): deferred exact neighbour updates, TMA-staged coupling slabs.

Same bar as test_gpu_parity.py -- states bytewise, energies bitwise, event counters equal to the oracle's -- on the paths
that are specific to this kernel: the replay -> push hand-over at every possible point (never, after the first sweep,
mid-schedule), CTAs of 1/2/4/8 warps, partially filled CTAs and tiles, several work items per CTA, batches."""
import os

import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _replay_ctx(permille=None, warps=None):
    env = {}
    if permille is not None:
        env["QA_REPLAY_SWITCH_PERMILLE"] = permille
    if warps is not None:
        env["QA_REPLAY_WARPS"] = warps
    with _Env(**env):   # development knobs are read when the context is created
        ctx = Context(0)
    ctx.set_kernel(_lib.QA_KERNEL_REPLAY)
    return ctx


def _check(ctx, model, R, sweeps, seed, beta_range=(0.05, 8.0), expect_replay=True, spb=1):
    n = model.num_variables
    groups = model.groups.astuple() if model.groups is not None else None
    betas, spb = schedule.make_beta_schedule(beta_range, sweeps, spb, "geometric")
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
    assert ctx.last_kernel == (_lib.QA_KERNEL_REPLAY if expect_replay else _lib.QA_KERNEL_LOCKSTEP_PUSH)
    bad = np.nonzero((states != ref).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {R} reads differ, first {bad[:8]}"
    assert np.array_equal(e.view(np.uint64), ref_e.view(np.uint64))
    for key in ("attempts", "candidates", "draws", "accepted", "nbr_updates"):
        assert getattr(st, key) == ref_st[key], key
    assert st.near_ties == 0


@pytest.fixture(scope="module")
def graph256():
    return snn.synthetic_snn(256, k=5, seed=3)[0]


@pytest.mark.parametrize("permille", [0, 20, 150, 1000])
def test_hand_over_point_does_not_change_the_result(built, graph256, permille):
    """0: the whole schedule replays; 1000: one replay sweep, catch-up, then push; in between: mid-schedule."""
    ctx = _replay_ctx(permille=permille)
    try:
        _check(ctx, models.subsampling_model(graph256, 7.0), 200, 100, 5)
        _check(ctx, models.cqm_model(graph256, 4, min_size=20), 96, 80, 10, beta_range=(0.02, 6.0))
    finally:
        ctx.close()


@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_cta_shapes_and_ragged_read_counts(built, graph256, warps):
    """R = 333 reads = 11 tiles (the last with 13 active lanes): partially filled CTAs for every CTA width."""
    ctx = _replay_ctx(warps=warps)
    try:
        _check(ctx, models.cut_linear_model(graph256, 0.01, 1.0), 333, 60, 6)
    finally:
        ctx.close()


def test_structured_models_with_groups(built, graph256):
    ctx = _replay_ctx()
    try:
        _check(ctx, models.cut_balance_model(graph256, 0.05, structured=True), 100, 100, 8, beta_range=(0.01, 4.0))
        _check(ctx, models.dqm_model(graph256, 4, 0.005, semantics="intended"), 64, 80, 9, beta_range=(0.02, 6.0))
    finally:
        ctx.close()


def test_dense_model_falls_back_to_the_push_kernel(built, graph256):
    ctx = _replay_ctx()
    try:
        m = models.cut_balance_model(graph256, 0.05, structured=False)   # K_256 does not fit the slab format
        _check(ctx, m, 64, 40, 7, beta_range=(0.01, 4.0), expect_replay=False)
    finally:
        ctx.close()


def test_unsorted_couplers_duplicates_and_isolated_variables(built):
    """Replay order is by variable index, neal's adjacency order is push_back order: shuffled and duplicated couplers."""
    rng = np.random.default_rng(0)
    n = 150
    pairs = [(u, v) for u in range(130) for v in range(u) if rng.random() < 0.06]   # variables 130..149 isolated
    pairs += pairs[:25]                                                               # the same coupler twice
    rng.shuffle(pairs)
    flip = rng.random(len(pairs)) < 0.5
    starts = np.array([p[1] if f else p[0] for p, f in zip(pairs, flip)], dtype=np.int32)
    ends = np.array([p[0] if f else p[1] for p, f in zip(pairs, flip)], dtype=np.int32)
    model = models.LoweredModel(rng.normal(size=n), starts, ends, rng.normal(size=len(pairs)), 0.0, list(range(n)))
    for permille in (0, 60):
        ctx = _replay_ctx(permille=permille)
        try:
            _check(ctx, model, 70, 80, 3, beta_range=(0.1, 5.0))
        finally:
            ctx.close()


def test_more_work_items_than_resident_ctas(built):
    """Tiny problem, many reads: every CTA loops over several work items and the TMA ring carries on across them."""
    g = snn.synthetic_snn(64, k=4, seed=5)[0]
    m = models.subsampling_model(g, 7.0)
    ctx = _replay_ctx(warps=2)
    try:
        _check(ctx, m, 148 * 2 * 32 * 9 + 77, 12, 21, beta_range=(0.1, 5.0))
    finally:
        ctx.close()


def test_batch_of_independent_problems(built):
    graphs = snn.subsample_problems(2000, 6, 200, k=6, dim=10, seed=4)
    ms = [models.subsampling_model(g, 7.0) for g in graphs]
    rpp = 70
    betas, spb = schedule.make_beta_schedule((0.05, 6.0), 50, 1, "geometric")
    voff = np.cumsum([0] + [m.num_variables for m in ms])
    coff = np.cumsum([0] + [m.num_couplers for m in ms])
    seeds = schedule.per_read_seeds(17, rpp * len(ms))
    inits = [schedule.random_spin_states(rpp, m.num_variables, 100 + i) for i, m in enumerate(ms)]
    states = np.concatenate([s.ravel() for s in inits]).copy()
    ctx = _replay_ctx()
    try:
        e, st, done = ctx.sample_ising_batch(voff, coff, np.concatenate([m.h for m in ms]),
                                             np.concatenate([m.starts for m in ms]), np.concatenate([m.ends for m in ms]),
                                             np.concatenate([m.weights for m in ms]), rpp, states, betas, spb, seeds)
        assert ctx.last_kernel == _lib.QA_KERNEL_REPLAY
    finally:
        ctx.close()
    assert done == rpp
    off = 0
    for i, m in enumerate(ms):
        ref = inits[i].copy()
        ref_e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, ref, betas, spb, seeds[i * rpp:(i + 1) * rpp])
        got = states[off:off + rpp * m.num_variables].reshape(rpp, m.num_variables)
        off += rpp * m.num_variables
        assert np.array_equal(got, ref)
        assert np.array_equal(e[i * rpp:(i + 1) * rpp].view(np.uint64), ref_e.view(np.uint64))


def test_large_group_coefficients_take_the_64_bit_path(built, graph256):
    """a * (a - s (M + kappa)) exceeds 2^31 for |a| = 50 000: the kernel instance with 64-bit group arithmetic runs."""
    base = models.subsampling_model(graph256, 7.0)
    n = base.num_variables
    rng = np.random.default_rng(3)
    grp = np.where(np.arange(n) % 3 == 0, 0, np.where(np.arange(n) % 3 == 1, 1, -1)).astype(np.int32)
    coef = rng.choice(np.array([1, -3, 50000, -47000], dtype=np.int32), size=n).astype(np.int32)
    coef[grp < 0] = 0
    groups = models.Groups(grp, coef, np.array([1e-9, 2.5e-10]), np.array([12345, -777], dtype=np.int64))
    m = models.LoweredModel(base.h, base.starts, base.ends, base.weights, 0.0, base.labels, groups=groups)
    ctx = _replay_ctx()
    try:
        _check(ctx, m, 100, 80, 4, beta_range=(0.02, 6.0))
    finally:
        ctx.close()


def test_several_sweeps_per_beta(built, graph256):
    ctx = _replay_ctx(permille=80)
    try:
        _check(ctx, models.subsampling_model(graph256, 7.0), 64, 30, 12, spb=3)
    finally:
        ctx.close()
