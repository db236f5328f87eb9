"""The CPU oracle (oracle/cpu_sa_ref.cpp) against independent known answers -- no GPU.

The real dwave-neal is absent (parity unpinned against it); what pins the restatement here:
  * xorshift128+ outputs against an independent pure-Python implementation of neal's FASTRAND,
  * energies against brute-force enumeration and numpy,
  * a pure-Python restatement of the Metropolis sweep (small cases) -- states, energies and RNG consumption equal,
  * SA reaching the true minimum on enumerable instances and the PROVABLE ground state of graph_noisy_circles.
"""
import itertools
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import models, schedule, snn

GOLD = Path(__file__).parent / "golden"
M64 = (1 << 64) - 1


def py_rng(seed):
    s0, s1 = (seed if seed else M64), 0
    while True:
        x, y = s0, s1
        s0 = y
        x ^= (x << 23) & M64
        s1 = x ^ y ^ (x >> 17) ^ (y >> 26)
        yield (s1 + y) & M64


def py_anneal(h, starts, ends, w, state, betas, spb, seed):
    """Pure-Python restatement of neal simulated_annealing_run + get_state_energy."""
    n = len(h)
    adj = [[] for _ in range(n)]
    for u, v, x in zip(starts, ends, w):
        adj[u].append((v, x))
        adj[v].append((u, x))
    rng = py_rng(seed)
    state = [int(s) for s in state]
    dE = []
    for v in range(n):
        e = h[v]
        for j, x in adj[v]:
            e += state[j] * x
        dE.append(-2 * state[v] * e)
    draws = 0
    for beta in betas:
        for _ in range(spb):
            thr = 44.36142 / beta
            for v in range(n):
                if dE[v] >= thr:
                    continue
                flip = False
                if dE[v] <= 0.0:
                    flip = True
                else:
                    r = next(rng)
                    draws += 1
                    if math.exp(-dE[v] * beta) * 18446744073709551616.0 > float(r):
                        flip = True
                if flip:
                    mult = 4 * state[v]
                    for j, x in adj[v]:
                        dE[j] += mult * x * state[j]
                    state[v] *= -1
                    dE[v] *= -1
    e = 0.0
    for v in range(n):
        e += state[v] * h[v]
    for u, v, x in zip(starts, ends, w):
        e += state[u] * x * state[v]
    return state, e, draws


def test_rng_known_answers():
    for seed in (0, 1, 1234, 2 ** 32 - 1, 2 ** 63 + 5):
        g = py_rng(seed)
        want = np.array([next(g) for _ in range(64)], dtype=np.uint64)
        assert np.array_equal(oracle.rng_stream(seed, 64), want)
    # seed 0 is remapped to 2^64 - 1 (neal: rng_state[0] = seed ? seed : RANDMAX)
    assert np.array_equal(oracle.rng_stream(0, 8), oracle.rng_stream(M64, 8))
    # first output of seed 1234: s1 = x ^ (x >> 17) with x = 1234 ^ (1234 << 23); y = 0
    x = 1234 ^ (1234 << 23)
    assert int(oracle.rng_stream(1234, 1)[0]) == (x ^ (x >> 17)) & M64


@pytest.mark.parametrize("seed", [0, 3, 99])
def test_matches_pure_python_restatement(seed):
    rng = np.random.default_rng(seed)
    n = 23
    pairs = [(u, v) for u in range(n) for v in range(u) if rng.random() < 0.3]
    rng.shuffle(pairs)
    starts = np.array([p[0] for p in pairs], dtype=np.int32)
    ends = np.array([p[1] for p in pairs], dtype=np.int32)
    w = rng.normal(size=len(pairs))
    h = rng.normal(size=n)
    betas = np.geomspace(0.1, 5.0, 30)
    init = schedule.random_spin_states(3, n, seed)
    seeds = np.array([seed, seed + 1, 0], dtype=np.uint64)
    states = init.copy()
    e, st = oracle.sample_ising(h, starts, ends, w, states, betas, 2, seeds)
    total_draws = 0
    for r in range(3):
        s_py, e_py, d = py_anneal(h.tolist(), starts.tolist(), ends.tolist(), w.tolist(), init[r], betas.tolist(), 2, int(seeds[r]))
        assert states[r].tolist() == s_py
        assert e[r] == e_py
        total_draws += d
    assert st["draws"] == total_draws and st["attempts"] == 3 * n * 60


def test_stream_mode_equals_chained_reads():
    g = snn.synthetic_snn(40, k=4, seed=1)[0]
    m = models.subsampling_model(g, 0.5)
    betas = np.geomspace(0.05, 1.0, 40)  # warm final beta: final states keep depending on the random draws
    init = schedule.random_spin_states(4, m.num_variables, 5)
    a = init.copy()
    e_stream, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, a, betas, 1, np.array([77], dtype=np.uint64), seed_mode=1)
    # read 0 of the stream == a per-read run with the same seed
    b = init[:1].copy()
    e0, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, b, betas, 1, np.array([77], dtype=np.uint64), seed_mode=0)
    assert np.array_equal(a[0], b[0]) and e_stream[0] == e0[0]
    # later reads continue the same RNG stream, so they differ from a re-seeded run
    c = init[1:2].copy()
    oracle.sample_ising(m.h, m.starts, m.ends, m.weights, c, betas, 1, np.array([77], dtype=np.uint64), seed_mode=0)
    assert not np.array_equal(a[1], c[0])


def test_energy_against_brute_force_and_minimum_found():
    g = snn.synthetic_snn(14, k=4, max_degree=None, seed=1)[0]
    for m in (models.cut_balance_model(g, 0.05, structured=False), models.cut_balance_model(g, 0.05, structured=True),
              models.subsampling_model(g, 0.4)):
        n = m.num_variables
        X = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=np.int8)
        groups = m.groups.astuple() if m.groups is not None else None
        e_all = oracle.state_energies(m.h, m.starts, m.ends, m.weights, X, groups=groups) + m.offset
        assert np.allclose(e_all, m.energies(X), rtol=1e-12, atol=1e-10)
        # QUBO-form check for the subsampling model: E = sum (1-w)(x_u x_v - x_u - x_v) + gamma sum x
        if m.meta["builder"] == "subsampling":
            _, eu, ev, w = models.graph_arrays(g)
            xb = (X + 1) // 2
            direct = ((1 - w) * (xb[:, eu] * xb[:, ev] - xb[:, eu] - xb[:, ev])).sum(axis=1) + 0.4 * xb.sum(axis=1)
            assert np.allclose(e_all, direct, rtol=1e-12, atol=1e-10)
        states = schedule.random_spin_states(200, n, 4)
        br = (0.02, 30.0)
        betas, spb = schedule.make_beta_schedule(br, 300, 1, "geometric")
        e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, states, betas, spb, schedule.per_read_seeds(4, 200), groups=groups)
        assert e.min() + m.offset == pytest.approx(e_all.min(), abs=1e-9)
        assert np.allclose(e + m.offset, m.energies(states), rtol=1e-12, atol=1e-9)


def test_structured_and_materialised_runs_take_the_same_decisions():
    """Rank-1 lazy evaluation changes roundings only: on a small model the two forms flip the same spins."""
    g = snn.synthetic_snn(60, k=4, seed=2)[0]
    a = models.cut_balance_model(g, 0.05, structured=True)
    b = models.cut_balance_model(g, 0.05, structured=False)
    betas, spb = schedule.make_beta_schedule((0.02, 20.0), 100, 1, "geometric")
    init = schedule.random_spin_states(50, 60, 6)
    seeds = schedule.per_read_seeds(6, 50)
    sa, sb = init.copy(), init.copy()
    ea, _ = oracle.sample_ising(a.h, a.starts, a.ends, a.weights, sa, betas, spb, seeds, groups=a.groups.astuple())
    eb, _ = oracle.sample_ising(b.h, b.starts, b.ends, b.weights, sb, betas, spb, seeds)
    same = (sa == sb).all(axis=1).mean()
    assert same >= 0.9  # identical unless an accept test is decided inside the rounding difference
    assert np.allclose(np.sort(ea + a.offset)[:5], np.sort(eb + b.offset)[:5], rtol=1e-9)


def test_provable_ground_state_of_noisy_circles_is_reached():
    known = json.loads((GOLD / "known_answers.json").read_text())["noisy_circles"]
    gz = np.load(GOLD / "graphs.npz")
    g = (256, gz["noisy_circles_eu"], gz["noisy_circles_ev"], gz["noisy_circles_w"])
    m = models.cut_balance_model(g, 0.05, structured=True)
    betas, spb = schedule.make_beta_schedule((0.01, 10.0), 400, 1, "geometric")
    states = schedule.random_spin_states(64, 256, 1)
    e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, states, betas, spb, schedule.per_read_seeds(1, 64),
                               groups=m.groups.astuple())
    best = e.min() + m.offset
    assert best >= known["lower_bound"] - 1e-8           # nothing beats the bound -gamma n^2/4
    assert best == pytest.approx(known["lower_bound"], rel=1e-10)  # ... and SA finds the component split


def test_input_validation():
    h = np.zeros(3)
    with pytest.raises(RuntimeError):
        oracle.sample_ising(h, [3], [0], [1.0], np.ones((1, 3), dtype=np.int8), [1.0], 1, [1])
    with pytest.raises(RuntimeError):
        oracle.sample_ising(h, [1], [1], [1.0], np.ones((1, 3), dtype=np.int8), [1.0], 1, [1])
    with pytest.raises(RuntimeError):
        oracle.sample_ising(h, [1], [0], [1.0], np.zeros((1, 3), dtype=np.int8), [1.0], 1, [1])


# ---- the rank-1 GROUP extension, checked by something other than oracle/cpu_sa_ref.cpp ------------------------
def py_anneal_groups_checked(model, state, betas, spb, seed, tol=1e-9):
    """Independent restatement of the structured loop: neal's sweep on the explicit couplers, plus for every variable in a
    group the lazily evaluated flip cost lambda * a * (a - s (M + kappa)) with integer M.  At EVERY attempt the float flip
    cost is also compared with the EXACT flip cost of the materialised energy function
        E(s) = sum h s + sum J s s + sum_g lambda_g / 4 (sum_{v in g} a_v s_v + kappa_g)^2
    evaluated in rational arithmetic (fractions.Fraction), and the decision taken is required to be the decision the exact
    value implies unless the margin is inside `tol` (relative)."""
    from fractions import Fraction as Fr
    h, starts, ends, w = model.h.tolist(), model.starts.tolist(), model.ends.tolist(), model.weights.tolist()
    grp, coef, lam, kap = (np.asarray(a).tolist() for a in model.groups.astuple())
    n = len(h)
    adj = [[] for _ in range(n)]
    for u, v, x in zip(starts, ends, w):
        adj[u].append((v, x))
        adj[v].append((u, x))
    s = [int(x) for x in state]
    M = [0] * len(lam)
    for v in range(n):
        if grp[v] >= 0:
            M[grp[v]] += coef[v] * s[v]
    dE = []
    for v in range(n):
        e = h[v]
        for j, x in adj[v]:
            e += s[j] * x
        dE.append(-2 * s[v] * e)

    def exact_flip_cost(v):
        f = Fr(h[v]) + sum(Fr(x) * s[j] for j, x in adj[v])
        d = -2 * s[v] * f
        g = grp[v]
        if g >= 0:
            t = M[g] + kap[g]
            d += Fr(lam[g]) * (Fr((t - 2 * coef[v] * s[v]) ** 2) - Fr(t ** 2)) / 4
        return d

    rng = py_rng(seed)
    checked = 0
    for beta in betas:
        thr = 44.36142 / beta
        for _ in range(spb):
            for v in range(n):
                d = dE[v]
                g = grp[v]
                if g >= 0:
                    a = coef[v]
                    d = d + lam[g] * float(a * (a - s[v] * (M[g] + kap[g])))
                ex = exact_flip_cost(v)
                scale = max(1.0, abs(float(ex)), max(abs(x) for _, x in adj[v]) if adj[v] else 0.0)
                assert abs(d - float(ex)) <= 1e-9 * scale * max(1, len(adj[v])), (v, d, float(ex))
                checked += 1
                if d >= thr:
                    assert float(ex) >= thr * (1 - tol)
                    continue
                flip = False
                if d <= 0.0:
                    assert float(ex) <= tol * scale
                    flip = True
                elif math.exp(-d * beta) * 18446744073709551616.0 > float(next(rng)):
                    flip = True
                if flip:
                    mult = 4 * s[v]
                    for j, x in adj[v]:
                        dE[j] += mult * x * s[j]
                    if g >= 0:
                        M[g] -= 2 * coef[v] * s[v]
                    s[v] *= -1
                    dE[v] *= -1
    return s, checked


@pytest.mark.parametrize("kind", ["cut_balance", "cqm", "dqm"])
def test_group_extension_against_an_independent_exact_restatement(kind):
    g = snn.synthetic_snn(24 if kind != "cut_balance" else 40, k=4, seed=2)[0]
    if kind == "cut_balance":
        m = models.cut_balance_model(g, 0.05, k=8.0, structured=True)
        br = (0.02, 6.0)
    elif kind == "cqm":
        m = models.cqm_model(g, 3, min_size=4)
        br = (0.02, 4.0)
    else:
        m = models.dqm_model(g, 3, 0.005, semantics="intended")
        br = (0.02, 6.0)
    assert m.groups is not None
    betas = np.geomspace(br[0], br[1], 25)
    R = 2
    init = schedule.random_spin_states(R, m.num_variables, 4)
    seeds = schedule.per_read_seeds(9, R)
    ref = init.copy()
    oracle.sample_ising(m.h, m.starts, m.ends, m.weights, ref, betas, 1, seeds, groups=m.groups.astuple())
    for r in range(R):
        s_py, checked = py_anneal_groups_checked(m, init[r], betas.tolist(), 1, int(seeds[r]))
        assert checked == m.num_variables * len(betas)
        assert ref[r].tolist() == s_py, "the C++ oracle's structured run differs from the independent restatement"
