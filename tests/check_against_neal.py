"""Close the loop oracle == dwave-neal wherever the real library is importable (it is NOT in this image: no network, no wheel).

The CPU oracle (oracle/cpu_sa_ref.cpp) restates neal's `cpu_sa.cpp` from the published algorithm; every parity claim of this
repo is GPU == oracle.  This script checks the remaining link on a machine that has `dimod` and `dwave-neal` (or
`dwave-samplers`) installed:

    python tests/check_against_neal.py            # exits 0 when every case matches bit for bit, 1 otherwise, 2 if neal is absent

For each case it builds a spin model, lets dimod produce neal's vectors (`to_numpy_vectors`, the order neal itself uses), runs

    neal.SimulatedAnnealingSampler().sample(bqm, num_reads=R, num_sweeps=S, beta_range=..., beta_schedule_type='geometric',
                                            seed=seed, initial_states=..., initial_states_generator='none')

and the oracle in stream-seeding mode (one xorshift128+ stream over all reads == what neal does with `seed`), and compares
final states bytewise and energies to the last bit.  It also prints both wall times (the CPU baseline of BASELINE.md).
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def _import_neal():
    try:
        import dimod  # noqa: F401
    except Exception as e:  # pragma: no cover - depends on the machine
        return None, None, f"dimod not importable: {e}"
    try:
        from neal import SimulatedAnnealingSampler
    except Exception:
        try:
            from dwave.samplers import SimulatedAnnealingSampler
        except Exception as e:  # pragma: no cover
            return None, None, f"neither neal nor dwave.samplers importable: {e}"
    import dimod
    return dimod, SimulatedAnnealingSampler, None


def cases():
    from scrna_seq_qannealing_clustering_b200 import models, snn
    g = snn.synthetic_snn(256, k=5, seed=3)[0]
    yield "subsampling 256 cells", models.subsampling_model(g, 7.0), (0.05, 8.0)
    yield "cut + linear 256 cells", models.cut_linear_model(g, 0.01, 1.0), (0.05, 8.0)
    yield "cut + balance K_256 (materialised)", models.cut_balance_model(g, 0.05, structured=False), (0.01, 4.0)
    yield "4-way DQM 256 cells (materialised)", models.dqm_model(g, 4, 0.005, semantics="intended", structured=False), (0.02, 6.0)
    rng = np.random.default_rng(0)
    n = 70
    pairs = [(u, v) for u in range(60) for v in range(u) if rng.random() < 0.2]
    starts = np.array([p[0] for p in pairs], dtype=np.int32)
    ends = np.array([p[1] for p in pairs], dtype=np.int32)
    yield "random 70 variables, isolated ones", models.LoweredModel(rng.normal(size=n), starts, ends, rng.normal(size=len(pairs)),
                                                                      0.0, list(range(n))), (0.1, 5.0)


def main() -> int:
    dimod, Sampler, why = _import_neal()
    if dimod is None:
        print(f"SKIP: {why}")
        return 2
    from oracle import oracle
    from scrna_seq_qannealing_clustering_b200 import schedule
    sampler = Sampler()
    R, sweeps, seed = 24, 200, 1234
    bad = 0
    for name, model, beta_range in cases():
        n = model.num_variables
        h = {i: float(model.h[i]) for i in range(n)}
        J = {(int(a), int(b)): float(w) for a, b, w in zip(model.starts, model.ends, model.weights)}
        bqm = dimod.BinaryQuadraticModel.from_ising(h, J)
        order = list(range(n))
        ldata, (irow, icol, qdata), _ = bqm.to_numpy_vectors(variable_order=order)   # the vectors neal hands to its C++ core
        init = schedule.random_spin_states(R, n, seed)
        t0 = time.perf_counter()
        ss = sampler.sample(bqm, num_reads=R, num_sweeps=sweeps, beta_range=beta_range, beta_schedule_type="geometric", seed=seed,
                            initial_states=(init.copy(), order), initial_states_generator="none")
        t_neal = time.perf_counter() - t0
        cols = [list(ss.variables).index(v) for v in order]
        neal_states = np.asarray(ss.record.sample)[:, cols].astype(np.int8)
        neal_e = np.asarray(ss.record.energy, dtype=np.float64)
        betas, spb = schedule.make_beta_schedule(beta_range, sweeps, 1, "geometric")
        states = init.copy()
        t0 = time.perf_counter()
        e, _ = oracle.sample_ising(np.asarray(ldata, dtype=np.float64), np.asarray(irow, dtype=np.int32), np.asarray(icol, dtype=np.int32),
                                   np.asarray(qdata, dtype=np.float64), states, betas, spb, np.array([seed], dtype=np.uint64),
                                   seed_mode=1, nthreads=1)
        t_oracle = time.perf_counter() - t0
        same_states = np.array_equal(states, neal_states)
        same_e = np.array_equal((e + bqm.offset).view(np.uint64), neal_e.view(np.uint64))
        close_e = np.allclose(e + bqm.offset, neal_e, rtol=1e-12, atol=1e-9)
        ok = same_states and (same_e or close_e)
        bad += not ok
        print(f"{'PASS' if ok else 'FAIL'}  {name}: states {'identical' if same_states else 'DIFFER'}, energies "
              f"{'bit-identical' if same_e else ('equal to 1e-12' if close_e else 'DIFFER')}; neal {t_neal:.3f} s, oracle {t_oracle:.3f} s "
              f"({n * sweeps * R / t_neal:.3e} vs {n * sweeps * R / t_oracle:.3e} attempts/s, 1 thread)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
