"""Host logic of the level-batched recursion (clustering.recursive_bipartition_batched) on CPU: the batched launch is
replaced by an oracle-backed stand-in with the same signature, so the level bookkeeping, the per-problem seeds / initial
states and the split rule are checked without a GPU (the GPU test runs the same driver through the C ABI)."""
import numpy as np

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import clustering, models, schedule, snn


class OracleBatchContext:
    """`Context.sample_ising_batch` semantics: independent problems, `reads_per_problem` reads each, states updated in place."""

    def __init__(self):
        self.calls = 0

    def sample_ising_batch(self, voff, coff, h, starts, ends, w, rpp, states, betas, spb, seeds):
        self.calls += 1
        P = len(voff) - 1
        e = np.empty(P * rpp)
        off = 0
        for p in range(P):
            n = int(voff[p + 1] - voff[p])
            sl = slice(int(coff[p]), int(coff[p + 1]))
            st = states[off:off + rpp * n].reshape(rpp, n).copy()
            ee, _ = oracle.sample_ising(h[voff[p]:voff[p + 1]], starts[sl], ends[sl], w[sl], st, betas, spb, seeds[p * rpp:(p + 1) * rpp])
            states[off:off + rpp * n] = st.ravel()
            e[p * rpp:(p + 1) * rpp] = ee
            off += rpp * n
        return e, None, rpp


def test_four_blobs_are_recovered_with_one_launch_per_level():
    X, truth = snn.gaussian_mixture_embedding(240, dim=8, centres=4, sep=9.0, seed=1)
    G = snn.to_networkx(snn.snn_graph(X, k=10))
    ctx = OracleBatchContext()
    labels, levels, energies = clustering.recursive_bipartition_batched(G, gamma_factor=0.05, k=8.0, size_limit=30, iter_limit=1,
                                                                        num_reads=96, num_sweeps=300, seed=3, context=ctx)
    assert ctx.calls == 2 and [len(lv) for lv in levels] == [1, 2]
    assert set(labels) == set(G.nodes) and len(set(labels.values())) == 4
    node_truth = {str(i): int(t) for i, t in enumerate(truth)}
    for leaf in set(labels.values()):
        assert len({node_truth[n] for n, l in labels.items() if l == leaf}) == 1      # every leaf is one planted blob
    # level 0 is exactly one neal-style call on the whole graph with the documented seeds / initial states
    m = models.cut_balance_model(G, 0.05, k=8.0, structured=False)
    br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights)
    betas, spb = schedule.make_beta_schedule(br, 300, 1, "geometric")
    st = schedule.random_spin_states(96, m.num_variables, 3)
    e, _ = oracle.sample_ising(m.h, m.starts, m.ends, m.weights, st, betas, spb, schedule.per_read_seeds(3, 96))
    assert energies[0][0] == float(e.min() + m.offset)


def test_size_limit_and_iteration_limit_stop_the_recursion():
    X, _ = snn.gaussian_mixture_embedding(120, dim=6, centres=2, sep=8.0, seed=2)
    G = snn.to_networkx(snn.snn_graph(X, k=8))
    ctx = OracleBatchContext()
    labels, levels, _ = clustering.recursive_bipartition_batched(G, 0.05, size_limit=1000, iter_limit=5, num_reads=32, num_sweeps=100,
                                                                 seed=1, context=ctx)
    assert ctx.calls == 1 and len(set(labels.values())) <= 2            # halves are below size_limit: no second level
    ctx = OracleBatchContext()
    labels, levels, _ = clustering.recursive_bipartition_batched(G, 0.05, size_limit=5, iter_limit=0, num_reads=32, num_sweeps=100,
                                                                 seed=1, context=ctx)
    assert ctx.calls == 1                                               # iteration limit reached at the first level
