"""Config 2 probe: 4-way DQM (one-hot expansion) on a 2048-cell SNN graph, 10 000 reads (development aid)."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scrna_seq_qannealing_clustering_b200 import models, schedule, snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=2048)
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--reads", type=int, default=10000)
ap.add_argument("--sweeps", type=int, default=200)
ap.add_argument("--kernels", default="0,1,2")
a = ap.parse_args()
g = snn.synthetic_snn(a.cells, k=5, seed=0)[0]
m = models.dqm_model(g, a.k, 0.005, semantics="intended")
br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
betas, spb = schedule.make_beta_schedule(br, a.sweeps, 1, "geometric")
seeds = schedule.per_read_seeds(1, a.reads)
states0 = schedule.random_spin_states(a.reads, m.num_variables, 1)
print("n", m.num_variables, "couplers", m.num_couplers, "beta", br, flush=True)
for k in [int(x) for x in a.kernels.split(",")]:
    with Context(0) as ctx:
        ctx.set_kernel(k)
        gm = IsingModel(ctx, m.h, m.starts, m.ends, m.weights)
        gm.set_groups(*m.groups.astuple())
        for it in range(2):
            s = states0.copy()
            e, st, done = gm.sample(s, betas, spb, seeds)
        print(f"kernel requested {k} ran {ctx.last_kernel}: anneal {st.ms_anneal:.1f} ms attempts/s {st.attempts / st.ms_anneal * 1e3:.3e} "
              f"acc {st.accepted / st.attempts:.4f} best {e.min() + m.offset:.4f}", flush=True)
        gm.close()
