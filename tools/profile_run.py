"""One short launch of the annealing kernel on the bench workload, for ncu (development aid)."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from scrna_seq_qannealing_clustering_b200 import schedule  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1184)
ap.add_argument("--sweeps", type=int, default=20)
ap.add_argument("--cells", type=int, default=16384)
ap.add_argument("--clusters", type=int, default=8)
ap.add_argument("--kernel", type=int, default=0)
a = ap.parse_args()
a.seed = 1234
model, beta_range, betas, spb = bench.build_workload(a)
seeds = schedule.per_read_seeds(a.seed, a.reads)
rng = np.random.default_rng(0)
states = (rng.integers(0, 2, size=(a.reads, model.num_variables), dtype=np.int8) * 2 - 1).astype(np.int8)
with Context(0) as ctx:
    if a.kernel and hasattr(ctx, "set_kernel"):
        ctx.set_kernel(a.kernel)
    gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
    gm.set_groups(*model.groups.astuple())
    e, st, done = gm.sample(states, betas, spb, seeds)
    print("anneal_ms", st.ms_anneal, "attempts/s %.3e" % (st.attempts / st.ms_anneal * 1e3), "acc", st.accepted / st.attempts,
          "nbr", st.nbr_updates, "best", e.min() + model.offset)
    gm.close()
