"""Replay-kernel experiments on the bench workload (development aid): several hand-over thresholds / sweep counts in one process."""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from scrna_seq_qannealing_clustering_b200 import schedule  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=75776)
ap.add_argument("--cells", type=int, default=16384)
ap.add_argument("--clusters", type=int, default=8)
ap.add_argument("--kernel", type=int, default=4)
ap.add_argument("--permille", default="20")
ap.add_argument("--sweeps", default="50")
ap.add_argument("--warps", default="0")
ap.add_argument("--prefix", type=int, default=0, help="anneal only the first PREFIX betas of the schedule (0: all)")
a = ap.parse_args()
a.seed = 1234
rng = np.random.default_rng(0)
states0 = None
for sweeps in [int(x) for x in a.sweeps.split(",")]:
    a.sweeps_n = sweeps
    ns = argparse.Namespace(cells=a.cells, clusters=a.clusters, sweeps=sweeps, seed=a.seed)
    model, beta_range, betas, spb = bench.build_workload(ns)
    if a.prefix:
        betas = betas[:a.prefix]
    if states0 is None:
        states0 = (rng.integers(0, 2, size=(a.reads, model.num_variables), dtype=np.int8) * 2 - 1).astype(np.int8)
    seeds = schedule.per_read_seeds(a.seed, a.reads)
    for warps in a.warps.split(","):
        for pm in a.permille.split(","):
            os.environ["QA_REPLAY_SWITCH_PERMILLE"] = pm
            os.environ["QA_REPLAY_WARPS"] = warps
            with Context(0) as ctx:
                ctx.set_kernel(a.kernel)
                gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
                gm.set_groups(*model.groups.astuple())
                states = states0.copy()
                e, st, done = gm.sample(states, betas, spb, seeds)
                print(f"sweeps {len(betas)} permille {pm} warps {warps} kernel {ctx.last_kernel}: anneal_ms {st.ms_anneal:.1f} "
                      f"attempts/s {st.attempts / st.ms_anneal * 1e3:.3e} acc {st.accepted / st.attempts:.4f} best {e.min() + model.offset:.3f}",
                      flush=True)
                gm.close()
