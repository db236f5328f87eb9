"""Quick throughput probe of the annealing kernel on a config-3-like model (development aid, not the bench)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scrna_seq_qannealing_clustering_b200 import models, schedule, snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=16384)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--reads", type=int, default=4736)
ap.add_argument("--sweeps", type=int, default=100)
ap.add_argument("--model", default="cqm")
ap.add_argument("--hot", type=float, default=None)
ap.add_argument("--cold", type=float, default=None)
ap.add_argument("--kernel", type=int, default=1)
ap.add_argument("--mode", type=int, default=0)
a = ap.parse_args()
t = time.time()
graph, _ = snn.synthetic_snn(a.cells, k=5, seed=0)
if a.model == "cqm":
    m = models.cqm_model(graph, a.k, min_size=20)
elif a.model == "cqm_sparse":
    m = models.cqm_model(graph, a.k, min_size=20)
    m.groups = None
elif a.model == "sub":
    m = models.subsampling_model(graph, 7.0)
else:
    m = models.cut_balance_model(graph, 0.05)
groups = m.groups.astuple() if m.groups is not None else None
br = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, groups)
br = (a.hot or br[0], a.cold or min(br[1], 50.0))
print("model", a.model, "n", m.num_variables, "m", m.num_couplers, "beta", br, "build_s", round(time.time() - t, 2), flush=True)
betas, spb = schedule.make_beta_schedule(br, a.sweeps, 1, "geometric")
seeds = schedule.per_read_seeds(1, a.reads)
rng = np.random.default_rng(0)
states = (rng.integers(0, 2, size=(a.reads, m.num_variables), dtype=np.int8) * 2 - 1).astype(np.int8)
with Context(0) as ctx:
    ctx.set_kernel(a.kernel)
    gm = IsingModel(ctx, m.h, m.starts, m.ends, m.weights)
    if groups is not None:
        gm.set_groups(*groups)
    for it in range(2):
        s = states.copy()
        e, st, done = gm.sample(s, betas, spb, seeds, mode=a.mode)
        att = st.attempts
        print(f"iter {it}: anneal {st.ms_anneal:.1f} ms  energy {st.ms_energy:.1f} ms  h2d {st.ms_h2d:.1f} d2h {st.ms_d2h:.1f}  "
              f"attempts/s {att / st.ms_anneal * 1e3:.3e}  acc {st.accepted / att:.4f} cand {st.candidates / att:.4f} "
              f"draws {st.draws / att:.4f} active_chunks {st.active_chunks / max(st.chunks, 1):.4f} "
              f"deg/acc {st.nbr_updates / max(st.accepted, 1):.2f} best {e.min() + m.offset:.4f}", flush=True)
    gm.close()
