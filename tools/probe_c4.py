"""Config 4 probe: QA_subsampling batch -- many independent 1000-cell sub-graph QUBOs in one launch (development aid)."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scrna_seq_qannealing_clustering_b200 import models, schedule, snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=512)
ap.add_argument("--cells", type=int, default=1000)
ap.add_argument("--reads", type=int, default=100)
ap.add_argument("--sweeps", type=int, default=1000)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--gamma", type=float, default=7.0)
a = ap.parse_args()
t = time.time()
graphs = snn.subsample_problems(a.problems * a.cells, a.problems, a.cells, k=10, dim=30, seed=0)
ms = [models.subsampling_model(g, a.gamma) for g in graphs]
voff = np.cumsum([0] + [m.num_variables for m in ms])
coff = np.cumsum([0] + [m.num_couplers for m in ms])
h = np.concatenate([m.h for m in ms])
st_ = np.concatenate([m.starts for m in ms])
en = np.concatenate([m.ends for m in ms])
w = np.concatenate([m.weights for m in ms])
br = schedule.default_ising_beta_range(ms[0].h, ms[0].starts, ms[0].ends, ms[0].weights)
betas, spb = schedule.make_beta_schedule(br, a.sweeps, 1, "geometric")
seeds = schedule.per_read_seeds(1, a.reads * a.problems)
rng = np.random.default_rng(0)
states = (rng.integers(0, 2, size=int(voff[-1]) * a.reads, dtype=np.int8) * 2 - 1).astype(np.int8)
print("problems", a.problems, "vars", int(voff[-1]), "couplers", int(coff[-1]), "beta", br, "build_s", round(time.time() - t, 1), flush=True)
with Context(0) as ctx:
    ctx.set_kernel(a.kernel)
    for it in range(2):
        s = states.copy()
        e, st, done = ctx.sample_ising_batch(voff, coff, h, st_, en, w, a.reads, s, betas, spb, seeds)
        print(f"iter {it}: anneal {st.ms_anneal:.1f} ms build {st.ms_build:.1f} energy {st.ms_energy:.1f} h2d {st.ms_h2d:.1f} "
              f"attempts/s {st.attempts / st.ms_anneal * 1e3:.3e} acc {st.accepted / st.attempts:.4f} "
              f"cand {st.candidates / st.attempts:.4f}", flush=True)
