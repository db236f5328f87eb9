"""Config 5 probe (development aid): dense Gaussian-affinity 4-way model, 4096 cells x 4 = 16 384 variables, on the tensor-core
kernel (QA_MODE_THROUGHPUT) and, for comparison, on the bit-exact eager kernel (reference mode) at a smaller read count."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=4096)
ap.add_argument("--K", type=int, default=4)
ap.add_argument("--reads", type=int, default=18944)
ap.add_argument("--sweeps", type=int, default=20)
ap.add_argument("--ref-reads", type=int, default=2368)
ap.add_argument("--ref-sweeps", type=int, default=4)
a = ap.parse_args()

t0 = time.time()
X, _ = snn.gaussian_mixture_embedding(a.cells, dim=15, centres=a.K, sep=4.0, seed=2)
m = models.dense_kway_model(snn.gaussian_affinity(X, k=10), a.K, 0.05)
n = m.num_variables
hot, cold = schedule.default_ising_beta_range(m.h, m.starts, m.ends, m.weights, None)
print(f"n {n} couplers {m.num_couplers} beta ({hot:.4g}, {cold:.4g}) host build {time.time() - t0:.1f} s", flush=True)
with Context(0) as ctx:
    gm = IsingModel(ctx, m.h, m.starts, m.ends, m.weights)
    t1 = time.time()
    ok = gm.enable_dense(a.K)
    print("dense form:", ok, f"{time.time() - t1:.3f} s", flush=True)
    cold = 1000.0 * hot       # (the default cold end comes from the smallest non-zero coupling: ~1e7 on a Gaussian affinity)
    for reads, sweeps, mode in ((a.reads, a.sweeps, _lib.QA_MODE_THROUGHPUT), (2 * a.reads, a.sweeps, _lib.QA_MODE_THROUGHPUT),
                                (a.ref_reads, a.ref_sweeps, _lib.QA_MODE_REFERENCE)):
        if reads <= 0:
            continue
        betas, spb = schedule.make_beta_schedule((hot, cold), sweeps, 1, "geometric")
        states = schedule.random_spin_states(reads, n, 1)
        e, st, done = gm.sample(states, betas, spb, schedule.per_read_seeds(1, reads), mode=mode)
        flops = 2.0 * a.cells * st.attempts        # 2 * n_cells flops per attempt (dense field contraction)
        print(f"mode {mode} kernel {ctx.last_kernel} reads {reads} sweeps {sweeps}: anneal {st.ms_anneal:.1f} ms "
              f"attempts/s {st.attempts / st.ms_anneal * 1e3:.3e} acc {st.accepted / st.attempts:.4f} "
              f"TFLOP/s {flops / st.ms_anneal / 1e9:.2f} best {e.min() + m.offset:.4f}", flush=True)
    gm.close()
