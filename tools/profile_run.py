"""One launch of an annealing kernel at a benched shape, for ncu (development aid; a number printed under ncu is never a bench value).

    --workload c3   config 3 model (bench.py's), replay kernel: --reads 75776 --sweeps 50 is the benched launch shape
    --workload c5   config 5 model (4096 cells x 4, dense), tensor-core kernel
Initial states are drawn on the device (torch) so that an application-replay pass costs seconds, not a 10 GB host fill."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3", choices=["c3", "c5"])
ap.add_argument("--reads", type=int, default=1184)
ap.add_argument("--sweeps", type=int, default=20)
ap.add_argument("--cells", type=int, default=0)
ap.add_argument("--clusters", type=int, default=0)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--size-penalty", type=float, default=None)
a = ap.parse_args()
a.seed = 1234
import torch  # noqa: E402

dev = torch.device("cuda", 0)
if a.workload == "c3":
    a.cells = a.cells or 16384
    a.clusters = a.clusters or 8
    model, beta_range, betas, spb = bench.build_workload(a)
    mode = _lib.QA_MODE_REFERENCE
else:
    a.cells = a.cells or 4096
    a.clusters = a.clusters or 4
    X, _ = snn.gaussian_mixture_embedding(a.cells, dim=15, centres=a.clusters, sep=4.0, seed=2)
    model = models.dense_kway_model(snn.gaussian_affinity(X, k=10), a.clusters, 0.05)
    beta_range = schedule.default_ising_beta_range(model.h, model.starts, model.ends, model.weights, None)
    betas, spb = schedule.make_beta_schedule(beta_range, a.sweeps, 1, "geometric")
    mode = _lib.QA_MODE_THROUGHPUT
n = model.num_variables
seeds = torch.from_numpy(schedule.per_read_seeds(a.seed, a.reads).view(np.int64)).to(dev)
g = torch.Generator(device=dev)
g.manual_seed(0)
states = torch.randint(0, 2, (a.reads, n), dtype=torch.int8, device=dev, generator=g)
states.mul_(2).sub_(1)
energies = torch.empty(a.reads, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
with Context(0) as ctx:
    if a.kernel:
        ctx.set_kernel(a.kernel)
    gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
    if model.groups is not None:
        gm.set_groups(*model.groups.astuple())
    if a.workload == "c5":
        assert gm.enable_dense(a.clusters)
    e, st, done = gm.sample(states, torch.from_numpy(betas).to(dev), spb, seeds, mode=mode, energies=energies)
    print("kernel", ctx.last_kernel, "anneal_ms", st.ms_anneal, "attempts/s %.3e" % (st.attempts / st.ms_anneal * 1e3), "acc",
          st.accepted / st.attempts, "best", float(energies.min().item()) + model.offset)
    gm.close()
