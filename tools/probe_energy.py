"""k_energy probe (development aid): neal-order energies of 75 776 random config-3 states, device time of the call."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context, IsingModel  # noqa: E402

import torch  # noqa: E402

a = argparse.Namespace(cells=16384, clusters=8, sweeps=50, seed=1234, size_penalty=None)
model, _, _, _ = bench.build_workload(a)
R, n = 75776, model.num_variables
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
states = torch.randint(0, 2, (R, n), dtype=torch.int8, device=dev, generator=g)
states.mul_(2).sub_(1)
energies = torch.empty(R, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
with Context(0) as ctx:
    gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
    gm.set_groups(*model.groups.astuple())
    for rep in range(3):
        e, be, bi, st = gm.energies(states, energies=energies)
        print(f"k_pack_states + k_energy + k_argmin: {st.ms_energy:.1f} ms (best {be:.3f} at read {bi})", flush=True)
    gm.close()
