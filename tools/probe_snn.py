"""SNN builder probe (development aid): device build times against the host recipe for config 3's graph (16 384 cells, k = 5) and
config 4's inputs (512 point sets of 1000 cells, k = 10, dim = 30, one call)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scrna_seq_qannealing_clustering_b200 import snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context  # noqa: E402

with Context(0) as ctx:
    X, _ = snn.gaussian_mixture_embedding(16384, dim=15, centres=8, seed=0)
    ctx.build_snn(X[:512], k=5).close()
    for rep in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        g = ctx.build_snn(X, k=5, prune=1.0 / 15.0, max_degree=15)
        dt = time.perf_counter() - t0
        m = g.num_edges()
        g.close()
        print(f"config 3 graph: 16384 cells, dim 15, k 5, trim 15: device {dt * 1e3:.1f} ms (host buffers in), {m} edges", flush=True)
    t0 = time.perf_counter()
    host = snn.snn_graph(X, 5, 1.0 / 15.0, 15)
    print(f"  host recipe (snn.py): {time.perf_counter() - t0:.2f} s, {len(host[1])} edges", flush=True)
    X4, _ = snn.gaussian_mixture_embedding(512000, dim=30, centres=8, seed=0)
    off = np.arange(0, 512001, 1000)
    for rep in range(2):
        ctx.synchronize()
        t0 = time.perf_counter()
        g = ctx.build_snn(X4, k=10, prune=1.0 / 15.0, max_degree=15, offsets=off)
        dt = time.perf_counter() - t0
        m = g.num_edges()
        g.close()
        print(f"config 4 inputs: 512 x 1000 cells, dim 30, k 10, trim 15, one call: device {dt * 1e3:.1f} ms, {m} edges", flush=True)
    t0 = time.perf_counter()
    snn.snn_graph(X4[:1000], 10, 1.0 / 15.0, 15)
    print(f"  host recipe, ONE of the 512 graphs: {time.perf_counter() - t0:.3f} s", flush=True)
