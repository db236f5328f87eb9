"""Recursion probe (development aid): the level-batched device driver against the call-by-call recursion of clustering_bqm on one
synthetic SNN graph, same seeds / reads / sweeps (the trees are identical: tests/test_gpu_recursion.py)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from scrna_seq_qannealing_clustering_b200 import clustering, snn  # noqa: E402
from scrna_seq_qannealing_clustering_b200.engine import Context  # noqa: E402
from scrna_seq_qannealing_clustering_b200.sampler import B200SimulatedAnnealingSampler  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
G = snn.to_networkx(snn.synthetic_snn(n, k=5, dim=15, centres=8, seed=0)[0])
kw = dict(terminate_on="min_size", size_limit=n // 16, iter_limit=3)
with Context(0) as ctx:
    smp = B200SimulatedAnnealingSampler(context=ctx)
    for rep in range(2):
        t0 = time.perf_counter()
        labels, levels, energies = clustering.recursive_bipartition_batched(G, 0.05, num_reads=64, num_sweeps=200, seed=5, context=ctx, **kw)
        t1 = time.perf_counter()
        H = G.copy()
        clustering.clustering_bqm(H, 0, {"name": "probe"}, "b200", 0.05, sampler=smp, num_reads=64, num_sweeps=200, seed=5, **kw)
        t2 = time.perf_counter()
        print(f"{n} cells: levels {[len(lv) for lv in levels]} leaves {len(set(labels.values()))}: batched driver {t1 - t0:.2f} s "
              f"({len(levels)} launches), call-by-call recursion {t2 - t1:.2f} s ({sum(len(lv) for lv in levels)} sampler calls)", flush=True)
