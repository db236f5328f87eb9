"""scrna_seq_qannealing_clustering_b200 -- B200-native simulated-annealing sampler behind the dimod Sampler API,
a drop-in for the sampling hot path of michal7kw/scRNA_seq_QAnnealing_Clustering.

Host side is Python over ctypes; all compute is in csrc/libqanneal.so (hand-written sm_100a CUDA, C ABI in
include/qanneal.h).  No CPU fallback.
"""
from .bqm import BINARY, SPIN, BinaryQuadraticModel, Vartype
from .cqm import Binary, ConstrainedQuadraticModel
from .dqm import DiscreteQuadraticModel
from .sampleset import SampleSet
from .sampler import B200SimulatedAnnealingSampler
from .clustering import (clustering_bqm, clustering_bqm_2, clustering_bqm_3, clustering_cqm, clustering_cqm_2,
                         clustering_dqm, graph_subsampling, disconnected_components, recursive_bipartition_batched)
from .graph_io import create_graph, create_graph_csv

# the name a neal user would look for
SimulatedAnnealingSampler = B200SimulatedAnnealingSampler

__all__ = [
    "BINARY", "SPIN", "Vartype", "BinaryQuadraticModel", "DiscreteQuadraticModel", "ConstrainedQuadraticModel", "Binary",
    "SampleSet", "B200SimulatedAnnealingSampler", "SimulatedAnnealingSampler", "clustering_bqm", "clustering_bqm_2",
    "clustering_bqm_3", "clustering_dqm", "clustering_cqm", "clustering_cqm_2", "graph_subsampling",
    "disconnected_components", "recursive_bipartition_batched", "create_graph", "create_graph_csv",
]
