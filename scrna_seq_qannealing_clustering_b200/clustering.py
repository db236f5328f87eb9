"""Same-signature stand-ins for the reference's clustering functions, with the sampler injected.

    clustering_bqm / _2 / _3     Python_Functions/BQM_clustering.py:25, :206, :353
    clustering_dqm               Python_Functions/DQM_clustering.py:24
    clustering_cqm / _2          Python_Functions/CQM_clustering.py:25, :57
    graph_subsampling            Python_Functions/QA_subsampling.py:24

Each builds the reference's model (models.py), hands it to ``sampler`` (default: B200SimulatedAnnealingSampler) through
the dimod call the reference makes, and consumes the SampleSet the way the reference does (labels on the graph,
recursion rules).  ``solver`` / ``dirs`` / ``chain_strength`` are accepted for signature compatibility; QPU-only
kwargs are forwarded and dropped by the sampler with a warning, as neal would.
Deviations, all deliberate: the reference's recursive calls omit ``chain_strength`` (TypeError on first recursion,
SURVEY.md 3.1) -- here the argument is passed; its ``conf`` rule reads ``record.energy[0]/[3]`` assuming an
energy-sorted record -- here the SampleSet is requested sorted.
"""
from __future__ import annotations

import random
import warnings
from typing import Optional

import numpy as np

from . import models
from .sampler import B200SimulatedAnnealingSampler

_default_sampler = None


def _sampler(sampler):
    global _default_sampler
    if sampler is not None:
        return sampler
    if _default_sampler is None:
        _default_sampler = B200SimulatedAnnealingSampler()
    return _default_sampler


def _sample(sampler, model, quiet_kwargs: dict, sa_kwargs: dict):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)  # label / chain_strength / return_embedding are QPU-only
        return sampler.sample(model, sorted=True, **quiet_kwargs, **sa_kwargs)


def _split(G, response):
    lut = response.first.sample
    S0 = [node for node in G.nodes if not lut[node]]
    S1 = [node for node in G.nodes if lut[node]]
    return S0, S1


def _label(G, nodes, label, value):
    for i in nodes:
        G.nodes[i][label] = value


def _recurse(fn, G, S0, S1, iteration, response, terminate_on, size_limit, iter_limit, color, args, kwargs):
    """Termination rules of BQM_clustering.py:113-203 (min_size / conf / once / iter_limit)."""
    label = "label" + str(iteration)
    go = False
    if terminate_on == "min_size":
        go = len(S0) > size_limit and len(S1) > size_limit and iteration < iter_limit
    elif terminate_on == "conf":
        e = response.record.energy
        conf = abs(e[0] - e[min(3, len(e) - 1)]) if len(e) > 1 else 0.0
        go = len(S0) > size_limit and len(S1) > size_limit and iteration < iter_limit and (conf > 0 or len(e) < 4)
    elif terminate_on == "iter_limit":
        go = iteration < iter_limit and len(S0) > 1 and len(S1) > 1
    _label(G, S0, label, random.randint(0, 100))
    _label(G, S1, label, random.randint(120, 220))
    if terminate_on == "once" or not go:
        return
    fn(G.subgraph(S0), iteration + 1, *args, color=color + 20, terminate_on=terminate_on, size_limit=size_limit,
       iter_limit=iter_limit, **kwargs)
    fn(G.subgraph(S1), iteration + 1, *args, color=color + 20, terminate_on=terminate_on, size_limit=size_limit,
       iter_limit=iter_limit, **kwargs)


def clustering_bqm(G, iteration, dirs, solver, gamma_factor, color=0, terminate_on="once", size_limit=40, iter_limit=2,
                   chain_strength=None, sampler=None, structured=True, **sa_kwargs):
    """2-way cut+balance partition (BQM_clustering.py:25-203); returns the SampleSet of this level."""
    model = models.cut_balance_model(G, gamma_factor, k=8, structured=structured)
    sa_kwargs.setdefault("num_reads", 500)  # BQM_clustering.py:52
    response = _sample(_sampler(sampler), model, {"label": str(dirs.get("name", "")) + "_" + str(solver),
                                                   "chain_strength": chain_strength}, sa_kwargs)
    S0, S1 = _split(G, response)
    _recurse(clustering_bqm, G, S0, S1, iteration, response, terminate_on, size_limit, iter_limit, color,
             (dirs, solver, gamma_factor), dict(chain_strength=chain_strength, sampler=sampler, structured=structured, **sa_kwargs))
    return response


def clustering_bqm_2(G, iteration, dirs, solver, gamma_factor, color=0, terminate_on="once", size_limit=40, k=1,
                     chain_strength=None, sampler=None, iter_limit=2, **sa_kwargs):
    """2-way cut + linear gamma (BQM_clustering.py:206-350)."""
    model = models.cut_linear_model(G, gamma_factor, k)
    sa_kwargs.setdefault("num_reads", 5000)  # BQM_clustering.py:240
    response = _sample(_sampler(sampler), model, {"label": str(dirs.get("name", "")) + "_" + str(solver)}, sa_kwargs)
    S0, S1 = _split(G, response)
    label = "label" + str(iteration)
    _label(G, S0, label, random.randint(0, 100))
    _label(G, S1, label, random.randint(120, 220))
    return response


def clustering_bqm_3(G, iteration, dirs, solver, gamma_factor, color=0, terminate_on="once", size_limit=40, sampler=None,
                     **sa_kwargs):
    """2-way cut + size window as a slack penalty, sampled with ``sampler.sample(bqm)`` (BQM_clustering.py:353-426)."""
    bqm = models.cut_inequality_bqm(G, gamma_factor, size_limit, k=8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        response = _sampler(sampler).sample(bqm, sorted=True, max_iter=1, qpu_reads=100, tabu_timeout=200, **sa_kwargs)
    S0, S1 = _split(G, response)
    label = "label" + str(iteration)
    _label(G, S0, label, random.randint(0, 100))
    _label(G, S1, label, random.randint(120, 220))
    return response


def clustering_dqm(G, num_of_clusters, gamma, sampler=None, penalty=None, semantics="as_written", structured=True,
                   **sa_kwargs):
    """k-way DQM clustering (DQM_clustering.py:24-47): returns a SampleSet of case indices per cell."""
    model = models.dqm_model(G, num_of_clusters, gamma, penalty=penalty, semantics=semantics, structured=structured)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        sampleset = _sampler(sampler).sample_dqm(model, label="DQM - scRAN-seq", **sa_kwargs)
    return sampleset


def clustering_cqm(G, num_of_clusters, sampler=None, onehot_penalty=None, size_penalty=None, min_size=20, **sa_kwargs):
    """k-way CQM clustering with one-hot and minimum-size constraints (CQM_clustering.py:25-55)."""
    model = models.cqm_model(G, num_of_clusters, min_size, onehot_penalty, size_penalty)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return _sampler(sampler).sample_cqm(model, label="CQM - scRAN-seq", **sa_kwargs)


def clustering_cqm_2(G, num_of_clusters, sampler=None, onehot_penalty=None, size_penalty=None, min_size=20, **sa_kwargs):
    """As ``clustering_cqm`` but variables are named by the node attribute ``subindex`` (CQM_clustering.py:57-91)."""
    sub = [G.nodes[i]["subindex"] for i in G.nodes]
    model = models.cqm_model(G, num_of_clusters, min_size, onehot_penalty, size_penalty, subindex=sub)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return _sampler(sampler).sample_cqm(model, label="CQM - scRAN-seq", **sa_kwargs)


def graph_subsampling(G, gamma, solver="hybrid", sampler=None, **sa_kwargs):
    """Pruning QUBO (QA_subsampling.py:24-97): labels nodes ``label1`` in {0, 1} and returns the SampleSet."""
    model = models.subsampling_model(G, gamma)
    sa_kwargs.setdefault("num_reads", 100)  # QA_subsampling.py:37
    response = _sample(_sampler(sampler), model, {"label": "prun_data", "chain_strength": 4}, sa_kwargs)
    S0, S1 = _split(G, response)
    _label(G, S0, "label1", 0)
    _label(G, S1, "label1", 1)
    return response


def disconnected_components(G):
    """``other_tools.disconnected_components`` (other_tools.py:71-86): ``subindex`` / ``valid`` node attributes."""
    import networkx as nx

    comps = list(nx.connected_components(G))
    lengths = [len(c) for c in sorted(comps, key=len, reverse=True)]
    S = [G.subgraph(c).copy() for c in comps]
    for s in S:
        if len(s.nodes()) > 15:
            for subindex, n in enumerate(s.nodes()):
                G.nodes[n]["subindex"] = subindex
                G.nodes[n]["valid"] = 1
        else:
            for n in s.nodes():
                G.nodes[n]["valid"] = 0
    return G, S, lengths
