#!/usr/bin/env python
"""Generate tests/golden/*.npz|json from the reference itself (run in the build container only; /root/reference does
not exist on the GPU box, the committed fixtures do).

1. graphs.npz   the six distinct SNN graphs of /root/reference/R/benchmarks/graph_*.gexf as index arrays, in the node
                and edge order ``nx.read_gexf`` yields (create_graphs.py:5-8).
2. q_*.npz      the Q dict / DQM / CQM coefficients built by the reference's OWN functions
                (Python_Functions/{BQM,DQM,CQM}_clustering.py, QA_subsampling.py), captured by importing those modules
                with the D-Wave packages stubbed out (they are not installable offline) and a sampler stub that records
                the model it is handed and aborts the call.  No reference source is copied.
3. known_answers.json   energies computed from the captured Q with math.fsum.
"""
from __future__ import annotations

import contextlib
import io
import json
import math
import os
import sys
import types
from pathlib import Path

import networkx as nx
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
OUT = Path(os.environ["QA_GOLDEN_OUT"]) if os.environ.get("QA_GOLDEN_OUT") else ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))

GRAPHS = ["noisy_circles", "noisy_moons", "varied", "aniso", "blobs", "no_structure"]


class Captured(Exception):
    def __init__(self, payload):
        self.payload = payload


class _CaptureSampler:
    def __init__(self, *a, **k):
        pass

    def sample_qubo(self, Q, **kw):
        raise Captured({"Q": dict(Q), "kwargs": kw})

    def sample(self, bqm, **kw):
        raise Captured({"bqm": bqm, "kwargs": kw})

    def sample_dqm(self, dqm, **kw):
        raise Captured({"dqm": dqm, "kwargs": kw})

    def sample_cqm(self, cqm, **kw):
        raise Captured({"cqm": cqm, "kwargs": kw})


class _RecBQM:
    """Records BinaryQuadraticModel.from_qubo + add_linear_inequality_constraint (BQM_clustering.py:371-380)."""

    def __init__(self):
        self.Q = None
        self.constraint = None

    @classmethod
    def from_qubo(cls, Q, offset=0.0):
        b = cls()
        b.Q = dict(Q)
        return b

    def add_linear_inequality_constraint(self, terms, **kw):
        self.constraint = {"terms": list(terms), **kw}
        return []


class _RecDQM:
    """Records the DiscreteQuadraticModel calls of DQM_clustering.py:29-43 with dimod's set (assign) semantics."""

    def __init__(self):
        self.cases = {}
        self.linear = {}
        self.quadratic = {}

    def add_variable(self, num_cases, label=None):
        self.cases[label] = num_cases
        self.linear[label] = [0.0] * num_cases
        return label

    def set_linear(self, v, biases):
        self.linear[v] = [float(b) for b in biases]

    def set_quadratic(self, u, v, biases):
        key = (u, v) if (v, u) not in self.quadratic else (v, u)
        blk = self.quadratic.setdefault(key, {})
        for (cu, cv), b in biases.items():
            blk[(cu, cv) if key == (u, v) else (cv, cu)] = float(b)


class _Sym:
    """Dumb stand-in for dimod's symbolic binary expressions (dimod.Binary and the QuadraticModel arithmetic that
    CQM_clustering.py:30-48 uses): linear / quadratic coefficient dicts in insertion order.  Deliberately independent of the
    product's cqm.py, so that the CQM fixture is not checked against the code that produced it."""

    def __init__(self, linear=None, quadratic=None, offset=0.0):
        self.linear = dict(linear or {})
        self.quadratic = dict(quadratic or {})
        self.offset = float(offset)

    @staticmethod
    def _of(x):
        return x if isinstance(x, _Sym) else _Sym(offset=float(x))

    def __add__(self, other):
        o = _Sym._of(other)
        r = _Sym(self.linear, self.quadratic, self.offset + o.offset)
        for k, b in o.linear.items():
            r.linear[k] = r.linear.get(k, 0.0) + b
        for (a, c), b in o.quadratic.items():
            key = (a, c) if (a, c) in r.quadratic or (c, a) not in r.quadratic else (c, a)
            r.quadratic[key] = r.quadratic.get(key, 0.0) + b
        return r

    __radd__ = __add__

    def __neg__(self):
        return _Sym({k: -b for k, b in self.linear.items()}, {k: -b for k, b in self.quadratic.items()}, -self.offset)

    def __sub__(self, other):
        return self + (-_Sym._of(other))

    def __mul__(self, other):
        if not isinstance(other, _Sym):
            c = float(other)
            return _Sym({k: c * b for k, b in self.linear.items()}, {k: c * b for k, b in self.quadratic.items()}, c * self.offset)
        if self.quadratic or other.quadratic:
            raise ValueError("degree > 2")
        r = _Sym(offset=self.offset * other.offset)
        for k, b in self.linear.items():
            if other.offset:
                r.linear[k] = r.linear.get(k, 0.0) + b * other.offset
        for k, b in other.linear.items():
            if self.offset:
                r.linear[k] = r.linear.get(k, 0.0) + b * self.offset
        for ka, ba in self.linear.items():
            for kb, bb in other.linear.items():
                if ka == kb:      # x * x = x for a binary variable
                    r.linear[ka] = r.linear.get(ka, 0.0) + ba * bb
                else:
                    key = (ka, kb) if (kb, ka) not in r.quadratic else (kb, ka)
                    r.quadratic[key] = r.quadratic.get(key, 0.0) + ba * bb
        return r

    __rmul__ = __mul__

    def __ge__(self, rhs):
        return ("ge", self, float(rhs))

    def __le__(self, rhs):
        return ("le", self, float(rhs))

    def __eq__(self, rhs):   # noqa: PLW1641  (recorder objects are never hashed)
        return ("eq", self, float(rhs))

    __hash__ = None


def _sym_binary(label):
    return _Sym({label: 1.0})


class _RecCQM:
    """Records ConstrainedQuadraticModel calls: add_discrete / set_objective / add_constraint (CQM_clustering.py:30-48)."""

    def __init__(self):
        self.discrete = {}
        self.objective = None
        self.constraints = {}

    def add_discrete(self, variables, label=None):
        self.discrete[label] = list(variables)
        return label

    def set_objective(self, expr):
        self.objective = _Sym._of(expr)

    def add_constraint(self, comparison, label=None):
        sense, lhs, rhs = comparison
        self.constraints[label] = {"sense": {"ge": ">=", "le": "<=", "eq": "=="}[sense], "lhs": lhs, "rhs": rhs}
        return label


def install_stubs():

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    dimod = mod("dimod", BinaryQuadraticModel=_RecBQM, DiscreteQuadraticModel=_RecDQM,
                ConstrainedQuadraticModel=_RecCQM, Binary=_sym_binary)
    mod("hybrid", KerberosSampler=_CaptureSampler)
    dw = mod("dwave")
    dw.inspector = mod("dwave.inspector", show=lambda *a, **k: None)
    mod("dwave_networkx")
    dw.system = mod("dwave.system", LeapHybridSampler=_CaptureSampler, LeapHybridDQMSampler=_CaptureSampler,
                    LeapHybridCQMSampler=_CaptureSampler, DWaveSampler=_CaptureSampler)
    mod("dwave.system.samplers", DWaveSampler=_CaptureSampler)
    mod("dwave.system.composites", EmbeddingComposite=_CaptureSampler, LazyFixedEmbeddingComposite=_CaptureSampler,
        FixedEmbeddingComposite=_CaptureSampler)
    mod("dwave.cloud")
    mod("dwave.cloud.client", Client=object)
    return dimod


def capture(fn, *args, **kw):
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            fn(*args, **kw)
    except Captured as c:
        return c.payload
    raise RuntimeError("the reference function returned without reaching its sampler call")


def q_arrays(Q, pos):
    keys = list(Q.keys())
    return (np.array([pos[a] for a, _ in keys], dtype=np.int32), np.array([pos[b] for _, b in keys], dtype=np.int32),
            np.array([Q[k] for k in keys], dtype=np.float64))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    install_stubs()
    sys.path.insert(0, str(REF))
    from Python_Functions import BQM_clustering, CQM_clustering, DQM_clustering, QA_subsampling  # the real reference code

    graphs = {}
    known = {}
    for name in GRAPHS:
        G = nx.read_gexf(REF / "R" / "benchmarks" / f"graph_{name}.gexf")
        nodes = list(G.nodes)
        pos = {v: i for i, v in enumerate(nodes)}
        eu = np.array([pos[u] for u, v in G.edges], dtype=np.int32)
        ev = np.array([pos[v] for u, v in G.edges], dtype=np.int32)
        w = np.array([G[u][v]["weight"] for u, v in G.edges], dtype=np.float64)
        graphs[f"{name}_labels"] = np.array([int(v) for v in nodes], dtype=np.int32)
        graphs[f"{name}_eu"], graphs[f"{name}_ev"], graphs[f"{name}_w"] = eu, ev, w

        dirs = {"name": name, "embedding": "/nonexistent", "embedding_pru": "/nonexistent"}
        out = {}
        # clustering_bqm (main.py:149: gamma_factor=0.05), hybrid branch -> sampler.sample_qubo(Q, label=...)
        p = capture(BQM_clustering.clustering_bqm, G, 1, dirs, "hybrid", 0.05, 0, "once", 40, 2, 20)
        Q = p["Q"]
        out["bqm_i"], out["bqm_j"], out["bqm_q"] = q_arrays(Q, pos)
        # known-answer energies: math.fsum over Q in insertion order
        def energy(x):
            return math.fsum(q * x[a] * x[b] for (a, b), q in Q.items())
        n = len(nodes)
        known[name] = {
            "n": n, "W": float(G.size(weight="weight")), "gamma": 0.05 * float(G.size(weight="weight")) / n,
            "E_parity": energy({v: int(v) % 2 for v in nodes}),
            "E_first_half": energy({v: int(int(v) < n // 2) for v in nodes}),
            "lower_bound": -(0.05 * float(G.size(weight="weight")) / n) * n * n / 4,
        }
        comps = sorted(nx.connected_components(G), key=len, reverse=True)
        if len(comps) > 1:
            big = comps[0]
            known[name]["E_largest_component"] = energy({v: int(v in big) for v in nodes})
            known[name]["component_sizes"] = [len(c) for c in comps]
        # clustering_bqm_2 (main.py:154: gamma_factor=0.01, k=1)
        p = capture(BQM_clustering.clustering_bqm_2, G, 1, dirs, "hybrid", 0.01, 0, "once", 40, 1, 20)
        out["bqm2_i"], out["bqm2_j"], out["bqm2_q"] = q_arrays(p["Q"], pos)
        # clustering_bqm_3: BQM.from_qubo(Q) + add_linear_inequality_constraint, then KerberosSampler().sample(bqm)
        p = capture(BQM_clustering.clustering_bqm_3, G, 1, dirs, "hybrid", 0.05, 0, "once", 40)
        out["bqm3_i"], out["bqm3_j"], out["bqm3_q"] = q_arrays(p["bqm"].Q, pos)
        con = p["bqm"].constraint
        out["bqm3_constraint"] = np.array([con["lb"], con["ub"], con["lagrange_multiplier"]], dtype=np.float64)
        out["bqm3_terms"] = np.array([pos[v] for v, _ in con["terms"]], dtype=np.int32)
        # graph_subsampling (main.py:127: gamma=7)
        p = capture(QA_subsampling.graph_subsampling, G, 7, "hybrid")
        out["sub_i"], out["sub_j"], out["sub_q"] = q_arrays(p["Q"], pos)
        # clustering_dqm (main.py:135: 3 clusters, gamma=0.005)
        p = capture(DQM_clustering.clustering_dqm, G, 3, 0.005)
        d = p["dqm"]
        out["dqm_linear"] = np.array([d.linear[v] for v in nodes], dtype=np.float64)
        keys = list(d.quadratic.keys())
        out["dqm_u"] = np.array([pos[a] for a, _ in keys], dtype=np.int32)
        out["dqm_v"] = np.array([pos[b] for _, b in keys], dtype=np.int32)
        out["dqm_diag"] = np.array([[d.quadratic[k].get((c, c), np.nan) for c in range(3)] for k in keys], dtype=np.float64)
        out["dqm_offdiag_count"] = np.array([sum(1 for k in keys for (a, b) in d.quadratic[k] if a != b)], dtype=np.int64)
        np.savez_compressed(OUT / f"q_{name}.npz", **out)
        print(name, "n", n, "m", len(w), "Q entries", len(Q), flush=True)

    # clustering_cqm on one fixture (symbolic sum() is quadratic in the number of terms): 3 clusters (main.py:139)
    name = "noisy_circles"
    G = nx.read_gexf(REF / "R" / "benchmarks" / f"graph_{name}.gexf")
    nodes = list(G.nodes)
    p = capture(CQM_clustering.clustering_cqm, G, 3)
    c = p["cqm"]
    var_pos = {f"v_{v},{k}": i * 3 + k for i, v in enumerate(nodes) for k in range(3)}
    lin = np.zeros(len(var_pos))
    for v, b in c.objective.linear.items():
        lin[var_pos[v]] = b
    qk = list(c.objective.quadratic.keys())
    cq = {
        "lin": lin,
        "qu": np.array([var_pos[a] for a, _ in qk], dtype=np.int32),
        "qv": np.array([var_pos[b] for _, b in qk], dtype=np.int32),
        "qq": np.array([c.objective.quadratic[k] for k in qk], dtype=np.float64),
        "offset": np.array([c.objective.offset]),
        "discrete": np.array([[var_pos[v] for v in grp] for grp in c.discrete.values()], dtype=np.int32),
        "size_rhs": np.array([con["rhs"] - con["lhs"].offset for con in c.constraints.values()], dtype=np.float64),
        "size_vars": np.array([[var_pos[v] for v in con["lhs"].linear] for con in c.constraints.values()], dtype=np.int32),
        "size_sense_ge": np.array([con["sense"] == ">=" for con in c.constraints.values()]),
    }
    np.savez_compressed(OUT / "cqm_noisy_circles.npz", **cq)
    # clustering_cqm_2 with subindex attrs from disconnected_components is only valid on a connected graph: noisy_moons
    np.savez_compressed(OUT / "graphs.npz", **graphs)
    (OUT / "known_answers.json").write_text(json.dumps(known, indent=1))
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
