"""Minimal dimod-compatible ``DiscreteQuadraticModel`` (the calls made at DQM_clustering.py:29-43).

``add_variable(num_cases, label=)``, ``set_linear(v, biases)``, ``set_quadratic(u, v, {(cu, cv): bias})`` with dimod's
*set* (overwrite) semantics, ``energies``; plus ``to_lowered`` = one-hot expansion into binary variables (v, case)
with the penalty  A * (sum_c x_vc - 1)^2  that stands in for LeapHybridDQMSampler's native one-hot handling.
For graphs of benchmark size use ``models.dqm_model`` (vectorised, rank-1 structured) instead of this object.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Mapping, Optional, Tuple

import numpy as np

from .bqm import BinaryQuadraticModel
from .models import LoweredModel, lowered_from_bqm


class DiscreteQuadraticModel:
    def __init__(self):
        self.variables: List[Hashable] = []
        self._index: Dict[Hashable, int] = {}
        self._cases: List[int] = []
        self._linear: List[np.ndarray] = []
        self._quad: Dict[Tuple[int, int], Dict[Tuple[int, int], float]] = {}  # (iu < iv) -> {(cu, cv): bias}

    def add_variable(self, num_cases: int, label: Optional[Hashable] = None):
        if label is None:
            label = len(self.variables)
        if label in self._index:
            raise ValueError(f"variable {label!r} already exists")
        if num_cases <= 0:
            raise ValueError("discrete variables must have at least one case")
        self._index[label] = len(self.variables)
        self.variables.append(label)
        self._cases.append(int(num_cases))
        self._linear.append(np.zeros(int(num_cases)))
        return label

    def num_variables(self) -> int:
        return len(self.variables)

    def num_cases(self, v: Optional[Hashable] = None) -> int:
        return sum(self._cases) if v is None else self._cases[self._index[v]]

    def set_linear(self, v, biases):
        i = self._index[v]
        b = np.asarray(biases, dtype=np.float64)
        if b.shape != (self._cases[i],):
            raise ValueError("one bias per case is required")
        self._linear[i] = b.copy()

    def get_linear(self, v) -> np.ndarray:
        return self._linear[self._index[v]].copy()

    def set_quadratic(self, u, v, biases):
        iu, iv = self._index[u], self._index[v]
        if iu == iv:
            raise ValueError("a variable cannot interact with itself")
        block = self._quad.setdefault((min(iu, iv), max(iu, iv)), {})
        if isinstance(biases, Mapping):
            items = biases.items()
        else:
            arr = np.asarray(biases, dtype=np.float64)
            items = (((cu, cv), arr[cu, cv]) for cu in range(arr.shape[0]) for cv in range(arr.shape[1]))
        for (cu, cv), b in items:
            if not (0 <= cu < self._cases[iu] and 0 <= cv < self._cases[iv]):
                raise ValueError("case index out of range")
            block[(cu, cv) if iu < iv else (cv, cu)] = float(b)

    def get_quadratic(self, u, v) -> Dict[Tuple[int, int], float]:
        iu, iv = self._index[u], self._index[v]
        block = self._quad.get((min(iu, iv), max(iu, iv)), {})
        return dict(block) if iu < iv else {(cv, cu): b for (cu, cv), b in block.items()}

    def energies(self, samples) -> np.ndarray:
        """samples [R][n] of case indices in ``self.variables`` order (or dicts)."""
        if isinstance(samples, Mapping):
            samples = [samples]
        if len(samples) and isinstance(samples[0], Mapping):
            samples = [[s[v] for v in self.variables] for s in samples]
        S = np.atleast_2d(np.asarray(samples, dtype=np.int64))
        e = np.zeros(S.shape[0])
        for i, lin in enumerate(self._linear):
            e += lin[S[:, i]]
        for (iu, iv), block in self._quad.items():
            for (cu, cv), b in block.items():
                e += b * ((S[:, iu] == cu) & (S[:, iv] == cv))
        return e

    def to_bqm(self, penalty: float) -> Tuple[BinaryQuadraticModel, List[Tuple[Hashable, int]]]:
        bqm = BinaryQuadraticModel({}, {}, 0.0, "BINARY")
        labels = []
        for i, v in enumerate(self.variables):
            for c in range(self._cases[i]):
                bqm.add_variable((v, c), float(self._linear[i][c]))
                labels.append((v, c))
        for (iu, iv), block in self._quad.items():
            for (cu, cv), b in block.items():
                bqm.add_quadratic((self.variables[iu], cu), (self.variables[iv], cv), b)
        for i, v in enumerate(self.variables):
            bqm.add_linear_equality_constraint([((v, c), 1) for c in range(self._cases[i])], penalty, -1)
        return bqm, labels

    def default_penalty(self) -> float:
        gain = [float(np.abs(lin).max()) for lin in self._linear]
        for (iu, iv), block in self._quad.items():
            mx = max((abs(b) for b in block.values()), default=0.0)
            gain[iu] += mx
            gain[iv] += mx
        return max(gain, default=0.0) + 1.0

    def to_lowered(self, penalty: Optional[float] = None) -> LoweredModel:
        if len(set(self._cases)) > 1:
            raise ValueError("to_lowered needs the same number of cases for every variable")
        A = self.default_penalty() if penalty is None else float(penalty)
        bqm, _ = self.to_bqm(A)
        model = lowered_from_bqm(bqm)
        model.meta.update({"kind": "dqm", "num_cases": self._cases[0] if self._cases else 0, "cells": list(self.variables),
                           "penalty": A})
        return model
