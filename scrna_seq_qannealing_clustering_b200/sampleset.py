"""Minimal dimod-compatible ``SampleSet`` covering what the reference consumes:

  response.data(fields=['sample','energy','num_occurrences'])   BQM_clustering.py:93,281,397; QA_subsampling.py:73
  response.first.sample / .energy / .num_occurrences            BQM_clustering.py:105,294,410; DQM_clustering.py:46;
                                                                CQM_clustering.py:54,90; main.py:175-177
  response.record.energy[i]                                     BQM_clustering.py:133-143,321-323
  sampleset.samples()[:k]                                       plot_and_save.py:106
  response.info[...]                                            BQM_clustering.py:79,267

When the real ``dimod`` is importable, ``SampleSet.to_dimod()`` converts losslessly.
"""
from __future__ import annotations

from collections import namedtuple
from collections.abc import Mapping
from typing import Iterable, Optional, Sequence

import numpy as np

from .bqm import BINARY, SPIN, Vartype, as_vartype


class SampleView(Mapping):
    """One row of the sample matrix as a read-only ``label -> value`` mapping."""

    __slots__ = ("_row", "_labels", "_index")

    def __init__(self, row: np.ndarray, labels: Sequence, index: dict):
        self._row = row
        self._labels = labels
        self._index = index

    def __getitem__(self, v):
        return int(self._row[self._index[v]])

    def __iter__(self):
        return iter(self._labels)

    def __len__(self):
        return len(self._labels)

    def __repr__(self):
        return repr(dict(self))


class SamplesArray:
    """``sampleset.samples()``: sequence of ``SampleView`` supporting ints and slices."""

    def __init__(self, sampleset: "SampleSet", order: np.ndarray):
        self._ss = sampleset
        self._order = order

    def __len__(self):
        return len(self._order)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SamplesArray(self._ss, self._order[i])
        if isinstance(i, tuple):  # samples()[row, label]
            row, v = i
            return self[row][v]
        return self._ss._view(int(self._order[i]))

    def __iter__(self):
        for r in self._order:
            yield self._ss._view(int(r))


class SampleSet:
    """Samples + energies + occurrences; record order = read order unless constructed sorted."""

    def __init__(self, record: np.recarray, variables: Sequence, info: dict, vartype):
        self.record = record
        self.variables = list(variables)
        self.info = dict(info)
        self.vartype = as_vartype(vartype)
        self._index = {v: i for i, v in enumerate(self.variables)}

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_samples(cls, samples_like, energy, vartype, info: Optional[dict] = None, num_occurrences=None,
                     aggregate_samples: bool = False, sort_labels: bool = False, **vectors) -> "SampleSet":
        if isinstance(samples_like, tuple) and len(samples_like) == 2:
            samples, labels = samples_like
        else:
            samples, labels = samples_like, None
        samples = np.atleast_2d(np.asarray(samples, dtype=np.int8))
        if labels is None:
            labels = list(range(samples.shape[1]))
        energy = np.atleast_1d(np.asarray(energy, dtype=np.float64))
        R = samples.shape[0]
        if energy.shape[0] != R:
            raise ValueError("one energy per sample row is required")
        if num_occurrences is None:
            num_occurrences = np.ones(R, dtype=np.int64)
        dtype = [("sample", np.int8, (samples.shape[1],)), ("energy", np.float64), ("num_occurrences", np.int64)]
        extra = {}
        for name, vec in vectors.items():
            vec = np.asarray(vec)
            dtype.append((name, vec.dtype, vec.shape[1:]))
            extra[name] = vec
        rec = np.recarray(R, dtype=dtype)
        rec.sample[:] = samples
        rec.energy[:] = energy
        rec.num_occurrences[:] = num_occurrences
        for name, vec in extra.items():
            rec[name][:] = vec
        ss = cls(rec, labels, info or {}, vartype)
        return ss.aggregate() if aggregate_samples else ss

    # ---- access ---------------------------------------------------------------------------------
    def __len__(self):
        return int(self.record.shape[0])

    def _view(self, r: int) -> SampleView:
        return SampleView(self.record.sample[r], self.variables, self._index)

    def _sorted_order(self, sorted_by: Optional[str] = "energy", reverse: bool = False) -> np.ndarray:
        if sorted_by is None:
            order = np.arange(len(self))
        else:
            order = np.argsort(self.record[sorted_by], kind="stable")
        return order[::-1] if reverse else order

    def samples(self, n: Optional[int] = None, sorted_by: Optional[str] = "energy") -> SamplesArray:
        order = self._sorted_order(sorted_by)
        return SamplesArray(self, order if n is None else order[:n])

    def __iter__(self):
        return iter(self.samples(sorted_by=None))

    def data(self, fields: Optional[Iterable[str]] = None, sorted_by: Optional[str] = "energy", name: str = "Sample",
             reverse: bool = False, sample_dict_cast: bool = True, index: bool = False):
        """Yield namedtuples of the requested fields, lowest energy first (dimod semantics)."""
        if fields is None:
            fields = [f for f in self.record.dtype.names]
        fields = list(fields)
        if index:
            fields = fields + ["idx"]
        tup = namedtuple(name, fields)
        for r in self._sorted_order(sorted_by, reverse):
            vals = []
            for f in fields:
                if f == "sample":
                    v = self._view(int(r))
                    vals.append(dict(v) if sample_dict_cast else v)
                elif f == "idx":
                    vals.append(int(r))
                else:
                    x = self.record[f][r]
                    vals.append(x.item() if np.ndim(x) == 0 else x)
            yield tup(*vals)

    @property
    def first(self):
        """Lowest-energy row (ties -> lowest record index)."""
        if len(self) == 0:
            raise ValueError("empty SampleSet has no first sample")
        return next(self.data(sorted_by="energy", sample_dict_cast=False))

    def lowest(self, rtol: float = 1e-5, atol: float = 1e-8) -> "SampleSet":
        if len(self) == 0:
            return self
        e = self.record.energy
        keep = np.isclose(e, e.min(), rtol=rtol, atol=atol)
        return SampleSet(self.record[keep].view(np.recarray), self.variables, self.info, self.vartype)

    def truncate(self, n: int, sorted_by: Optional[str] = "energy") -> "SampleSet":
        order = self._sorted_order(sorted_by)[:n]
        return SampleSet(self.record[order].view(np.recarray), self.variables, self.info, self.vartype)

    def aggregate(self) -> "SampleSet":
        """Merge identical samples, summing ``num_occurrences`` (QPU-histogram-like result)."""
        if len(self) == 0:
            return self
        s = np.ascontiguousarray(self.record.sample)
        _, first_idx, inverse = np.unique(s, axis=0, return_index=True, return_inverse=True)
        inverse = inverse.reshape(-1)
        occ = np.zeros(len(first_idx), dtype=np.int64)
        np.add.at(occ, inverse, self.record.num_occurrences)
        order = np.argsort(first_idx, kind="stable")  # keep first-appearance order
        rec = self.record[first_idx[order]].view(np.recarray).copy()
        rec.num_occurrences[:] = occ[order]
        return SampleSet(rec, self.variables, self.info, self.vartype)

    def sorted(self) -> "SampleSet":
        """Energy-sorted copy, so ``record.energy[0]`` is the best energy as with QPU answers
        (the reference's ``conf`` termination rule, BQM_clustering.py:133-146, relies on that)."""
        order = self._sorted_order("energy")
        return SampleSet(self.record[order].view(np.recarray), self.variables, self.info, self.vartype)

    def change_vartype(self, vartype, energy_offset: float = 0.0, inplace: bool = True) -> "SampleSet":
        vartype = as_vartype(vartype)
        ss = self if inplace else SampleSet(self.record.copy(), self.variables, self.info, self.vartype)
        if energy_offset:
            ss.record.energy[:] = ss.record.energy + energy_offset
        if vartype is ss.vartype:
            return ss
        if vartype is BINARY:
            ss.record.sample[:] = (ss.record.sample + 1) // 2
        else:
            ss.record.sample[:] = 2 * ss.record.sample - 1
        ss.vartype = vartype
        return ss

    def relabel_variables(self, mapping: Mapping, inplace: bool = True) -> "SampleSet":
        ss = self if inplace else SampleSet(self.record.copy(), self.variables, self.info, self.vartype)
        ss.variables = [mapping.get(v, v) for v in ss.variables]
        ss._index = {v: i for i, v in enumerate(ss.variables)}
        return ss

    def to_dimod(self):
        """Convert to a real ``dimod.SampleSet`` (only where dimod is installed)."""
        import dimod  # noqa: F401  (absent offline)
        return dimod.SampleSet.from_samples((np.asarray(self.record.sample), self.variables), energy=self.record.energy,
                                            num_occurrences=self.record.num_occurrences, vartype=self.vartype.name,
                                            info=self.info)

    def __repr__(self):
        return f"SampleSet({len(self)} rows, {len(self.variables)} variables, {self.vartype.name})"
