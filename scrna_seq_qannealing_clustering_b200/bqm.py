"""Minimal dimod-compatible ``BinaryQuadraticModel`` (dimod is not installable offline).

Only the surface the reference touches is mirrored:
  * ``BinaryQuadraticModel.from_qubo(Q)``                      BQM_clustering.py:371, and implicitly every
    ``sampler.sample_qubo(Q, ...)`` (BQM_clustering.py:57,75,85,245,263,273; QA_subsampling.py:42,56,65)
  * ``bqm.add_linear_inequality_constraint(terms, lb=, ub=, lagrange_multiplier=, label=)``  BQM_clustering.py:376-380
  * what neal's sampler needs from it: ``change_vartype(SPIN)``, ``to_numpy_vectors``, ``energies``.

Conventions restated from dimod (from memory -- see SURVEY.md 8c, parity unpinned):
  * variable order = first appearance while iterating Q;
  * QUBO -> Ising: J = Q_ij/4, h_i = Q_ii/2 + sum_j Q_ij/4 (accumulated in ascending neighbour order),
    offset += sum_i Q_ii/2 + sum_{i<j} Q_ij/4;
  * ``to_numpy_vectors``: couplers sorted by (row = larger index, col = smaller index), which makes
    neal's per-vertex push_back adjacency ascending in neighbour index.
"""
from __future__ import annotations

import enum
import math
from typing import Dict, Hashable, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np


class Vartype(enum.Enum):
    SPIN = frozenset({-1, 1})
    BINARY = frozenset({0, 1})


SPIN = Vartype.SPIN
BINARY = Vartype.BINARY


def as_vartype(v) -> Vartype:
    if isinstance(v, Vartype):
        return v
    if isinstance(v, str):
        return Vartype[v.upper()]
    if hasattr(v, "name") and str(v.name) in ("SPIN", "BINARY"):  # a real dimod.Vartype
        return Vartype[str(v.name)]
    s = frozenset(v)
    for vt in Vartype:
        if vt.value == s:
            return vt
    raise TypeError(f"unknown vartype {v!r}")


class BinaryQuadraticModel:
    """E(x) = offset + sum_i linear[i] x_i + sum_{i<j} quadratic[i,j] x_i x_j over SPIN or BINARY x."""

    def __init__(self, linear: Optional[Mapping] = None, quadratic: Optional[Mapping] = None, offset: float = 0.0,
                 vartype=BINARY):
        if linear is not None and not isinstance(linear, Mapping) and quadratic is None:
            # BinaryQuadraticModel(vartype)
            vartype, linear = linear, None
        self.vartype = as_vartype(vartype)
        self.offset = float(offset)
        self._index: Dict[Hashable, int] = {}
        self._labels = []
        self._linear = []
        self._quad: Dict[Tuple[int, int], float] = {}  # key (lo, hi) index pair, insertion ordered
        self._vec_cache = None
        if linear:
            for v, b in linear.items():
                self.add_linear(v, b)
        if quadratic:
            for (u, v), b in quadratic.items():
                self.add_quadratic(u, v, b)

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_qubo(cls, Q: Mapping, offset: float = 0.0) -> "BinaryQuadraticModel":
        bqm = cls({}, {}, offset, BINARY)
        for (u, v), bias in Q.items():
            if u == v:
                bqm.add_linear(u, bias)
            else:
                bqm.add_quadratic(u, v, bias)
        return bqm

    @classmethod
    def from_ising(cls, h, J: Mapping, offset: float = 0.0) -> "BinaryQuadraticModel":
        bqm = cls({}, {}, offset, SPIN)
        items = h.items() if isinstance(h, Mapping) else enumerate(h)
        for v, bias in items:
            bqm.add_linear(v, bias)
        for (u, v), bias in J.items():
            bqm.add_quadratic(u, v, bias)
        return bqm

    @classmethod
    def from_numpy_vectors(cls, linear, quadratic, offset, vartype, variable_order: Optional[Sequence] = None):
        irow, icol, qdata = quadratic
        linear = np.asarray(linear, dtype=np.float64)
        labels = list(range(len(linear))) if variable_order is None else list(variable_order)
        bqm = cls({}, {}, offset, vartype)
        for v, b in zip(labels, linear.tolist()):
            bqm.add_linear(v, b)
        for r, c, q in zip(np.asarray(irow).tolist(), np.asarray(icol).tolist(), np.asarray(qdata).tolist()):
            bqm.add_quadratic(labels[r], labels[c], q)
        return bqm

    def add_variable(self, v=None, bias: float = 0.0):
        if v is None:
            v = len(self._labels)
            while v in self._index:
                v += 1
        i = self._index.get(v)
        if i is None:
            self._index[v] = len(self._labels)
            self._labels.append(v)
            self._linear.append(0.0)
            i = len(self._labels) - 1
        self._linear[i] += bias
        self._vec_cache = None
        return v

    add_linear = add_variable

    def set_linear(self, v, bias: float):
        self.add_variable(v, 0.0)
        self._linear[self._index[v]] = float(bias)
        self._vec_cache = None

    def add_quadratic(self, u, v, bias: float):
        if u == v:
            raise ValueError(f"{u!r} cannot have an interaction with itself")
        self.add_variable(u, 0.0)
        self.add_variable(v, 0.0)
        iu, iv = self._index[u], self._index[v]
        key = (iu, iv) if iu < iv else (iv, iu)
        self._quad[key] = self._quad.get(key, 0.0) + bias
        self._vec_cache = None

    add_interaction = add_quadratic

    def set_quadratic(self, u, v, bias: float):
        self.add_quadratic(u, v, 0.0)
        iu, iv = self._index[u], self._index[v]
        self._quad[(iu, iv) if iu < iv else (iv, iu)] = float(bias)
        self._vec_cache = None

    # ---- views ----------------------------------------------------------------------------------
    @property
    def variables(self):
        return list(self._labels)

    @property
    def num_variables(self) -> int:
        return len(self._labels)

    @property
    def num_interactions(self) -> int:
        return len(self._quad)

    def __len__(self):
        return len(self._labels)

    @property
    def shape(self):
        return (self.num_variables, self.num_interactions)

    @property
    def linear(self) -> Dict:
        return {v: self._linear[i] for v, i in self._index.items()}

    @property
    def quadratic(self) -> Dict:
        L = self._labels
        return {(L[i], L[j]): b for (i, j), b in self._quad.items()}

    def get_linear(self, v) -> float:
        return self._linear[self._index[v]]

    def get_quadratic(self, u, v, default=None) -> float:
        iu, iv = self._index[u], self._index[v]
        key = (iu, iv) if iu < iv else (iv, iu)
        if key not in self._quad:
            if default is None:
                raise ValueError(f"no interaction between {u!r} and {v!r}")
            return default
        return self._quad[key]

    def copy(self) -> "BinaryQuadraticModel":
        new = BinaryQuadraticModel({}, {}, self.offset, self.vartype)
        new._index = dict(self._index)
        new._labels = list(self._labels)
        new._linear = list(self._linear)
        new._quad = dict(self._quad)
        return new

    # ---- vectors (dimod to_numpy_vectors) -------------------------------------------------------
    def to_numpy_vectors(self, variable_order: Optional[Sequence] = None, dtype=np.float64, index_dtype=np.int32,
                         sort_indices: bool = True, return_labels: bool = False):
        """(ldata, (irow, icol, qdata), offset[, labels]); couplers sorted by (row=larger, col=smaller)."""
        if variable_order is None:
            perm = None
            labels = list(self._labels)
        else:
            labels = list(variable_order)
            if len(labels) != len(self._labels) or set(labels) != set(self._labels):
                raise ValueError("variable_order must be a permutation of the variables")
            perm = np.empty(len(labels), dtype=np.int64)
            for new, v in enumerate(labels):
                perm[self._index[v]] = new
        n = len(labels)
        ldata = np.zeros(n, dtype=dtype)
        lin = np.asarray(self._linear, dtype=dtype)
        if perm is None:
            ldata[:] = lin
        else:
            ldata[perm] = lin
        m = len(self._quad)
        if m:
            keys = np.fromiter((k for ij in self._quad.keys() for k in ij), dtype=np.int64, count=2 * m).reshape(m, 2)
            qdata = np.fromiter(self._quad.values(), dtype=dtype, count=m)
            if perm is not None:
                keys = perm[keys]
            irow = keys.max(axis=1)
            icol = keys.min(axis=1)
            if sort_indices:
                order = np.lexsort((icol, irow))
                irow, icol, qdata = irow[order], icol[order], qdata[order]
        else:
            irow = np.zeros(0, dtype=np.int64)
            icol = np.zeros(0, dtype=np.int64)
            qdata = np.zeros(0, dtype=dtype)
        out = (ldata, (irow.astype(index_dtype), icol.astype(index_dtype), qdata), self.offset)
        return out + (labels,) if return_labels else out

    # ---- vartype --------------------------------------------------------------------------------
    def change_vartype(self, vartype, inplace: bool = True) -> "BinaryQuadraticModel":
        vartype = as_vartype(vartype)
        bqm = self if inplace else self.copy()
        if vartype is bqm.vartype:
            return bqm
        ldata, (irow, icol, qdata), offset = bqm.to_numpy_vectors()
        n = len(ldata)
        if vartype is SPIN:  # x = (s+1)/2
            h = ldata / 2.0
            q4 = qdata / 4.0
            # ascending-neighbour accumulation per variable (sequential, deterministic)
            h = _accumulate_rows(h, irow, icol, q4)
            new_offset = offset + _seq_sum(ldata / 2.0) + _seq_sum(q4)
            new_q = q4
        else:  # s = 2x - 1
            lin = 2.0 * ldata
            q2 = 2.0 * qdata
            lin = _accumulate_rows(lin, irow, icol, -q2)
            new_offset = offset - _seq_sum(ldata) + _seq_sum(qdata)
            new_q = 4.0 * qdata
        bqm._linear = h.tolist() if vartype is SPIN else lin.tolist()
        bqm._quad = {}
        for r, c, q in zip(irow.tolist(), icol.tolist(), new_q.tolist()):
            bqm._quad[(c, r)] = q
        bqm.offset = float(new_offset)
        bqm.vartype = vartype
        bqm._vec_cache = None
        return bqm

    @property
    def spin(self) -> "BinaryQuadraticModel":
        return self.change_vartype(SPIN, inplace=False)

    @property
    def binary(self) -> "BinaryQuadraticModel":
        return self.change_vartype(BINARY, inplace=False)

    # ---- energies -------------------------------------------------------------------------------
    def energies(self, samples_like, dtype=np.float64) -> np.ndarray:
        """Energies of samples given as an array [R][n] in variable order, (array, labels), or dict(s)."""
        labels = None
        if isinstance(samples_like, tuple) and len(samples_like) == 2:
            samples_like, labels = samples_like
        if isinstance(samples_like, Mapping):
            samples_like = [samples_like]
        if isinstance(samples_like, (list, tuple)) and samples_like and isinstance(samples_like[0], Mapping):
            arr = np.array([[s[v] for v in self._labels] for s in samples_like], dtype=dtype)
        else:
            arr = np.atleast_2d(np.asarray(samples_like)).astype(dtype)
            if labels is not None:
                pos = {v: i for i, v in enumerate(labels)}
                arr = arr[:, [pos[v] for v in self._labels]]
        ldata, (irow, icol, qdata), offset = self.to_numpy_vectors()
        e = arr @ ldata + offset
        if len(qdata):
            e = e + (arr[:, irow] * arr[:, icol]) @ qdata
        return e

    def energy(self, sample, dtype=np.float64) -> float:
        return float(self.energies(sample, dtype=dtype)[0])

    # ---- constraints (dimod add_linear_(in)equality_constraint) ---------------------------------
    def add_linear_equality_constraint(self, terms: Iterable[Tuple[Hashable, float]], lagrange_multiplier: float,
                                       constant: float):
        """Add lagrange_multiplier * (sum_i a_i x_i + constant)^2 (BINARY or SPIN variables)."""
        terms = list(terms)
        lam = lagrange_multiplier
        for idx, (v, a) in enumerate(terms):
            if self.vartype is BINARY:
                self.add_linear(v, lam * a * (2 * constant + a))
            else:
                self.add_linear(v, lam * a * 2 * constant)
                self.offset += lam * a * a
            for u, b in terms[idx + 1:]:
                if u == v:
                    if self.vartype is BINARY:
                        self.add_linear(v, 2 * lam * a * b)
                    else:
                        self.offset += 2 * lam * a * b
                else:
                    self.add_quadratic(v, u, 2 * lam * a * b)
        self.offset += lam * constant * constant

    def add_linear_inequality_constraint(self, terms: Iterable[Tuple[Hashable, int]], lagrange_multiplier: float,
                                         label: str, constant: int = 0, lb: int = -(2 ** 53), ub: int = 0,
                                         cross_zero: bool = False, penalization_method: str = "slack"):
        """lb <= sum_i a_i x_i + constant <= ub as a penalty with binary-encoded slack variables.

        Restates dimod's ``add_linear_inequality_constraint`` (slack method; BQM_clustering.py:376-380):
        slack variables ``slack_{label}_{j}`` with coefficients 2^j and a final remainder so that the
        slack spans exactly 0..(ub-lb); penalty  lagrange * (sum a_i x_i + sum c_j slack_j + constant - ub)^2.
        Returns the list of slack (label, coefficient) terms.
        """
        if self.vartype is not BINARY:
            raise ValueError("inequality constraints are implemented for BINARY models")
        if penalization_method != "slack":
            raise ValueError("only the 'slack' penalization method is implemented")
        terms = list(terms)
        if int(constant) != constant or int(lb) != lb or int(ub) != ub or any(int(a) != a for _, a in terms):
            import warnings
            warnings.warn("For constraints with fractional coefficients, multiply both sides of the inequality by an "
                          "appropriate factor to attain integer coefficients (dimod semantics: the slack range is "
                          "int(ub - lb), the penalty keeps the fractional bound)", stacklevel=2)
        hi = sum(a for _, a in terms if a > 0)
        lo = sum(a for _, a in terms if a < 0)
        ub_c = min(hi, ub - constant)
        lb_c = max(lo, lb - constant)
        if hi <= ub_c and lo >= lb_c:
            import warnings
            warnings.warn(f"Did not add constraint {label}. This constraint is feasible with any value for state "
                          "variables.", stacklevel=2)
            return []
        if ub_c < lb_c:
            raise ValueError(f"The given constraint ({label}) is infeasible with any value for state variables.")
        slack_upper = int(ub_c - lb_c)
        slack_terms = []
        if slack_upper > 0:
            nbits = int(math.floor(math.log2(slack_upper)))
            coeffs = [2 ** j for j in range(nbits)]
            if slack_upper - 2 ** nbits >= 0:
                coeffs.append(slack_upper - 2 ** nbits + 1)
            for j, c in enumerate(coeffs):
                sv = f"slack_{label}_{j}"
                self.add_variable(sv, 0.0)
                slack_terms.append((sv, int(c)))
        self.add_linear_equality_constraint(terms + slack_terms, lagrange_multiplier, -ub_c)
        return slack_terms

    def __repr__(self):
        return (f"BinaryQuadraticModel({self.num_variables} variables, {self.num_interactions} interactions, "
                f"offset={self.offset}, {self.vartype.name})")


def _seq_sum(a: np.ndarray) -> float:
    """Left-to-right fp64 sum (np.sum is pairwise; the offset should not depend on numpy's blocking)."""
    return float(np.cumsum(a, dtype=np.float64)[-1]) if len(a) else 0.0


def _accumulate_rows(base: np.ndarray, irow: np.ndarray, icol: np.ndarray, contrib: np.ndarray) -> np.ndarray:
    """base[v] += contrib over the couplers touching v, in coupler order (ascending neighbour for sorted couplers)."""
    out = np.array(base, dtype=np.float64, copy=True)
    if len(contrib) == 0:
        return out
    m = len(contrib)
    # entry sequence exactly like the adjacency build: coupler c contributes to irow[c] then icol[c]
    idx = np.empty(2 * m, dtype=np.int64)
    idx[0::2] = irow
    idx[1::2] = icol
    val = np.repeat(contrib, 2)
    np.add.at(out, idx, val)  # unbuffered, sequential in index order
    return out


def qubo_to_ising_vectors(ldata: np.ndarray, irow: np.ndarray, icol: np.ndarray, qdata: np.ndarray, offset: float = 0.0):
    """Vector form of the QUBO -> Ising map (same arithmetic as ``change_vartype(SPIN)``)."""
    q4 = np.asarray(qdata, dtype=np.float64) / 4.0
    h = _accumulate_rows(np.asarray(ldata, dtype=np.float64) / 2.0, irow, icol, q4)
    new_offset = offset + _seq_sum(np.asarray(ldata, dtype=np.float64) / 2.0) + _seq_sum(q4)
    return h, q4, float(new_offset)
