"""ctypes wrapper of oracle/liboracle_sa.so (built from oracle/cpu_sa_ref.cpp by oracle/Makefile).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  Parity is UNPINNED against the real dwave-neal (absent offline) -- see cpu_sa_ref.cpp.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle_sa.so"


class OracleStats(C.Structure):
    _fields_ = [("attempts", C.c_uint64), ("candidates", C.c_uint64), ("draws", C.c_uint64), ("accepted", C.c_uint64),
                ("nbr_updates", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def build(force: bool = False) -> Path:
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (_HERE / "cpu_sa_ref.cpp").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        lib = C.CDLL(str(LIB_PATH))
        p = C.c_void_p
        lib.oracle_sa_sample_ising.restype = C.c_int
        lib.oracle_sa_sample_ising.argtypes = [C.c_int32, p, C.c_int64, p, p, p, C.c_int32, p, p, C.c_int32, p, C.c_int32, p,
                                               C.c_int32, C.c_int32, p, p, p, p, C.c_int32, C.POINTER(OracleStats)]
        lib.oracle_state_energies.restype = C.c_int
        lib.oracle_state_energies.argtypes = [C.c_int32, p, C.c_int64, p, p, p, C.c_int32, p, p, C.c_int32, p, p, p, p]
        lib.oracle_rng_stream.restype = None
        lib.oracle_rng_stream.argtypes = [C.c_uint64, C.c_int32, p]
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _groups(groups, n):
    if groups is None:
        return 0, None, None, None, None
    grp, coef, lam, kappa = groups
    return (len(lam), np.ascontiguousarray(grp, dtype=np.int32), np.ascontiguousarray(coef, dtype=np.int32),
            np.ascontiguousarray(lam, dtype=np.float64), np.ascontiguousarray(kappa, dtype=np.int64))


def sample_ising(h, starts, ends, weights, states, beta_schedule, sweeps_per_beta, seeds, seed_mode=0, groups=None,
                 nthreads=0):
    """neal general_simulated_annealing restated; ``states`` int8 [R][n] is updated in place.  Returns (energies, stats)."""
    lib = load()
    h = np.ascontiguousarray(h, dtype=np.float64)
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    ends = np.ascontiguousarray(ends, dtype=np.int32)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    betas = np.ascontiguousarray(beta_schedule, dtype=np.float64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    if states.dtype != np.int8 or not states.flags.c_contiguous:
        raise ValueError("states must be C-contiguous int8")
    R, n = states.shape
    energies = np.empty(R, dtype=np.float64)
    ng, grp, coef, lam, kappa = _groups(groups, n)
    st = OracleStats()
    rc = lib.oracle_sa_sample_ising(n, _ptr(h), len(starts), _ptr(starts), _ptr(ends), _ptr(weights), R, _ptr(states),
                                    _ptr(energies), len(betas), _ptr(betas), int(sweeps_per_beta), _ptr(seeds), int(seed_mode),
                                    ng, _ptr(grp), _ptr(coef), _ptr(lam), _ptr(kappa), int(nthreads), C.byref(st))
    if rc < 0:
        raise RuntimeError(f"oracle error {rc}: {lib.oracle_last_error().decode()}")
    return energies, st.as_dict()


def state_energies(h, starts, ends, weights, states, groups=None):
    lib = load()
    h = np.ascontiguousarray(h, dtype=np.float64)
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    ends = np.ascontiguousarray(ends, dtype=np.int32)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    states = np.ascontiguousarray(states, dtype=np.int8)
    R, n = states.shape
    energies = np.empty(R, dtype=np.float64)
    ng, grp, coef, lam, kappa = _groups(groups, n)
    rc = lib.oracle_state_energies(n, _ptr(h), len(starts), _ptr(starts), _ptr(ends), _ptr(weights), R, _ptr(states),
                                   _ptr(energies), ng, _ptr(grp), _ptr(coef), _ptr(lam), _ptr(kappa))
    if rc < 0:
        raise RuntimeError(f"oracle error {rc}: {lib.oracle_last_error().decode()}")
    return energies


def rng_stream(seed: int, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.uint64)
    load().oracle_rng_stream(C.c_uint64(seed), count, _ptr(out))
    return out


def num_threads() -> int:
    return int(load().oracle_num_threads())
