"""``B200SimulatedAnnealingSampler``: dimod ``Sampler`` surface over libqanneal.so.

Drop-in for the sampler objects the reference constructs (``LeapHybridSampler``, ``EmbeddingComposite(DWaveSampler())``,
``LeapHybridDQMSampler``, ``LeapHybridCQMSampler``, ``hybrid.KerberosSampler``; BQM_clustering.py:56-85,386;
DQM_clustering.py:45; CQM_clustering.py:52-53,88-89; QA_subsampling.py:40-65) with the argument semantics of
dwave-neal's ``SimulatedAnnealingSampler.sample``:

    sample(bqm, beta_range=None, num_reads=None, num_sweeps=None, num_sweeps_per_beta=1,
           beta_schedule_type='geometric', seed=None, interrupt_function=None, beta_schedule=None,
           initial_states=None, initial_states_generator='random', **kwargs) -> SampleSet

All annealing, energy evaluation and argmin run in the CUDA library; this file only parses arguments,
lowers models to vectors and wraps results.  There is no CPU path: without the library or a GPU it raises.
"""
from __future__ import annotations

import time
import warnings
from typing import Mapping, Optional

import numpy as np

from . import _lib, schedule
from .bqm import BINARY, SPIN, BinaryQuadraticModel, as_vartype
from .engine import Context, IsingModel
from .models import LoweredModel, decode_onehot, lowered_from_bqm
from .sampleset import SampleSet

_KNOWN_QPU_KWARGS = ("label", "chain_strength", "return_embedding", "time_limit", "max_iter", "qpu_reads", "tabu_timeout",
                     "qpu_params", "annealing_time", "answer_mode")


class B200SimulatedAnnealingSampler:
    """Simulated annealing on one B200 (reads shard over ranks when torch.distributed is initialised)."""

    parameters = None
    properties = None

    def __init__(self, device: Optional[int] = None, context: Optional[Context] = None):
        self.parameters = {
            "beta_range": [], "num_reads": [], "num_sweeps": [], "num_sweeps_per_beta": [],
            "beta_schedule_type": ["beta_schedule_options"], "seed": [], "interrupt_function": [], "beta_schedule": [],
            "initial_states": [], "initial_states_generator": [],
            # extensions
            "mode": [], "seed_mode": [], "sorted": [], "aggregate": [],
        }
        self.properties = {"beta_schedule_options": schedule.BETA_SCHEDULE_OPTIONS}
        self._ctx = context
        self._device = device
        self.last_stats: Optional[dict] = None

    # ---- plumbing -------------------------------------------------------------------------------
    @property
    def context(self) -> Context:
        if self._ctx is None:
            dev = self._device
            if dev is None:
                import os
                dev = int(os.environ.get("LOCAL_RANK", "0"))
                if dev >= max(_lib.load().qa_device_count(), 1):
                    dev = 0
            self._ctx = Context(dev)
        return self._ctx

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    # ---- dimod Sampler API ----------------------------------------------------------------------
    def sample(self, bqm, beta_range=None, num_reads=None, num_sweeps=None, num_sweeps_per_beta=1,
               beta_schedule_type="geometric", seed=None, interrupt_function=None, beta_schedule=None,
               initial_states=None, initial_states_generator="random", mode="reference", seed_mode="per_read",
               sorted=False, aggregate=False, **kwargs) -> SampleSet:
        """Anneal a ``BinaryQuadraticModel`` (ours or dimod's) or a ``LoweredModel``."""
        schedule.warn_unknown_kwargs(type(self).__name__, kwargs)
        t0 = time.perf_counter_ns()
        if isinstance(bqm, LoweredModel):
            model, vartype = bqm, BINARY
        else:
            if not isinstance(bqm, BinaryQuadraticModel):
                bqm = _coerce_bqm(bqm)
            model, vartype = lowered_from_bqm(bqm), bqm.vartype
        return self._run(model, vartype, t0, beta_range, num_reads, num_sweeps, num_sweeps_per_beta, beta_schedule_type, seed,
                         interrupt_function, beta_schedule, initial_states, initial_states_generator, mode, seed_mode,
                         sorted, aggregate)

    def sample_qubo(self, Q: Mapping, **parameters) -> SampleSet:
        """dimod ``Sampler.sample_qubo``: ``BQM.from_qubo(Q)`` then ``sample``."""
        return self.sample(BinaryQuadraticModel.from_qubo(Q), **parameters)

    def sample_ising(self, h, J, **parameters) -> SampleSet:
        return self.sample(BinaryQuadraticModel.from_ising(h, J), **parameters)

    def sample_dqm(self, dqm, penalty: Optional[float] = None, **parameters) -> SampleSet:
        """``LeapHybridDQMSampler().sample_dqm`` stand-in (DQM_clustering.py:45): one-hot expansion + anneal.

        Returns one row per read with a case index per discrete variable; ``energy`` is the DQM energy of the
        decoded assignment and the extra record field ``is_feasible`` says whether every cell was one-hot.
        """
        from .dqm import DiscreteQuadraticModel
        if isinstance(dqm, LoweredModel):
            model = dqm
        elif isinstance(dqm, DiscreteQuadraticModel):
            model = dqm.to_lowered(penalty)
        else:
            raise TypeError("sample_dqm expects a DiscreteQuadraticModel or a LoweredModel built by models.dqm_model")
        ss = self.sample(model, **parameters)
        return _decode_discrete(ss, model, dqm if isinstance(dqm, DiscreteQuadraticModel) else None)

    def sample_cqm(self, cqm, onehot_penalty: Optional[float] = None, constraint_penalty: Optional[float] = None,
                   **parameters) -> SampleSet:
        """``LeapHybridCQMSampler().sample_cqm`` stand-in (CQM_clustering.py:53,89): constraints lowered to penalties.

        Rows keep the binary variables ('v_{i},{k}' and slack bits); record fields ``is_feasible`` (one-hot and
        minimum-size constraints satisfied by the cell variables) and ``objective`` are added.
        """
        from .cqm import ConstrainedQuadraticModel
        if isinstance(cqm, LoweredModel):
            model = cqm
        elif isinstance(cqm, ConstrainedQuadraticModel):
            model = cqm.to_lowered(onehot_penalty, constraint_penalty)
        else:
            raise TypeError("sample_cqm expects a ConstrainedQuadraticModel or a LoweredModel built by models.cqm_model")
        ss = self.sample(model, **parameters)
        return _annotate_cqm(ss, model)

    # ---- core -----------------------------------------------------------------------------------
    def _run(self, model: LoweredModel, vartype, t0, beta_range, num_reads, num_sweeps, num_sweeps_per_beta,
             beta_schedule_type, seed, interrupt_function, beta_schedule, initial_states, initial_states_generator, mode,
             seed_mode, sort_result, aggregate) -> SampleSet:
        if mode not in ("reference", "throughput"):
            raise ValueError("mode must be 'reference' (bit-exact neal order) or 'throughput' (fields recomputed from spins)")
        if seed_mode not in ("per_read", "stream"):
            raise ValueError("seed_mode must be 'per_read' or 'stream'")
        if interrupt_function is not None and not callable(interrupt_function):
            raise TypeError("'interrupt_function' should be a callable")
        seed = schedule.resolve_seed(seed)
        n = model.num_variables
        labels = model.labels
        vartype = as_vartype(vartype)
        info = {"beta_schedule_type": beta_schedule_type}
        if n == 0:
            ss = SampleSet.from_samples((np.empty((0, 0), dtype=np.int8), []), energy=[], vartype=vartype, info=info)
            return ss
        states = schedule.parse_initial_states(n, labels, vartype is SPIN, initial_states, initial_states_generator,
                                               num_reads, seed)
        num_reads = states.shape[0]
        groups = model.groups.astuple() if model.groups is not None else None
        if beta_schedule_type != "custom":
            if beta_range is None:
                beta_range = schedule.default_ising_beta_range(model.h, model.starts, model.ends, model.weights, groups)
            elif len(beta_range) != 2 or min(beta_range) < 0:
                raise ValueError("'beta_range' should be a 2-tuple of positive numbers")
            info["beta_range"] = [float(beta_range[0]), float(beta_range[1])]
        betas, spb = schedule.make_beta_schedule(beta_range, num_sweeps, num_sweeps_per_beta, beta_schedule_type, beta_schedule)

        # read sharding over ranks (one process per GPU); seeds depend on the global read index only
        rank, world = _dist_rank_world()
        lo, hi = _shard(num_reads, rank, world)
        if seed_mode == "stream":
            if world > 1:
                raise ValueError("seed_mode='stream' is one serial RNG chain and cannot shard over GPUs")
            seeds = np.array([seed], dtype=np.uint64)
        else:
            seeds = schedule.per_read_seeds(seed, hi - lo, first_read=lo)
        local_states = np.ascontiguousarray(states[lo:hi])

        ctx = self.context
        gm = IsingModel(ctx, model.h, model.starts, model.ends, model.weights)
        try:
            if groups is not None:
                gm.set_groups(*groups)
            t1 = time.perf_counter_ns()
            energies, st, done = gm.sample(local_states, betas, spb, seeds,
                                           seed_mode=_lib.QA_SEED_STREAM if seed_mode == "stream" else _lib.QA_SEED_PER_READ,
                                           mode=_lib.QA_MODE_THROUGHPUT if mode == "throughput" else _lib.QA_MODE_REFERENCE,
                                           interrupt_function=interrupt_function)
            t2 = time.perf_counter_ns()
        finally:
            gm.close()
        self.last_stats = st.as_dict()
        local_states, energies = local_states[:done], energies[:done]
        if world > 1:
            local_states, energies = _gather_reads(local_states, energies, num_reads, world)
        energies = energies + model.offset
        samples = local_states if vartype is SPIN else ((local_states + 1) // 2).astype(np.int8)
        info["timing"] = {"preprocessing_ns": t1 - t0, "sampling_ns": t2 - t1, "postprocessing_ns": 0}
        info["b200"] = {"mode": mode, "seed_mode": seed_mode, "world_size": world, **{k: self.last_stats[k] for k in
                        ("attempts", "accepted", "draws", "candidates", "nbr_updates", "near_ties", "ms_anneal")}}
        ss = SampleSet.from_samples((samples, labels), energy=energies, vartype=vartype, info=info)
        if aggregate:
            ss = ss.aggregate()
        if sort_result:
            ss = ss.sorted()
        ss.info["timing"]["postprocessing_ns"] = time.perf_counter_ns() - t2
        return ss


# ------------------------------------------------------------------------------------------------
def _coerce_bqm(obj) -> BinaryQuadraticModel:
    """Accept a real dimod BQM (anything with to_numpy_vectors / vartype / variables)."""
    if hasattr(obj, "to_numpy_vectors") and hasattr(obj, "vartype"):
        labels = list(obj.variables)
        vec = obj.to_numpy_vectors(variable_order=labels)
        ldata, (irow, icol, qdata), offset = vec[0], vec[1], vec[2]
        return BinaryQuadraticModel.from_numpy_vectors(ldata, (irow, icol, qdata), offset, as_vartype(obj.vartype), labels)
    raise TypeError(f"cannot sample an object of type {type(obj)!r}")


def _dist_rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _shard(num_reads: int, rank: int, world: int):
    """Contiguous block of reads for ``rank`` (first ``num_reads % world`` ranks take one more)."""
    base, rem = divmod(num_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _gather_reads(states: np.ndarray, energies: np.ndarray, num_reads: int, world: int):
    """All ranks end with every read (small problems); large runs should keep shards and gather only the best."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = states.shape[1]
    per = -(-num_reads // world)
    s_pad = torch.zeros((per, n), dtype=torch.int8, device=dev)
    e_pad = torch.full((per,), float("inf"), dtype=torch.float64, device=dev)
    s_pad[: states.shape[0]] = torch.from_numpy(states).to(dev)
    e_pad[: energies.shape[0]] = torch.from_numpy(energies).to(dev)
    s_all = [torch.empty_like(s_pad) for _ in range(world)]
    e_all = [torch.empty_like(e_pad) for _ in range(world)]
    dist.all_gather(s_all, s_pad)
    dist.all_gather(e_all, e_pad)
    out_s, out_e = [], []
    for r in range(world):
        lo, hi = _shard(num_reads, r, world)
        out_s.append(s_all[r][: hi - lo].cpu().numpy())
        out_e.append(e_all[r][: hi - lo].cpu().numpy())
    return np.concatenate(out_s), np.concatenate(out_e)


def _decode_discrete(ss: SampleSet, model: LoweredModel, dqm=None) -> SampleSet:
    K = model.meta["num_cases"]
    cells = model.meta["cells"]
    n = len(cells)
    case, ok = decode_onehot(ss.record.sample, n, K)
    feasible = ok.all(axis=1)
    # energy of the decoded (repaired) assignment under the lowered model
    bits = np.zeros((len(ss), model.num_variables), dtype=np.int8)
    rows = np.repeat(np.arange(len(ss)), n)
    bits[rows, (np.arange(n)[None, :] * K + case).ravel()] = 1
    energy = model.energies(2 * bits.astype(np.int16) - 1)
    out = SampleSet.from_samples((case.astype(np.int8) if K < 128 else case, cells), energy=energy, vartype=BINARY,
                                 info=dict(ss.info), is_feasible=feasible)
    out.vartype_name = "DISCRETE"
    out.info["onehot_penalty"] = model.meta.get("penalty")
    return out


def _annotate_cqm(ss: SampleSet, model: LoweredModel) -> SampleSet:
    if model.meta.get("kind") == "cqm_generic":
        cqm = model.meta["cqm"]
        feasible = cqm.check_feasible(np.asarray(ss.record.sample), ss.variables)
        out = SampleSet.from_samples((ss.record.sample, ss.variables), energy=ss.record.energy, vartype=BINARY,
                                     info=dict(ss.info), is_feasible=feasible)
        out.info["onehot_penalty"] = model.meta.get("onehot_penalty")
        out.info["size_penalty"] = model.meta.get("size_penalty")
        return out
    K = model.meta["num_cases"]
    n = len(model.meta["cells"])
    case, ok = decode_onehot(ss.record.sample, n, K)
    x = np.asarray(ss.record.sample)[:, : n * K].reshape(-1, n, K)
    sizes = x.sum(axis=1)
    feasible = ok.all(axis=1) & (sizes >= model.meta["min_size"]).all(axis=1)
    out = SampleSet.from_samples((ss.record.sample, ss.variables), energy=ss.record.energy, vartype=BINARY, info=dict(ss.info),
                                 is_feasible=feasible, cluster_sizes=sizes.astype(np.int32))
    out.info["onehot_penalty"] = model.meta.get("onehot_penalty")
    out.info["size_penalty"] = model.meta.get("size_penalty")
    return out
