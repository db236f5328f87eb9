"""World-size-2 host-side logic on CPU (gloo): read sharding, sharding-invariant seeds, gather of read shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scrna_seq_qannealing_clustering_b200 import schedule
from scrna_seq_qannealing_clustering_b200.sampler import _broadcast_int, _gather_best, _gather_reads, _shard


def test_shards_partition_reads():
    for R in (0, 1, 7, 8, 100000):
        for world in (1, 2, 4, 8):
            spans = [_shard(R, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, R, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = _shard(R, rank, world)
    seeds = schedule.per_read_seeds(99, hi - lo, first_read=lo)
    # stand-in for the per-rank anneal: states/energies are a function of the (global) per-read seed only
    states = ((seeds[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int8) * 2 - 1
    energies = (seeds % np.uint64(1000)).astype(np.float64)
    s_all, e_all = _gather_reads(states, energies, world)
    best = torch.tensor([float(energies.min()) if len(energies) else float("inf"), float(lo + int(np.argmin(energies)))],
                        dtype=torch.float64)
    gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, best)
    if rank == 0:
        np.save(out + "_s.npy", s_all)
        np.save(out + "_e.npy", e_all)
        np.save(out + "_b.npy", torch.stack(gathered).numpy())
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(tmp_path):
    R, n, world = 37, 12, 2
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "g")
    mp.spawn(_worker, args=(world, port, R, n, out), nprocs=world, join=True)
    seeds = schedule.per_read_seeds(99, R)
    want_s = ((seeds[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int8) * 2 - 1
    want_e = (seeds % np.uint64(1000)).astype(np.float64)
    assert np.array_equal(np.load(out + "_s.npy"), want_s)
    assert np.array_equal(np.load(out + "_e.npy"), want_e)
    b = np.load(out + "_b.npy")
    winner = b[np.argmin(b[:, 0])]
    assert winner[0] == want_e.min() and want_e[int(winner[1])] == want_e.min()


def _worker_best(rank, world, port, R, n, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seed = _broadcast_int(1000 + rank)          # every rank continues with rank 0's value
    lo, hi = _shard(R, rank, world)
    seeds = schedule.per_read_seeds(seed, hi - lo, first_read=lo)
    states = ((seeds[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int8) * 2 - 1
    energies = (seeds % np.uint64(50)).astype(np.float64)          # many ties: the tie-break must be the global read index
    # what one rank's device post-processing hands over: its k best rows (ascending energy, ties in read order)
    order = np.argsort(energies, kind="stable")[:k]
    e_all, rows, index, occ = _gather_best(energies, states[order], order.astype(np.int64) + lo, energies[order],
                                           np.ones(len(order), dtype=np.int64), k, world, False)
    # interrupted rank 1 completed only 3 reads: lengths differ, nothing un-annealed may come back
    done = (hi - lo) if rank == 0 else 3
    s_part, e_part = _gather_reads(states[:done], energies[:done], world)
    if rank == 0:
        np.savez(out, e_all=e_all, rows=rows, index=index, occ=occ, seed=seed, s_part=s_part, e_part=e_part)
    dist.destroy_process_group()


def test_two_rank_best_k_gather_and_ragged_shards(tmp_path):
    R, n, k, world = 41, 16, 5, 2
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    out = str(tmp_path / "b.npz")
    mp.spawn(_worker_best, args=(world, port, R, n, k, out), nprocs=world, join=True)
    d = np.load(out)
    assert int(d["seed"]) == 1000
    seeds = schedule.per_read_seeds(1000, R)
    want_s = ((seeds[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.int8) * 2 - 1
    want_e = (seeds % np.uint64(50)).astype(np.float64)
    assert np.array_equal(d["e_all"], want_e)                     # every energy, 8 bytes per read
    best = np.argsort(want_e, kind="stable")[:k]                   # global top-k, ties to the lower read index
    assert np.array_equal(d["index"], best) and np.array_equal(d["rows"], want_s[best]) and (d["occ"] == 1).all()
    lo1 = _shard(R, 1, world)[0]
    keep = np.r_[0:lo1, lo1:lo1 + 3]
    assert np.array_equal(d["s_part"], want_s[keep]) and np.array_equal(d["e_part"], want_e[keep])
