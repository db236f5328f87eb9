"""CPU tests of the replay kernel's host logic (csrc/anneal_replay.cu::pack_replay_slabs through the host-only C-ABI hook
qa_debug_pack_slabs): block invariants of the coupling slabs, and -- the strong one -- a pure-Python replay that consumes
ONLY the packed slabs in the kernel's two-phase order (pre parts of a block against the state at block start, then seq
parts + decisions) and must reproduce the oracle's final states bit for bit.
No GPU is needed: the hook makes no CUDA call."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import oracle
from scrna_seq_qannealing_clustering_b200 import _lib, models, schedule, snn

RP_D, RP_SLOTS, RP_CAP = 16, 64, 448
RP_MAXBW = RP_SLOTS - 1
HDR = np.dtype([("nent", "<i4"), ("nbw", "<i4"), ("v0", "<i4"), ("nv", "<i4"), ("prev_slot", "<i4"), ("seq_off", "<i4"),
                ("nbw_next", "<i4"), ("pad", "<i4"), ("rowa", "<u4", RP_D), ("rowb", "<u4", RP_D), ("ga", "<i4", RP_D),
                ("bw", "<i4", RP_SLOTS), ("bw_next", "<i4", RP_SLOTS)])
ENT = np.dtype([("J", "<f8"), ("zero", "<u4"), ("B", "<u4")])
assert HDR.itemsize == 736 and ENT.itemsize == 16


def adjacency(n, starts, ends, weights):
    """neal's adjacency lists: every coupler appended to both endpoints in coupler order."""
    rows = [[] for _ in range(n)]
    for a, b, w in zip(starts, ends, weights):
        rows[a].append((int(b), float(w)))
        rows[b].append((int(a), float(w)))
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.array([j for r in rows for j, _ in r], dtype=np.int32)
    val = np.array([w for r in rows for _, w in r], dtype=np.float64)
    return rows, rowptr, col, val


def pack(model):
    n = model.num_variables
    rows, rowptr, col, val = adjacency(n, model.starts, model.ends, model.weights)
    lib = _lib.load()
    ns, nb, uni = C.c_int64(), C.c_int64(), C.c_int32()
    grp = coef = None
    ng = 0
    if model.groups is not None:
        grp = np.ascontiguousarray(model.groups.grp, dtype=np.int32)
        coef = np.ascontiguousarray(model.groups.coef, dtype=np.int32)
        ng = len(model.groups.lam)
    args = [n, _lib.ptr(rowptr), _lib.ptr(col if len(col) else np.zeros(1, np.int32)), _lib.ptr(val if len(val) else np.zeros(1)), ng,
            _lib.ptr(grp), _lib.ptr(coef), C.byref(ns), C.byref(nb), C.byref(uni)]
    rc = lib.qa_debug_pack_slabs(*args, None, None)
    assert rc >= 0
    if rc == 0:
        return None
    slabs = np.zeros(nb.value, dtype=np.uint8)
    off = np.zeros(ns.value + 1, dtype=np.uint32)
    assert lib.qa_debug_pack_slabs(*args, _lib.ptr(slabs), _lib.ptr(off)) == 1
    blocks = []
    for b in range(ns.value):
        o = int(off[b]) * 16
        hdr = np.frombuffer(slabs, dtype=HDR, count=1, offset=o)[0]
        ent = np.frombuffer(slabs, dtype=ENT, count=int(hdr["nent"]), offset=o + HDR.itemsize)
        assert int(off[b + 1]) * 16 == o + HDR.itemsize + ENT.itemsize * int(hdr["nent"])
        blocks.append((hdr, ent))
    return rows, blocks, bool(uni.value)


def row_parts(hdr, ent, i):
    """(pre entries without padding, padding entries, seq entries) of row i of a block, in slab order."""
    pre0 = 4 * sum(int(hdr["rowa"][k]) & 255 for k in range(i))
    rounds = int(hdr["rowa"][i]) & 255
    npre = int(hdr["rowb"][i]) >> 16
    seq0 = int(hdr["seq_off"]) + sum((int(hdr["rowa"][k]) >> 8) & 255 for k in range(i))
    nseq = (int(hdr["rowa"][i]) >> 8) & 255
    return ent[pre0:pre0 + npre], ent[pre0 + npre:pre0 + 4 * rounds], ent[seq0:seq0 + nseq]


def check_invariants(model, rows, blocks):
    n = model.num_variables
    npad = (n + 15) // 16 * 16
    v = 0
    for b, (hdr, ent) in enumerate(blocks):
        v0, nv = int(hdr["v0"]), int(hdr["nv"])
        assert v0 == v and 1 <= nv <= RP_D and (v0 >> 4) == ((v0 + nv - 1) >> 4), "blocks tile the variables inside one half-word"
        assert int(hdr["nbw"]) <= RP_MAXBW and int(hdr["nent"]) <= RP_CAP and int(hdr["seq_off"]) % 4 == 0
        own = v0 >> 4
        words = list(hdr["bw"][: int(hdr["nbw"])])
        assert own not in words and len(set(words)) == len(words)
        nxt = blocks[(b + 1) % len(blocks)][0]
        assert int(hdr["nbw_next"]) == int(nxt["nbw"]) and np.array_equal(hdr["bw_next"], nxt["bw"])
        # the slot of the previous block's half-word (the one word asynchronous staging cannot have up to date)
        prev_hw = (int(blocks[b - 1][0]["v0"]) >> 4) if b > 0 else None
        want_prev = words.index(prev_hw) + 1 if (prev_hw is not None and prev_hw in words) else 0
        assert int(hdr["prev_slot"]) == want_prev
        total_pre = 0
        for i in range(nv):
            u = v0 + i
            row = rows[u] if u < n else []
            later = sorted([(j, k) for k, (j, _) in enumerate(row) if j > u])      # stable: ties keep adjacency order
            early = sorted([(j, k) for k, (j, _) in enumerate(row) if j < v0])
            inblk = sorted([(j, k) for k, (j, _) in enumerate(row) if v0 <= j < u])
            pre, pad, seq = row_parts(hdr, ent, i)
            assert [(int(e["B"]) >> 13, float(e["J"])) for e in pre] == [(j, row[k][1]) for j, k in later + early], f"row {u}: pre part"
            assert [(int(e["B"]) >> 13, float(e["J"])) for e in seq] == [(j, row[k][1]) for j, k in inblk], f"row {u}: seq part"
            assert all(float(e["J"]) == 0.0 for e in pad) and len(pre) + len(pad) == 4 * (int(hdr["rowa"][i]) & 255)
            assert int(hdr["rowa"][i]) >> 16 == len(row) and (int(hdr["rowb"][i]) & 0xFFFF) == len(later)
            for e in list(pre) + list(seq):
                j, B = int(e["B"]) >> 13, int(e["B"])
                assert int(e["zero"]) == 0 and (B & 31) == 30 - 2 * (j & 15)
                slot = (B >> 7) & 63
                assert (slot == 0 and (j >> 4) == own) or (slot > 0 and words[slot - 1] == (j >> 4))
            for e in pad:   # zero coupling: any readable slot will do
                assert int(e["zero"]) == 0 and ((int(e["B"]) >> 7) & 63) == 0
            total_pre += 4 * (int(hdr["rowa"][i]) & 255)
            if model.groups is not None and u < n and model.groups.grp[u] >= 0:
                ga = int(hdr["ga"][i])
                assert (ga & 255) == model.groups.grp[u] and (ga >> 8) == model.groups.coef[u]
            else:
                assert (int(hdr["ga"][i]) & 255) == 255
        assert total_pre == int(hdr["seq_off"])
        v += nv
    assert v == npad


MASK = (1 << 64) - 1


def _rng_next(s):
    x, y = s
    s[0] = y
    x ^= (x << 23) & MASK
    s[1] = x ^ y ^ (x >> 17) ^ (y >> 26)
    return (s[1] + y) & MASK


def replay_from_slabs(model, rows, blocks, state, betas, spb, seed):
    """Deferred exact replay driven by the slab entries only (fields start from neal's get_flip_energy order)."""
    n = model.num_variables
    s = state.astype(np.int64).copy()
    f = np.empty(n)
    for v in range(n):
        acc = model.h[v]
        for j, J in rows[v]:
            acc = acc + (J if s[j] > 0 else -J)
        f[v] = acc
    F = np.zeros(n, dtype=bool)
    st = [seed if seed else MASK, 0]
    for beta in betas:
        thr = 44.36142 / beta
        for _ in range(spb):
            for hdr, ent in blocks:
                v0, nv = int(hdr["v0"]), int(hdr["nv"])
                # batch phase: the pre parts of all rows see the state as of the START of the block
                Fb, sb = F.copy(), s.copy()
                part = []
                for i in range(nv):
                    pre, pad, _ = row_parts(hdr, ent, i)
                    acc = f[v0 + i] if v0 + i < n else 0.0
                    for e in list(pre) + list(pad):
                        u = int(e["B"]) >> 13
                        sigma = (2.0 if sb[u] > 0 else -2.0) if Fb[u] else 0.0
                        acc = acc + float(e["J"]) * sigma
                    part.append(acc)
                # sequential phase: in-block earlier neighbours (decided in this block), then the decision
                for i in range(nv):
                    v = v0 + i
                    if v >= n:
                        break
                    fv = part[i]
                    for e in row_parts(hdr, ent, i)[2]:
                        u = int(e["B"]) >> 13
                        if F[u]:
                            fv = fv + float(e["J"]) * (2.0 if s[u] > 0 else -2.0)
                    f[v] = fv
                    dE = -2.0 * s[v] * f[v]
                    flip = False
                    if not (dE >= thr):
                        if dE <= 0.0:
                            flip = True
                        else:
                            flip = math.exp(-dE * beta) * 18446744073709551616.0 > float(_rng_next(st))
                    F[v] = flip
                    if flip:
                        s[v] = -s[v]
    return s.astype(np.int8)


@pytest.fixture(scope="module")
def graph256():
    return snn.synthetic_snn(256, k=5, seed=3)[0]


def test_sparse_models_pack_into_full_blocks(graph256):
    for model in (models.subsampling_model(graph256, 7.0), models.cut_linear_model(graph256, 0.01, 1.0)):
        rows, blocks, uniform = pack(model)
        assert uniform and len(blocks) == 256 // RP_D
        check_invariants(model, rows, blocks)


def test_cqm_blocks_carry_the_group_metadata(graph256):
    model = models.cqm_model(graph256, 8, min_size=20)
    rows, blocks, uniform = pack(model)
    check_invariants(model, rows, blocks)
    assert uniform


def test_four_way_model_packs_into_smaller_blocks():
    """Config 2: 4 bits per cell -> 16 variables span 4 cells and up to 60 neighbour words; the packer closes blocks early."""
    g = snn.synthetic_snn(2048, k=5, seed=0)[0]
    model = models.dqm_model(g, 4, 0.005, semantics="intended")
    rows, blocks, uniform = pack(model)
    check_invariants(model, rows, blocks)
    sizes = np.array([int(h["nv"]) for h, _ in blocks])
    assert sizes.min() >= 1 and sizes.mean() >= 4


def test_dense_models_are_rejected_and_scattered_ones_use_the_wide_slot_format(graph256):
    assert pack(models.cut_balance_model(graph256, 0.05, structured=False)) is None          # K_256: one variable per block
    # 15 neighbours per cell scattered over 256 half-words: too many for 31 foreign slots, packs with 63 (smaller blocks)
    g = snn.synthetic_snn(4096, k=5, seed=1)[0]
    m = models.cut_linear_model(g, 0.01, 1.0)
    rows, blocks, uniform = pack(m)
    check_invariants(m, rows, blocks)
    assert not uniform and max(int(h["nbw"]) for h, _ in blocks) > 31
    # random graph of degree ~60: not even four rows share 63 half-words
    rng = np.random.default_rng(5)
    n = 4096
    pairs = {(int(max(u, v)), int(min(u, v))) for u, v in rng.integers(0, n, size=(30 * n, 2)) if u != v}
    pairs = sorted(pairs)
    dense = models.LoweredModel(np.zeros(n), np.array([p[0] for p in pairs], dtype=np.int32), np.array([p[1] for p in pairs], dtype=np.int32),
                                rng.normal(size=len(pairs)), 0.0, list(range(n)))
    assert pack(dense) is None


def test_shuffled_and_duplicated_couplers_keep_adjacency_order_among_ties():
    rng = np.random.default_rng(0)
    n = 150
    pairs = [(u, v) for u in range(130) for v in range(u) if rng.random() < 0.06]
    pairs += pairs[:25]
    rng.shuffle(pairs)
    flip = rng.random(len(pairs)) < 0.5
    starts = np.array([p[1] if f else p[0] for p, f in zip(pairs, flip)], dtype=np.int32)
    ends = np.array([p[0] if f else p[1] for p, f in zip(pairs, flip)], dtype=np.int32)
    model = models.LoweredModel(rng.normal(size=n), starts, ends, rng.normal(size=len(pairs)), 0.0, list(range(n)))
    rows, blocks, _ = pack(model)
    check_invariants(model, rows, blocks)


@pytest.mark.parametrize("case", ["subsampling", "shuffled"])
def test_python_replay_of_the_packed_slabs_equals_the_oracle(case):
    rng = np.random.default_rng(7)
    if case == "subsampling":
        model = models.subsampling_model(snn.synthetic_snn(96, k=5, seed=4)[0], 7.0)
    else:
        n = 70
        pairs = [(u, v) for u in range(64) for v in range(u) if rng.random() < 0.12]
        pairs += pairs[:10]
        rng.shuffle(pairs)
        starts = np.array([p[0] for p in pairs], dtype=np.int32)
        ends = np.array([p[1] for p in pairs], dtype=np.int32)
        model = models.LoweredModel(rng.normal(size=n) / 4, starts, ends, rng.normal(size=len(pairs)) / 3, 0.0, list(range(n)))
    rows, blocks, _ = pack(model)
    betas, spb = schedule.make_beta_schedule((0.05, 8.0), 40, 1, "geometric")
    R = 3
    seeds = schedule.per_read_seeds(21, R)
    init = schedule.random_spin_states(R, model.num_variables, 5)
    ref = init.copy()
    oracle.sample_ising(model.h, model.starts, model.ends, model.weights, ref, betas, spb, seeds)
    for r in range(R):
        out = replay_from_slabs(model, rows, blocks, init[r], betas, spb, int(seeds[r]))
        assert np.array_equal(out, ref[r]), r


@pytest.mark.parametrize("seed", range(6))
def test_random_graphs_invariants_and_replay(seed):
    """Random sparse models (ragged sizes, shuffled / duplicated couplers, isolated variables, zero weights): the packed slabs
    satisfy the block invariants and the Python replay driven by them reproduces the oracle."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 140))
    m = int(min(n * (n - 1) // 2, rng.integers(0, 4 * n + 1)))
    pairs = set()
    while len(pairs) < m:
        u, v = rng.integers(0, n, 2)
        if u != v:
            pairs.add((int(max(u, v)), int(min(u, v))))
    pairs = sorted(pairs)
    if pairs and seed % 2:
        pairs += [pairs[i] for i in rng.integers(0, len(pairs), size=min(5, len(pairs)))]   # duplicated couplers
        rng.shuffle(pairs)
    flip = rng.random(len(pairs)) < 0.5
    starts = np.array([p[1] if f else p[0] for p, f in zip(pairs, flip)], dtype=np.int32)
    ends = np.array([p[0] if f else p[1] for p, f in zip(pairs, flip)], dtype=np.int32)
    w = rng.normal(size=len(pairs)) / 2
    if len(w):
        w[rng.integers(0, len(w))] = 0.0
    model = models.LoweredModel(rng.normal(size=n) / 3, starts, ends, w, 0.0, list(range(n)))
    packed = pack(model)
    if packed is None:
        pytest.skip("model does not fit the slab format (fewer than 4 variables per block)")
    rows, blocks, _ = packed
    check_invariants(model, rows, blocks)
    betas, spb = schedule.make_beta_schedule((0.05, 6.0), 12, 2, "geometric")
    seeds = schedule.per_read_seeds(seed, 2)
    init = schedule.random_spin_states(2, n, seed)
    ref = init.copy()
    oracle.sample_ising(model.h, model.starts, model.ends, model.weights, ref, betas, spb, seeds)
    for r in range(2):
        assert np.array_equal(replay_from_slabs(model, rows, blocks, init[r], betas, spb, int(seeds[r])), ref[r])
