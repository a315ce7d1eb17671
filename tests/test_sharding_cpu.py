"""CPU, world_size 2 over gloo: the N>1 protocol of bench.py -- shard the sequences, score the same pairs per
shard, all-gather the shards' (any, pass-1) bitsets, splice + re-sum -- with the CPU oracle standing in for the
GPU scoring call and a numpy restatement of merge_shards_kernel.  The merged result must equal the unsharded one
bit for bit, coverage included (non-uniform weights)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcramp_b200 import synth
from pcramp_b200.sharding import shard_bounds, shard_sizes, shard_words


def merge_model(gathered_any, gathered_p1, sizes, weights, n_pairs):
    """numpy restatement of merge_shards_kernel: splice LSB-first shard bitsets, sum weights in the reference's order"""
    n = int(sum(sizes))
    bits = np.zeros((n_pairs, n), np.uint8)
    cov = np.zeros(n_pairs, np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
    for p in range(n_pairs):
        acc = 0.0
        for which in (0, 1):
            for s, ns in enumerate(sizes):
                a = np.unpackbits(gathered_any[s][p].view(np.uint8), bitorder="little")[:ns]
                b = np.unpackbits(gathered_p1[s][p].view(np.uint8), bitorder="little")[:ns]
                m = b if which == 0 else (a & (1 - b))
                for i in np.nonzero(m)[0]:
                    bits[p, off[s] + i] = 1
                    acc += float(weights[off[s] + i])
        cov[p] = np.float32(acc)
    return cov, bits


def pack_lsb(bits01):
    n_pairs, n = bits01.shape
    words = (n + 31) // 32
    padded = np.zeros((n_pairs, words * 32), np.uint8)
    padded[:, :n] = bits01
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32)


def scenario():
    coll = synth.make_targets(901, 11, 900, n_clades=2, between=0.1, within=0.05)
    coll.weight = np.random.default_rng(3).uniform(0.2, 2.5, size=coll.n).astype(np.float32)
    f, r = synth.make_pairs(902, coll, 24)
    return coll, f, r


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.harness import OracleLib
    coll, f, r = scenario()
    thr = float(np.float32(1.0) * np.float32(0.9))
    b = shard_bounds(coll.n, world)
    shard = coll.subset(range(b[rank], b[rank + 1]))
    o = OracleLib()
    o.set_sequences(shard)
    o.select_words(f, r, thr)
    _, bits = o.score_pairs(f, r, thr, 1.0, 80, 200, False)
    # the oracle exposes the union only; for unit tests of the exchange use pass-1 := any (the order of summation
    # is then "ascending over all shards", which is also what the unsharded oracle does when pass 2 finds nothing new)
    words = shard_words(coll.n, world)
    mine = torch.from_numpy(pack_lsb(bits).astype(np.int64))
    max_words = max(words)
    padded = torch.zeros((len(f), max_words), dtype=torch.int64)
    padded[:, :words[rank]] = mine
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    if rank == 0:
        g = [gathered[s][:, :words[s]].numpy().astype(np.uint32) for s in range(world)]
        cov, merged = merge_model(g, g, shard_sizes(coll.n, world), coll.weight, len(f))
        np.savez(out, cov=cov, bits=merged)
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 20000):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            assert int(shard_sizes(n, w).sum()) == n


def test_two_rank_exchange_equals_unsharded(tmp_path, oracle):
    out = str(tmp_path / "merged.npz")
    mp.spawn(worker, args=(2, free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    coll, f, r = scenario()
    thr = float(np.float32(1.0) * np.float32(0.9))
    oracle.set_sequences(coll)
    oracle.select_words(f, r, thr)
    cov, bits = oracle.score_pairs(f, r, thr, 1.0, 80, 200, False)
    assert bits.sum() > 0
    assert np.array_equal(got["bits"], bits)
    # select_words is per sequence and pairing never crosses sequences, so sharding changes nothing; with
    # pass-1 := any the merged sum visits detected sequences in ascending order, as the oracle does here
    want = np.array([np.float32(sum(float(coll.weight[i]) for i in np.nonzero(bits[p])[0])) for p in range(len(f))], np.float32)
    assert np.array_equal(got["cov"], want)


# ---- the best assay over ranks (reduce_best_assay, main.cpp:1421-1601) ----------------------------------------------------------
def fold_best(records):
    """the reference's update rule (main.cpp:829-858) folded over (accuracy, overlap, degeneracy, index) records in order"""
    from pcramp_b200.sharding import better
    best = None
    for rec in records:
        if better(rec, best):
            best = rec
    return best


def best_scenario():
    rng = np.random.default_rng(11)
    n = 64
    acc = rng.integers(0, 4, size=n).astype(np.float32)          # many ties
    ov = (rng.integers(0, 3, size=n) * 0.5).astype(np.float32)
    deg = rng.integers(2, 5, size=n).astype(np.float64)
    return [(acc[i], ov[i], deg[i], i) for i in range(n)]


def best_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcramp_b200.sharding import reduce_best
    recs = best_scenario()
    mine = recs[rank::world]                                     # the pair-sharded sweep: every rank scores its own trials
    local = fold_best(mine) if rank != 1 or world == 1 else fold_best(mine)
    best, owner = reduce_best(dist, local)
    none_best, none_owner = reduce_best(dist, None if rank == 0 else local)   # a rank without a candidate
    if rank == 0:
        np.savez(out, best=np.array(best, np.float64), owner=owner, none_best=np.array(none_best, np.float64), none_owner=none_owner)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_best_assay_equals_single_fold(tmp_path):
    out = str(tmp_path / "best.npz")
    mp.spawn(best_worker, args=(2, free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    recs = best_scenario()
    want = fold_best(recs)                                       # one process, trials in order
    assert tuple(got["best"]) == tuple(float(x) for x in want)
    assert int(got["owner"]) == want[3] % 2
    want1 = fold_best(recs[1::2])
    assert tuple(got["none_best"]) == tuple(float(x) for x in want1) and int(got["none_owner"]) == 1


def test_rank_fold_equals_reference_golden():
    """sharding.better (the rule sharding.reduce_best applies to the gathered records) against the reference's own fold
    (main.cpp:1455-1480 through oracle/ref_driver.cpp::ref_reduce_best, golden kat_best_assay.npz): records in rank order,
    trial index := rank, so a full tie keeps the lower rank as the root's receive loop does"""
    from tests import best_assay_cases
    from pcramp_b200.sharding import better
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_best_assay.npz"))
    n = 0
    for name, score, deg, valid in best_assay_cases.rank_cases():
        best, owner = None, 0
        for k in range(len(deg)):
            rec = (np.float32(score[k, 0] - score[k, 1]), score[k, 2], deg[k], k) if valid[k] else None
            if better(rec, best):
                best, owner = rec, k
        if valid.any():
            assert owner == int(gold["rank_" + name][0]), name
            n += 1
    assert n > 20


def test_rank_fold_golden_matches_live_reference(ref):
    from tests import best_assay_cases
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_best_assay.npz"))
    for name, score, deg, valid in best_assay_cases.rank_cases():
        assert ref.reduce_best(score, deg) == int(gold["rank_" + name][0])
    for name, tgt, bg, ov, f, r, max_bg in best_assay_cases.trial_cases():
        assert np.array_equal(np.array(ref.best_assay(tgt, bg, ov, f, r, max_bg), np.float64), gold["trial_" + name])
