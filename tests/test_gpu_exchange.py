"""GPU (-m gpu): the peer-memory exchange of the multi-GPU path (pcramp_b200/csrc/xchg.cuh).  Ranks are played by separate
library contexts -- in one process (peer pointers) and in two processes (cudaIpc handles, the torchrun arrangement) -- on
whatever GPUs the box has (all on cuda:0 when there is one).  The merged bitsets and coverage must equal the unsharded
pcramp_gpu_score_pairs result bit for bit, weighted and unweighted, over several steps (double buffering)."""
import os

import numpy as np
import pytest

from pcramp_b200 import TARGET, synth
from pcramp_b200.sharding import shard_bounds, shard_sizes

pytestmark = pytest.mark.gpu

THR = float(np.float32(1.0) * np.float32(0.9))


def scenario(weighted, n=150):
    coll = synth.make_targets(961, n, 2500, n_clades=3, between=0.12, within=0.05)
    if weighted:
        coll.weight = np.random.default_rng(6).uniform(0.1, 3.0, size=coll.n).astype(np.float32)
    f, r = synth.make_pairs(962, coll, 120)
    return coll, f, r


def unsharded(gpu, coll, f, r):
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length, coll.weight)
    gpu.select_words(TARGET, f, r, THR)
    return gpu.score_pairs(TARGET, f, r, THR, 0.9)


@pytest.mark.parametrize("world,weighted", [(1, False), (2, False), (3, True), (4, True)])
def test_exchange_between_contexts_equals_unsharded(gpu, world, weighted):
    import torch
    from pcramp_b200 import PcrampGpu
    coll, f, r = scenario(weighted)
    n_dev = torch.cuda.device_count()
    b = shard_bounds(coll.n, world, align=32)
    sizes = shard_sizes(coll.n, world, align=32)
    assert all(int(x) % 32 == 0 for x in sizes[:-1]) and int(sizes.sum()) == coll.n
    ranks = [PcrampGpu(k % n_dev) for k in range(world)]
    try:
        for k, g in enumerate(ranks):
            sh = coll.subset(range(b[k], b[k + 1]))
            g.upload_sequences(TARGET, sh.nibbles, sh.byte_off, sh.length, sh.weight)
            g.exchange_create(k, world, sizes, 64, coll.weight)
        ptrs = [g.exchange_buffer() for g in ranks]
        for g in ranks:
            g.exchange_connect_pointers(ptrs)
        for step, (lo, hi) in enumerate([(0, 64), (64, 120), (30, 94)]):      # three batches: both buffers get re-used
            fb, rb = f[lo:hi], r[lo:hi]
            for g in ranks:
                g.stage_pairs(fb, rb)
                g.select_words_staged(TARGET, THR, want_keys=False)
                g.score_pairs_staged(TARGET, THR, 0.9)
            # one host thread plays every rank here, so the (asynchronous, allocation-free) exchange launches go last: a
            # cudaMalloc of the next rank's scoring call would otherwise wait for the previous rank's spinning wait_kernel
            for g in ranks:
                g.exchange_step(TARGET)
            cov_all, bits_all = unsharded(gpu, coll, fb, rb)
            for g in ranks:
                cov, bits = g.exchange_fetch(hi - lo)
                assert np.array_equal(bits, bits_all), (step, world)
                assert np.array_equal(cov.view(np.uint32), cov_all.view(np.uint32)), (step, world)
            assert bits_all.any()
    finally:
        for g in ranks:
            g.close()


def test_exchange_rejects_unaligned_shards(gpu):
    with pytest.raises(RuntimeError, match="multiple of 32"):
        gpu.exchange_create(0, 2, np.array([50, 100], np.uint32), 16)


def _ipc_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcramp_b200 import PcrampGpu
    dev = rank % torch.cuda.device_count()
    coll, f, r = scenario(True)
    b = shard_bounds(coll.n, world, align=32)
    sizes = shard_sizes(coll.n, world, align=32)
    g = PcrampGpu(dev)
    sh = coll.subset(range(b[rank], b[rank + 1]))
    g.upload_sequences(TARGET, sh.nibbles, sh.byte_off, sh.length, sh.weight)
    g.exchange_create(rank, world, sizes, len(f), coll.weight)
    handles = [None] * world
    dist.all_gather_object(handles, g.exchange_ipc_handle())
    g.exchange_connect_ipc(handles)
    dist.barrier()
    res = []
    for lo, hi in [(0, 60), (60, 120), (20, 100)]:
        g.stage_pairs(f[lo:hi], r[lo:hi])
        g.select_words_staged(TARGET, THR, want_keys=False)
        g.score_pairs_staged(TARGET, THR, 0.9)
        g.exchange_step(TARGET)
        res.append(g.exchange_fetch(hi - lo))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{"cov%d" % i: c for i, (c, _) in enumerate(res)},
             **{"bits%d" % i: x for i, (_, x) in enumerate(res)})
    dist.barrier()          # nobody frees its buffer while a peer may still push into it
    g.close()
    dist.destroy_process_group()


def test_exchange_between_processes_over_ipc(gpu, tmp_path):
    import torch.multiprocessing as mp
    from tests.test_sharding_cpu import free_port
    world = 2
    mp.spawn(_ipc_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    coll, f, r = scenario(True)
    for i, (lo, hi) in enumerate([(0, 60), (60, 120), (20, 100)]):
        cov_all, bits_all = unsharded(gpu, coll, f[lo:hi], r[lo:hi])
        for rank in range(world):
            got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
            assert np.array_equal(got["bits%d" % i], bits_all)
            assert np.array_equal(got["cov%d" % i].view(np.uint32), cov_all.view(np.uint32))


def test_reduce_best_over_peer_memory_equals_reference_fold(gpu):
    """pcramp_gpu_reduce_best (one kernel: push the 32-byte record to every rank, wait, fold) against the reference's receive loop
    (main.cpp:1455-1480 via ref_reduce_best goldens): every rank must report the same owner and the owner's record"""
    import threading
    import torch
    from pcramp_b200 import PcrampGpu
    from tests import best_assay_cases
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_best_assay.npz"))
    n_dev = torch.cuda.device_count()
    for world in (1, 2, 3, 4, 8):
        cases = [c for c in best_assay_cases.rank_cases() if len(c[2]) == world]
        ranks = [PcrampGpu(k % n_dev) for k in range(world)]
        try:
            sizes = np.full(world, 32, np.uint32)
            for k, g in enumerate(ranks):
                g.exchange_create(k, world, sizes, 4)
            ptrs = [g.exchange_buffer() for g in ranks]
            for g in ranks:
                g.exchange_connect_pointers(ptrs)
            for name, score, deg, valid in cases:                      # several rounds on one exchange: both record buffers re-used
                out = [None] * world

                def call(k):
                    out[k] = ranks[k].reduce_best(score[k, 0], score[k, 1], score[k, 2], deg[k], 1000 + k if valid[k] else -1)
                th = [threading.Thread(target=call, args=(k,)) for k in range(world)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                owner = int(gold["rank_" + name][0])
                for k in range(world):
                    assert out[k] is not None and out[k][0] == owner, (name, k, out[k])
                    assert (out[k][1], out[k][2], out[k][3], out[k][4]) == (score[owner, 0], score[owner, 1], score[owner, 2], deg[owner]), name
                    assert out[k][5] == (1000 + owner if valid[owner] else -1)
        finally:
            for g in ranks:
                g.close()


def test_exchange_guards(gpu):
    """advisor findings: create refuses while an exchange exists, connect refuses a second call, fetch sizes from the library"""
    sizes = np.array([32], np.uint32)
    gpu.exchange_create(0, 1, sizes, 8)
    try:
        with pytest.raises(RuntimeError, match="exists"):
            gpu.exchange_create(0, 1, sizes, 8)
        gpu.exchange_connect_pointers([gpu.exchange_buffer()])
        with pytest.raises(RuntimeError, match="already connected"):
            gpu.exchange_connect_pointers([gpu.exchange_buffer()])
        assert gpu.exchange_status() == 0
        cov, bits = gpu.exchange_fetch()
        assert len(cov) == 0
    finally:
        gpu.exchange_destroy()
