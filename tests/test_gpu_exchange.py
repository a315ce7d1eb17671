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
