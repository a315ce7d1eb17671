"""Host-side sharding plan for the multi-GPU path: sequences are split into contiguous
index ranges, one per rank; every rank scores the same pairs on its shard; the shards' bitsets are
all-gathered and spliced by pcramp_gpu_merge_shards."""
import numpy as np


def shard_bounds(n_seq, world, align=1):
    """[lo_0, lo_1, ..., lo_world]: rank k owns sequences [lo_k, lo_k+1).  align = 32 puts every inner boundary on a bitset
    word, which the peer-memory exchange (xchg.cuh) needs: the shards then own disjoint words of a pair's bitset."""
    b = [min(n_seq, (n_seq * k // world + align // 2) // align * align) for k in range(world + 1)]
    b[0], b[-1] = 0, n_seq
    return b


def shard_sizes(n_seq, world, align=1):
    b = shard_bounds(n_seq, world, align)
    return np.array([b[k + 1] - b[k] for k in range(world)], dtype=np.uint32)




def better(a, b):
    """would reduce_best_assay (main.cpp:1455-1480) replace b by a?  a, b = (accuracy, overlap, degeneracy, global trial index) or None.
    Score::operator< then the smaller total degeneracy; equal on all three: the lower trial index (what one thread would have kept)."""
    if a is None:
        return False
    if b is None:
        return True
    if a[0] != b[0]:
        return a[0] > b[0]
    if a[1] != b[1]:
        return a[1] > b[1]
    if a[2] != b[2]:
        return a[2] < b[2]
    return a[3] < b[3]


def reduce_best(dist, local):
    """the best assay over all ranks: all-gather of every rank's winner (accuracy, overlap, degeneracy, global trial index; None = the
    rank has none) and the same rule on every rank -- replaces the MPI_Recv loop of reduce_best_assay.  Returns (winner, owning rank)."""
    world = dist.get_world_size() if dist is not None else 1
    if world == 1:
        return local, 0
    import torch
    rec = torch.tensor([0.0, 0.0, 0.0, -1.0] if local is None else [float(local[0]), float(local[1]), float(local[2]), float(local[3])],
                       dtype=torch.float64)
    if dist.get_backend() == "nccl":
        rec = rec.cuda()
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    best, owner = None, -1
    for k, t in enumerate(out):
        v = t.cpu().tolist()
        cand = None if v[3] < 0 else (np.float32(v[0]), np.float32(v[1]), v[2], int(v[3]))
        if better(cand, best):
            best, owner = cand, k
    return best, owner


def shard_words(n_seq, world, align=1):
    """uint32 words per pair in each shard's LSB-first bitset"""
    return [(int(n) + 31) // 32 for n in shard_sizes(n_seq, world, align)]
