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




def shard_words(n_seq, world, align=1):
    """uint32 words per pair in each shard's LSB-first bitset"""
    return [(int(n) + 31) // 32 for n in shard_sizes(n_seq, world, align)]
