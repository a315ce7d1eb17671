"""Host-side sharding plan for the multi-GPU path: sequences are split into contiguous
index ranges, one per rank; every rank scores the same pairs on its shard; the shards' bitsets are
all-gathered and spliced by pcramp_gpu_merge_shards."""
import numpy as np


def shard_bounds(n_seq, world):
    """[lo_0, lo_1, ..., lo_world]: rank k owns sequences [lo_k, lo_k+1)"""
    return [n_seq * k // world for k in range(world + 1)]


def shard_sizes(n_seq, world):
    b = shard_bounds(n_seq, world)
    return np.array([b[k + 1] - b[k] for k in range(world)], dtype=np.uint32)


def shard_words(n_seq, world):
    """uint32 words per pair in each shard's LSB-first bitset"""
    return [(int(n) + 31) // 32 for n in shard_sizes(n_seq, world)]
