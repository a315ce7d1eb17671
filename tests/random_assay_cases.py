"""Seeded cases for candidate generation: PCR::random_assay per seed stream; shared by the golden generator
and the GPU tier."""
import numpy as np

from pcramp_b200 import synth
from pcramp_b200.api import RandomAssayOptions


class RaCase:
    def __init__(self, name, coll, seeds, per_stream, opt, splits=(), inactive=()):
        self.name, self.coll, self.opt = name, coll, opt
        self.seeds = np.array(seeds, np.uint32)
        self.per = np.array(per_stream, np.uint32)
        self.splits, self.inactive = list(splits), list(inactive)

    @property
    def active(self):
        a = np.ones(self.coll.n, np.uint8)
        a[self.inactive] = 0
        return a


def ra_cases():
    out = []
    rng = np.random.default_rng(101)
    # the reference's defaults; many short streams and one long one (a long chain carries the NucCruc ring-buffer history and
    # the seed through hundreds of rejected candidates)
    coll = synth.make_targets(102, 16, 2000, n_clades=2, between=0.10, within=0.02)
    seeds = rng.integers(0, 2**32, size=41, dtype=np.uint64).astype(np.uint32)
    per = np.concatenate([rng.integers(1, 12, size=40), [300]])
    out.append(RaCase("defaults", coll, seeds, per, RandomAssayOptions()))
    # degenerate targets and primers (degen 4: expansions in Word::next order, strand concentration / degeneracy), other filter
    # settings, inactive / split / odd-length / too-short-for-two-primers sequences
    base = synth.make_targets(103, 10, 901, n_clades=2, between=0.08, within=0.02)
    codes = [base.codes(i).copy() for i in range(base.n)]
    for c in codes:
        for p in rng.integers(0, len(c), size=60):
            c[int(p)] = int(rng.choice([3, 5, 6, 9, 10, 12, 7, 15]))
    codes.append(base.codes(0)[:131].copy())
    coll2 = synth.Collection(codes)
    seeds2 = rng.integers(0, 2**32, size=24, dtype=np.uint64).astype(np.uint32)
    per2 = rng.integers(1, 20, size=24)
    opt2 = RandomAssayOptions(primer_range=(17, 28), amplicon_range=(70, 160), degen=4, salt=0.1, primer_strand=2.0e-7, primer_tm_range=(48.0, 68.0),
                              max_hairpin=35.0, max_dimer=30.0)
    out.append(RaCase("degenerate", coll2, seeds2, per2, opt2, splits=[(0, 300), (0, 301), (2, 450), (5, 100), (5, 700)], inactive=[1, 7]))
    # one stream = the reference at --thread 1
    out.append(RaCase("single_stream", coll, [20261018], [500], RandomAssayOptions(primer_tm_range=(52.0, 70.0), max_hairpin=30.0, max_dimer=25.0)))
    return out
