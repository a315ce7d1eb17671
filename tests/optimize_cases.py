"""Seeded cases for the local search (optimize() + the six moves), shared by the golden generator and the GPU tier."""
import numpy as np

from pcramp_b200 import synth
from pcramp_b200.api import MOVES, OptimizeOptions

ALL_MOVES = [MOVES[m] for m in ("IncreaseDegeneracy", "DecreaseDegeneracy", "Trim5", "Grow5", "Trim3", "Grow3")]   # main.cpp:77-96
LENGTH_MOVES = [MOVES[m] for m in ("Trim5", "Grow5", "Trim3", "Grow3")]


class OptCase:
    def __init__(self, name, targets, f, r, moves, options, background=None):
        self.name, self.targets, self.f, self.r, self.moves, self.options, self.background = name, targets, f, r, moves, options, background
        self.optimize_5 = MOVES["Trim5"] in moves
        self.optimize_3 = MOVES["Trim3"] in moves

    @property
    def target_search(self):
        return float(np.float32(self.options.target_threshold) * np.float32(self.options.target_search_multiplier))

    @property
    def background_search(self):
        return float(np.float32(self.options.background_threshold) * np.float32(self.options.background_search_multiplier))


def cases():
    out = []
    # a diverged family: the local search trades primer length / degeneracy for coverage
    tg = synth.make_targets(81, 24, 1200, n_clades=3, between=0.05, within=0.015)
    f, r = synth.make_pairs(82, tg, 48, primer_range=(19, 24))
    out.append(OptCase("length_moves", tg, f, r, LENGTH_MOVES, OptimizeOptions()))
    out.append(OptCase("all_moves_degen4", tg, f[:32], r[:32], ALL_MOVES, OptimizeOptions(degen=4)))
    bg = synth.make_targets(83, 10, 1200, n_clades=2, between=0.12, within=0.05)
    codes = [bg.codes(i).copy() for i in range(bg.n)]
    for i in range(0, bg.n, 2):
        codes[i][100:1000] = tg.codes(i % tg.n)[100:1000]      # near-neighbour backgrounds share most of the target
    bgc = synth.Collection(codes)
    out.append(OptCase("background_taq", tg, f[:32], r[:32], ALL_MOVES, OptimizeOptions(degen=2, use_taq_mama=1), background=bgc))
    f2, r2 = synth.make_pairs(84, tg, 32, primer_range=(18, 25), degenerate_fraction=0.6)
    out.append(OptCase("degenerate_start", tg, f2, r2, ALL_MOVES, OptimizeOptions(degen=8, primer_tm_min=45.0, max_hairpin=50.0)))
    return out
