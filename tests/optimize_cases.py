"""Seeded cases for the local search (optimize() + the six moves), shared by the golden generator and the GPU tier."""
import numpy as np

from pcramp_b200 import synth
from pcramp_b200.api import MOVES, OptimizeOptions

ALL_MOVES = [MOVES[m] for m in ("IncreaseDegeneracy", "DecreaseDegeneracy", "Trim5", "Grow5", "Trim3", "Grow3")]   # main.cpp:77-96
LENGTH_MOVES = [MOVES[m] for m in ("Trim5", "Grow5", "Trim3", "Grow3")]


class OptCase:
    def __init__(self, name, targets, f, r, moves, options, background=None, multiplex=None, pool=None):
        self.name, self.targets, self.f, self.r, self.moves, self.options, self.background = name, targets, f, r, moves, options, background
        self.multiplex, self.pool = multiplex, pool       # multiplex background amplicons (Collection), assay pool (pool_f, pool_r)
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


def multiplex_cases():
    """optimize() for a LATER assay of a multiplex run (main.cpp:989-1017): amplicons of earlier assays form the multiplex
    background (packed whole), their oligos the pool (oligo re-use bonus, Word::max_overlap)."""
    out = []
    tg = synth.make_targets(81, 24, 1200, n_clades=3, between=0.05, within=0.015)
    f, r = synth.make_pairs(82, tg, 40, primer_range=(19, 24))
    amps = [tg.codes(0)[0:420].copy(), tg.codes(5)[380:830].copy(), tg.codes(10)[760:1200].copy(), tg.codes(7)[200:231].copy(),
            tg.codes(13)[500:760].copy(), tg.codes(2)[30:47].copy()]
    amps[1][200] = 0                      # an EOS inside an amplicon (a split that was applied before the amplicon was cut)
    amps[4][17] = 1 | 4                   # a degenerate base
    mpx = synth.Collection(amps)
    short = lambda w, a, b: np.array(synth.word_from_codes(_codes(w)[a:b]), np.uint64)   # noqa: E731
    pool_f = np.stack([f[0], short(f[3], 1, 99), f[11], short(r[20], 0, -2)])
    pool_r = np.stack([r[1], r[3], f[7], short(f[20], 2, 99)])
    out.append(OptCase("mpx_length_moves", tg, f, r, LENGTH_MOVES, OptimizeOptions(use_multiplex=1), multiplex=mpx, pool=(pool_f, pool_r)))
    out.append(OptCase("mpx_all_moves", tg, f[:24], r[:24], ALL_MOVES, OptimizeOptions(degen=4, use_multiplex=1, use_taq_mama=1),
                       multiplex=mpx, pool=(pool_f, pool_r)))
    out.append(OptCase("mpx_pool_only", tg, f[:24], r[:24], ALL_MOVES, OptimizeOptions(degen=2, use_multiplex=1), pool=(pool_f, pool_r)))
    empty = (np.zeros((0, 2), np.uint64), np.zeros((0, 2), np.uint64))
    out.append(OptCase("mpx_keys_only", tg, f[:24], r[:24], LENGTH_MOVES, OptimizeOptions(use_multiplex=1), multiplex=mpx, pool=empty))
    return out


def _codes(w):
    """the non-EOS nibbles of a word, 5' -> 3'"""
    hi, lo = int(w[0]), int(w[1])
    c = [((hi if i < 16 else lo) >> ((15 - (i % 16)) * 4)) & 15 for i in range(32)]
    return np.array([x for x in c if x], np.uint8)


def overlap_words(seed=17, n=600):
    """word pairs for Word::max_overlap: related (shifted / trimmed / mutated copies) and unrelated, degenerate bases, short words"""
    rng = np.random.default_rng(seed)
    a = np.zeros((n, 2), np.uint64)
    b = np.zeros((n, 2), np.uint64)
    for i in range(n):
        la = int(rng.integers(1, 33))
        ca = rng.integers(1, 16, size=la).astype(np.uint8) if i % 5 == 0 else synth.CODE[rng.integers(0, 4, size=la)]
        kind = i % 4
        if kind == 0:
            cb = synth.CODE[rng.integers(0, 4, size=int(rng.integers(1, 33)))]
        else:
            lo = int(rng.integers(0, max(1, la // 2)))
            cb = ca[lo:la - int(rng.integers(0, max(1, la // 3)))].copy()
            if len(cb) == 0:
                cb = ca.copy()
            if kind == 2:
                k = int(rng.integers(0, len(cb)))
                cb[k] = synth.CODE[int(rng.integers(0, 4))]
        a[i] = synth.word_from_codes(ca, centre=bool(rng.integers(0, 2)))
        b[i] = synth.word_from_codes(cb, centre=bool(rng.integers(0, 2)))
    return a, b
