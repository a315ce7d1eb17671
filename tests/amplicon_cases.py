"""Seeded cases for the multiplex bookkeeping path: PCR::collect_unique_amplicons, the pool x amplicon
Smith-Waterman coverage of main.cpp:783-803 and the accept step of main.cpp:989-1017; shared by the golden generator and
the GPU tier."""
import numpy as np

from pcramp_b200 import synth

TARGET_AMP = (80, 200)              # DEFAULT_MIN/MAX_TARGET_AMPLICON (pcramp.h:14-15)
BG_THRESHOLD = np.float32(0.8)      # DEFAULT_BACKGROUND_THRESHOLD (pcramp.h:40)
MULT = np.float32(0.9)              # DEFAULT_SEARCH_THRESHOLD_MULTIPLIER (pcramp.h:51)


class AmpCase:
    def __init__(self, name, coll, f, r, threshold, amp=TARGET_AMP, splits=(), inactive=(), pool=4, taq=False):
        self.name, self.coll, self.f, self.r = name, coll, f, r
        self.threshold = np.float32(threshold)
        self.amp, self.splits, self.inactive, self.pool, self.taq = amp, list(splits), list(inactive), pool, taq

    @property
    def search_threshold(self):
        return float(self.threshold * MULT)    # main.cpp:601: float product

    @property
    def active(self):
        a = np.ones(self.coll.n, np.uint8)
        a[self.inactive] = 0
        return a


def amp_cases():
    out = []
    # related targets: the same amplicon string comes out of many sequences (made unique), clades give several strings per assay
    coll = synth.make_targets(71, 30, 2500, n_clades=3, between=0.12, within=0.01)
    f, r = synth.make_pairs(72, coll, 48)
    out.append(AmpCase("clades", coll, f, r, 1.0))
    # mismatches allowed (0.9^2 of the primer must match), splits inside / next to amplicons, inactive and odd-length
    # sequences, tandem copies of an amplified stretch (several amplicons per sequence, inner loops that break on length)
    src = synth.make_targets(73, 10, 1801, n_clades=2, between=0.08, within=0.02)
    codes = [src.codes(i).copy() for i in range(src.n)]
    unit = src.codes(0)[300:520]
    codes.append(np.concatenate([src.codes(1)[:40], unit, unit, unit, src.codes(1)[40:77]]))
    codes.append(np.concatenate([unit, src.codes(2)[:33], unit]))
    coll2 = synth.Collection(codes)
    f2, r2 = synth.make_pairs(74, synth.Collection([c for c in codes[:10]] + [unit]), 64)
    out.append(AmpCase("mismatch_splits", coll2, f2, r2, 0.9, splits=[(0, 400), (0, 401), (3, 900), (5, 1200), (10, 300), (11, 100)],
                       inactive=[4, 7], taq=True))
    # wide amplicon window, degenerate bases in the targets (the letters of the returned strings are IUPAC codes)
    coll3 = synth.make_targets(75, 12, 1500, n_clades=2, between=0.10, within=0.02)
    codes3 = [coll3.codes(i).copy() for i in range(coll3.n)]
    rng = np.random.default_rng(76)
    for c in codes3:
        for p in rng.integers(0, len(c), size=12):
            c[int(p)] = int(rng.choice([3, 5, 6, 9, 10, 12, 7, 11, 13, 14, 15]))
    coll3 = synth.Collection(codes3)
    f3, r3 = synth.make_pairs(77, synth.make_targets(75, 12, 1500, n_clades=2, between=0.10, within=0.02), 40, amplicon_range=(60, 400))
    out.append(AmpCase("degenerate_wide", coll3, f3, r3, 0.95, amp=(60, 400)))
    return out


def flatten(name, per_pair):
    """[(strings, bounds)] per pair -> arrays for a fixture / a comparison: counts per pair, the strings joined by newlines, the bounds"""
    n_amp = np.array([len(a) for a, _ in per_pair], np.uint32)
    n_bounds = np.array([len(b) for _, b in per_pair], np.uint32)
    text = "\n".join(s for a, _ in per_pair for s in a).encode()
    bounds = np.concatenate([np.asarray(b, np.uint32).reshape(-1, 3) for _, b in per_pair] + [np.zeros((0, 3), np.uint32)])
    return {"%s_n_amp" % name: n_amp, "%s_n_bounds" % name: n_bounds, "%s_text" % name: np.frombuffer(text, np.uint8).copy(),
            "%s_bounds" % name: bounds}
