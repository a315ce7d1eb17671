"""Seeded inputs for the a15 pins (Score ordering and the best-assay update rule, pcramp.h:180-201, main.cpp:829-858,1455-1480):
small integer-ish scores so that ties on accuracy, on overlap and on total degeneracy all occur."""
import numpy as np


def trial_cases():
    """-> list of (name, target, background, overlap, f, r, max_background_cover)"""
    from pcramp_b200 import synth
    coll = synth.make_targets(22, 4, 800, n_clades=1, within=0.02)
    rng = np.random.default_rng(21)
    out = []
    for n, degfrac in ((1, 0.0), (300, 0.0), (5000, 0.3), (2000, 1.0), (20000, 0.5)):
        f, r = synth.make_pairs(23 + n, coll, n, degenerate_fraction=degfrac)
        tgt = rng.integers(0, 6, size=n).astype(np.float32)
        bg = (rng.integers(0, 3, size=n) * 0.5).astype(np.float32)
        ov = (rng.integers(0, 3, size=n) * 0.25).astype(np.float32)
        for max_bg in (0.0, 0.5, 10.0):
            out.append(("n%d_bg%g" % (n, max_bg), tgt, bg, ov, f, r, max_bg))
    f, r = synth.make_pairs(5, coll, 3)
    out.append(("none_competes", np.ones(3, np.float32), np.full(3, 5.0, np.float32), np.zeros(3, np.float32), f, r, 0.0))
    return out


def rank_cases():
    """-> list of (name, score (n_ranks, 3) float32 {target, background, overlap}, degeneracy (n_ranks,) float64, valid (n_ranks,) bool);
    an invalid rank carries the reference's default Score (pcramp.h:176-179)"""
    rng = np.random.default_rng(77)
    out = []
    for world in (1, 2, 3, 4, 8, 16):
        for rep in range(6):
            score = np.stack([rng.integers(0, 4, size=world), rng.integers(0, 2, size=world) * 0.5, rng.integers(0, 3, size=world) * 0.25],
                             axis=1).astype(np.float32)
            deg = rng.choice([2.0, 3.0, 5.0, 17.0], size=world).astype(np.float64)
            valid = rng.random(world) < 0.8
            if rep == 0:
                valid[:] = False
            score[~valid] = (-1.0e6, 1.0e6, 0.0)
            out.append(("w%d_%d" % (world, rep), score, deg, valid))
    return out
