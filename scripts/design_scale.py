"""Whole design iterations at the scale of BASELINE configs 2 / 3 through pcramp_gpu_design_iteration (no reference beside it: the stock
program needs minutes per iteration here); prints the stage times of every iteration.
usage: python scripts/design_scale.py c2|c3 [iterations] [n_streams]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcramp_b200 import BACKGROUND, MULTIPLEX, TARGET, PcrampGpu, synth  # noqa: E402
from pcramp_b200.api import DesignLoop  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
streams = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
t0 = time.perf_counter()
if which == "c2":      # 10 000 x 5 000 nt targets at 2 %, 1 000 x 5 000 nt backgrounds from a sister ancestor 10 % away
    tg = synth.TargetFactory(2, 10000, 5000, n_clades=1, between=0.0, within=0.02).collection()
    bg = synth.TargetFactory(2, 1000, 5000, n_clades=1, between=0.10, within=0.02).collection()
    opts = dict(num_trial=1000, n_streams=streams)
else:                  # 20 000 x 30 000 nt in 20 clades, -d 16
    tg = synth.TargetFactory(3, 20000, 30000, n_clades=20, between=0.15, within=0.05).collection()
    bg = None
    opts = dict(num_trial=1000, n_streams=streams, degen=16)
print("collections: %.1f s" % (time.perf_counter() - t0), flush=True)
g = PcrampGpu(0)
g.upload_sequences(TARGET, tg.nibbles, tg.byte_off, tg.length)
empty = (np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
if bg is None:
    g.upload_sequences(BACKGROUND, *empty)
else:
    g.upload_sequences(BACKGROUND, bg.nibbles, bg.byte_off, bg.length)
g.upload_sequences(MULTIPLEX, *empty)
g.multiplex_keys()
g.set_pool(np.zeros((0, 2), np.uint64), np.zeros((0, 2), np.uint64))
loop = DesignLoop(g, 42, **opts)
try:
    for k in range(iters):
        t0 = time.perf_counter()
        res = loop.iteration()
        dt = time.perf_counter() - t0
        print("iteration %d: %.1f ms wall; candidates %.1f  background db %.1f  target db %.1f  optimize %.1f  screen %.1f  accept %.1f | found %d "
              "coverage %.0f (background %.0f) remaining %d splits %d entries T %d B %d" % (
                  k + 1, dt * 1e3, res.ms_candidates, res.ms_select_background, res.ms_select_target, res.ms_optimize, res.ms_screen, res.ms_accept,
                  res.found, res.target_coverage, res.background_coverage, res.targets_remaining, res.n_splits, res.n_target_entries,
                  res.n_background_entries), flush=True)
        if not res.found:
            break
finally:
    loop.close()
    g.close()
