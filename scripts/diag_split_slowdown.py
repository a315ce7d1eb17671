import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pcramp_b200 import PcrampGpu, TARGET, synth
from pcramp_b200.api import DesignLoop
coll = synth.make_targets(1, 100, 10000, within=0.03)
g = PcrampGpu(0)
g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
loop = DesignLoop(g, 42, num_trial=1000)
for it in range(4):
    res = loop.iteration()
    st = g.stats()
    print(it, "found", res.found, "ms sel_t %.1f cand %.1f screen %.1f accept %.1f" % (res.ms_select_target, res.ms_candidates, res.ms_screen, res.ms_accept))
    print("   stats", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
# now time select_words alone repeatedly on the split collection
f, r = synth.make_pairs(3, coll, 1000)
thr = float(np.float32(1.0) * np.float32(0.9))
for rep in range(3):
    t0 = time.perf_counter(); g.select_words(TARGET, f, r, thr, want_keys=False); g.synchronize(); dt = time.perf_counter() - t0
    st = g.stats()
    print("select_words again: %.1f ms" % (dt * 1e3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k.startswith("ms") or k in ("n_hits", "n_entries", "n_indexed", "n_seeded", "kernel_launches")})
for opt in ("use_index",):
    g.set_option(opt, 0)
    t0 = time.perf_counter(); g.select_words(TARGET, f, r, thr, want_keys=False); g.synchronize(); dt = time.perf_counter() - t0
    print("without", opt, "%.1f ms" % (dt * 1e3), {k: round(v, 3) for k, v in g.stats().items() if k.startswith("ms")})
    g.set_option(opt, 1)
