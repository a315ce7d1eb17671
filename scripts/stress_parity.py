"""One-off parity stress on the GPU box (larger than the test suite's cases): K4 with coordinates against the live reference on 4 x 10^5
problems, find_background_match by units against the record form and the live reference on a few hundred sequences, degenerate primers
through the index against the table scan at 2000 x 8 kb."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcramp_b200 import BACKGROUND, TARGET, PcrampGpu, synth  # noqa: E402
from pcramp_b200.api import unpack_bits  # noqa: E402
from tests import background_cases as bc  # noqa: E402
from tests.harness import RefLib  # noqa: E402
from bench_legs import widen  # noqa: E402

g = PcrampGpu(0)
ref = RefLib()
ref.set_threads(0)
t0 = time.time()
q, t = bc.sw_problems(123, 40000 if "--quick" in sys.argv else 400000, ref.word_from_string)
got = g.sw_batch(q, t)
want = ref.sw_batch(q, t)
assert np.array_equal(got[:, 0], want[:, 0])
ok = got[:, 2] >= 0
assert np.array_equal(got[ok], want[ok]), "sw coordinates"
print("sw: %d problems identical (%d with an alignment), %.1f s" % (len(q), int(ok.sum()), time.time() - t0), flush=True)

t0 = time.time()
rng = np.random.default_rng(7)
coll = synth.make_targets(901, 300, 3000, n_clades=3, between=0.08, within=0.03)
f, r = synth.make_pairs(902, coll, 400)
pal = [synth.word_from_string(s) for s in ("ACGTTGCAACGTTGCAACGT", "GGATCCGGATCCGGATCCGG", "AATTCCGGAATTCCGGAATT")]
f[:3] = np.array(pal, np.uint64)
r[:3] = np.array(pal, np.uint64)
thr = float(np.float32(0.8) * np.float32(0.9))
g.upload_sequences(BACKGROUND, coll.nibbles, coll.byte_off, coll.length, coll.weight)
ref.set_sequences(coll)
for taq in (False, True):
    g.select_words(BACKGROUND, f, r, thr, min_oligo_length=bc.BG_MIN_LEN)
    out = {}
    for units in (1, 0):
        g.set_option("use_background_units", units)
        out[units] = g.background_match(BACKGROUND, f, r, thr, 0.8, 0, 2000, taq)
    g.set_option("use_background_units", 1)
    assert out[1][1] == out[0][1] and np.array_equal(out[1][0], out[0][0]), "units vs records"
    ref.select_words(f, r, thr, min_oligo_length=bc.BG_MIN_LEN)
    want, cnt = ref.background_match(f, r, 0.8, 0.9, 0, 2000, taq)
    got = unpack_bits(out[1][0], coll.n)
    defined = ~(want == 255).any(1)
    assert out[1][1] == int(cnt.sum()) and np.array_equal(got[defined], want[defined]), "background vs reference"
    print("background taq=%d: %d candidate amplicons, %d pairs compared with the reference, %d bits set, %.1f s" % (
        taq, out[1][1], int(defined.sum()), int(want[defined].sum()), time.time() - t0), flush=True)
    # pair scoring at the background thresholds: units vs key matrix
    res = {}
    for use in (1, 0):
        g.set_option("use_unit_score", use)
        res[use] = g.score_pairs(BACKGROUND, f, r, thr, 0.8, 0, 2000, taq)
    g.set_option("use_unit_score", 1)
    assert np.array_equal(res[1][0], res[0][0]) and np.array_equal(res[1][1], res[0][1]), "unit scoring"
    cov_ref, bits_ref = ref.score_pairs(f, r, 0.8, 0.9, 0, 2000, taq, want_bits=False)
    assert np.array_equal(res[1][0], cov_ref), "background coverage vs reference"
    print("  scoring at 0.72 / 0.8: coverage identical to the reference (sum %.0f)" % float(cov_ref.sum()), flush=True)

t0 = time.time()
big = synth.make_targets(903, 2000, 8000, n_clades=8, between=0.12, within=0.04)
f, r = synth.make_pairs(904, big, 1000)
f, r = widen(f, rng), widen(r, rng)
thr9 = float(np.float32(1.0) * np.float32(0.9))
g.upload_sequences(TARGET, big.nibbles, big.byte_off, big.length)
outs = []
for use in (1, 0):
    g.set_option("use_index", use)
    ne, nk = g.select_words(TARGET, f, r, thr9)
    cov, bits = g.score_pairs(TARGET, f, r, thr9, 1.0)
    outs.append((ne, nk, g.stats()["n_hits"], cov.copy(), bits.copy(), g.stats()["n_indexed"]))
g.set_option("use_index", 1)
print("index:", outs[0][:3], "indexed patterns", outs[0][5], "| table:", outs[1][:3], "| coverage equal", np.array_equal(outs[0][3], outs[1][3]),
      "bits equal", np.array_equal(outs[0][4], outs[1][4]), flush=True)
assert outs[0][:3] == outs[1][:3] and np.array_equal(outs[0][3], outs[1][3]) and np.array_equal(outs[0][4], outs[1][4]), "degenerate index vs table"
print("degenerate primers: %d entries, %d keys, %d hits identical through the index and the table scan, %.1f s" % (outs[0][0], outs[0][1], outs[0][2], time.time() - t0))
g.close()
print("stress ok")
