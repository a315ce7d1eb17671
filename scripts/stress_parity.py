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

# 5'/3' shift families (optimize_5 / optimize_3) with the same degenerate primers: index against table scan
t0 = time.time()
outs = []
for use in (1, 0):
    g.set_option("use_index", use)
    ne, nk = g.select_words(TARGET, f[:400], r[:400], thr9, optimize_5=True, optimize_3=True)
    cov, bits = g.score_pairs(TARGET, f[:400], r[:400], thr9, 1.0)
    outs.append((ne, nk, g.stats()["n_hits"], cov.copy(), bits.copy()))
g.set_option("use_index", 1)
print("families index:", outs[0][:3], "| table:", outs[1][:3], flush=True)
assert outs[0][:3] == outs[1][:3] and np.array_equal(outs[0][3], outs[1][3]) and np.array_equal(outs[0][4], outs[1][4]), "families index vs table"
print("shift families: %d entries, %d keys, %d hits identical, %.1f s" % (outs[0][0], outs[0][1], outs[0][2], time.time() - t0), flush=True)

# the index through the design loop's lifecycle at this size: sequences split (some several times), a third switched off and on again;
# after every stage index == table scan, and the first 150 sequences (all of them split) == the live reference fed the same splits
t0 = time.time()
srng = np.random.default_rng(11)
sub = big.subset(np.arange(150))
ref.set_sequences(sub)
gs = PcrampGpu(0)   # the same 150 sequences alone: per-sequence independence (select_words.cpp:131-138) ties it to the big run
gs.upload_sequences(TARGET, sub.nibbles, sub.byte_off, sub.length)
for stage in range(3):
    # stage 0 stays below 1/16 of the text (the split sequences keep their stale index entries and take the table scan), the later
    # stages go past it (the index is rebuilt)
    n_own, n_other = ((60, 40), (150, 400), (150, 1500))[stage]
    seqs = np.concatenate([np.arange(n_own), srng.integers(150, big.n, n_other)]).astype(np.uint32)
    pos = srng.integers(40, 7900, len(seqs)).astype(np.uint32)
    g.split_sequences(TARGET, seqs, pos)
    gs.split_sequences(TARGET, seqs[:n_own], pos[:n_own])
    for s_, p_ in zip(seqs[:n_own], pos[:n_own]):
        ref.split_sequence(int(s_), int(p_))
    active = np.ones(big.n, np.uint8)
    if stage == 1:
        active[srng.integers(150, big.n, 600)] = 0
    g.set_active(TARGET, active)
    outs = []
    for use in (1, 0):
        g.set_option("use_index", use)
        ne, nk = g.select_words(TARGET, f, r, thr9)
        st = g.stats()
        cov, bits = g.score_pairs(TARGET, f, r, thr9, 1.0)
        outs.append((ne, nk, st["n_hits"], cov.copy(), unpack_bits(bits, big.n), st["n_index_builds"], st["n_index_stale"]))
    g.set_option("use_index", 1)
    assert outs[0][:3] == outs[1][:3] and np.array_equal(outs[0][3], outs[1][3]) and np.array_equal(outs[0][4], outs[1][4]), "split lifecycle index vs table"
    # the reference's coverage is optimize()'s (search 0.9, detect 1.0), its bits find_target_match's (search = detect = 1.0)
    g.select_words(TARGET, f, r, thr9)
    bits_big = unpack_bits(g.score_pairs(TARGET, f, r, 1.0, 1.0)[1], big.n)
    gs.select_words(TARGET, f, r, thr9)
    cov_s = gs.score_pairs(TARGET, f, r, thr9, 1.0)[0].copy()
    bits_s = unpack_bits(gs.score_pairs(TARGET, f, r, 1.0, 1.0)[1], sub.n)
    ref.select_words(f, r, thr9)
    cov_r, bits_r = ref.score_pairs(f, r, 1.0, 0.9)
    assert np.array_equal(bits_s, bits_r) and np.array_equal(cov_s, cov_r), "split sequences vs reference"
    assert np.array_equal(bits_big[:, :150], bits_r), "split sequences inside the big collection vs reference"
    print("splits stage %d: %d entries, %d hits, index builds %d, stale sequences %d: index == table, 150 split sequences == reference (%d bits), %.1f s" % (
        stage, outs[0][0], outs[0][2], outs[0][5], outs[0][6], int(bits_r.sum()), time.time() - t0), flush=True)
gs.close()
g.close()
print("stress ok")
