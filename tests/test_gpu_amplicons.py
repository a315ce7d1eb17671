"""GPU (-m gpu): the multiplex bookkeeping path through the C ABI -- PCR::collect_unique_amplicons
(amplicon strings in the returned order + AmpliconBounds in push order), the pool x amplicon coverage of main.cpp:783-803
and the accept step of main.cpp:989-1017 (amplicons appended to the multiplex background, keys() rebuilt, targets split) --
against goldens of the UNMODIFIED reference and, when the compiled reference travelled with the snapshot, the live
reference.  Everything is integer / byte work: bit-exact."""
import os

import numpy as np
import pytest

from pcramp_b200 import MULTIPLEX, TARGET
from tests import amplicon_cases as ac
from tests.harness import REF_PATH, RefLib

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_amplicons.npz")
EMPTY = (np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))


def load_targets(gpu, case):
    gpu.upload_sequences(TARGET, case.coll.nibbles, case.coll.byte_off, case.coll.length, case.coll.weight)
    gpu.set_active(TARGET, case.active)
    for seq, pos in case.splits:
        gpu.split_sequence(TARGET, seq, pos)
    gpu.select_words(TARGET, case.f, case.r, case.search_threshold)


def unpacked(gpu, kind):
    """the collection as (lengths, one nibble per byte) like RefLib.sequences()"""
    off, ln, nib = gpu.sequences_copy(kind)
    out = []
    for o, n in zip(off, ln):
        b = nib[int(o):int(o) + (int(n) + 1) // 2]
        two = np.stack([b >> 4, b & 15], 1).reshape(-1)[:int(n)]
        out.append(two)
    return ln, np.concatenate(out + [np.zeros(0, np.uint8)])


def same(got, want, name):
    for k, v in got.items():
        assert np.array_equal(v, want[k]), (name, k)


@pytest.mark.parametrize("case", ac.amp_cases(), ids=lambda c: c.name)
def test_unique_amplicons_match_reference_golden(gpu, case):
    g = np.load(GOLD)
    load_targets(gpu, case)
    got = gpu.unique_amplicons(TARGET, case.f, case.r, float(case.threshold), *case.amp)
    same(ac.flatten(case.name, got), g, case.name)
    assert g["%s_n_amp" % case.name].sum() < g["%s_n_bounds" % case.name].sum()     # making the strings unique removed something
    # one pair at a time gives the same lists as the batch
    p = int(np.argmax(g["%s_n_amp" % case.name]))
    one = gpu.unique_amplicons(TARGET, case.f[p:p + 1], case.r[p:p + 1], float(case.threshold), *case.amp)
    assert one[0][0] == got[p][0] and np.array_equal(one[0][1], got[p][1])


@pytest.mark.parametrize("case", ac.amp_cases(), ids=lambda c: c.name)
def test_pool_amplicon_coverage_matches_reference_golden(gpu, case):
    g = np.load(GOLD)
    load_targets(gpu, case)
    pool = g["%s_pool" % case.name]
    gpu.set_pool(case.f[pool], case.r[pool])
    cov = gpu.pool_amplicon_coverage(TARGET, case.f, case.r, float(case.threshold), case.amp[0], case.amp[1], float(ac.BG_THRESHOLD), case.taq)
    assert np.array_equal(cov, g["%s_pool_cov" % case.name])
    gpu.set_pool(case.f[:0], case.r[:0])
    assert not gpu.pool_amplicon_coverage(TARGET, case.f, case.r, float(case.threshold), case.amp[0], case.amp[1], float(ac.BG_THRESHOLD)).any()


def accept_steps(gpu, case, order):
    """two accept steps on the device; yields what the reference's state would have to equal after each"""
    gpu.upload_sequences(MULTIPLEX, *EMPTY)
    gpu.set_pool(case.f[:0], case.r[:0])
    load_targets(gpu, case)
    for step, p in enumerate(order[:2]):
        if step:
            gpu.select_words(TARGET, case.f, case.r, case.search_threshold)
        gpu.unique_amplicons(TARGET, case.f[p:p + 1], case.r[p:p + 1], float(case.threshold), *case.amp, want_bounds=True, copy=False)
        n_added, n_keys = gpu.accept_assay(0)
        yield step, n_added, n_keys, gpu.multiplex_keys(), unpacked(gpu, MULTIPLEX), unpacked(gpu, TARGET)


@pytest.mark.parametrize("case", ac.amp_cases(), ids=lambda c: c.name)
def test_accept_assay_matches_reference_golden(gpu, case):
    g = np.load(GOLD)
    order = np.argsort(-g["%s_n_amp" % case.name].astype(np.int64), kind="stable")
    for step, n_added, n_keys, keys, mpx, tgt in accept_steps(gpu, case, order):
        pre = "%s_accept%d_" % (case.name, step)
        assert [n_added, n_keys] == list(g[pre + "n"])
        assert np.array_equal(keys, g[pre + "keys"])
        assert np.array_equal(mpx[0], g[pre + "mpx_len"]) and np.array_equal(mpx[1], g[pre + "mpx_nib"])
        assert np.array_equal(tgt[0], g[pre + "tgt_len"]) and np.array_equal(tgt[1], g[pre + "tgt_nib"])
    # the accepted assays are in the pool now: the overlap term sees an exact re-use (MULTIPLEX_OLIGO_REUSE_BONUS)
    p = order[0]
    assert gpu.oligo_overlap(case.f[p:p + 1], case.r[p:p + 1])[0] > 1.0
    gpu.upload_sequences(MULTIPLEX, *EMPTY)
    gpu.set_pool(case.f[:0], case.r[:0])


def test_split_sequences_equals_one_by_one(gpu):
    case = ac.amp_cases()[1]
    rng = np.random.default_rng(5)
    seq = rng.integers(0, case.coll.n, size=200).astype(np.uint32)
    pos = np.array([rng.integers(0, case.coll.length[s]) for s in seq], np.uint32)
    seq[10:20], pos[10:20] = seq[0:10], pos[0:10] ^ 1              # both nibbles of a byte, and repeats
    seq[20:25], pos[20:25] = seq[0:5], pos[0:5]
    res = []
    for batch in (True, False):
        gpu.upload_sequences(TARGET, case.coll.nibbles, case.coll.byte_off, case.coll.length, case.coll.weight)
        if batch:
            gpu.split_sequences(TARGET, seq, pos)
        else:
            for s, p in zip(seq, pos):
                gpu.split_sequence(TARGET, int(s), int(p))
        ne, nk = gpu.select_words(TARGET, case.f, case.r, case.search_threshold)
        res.append((unpacked(gpu, TARGET)[1], gpu.db_copy(TARGET)[:4], (ne, nk)))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][2] == res[1][2]
    for a, b in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, b)
    nib = res[0][0]
    cum = np.concatenate([[0], np.cumsum(case.coll.length.astype(np.int64))])
    assert all(nib[cum[s] + p] == 0 for s, p in zip(seq, pos))


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_against_live_reference(gpu):
    for case in ac.amp_cases():
        ref, mref = RefLib(), RefLib()
        ref.set_sequences(case.coll, case.active)
        for seq, pos in case.splits:
            ref.split_sequence(seq, pos)
        ref.select_words(case.f, case.r, case.search_threshold)
        thr = float(case.threshold)
        want = [ref.unique_amplicons(case.f[p], case.r[p], thr, *case.amp) for p in range(len(case.f))]
        load_targets(gpu, case)
        same(ac.flatten("x", gpu.unique_amplicons(TARGET, case.f, case.r, thr, *case.amp)), ac.flatten("x", want), case.name)
        order = np.argsort(-np.array([len(a) for a, _ in want]), kind="stable")
        for step, n_added, n_keys, keys, mpx, tgt in accept_steps(gpu, case, order):
            if step:
                ref.select_words(case.f, case.r, case.search_threshold)
            p = order[step]
            assert n_added == ref.accept_assay(mref, case.f[p], case.r[p], thr, *case.amp)
            assert np.array_equal(keys, mref.keys())
            for got, ctx in ((mpx, mref), (tgt, ref)):
                seqs = ctx.sequences()
                assert np.array_equal(got[0], np.array([q[0] for q in seqs], np.uint32))
                assert np.array_equal(got[1], np.concatenate([q[2] for q in seqs]))
        # after the two accepts: the pool on the device holds the two assays; trial amplicons of the split targets against it
        ref.select_words(case.f, case.r, case.search_threshold)
        gpu.select_words(TARGET, case.f, case.r, case.search_threshold)
        pool = order[:2]
        want_cov = ref.pool_amplicon_coverage(case.f, case.r, case.f[pool], case.r[pool], thr, case.amp[0], case.amp[1], float(ac.BG_THRESHOLD),
                                              case.taq)
        got_cov = gpu.pool_amplicon_coverage(TARGET, case.f, case.r, thr, case.amp[0], case.amp[1], float(ac.BG_THRESHOLD), case.taq)
        assert np.array_equal(got_cov, want_cov), case.name
        gpu.upload_sequences(MULTIPLEX, *EMPTY)
        gpu.set_pool(case.f[:0], case.r[:0])


def test_bounds_before_the_sequence_start_fail_like_the_reference(gpu):
    """a forward primer hanging over the 5' end of its target (two leading bases off the sequence, allowed at 0.9^2): the
    reference throws from AmpliconBounds() (assay.h:82-89) when bounds are asked for, and returns the amplicon when they are not"""
    from pcramp_b200 import synth
    coll = synth.make_targets(81, 3, 600, n_clades=1, within=0.0)
    c = coll.codes(0)
    # 25-mer = 2 bases off the sequence + its first 23: in frame with the centred 22-base partial word pack() emits at the start
    F = np.concatenate([np.array([1, 2], np.uint8), c[:23]])
    R = synth.revcomp_codes(c[130:150])
    f = np.array([synth.word_from_codes(F)], dtype=np.uint64)
    r = np.array([synth.word_from_codes(R)], dtype=np.uint64)
    thr = np.float32(0.9)
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length, coll.weight)
    gpu.select_words(TARGET, f, r, float(thr * ac.MULT))
    got = gpu.unique_amplicons(TARGET, f, r, float(thr), 80, 200, want_bounds=False)
    assert len(got[0][0]) == 1
    with pytest.raises(RuntimeError, match="AmpliconBounds"):
        gpu.unique_amplicons(TARGET, f, r, float(thr), 80, 200, want_bounds=True)
    if os.path.exists(REF_PATH):
        ref = RefLib()
        ref.set_sequences(coll)
        ref.select_words(f, r, float(thr * ac.MULT))
        assert ref.unique_amplicons(f[0], r[0], float(thr), 80, 200, want_bounds=False)[0] == got[0][0]
        assert ref.unique_amplicons(f[0], r[0], float(thr), 80, 200, want_bounds=True) is None and "AmpliconBounds" in ref.last_error()


def test_best_assay_equals_the_reference_update_rule(gpu):
    """pcramp_gpu_best_assay == main.cpp:829-858 folded over the trials in order with the reference's own Score::operator< / == and
    PCR::total_degeneracy (pcramp.h:180-201, assay.h:536-539): goldens written by oracle/ref_driver.cpp::ref_best_assay, and the live
    reference when it travelled"""
    from tests import best_assay_cases
    gold = np.load(os.path.join(os.path.dirname(GOLD), "kat_best_assay.npz"))
    ref = RefLib() if os.path.exists(REF_PATH) else None
    for name, tgt, bg, ov, f, r, max_bg in best_assay_cases.trial_cases():
        idx, acc, o, dg = gpu.best_assay(tgt, bg, ov, f, r, max_bg)
        want = gold["trial_" + name]
        assert idx == int(want[0]), name
        if idx >= 0:
            assert (np.float32(acc), np.float32(o), float(dg)) == (np.float32(want[1]), np.float32(want[2]), float(want[3])), name
        if ref is not None:
            live = ref.best_assay(tgt, bg, ov, f, r, max_bg)
            assert live[0] == idx and (idx < 0 or (np.float32(live[1]), np.float32(live[2]), live[3]) == (np.float32(acc), np.float32(o), float(dg))), name
