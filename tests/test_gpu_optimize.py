"""GPU (-m gpu): the local search -- one move evaluation (pcramp_gpu_score_variants) and optimize() with the six moves
(pcramp_gpu_optimize) -- against goldens of the UNMODIFIED reference's optimize() (tests/golden/make_golden.py) and the
live reference when it travelled.  Selected oligos bit-exact, scores bit-exact."""
import os

import numpy as np
import pytest

from pcramp_b200 import BACKGROUND, TARGET
from tests import background_cases as bc
from tests import optimize_cases as oc
from tests.harness import REF_PATH, RefLib

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def prepare(gpu, case):
    # the multiplex terms of optimize() read library state (multiplex key list, assay pool): start every case from "first assay of a run"
    from pcramp_b200 import MULTIPLEX
    none = np.zeros((0, 2), np.uint64)
    gpu.upload_sequences(MULTIPLEX, np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
    gpu.multiplex_keys()
    gpu.set_pool(none, none)
    gpu.upload_sequences(TARGET, case.targets.nibbles, case.targets.byte_off, case.targets.length, case.targets.weight)
    gpu.select_words(TARGET, case.f, case.r, case.target_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3)
    if case.background is not None:
        b = case.background
        gpu.upload_sequences(BACKGROUND, b.nibbles, b.byte_off, b.length, b.weight)
        gpu.select_words(BACKGROUND, case.f, case.r, case.background_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3,
                         min_oligo_length=bc.BG_MIN_LEN)
    else:
        empty = np.zeros(0, np.uint8)
        gpu.upload_sequences(BACKGROUND, empty, np.zeros(0, np.uint64), np.zeros(0, np.uint32))
        gpu.select_words(BACKGROUND, case.f, case.r, case.background_search)


@pytest.mark.parametrize("case", oc.cases(), ids=lambda c: c.name)
def test_optimize_matches_reference_golden(gpu, case):
    g = np.load(os.path.join(GOLD, "kat_optimize.npz"))
    prepare(gpu, case)
    o = case.options
    cov, _ = gpu.score_variants(TARGET, case.f, case.r, g["opt_%s_f" % case.name], g["opt_%s_r" % case.name], case.target_search,
                                float(o.target_threshold), o.target_amplicon_min, o.target_amplicon_max, bool(o.use_taq_mama))
    assert np.array_equal(cov.view(np.uint32), g["var_%s_cov" % case.name].view(np.uint32))
    f, r, tc, bcov, ov, it = gpu.optimize(case.f, case.r, case.moves, o)
    want = g["opt_%s_score" % case.name]
    bad = np.where((f != g["opt_%s_f" % case.name]).any(1) | (r != g["opt_%s_r" % case.name]).any(1))[0]
    assert len(bad) == 0, "trials with different oligos: %s" % bad[:10]
    got = np.stack([tc, bcov, ov], 1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert it.min() >= 1


def test_variants_with_base_equal_score_pairs(gpu):
    case = oc.cases()[0]
    prepare(gpu, case)
    o = case.options
    cov_a, bits_a = gpu.score_pairs(TARGET, case.f, case.r, case.target_search, float(o.target_threshold))
    cov_b, bits_b = gpu.score_variants(TARGET, case.f, case.r, case.f, case.r, case.target_search, float(o.target_threshold))
    assert np.array_equal(cov_a, cov_b) and np.array_equal(bits_a, bits_b)


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_optimize_matches_live_reference(gpu):
    case = oc.cases()[2]
    ref = RefLib()
    ref.set_sequences(case.targets)
    ref.select_words(case.f, case.r, case.target_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3)
    bg = RefLib()
    bg.set_sequences(case.background)
    bg.select_words(case.f, case.r, case.background_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3, min_oligo_length=bc.BG_MIN_LEN)
    wf, wr, wscore = ref.optimize(case.f, case.r, case.moves, case.options, bg)
    prepare(gpu, case)
    f, r, tc, bcov, ov, it = gpu.optimize(case.f, case.r, case.moves, case.options)
    assert np.array_equal(f, wf) and np.array_equal(r, wr)
    assert np.array_equal(np.stack([tc, bcov, ov], 1).view(np.uint32), wscore.view(np.uint32))


# ---- the multiplex terms (optimize.cpp:76-96; multiplex background keys, pool overlap) --------------------------------------
def prepare_multiplex(gpu, case):
    from pcramp_b200 import MULTIPLEX
    prepare(gpu, case)
    m = case.multiplex
    if m is not None:
        gpu.upload_sequences(MULTIPLEX, m.nibbles, m.byte_off, m.length, m.weight)
    else:
        gpu.upload_sequences(MULTIPLEX, np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
    keys = gpu.multiplex_keys()
    gpu.set_pool(*case.pool)
    return keys


@pytest.mark.parametrize("case", oc.multiplex_cases(), ids=lambda c: c.name)
def test_multiplex_optimize_matches_reference_golden(gpu, case):
    g = np.load(os.path.join(GOLD, "kat_optimize_multiplex.npz"))
    keys = prepare_multiplex(gpu, case)
    o = case.options
    wf, wr = g["opt_%s_f" % case.name], g["opt_%s_r" % case.name]
    if case.multiplex is not None:
        assert np.array_equal(keys, g["mpx_%s_keys" % case.name])          # keys() order: lexicographic on (buffer[0], buffer[1])
        for taq in (0, 1):
            cov = gpu.multiplex_coverage(case.f, case.r, wf, wr, float(o.background_threshold), bool(taq))
            assert np.array_equal(cov.view(np.uint32), g["mpx_%s_cov_taq%d" % (case.name, taq)].view(np.uint32))
    else:
        assert len(keys) == 0
    ov = gpu.oligo_overlap(wf, wr)
    assert np.array_equal(ov.view(np.uint32), g["mpx_%s_overlap" % case.name].view(np.uint32))
    f, r, tc, bcov, ovl, it = gpu.optimize(case.f, case.r, case.moves, o)
    bad = np.where((f != wf).any(1) | (r != wr).any(1))[0]
    assert len(bad) == 0, "trials with different oligos: %s" % bad[:10]
    got = np.stack([tc, bcov, ovl], 1)
    assert np.array_equal(got.view(np.uint32), g["opt_%s_score" % case.name].view(np.uint32))


def test_multiplex_needs_keys(gpu):
    from pcramp_b200 import MULTIPLEX
    case = oc.multiplex_cases()[0]
    prepare(gpu, case)
    m = case.multiplex
    gpu.upload_sequences(MULTIPLEX, m.nibbles, m.byte_off, m.length, m.weight)     # no multiplex_keys() after the upload
    with pytest.raises(RuntimeError, match="pcramp_gpu_multiplex_keys"):
        gpu.optimize(case.f, case.r, case.moves, case.options)


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_multiplex_pieces_match_live_reference(gpu):
    """random trial oligos: multiplex coverage, overlap and Word::max_overlap against the live reference"""
    case = oc.multiplex_cases()[1]
    prepare_multiplex(gpu, case)
    mp = RefLib()
    mp.set_sequences(case.multiplex)
    mp.pack_all()
    rng = np.random.default_rng(5)
    n = 200
    f2, r2 = synth_pairs(case, n)
    base_f, base_r = f2.copy(), r2.copy()
    perm = rng.permutation(n)
    var_f, var_r = f2[perm], r2.copy()                                          # trial oligos unrelated to the collected lists too
    var_f[: n // 2] = f2[: n // 2]
    for thr in (0.8, 0.6):
        for taq in (False, True):
            want = mp.multiplex_coverage(base_f, base_r, var_f, var_r, thr, taq)
            got = gpu.multiplex_coverage(base_f, base_r, var_f, var_r, thr, taq)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (thr, taq)
    ref = RefLib()
    pf, pr = case.pool
    pool_f = np.concatenate([pf, f2[:5]])
    pool_r = np.concatenate([pr, r2[5:10]])
    gpu.set_pool(pool_f, pool_r)
    want = ref.oligo_overlap(f2, r2, pool_f, pool_r)
    got = gpu.oligo_overlap(f2, r2)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (want >= 10.0).any()                                                # the re-use bonus was exercised


def synth_pairs(case, n):
    from pcramp_b200 import synth
    coll = synth.Collection([case.multiplex.codes(i)[case.multiplex.codes(i) != 0] for i in (0, 1, 2, 4)])
    return synth.make_pairs(9, coll, n, primer_range=(18, 25), degenerate_fraction=0.3)


def test_grouped_variants_equal_one_by_one(gpu):
    """score_variants by groups that share their base assay (score_entries_groups_kernel) == variant by variant"""
    case = oc.cases()[1]
    prepare(gpu, case)
    o = case.options
    rng = np.random.default_rng(8)
    reps = rng.integers(1, 9, size=len(case.f))
    base_f, base_r = np.repeat(case.f, reps, axis=0), np.repeat(case.r, reps, axis=0)
    g = np.load(os.path.join(GOLD, "kat_optimize.npz"))
    pool_f = np.concatenate([case.f, g["opt_%s_f" % case.name]])
    pool_r = np.concatenate([case.r, g["opt_%s_r" % case.name]])
    var_f, var_r = base_f.copy(), base_r.copy()
    pick = rng.integers(0, len(pool_f), size=len(base_f))
    side = rng.integers(0, 2, size=len(base_f)).astype(bool)
    var_f[side] = pool_f[pick[side]]
    var_r[~side] = pool_r[pick[~side]]
    out = {}
    for groups in (1, 0):
        gpu.set_option("use_variant_groups", groups)
        for taq in (False, True):
            out[groups, taq] = gpu.score_variants(TARGET, base_f, base_r, var_f, var_r, case.target_search, float(o.target_threshold),
                                                  o.target_amplicon_min, o.target_amplicon_max, taq)
    gpu.set_option("use_variant_groups", 1)
    for taq in (False, True):
        assert np.array_equal(out[1, taq][0].view(np.uint32), out[0, taq][0].view(np.uint32))
        assert np.array_equal(out[1, taq][1], out[0, taq][1])
    assert out[1, False][0].max() > 0
