"""GPU (-m gpu): K4 through the C ABI -- raw SeqOverlap alignments, PCR::find_background_match and
PCR::find_multiplex_background_match -- against goldens of the UNMODIFIED reference and, when the compiled reference
travelled with the snapshot, the live reference.  Bit-exact (integer alignments; the float score only feeds a threshold).

Pairs whose candidate-amplicon count is odd and below the number of sequences are excluded from the bit comparison:
the reference indexes one element past the end of its list there (background_match.cpp:122, undefined behaviour --
it crashes on some of these inputs); oracle/ref_driver.cpp marks them 255."""
import os

import numpy as np
import pytest

from pcramp_b200 import BACKGROUND, MULTIPLEX
from pcramp_b200.api import unpack_bits
from tests import background_cases as bc
from tests.harness import REF_PATH, RefLib

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold():
    return np.load(os.path.join(GOLD, "kat_background.npz"))


def compare_sw(got, want):
    assert np.array_equal(got[:, 0], want[:, 0])
    ok = got[:, 2] >= 0
    assert np.array_equal(got[ok], want[ok])


def test_sw_batch_matches_reference_golden(gpu):
    g = gold()
    compare_sw(gpu.sw_batch(g["sw_query"], g["sw_target"]), g["sw_out"])


def run_case(gpu, case):
    gpu.upload_sequences(BACKGROUND, case.coll.nibbles, case.coll.byte_off, case.coll.length, case.coll.weight)
    for seq, pos in case.splits:
        gpu.split_sequence(BACKGROUND, seq, pos)
    gpu.select_words(BACKGROUND, case.f, case.r, case.search_threshold, min_oligo_length=bc.BG_MIN_LEN)
    bits, n_amp = gpu.background_match(BACKGROUND, case.f, case.r, case.search_threshold, float(bc.BG_THRESHOLD), bc.BG_AMP[0], bc.BG_AMP[1],
                                       case.taq)
    return unpack_bits(bits, case.coll.n), n_amp


@pytest.mark.parametrize("case", bc.bg_cases(), ids=lambda c: c.name)
def test_background_match_matches_reference_golden(gpu, case):
    g = gold()
    want, cnt = g["bg_%s_bits" % case.name], g["bg_%s_count" % case.name]
    got, n_amp = run_case(gpu, case)
    assert n_amp == int(cnt.sum())               # the candidate amplicon lists have the reference's sizes
    defined = ~(want == 255).any(1)
    assert defined.sum() >= len(want) // 3
    assert np.array_equal(got[defined], want[defined])
    if case.name == "repeats":                   # more candidates than sequences: the odd-index guard is live
        assert cnt.max() > case.coll.n


def test_multiplex_background_match_matches_reference_golden(gpu):
    g = gold()
    coll, f, r = bc.multiplex_case()
    gpu.upload_sequences(MULTIPLEX, coll.nibbles, coll.byte_off, coll.length, coll.weight)
    for taq in (0, 1):
        bits = gpu.multiplex_background_match(MULTIPLEX, f, r, float(bc.BG_THRESHOLD), bool(taq))
        assert np.array_equal(unpack_bits(bits, coll.n), g["multiplex_bits_taq%d" % taq])


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_against_live_reference(gpu):
    ref = RefLib()
    q, t = bc.sw_problems(9, 50000, ref.word_from_string)
    compare_sw(gpu.sw_batch(q, t), ref.sw_batch(q, t))
    for case in bc.bg_cases():
        ref.set_sequences(case.coll)
        for seq, pos in case.splits:
            ref.split_sequence(seq, pos)
        ref.select_words(case.f, case.r, case.search_threshold, min_oligo_length=bc.BG_MIN_LEN)
        want, cnt = ref.background_match(case.f, case.r, float(bc.BG_THRESHOLD), float(bc.BG_MULT), bc.BG_AMP[0], bc.BG_AMP[1], case.taq)
        got, n_amp = run_case(gpu, case)
        defined = ~(want == 255).any(1)
        assert n_amp == int(cnt.sum()) and np.array_equal(got[defined], want[defined]), case.name


@pytest.mark.parametrize("case", bc.bg_cases(), ids=lambda c: c.name)
def test_background_units_equal_amplicon_records(gpu, case):
    """aligning once per (pair, matching entry) and walking the combinations in the reference's order == one record and four
    alignments per candidate amplicon"""
    out = {}
    for units in (1, 0):
        gpu.set_option("use_background_units", units)
        out[units] = run_case(gpu, case)
    gpu.set_option("use_background_units", 1)
    assert out[1][1] == out[0][1] and out[1][1] > 0
    assert np.array_equal(out[1][0], out[0][0])
