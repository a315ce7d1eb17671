"""CPU: the oracle restatement (oracle/pcramp_oracle.cpp) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from tests import scenarios

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = {fn.__name__[2:]: fn for fn in scenarios.ALL}


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario_matches_reference_golden(oracle, name):
    sc = SCENARIOS[name]()
    got = scenarios.run_checker(oracle, sc, "oracle")
    gold = load("scenario_%s.npz" % name)
    assert sorted(got) == sorted(gold.files)
    for k in gold.files:
        assert got[k].shape == gold[k].shape, k
        assert np.array_equal(got[k], gold[k]), k  # bit-exact, floats included


def test_word_primitives_match_reference(oracle):
    g = load("kat_words.npz")
    for row in g["word_rows"]:
        w = (int(row[0]), int(row[1]))
        assert oracle.word_size(w) == int(row[2])
        assert oracle.word_start(w) & 0xFFFFFFFF == int(row[3])
        assert oracle.word_stop(w) & 0xFFFFFFFF == int(row[4])
        assert int(min(oracle.word_degeneracy(w), 2.0 ** 62)) == int(row[5])
        assert oracle.word_complement(w) == (int(row[6]), int(row[7]))
        assert oracle.word_center(w) == (int(row[8]), int(row[9]))
        assert oracle.word_shift(w, 1) == (int(row[10]), int(row[11]))
        assert oracle.word_shift(w, 0) == (int(row[12]), int(row[13]))
    for row in g["and_rows"]:
        assert oracle.word_and((int(row[0]), int(row[1])), (int(row[2]), int(row[3]))) == int(row[4])
    taq = np.array([[oracle.taq_mama(a, b, c, d) for c in (1, 2, 4, 8, 5) for d in (1, 2, 4, 8, 0)] for a in (1, 2, 4, 8, 15) for b in (1, 2, 4, 8, 3)],
                   dtype=np.float32)
    assert np.array_equal(taq, g["taq"])
    exp = []
    for s in ("ACGT", "RAGY", "NAN", "ACRTNGB", "SWKM"):
        n, ws = oracle.word_expand(oracle.word_from_string(s, True))
        exp.append(np.concatenate([[n], ws.reshape(-1)]).astype(np.uint64))
    assert np.array_equal(np.concatenate(exp), g["expand"])


def test_word_strings_round_trip(oracle):
    g = load("kat_words.npz")
    rows = g["word_rows"]
    for i, s in enumerate(g["oligos"]):
        s = str(s)
        assert oracle.word_from_string(s, False) == (int(rows[2 * i][0]), int(rows[2 * i][1]))
        assert oracle.word_from_string(s, True) == (int(rows[2 * i + 1][0]), int(rows[2 * i + 1][1]))


def test_appendix_b_known_answers(oracle):
    """values read from the compiled reference in SURVEY.md Appendix B"""
    w = oracle.word_from_string("CAGCCACTGCACCTCTTCAT", True)
    assert w[0] ^ w[1] == 1308340373808031810          # Word::hash of the centred word
    assert oracle.word_start(w) == 6 and oracle.word_stop(w) == 25
    assert oracle.word_and(w, w) == 20
    assert oracle.word_and(w, oracle.word_from_string("CAGCCTCTGCACCTNTTCAT", True)) == 19
    assert oracle.word_and(w, oracle.word_from_string("RAGCCACTGCACCTCTTCAT", True)) == 19
    assert oracle.word_degeneracy(oracle.word_from_string("CAGCCTCTGCACCTNTTCAT", True)) == 4
    assert oracle.word_degeneracy(oracle.word_from_string("RAGCCACTGCACCTCTTCAT", True)) == 2
    assert oracle.taq_mama(1, 8, 1, 8) == 1.0            # primer AT on template AT
    assert abs(oracle.taq_mama(2, 4, 8, 8) - 0.364) < 1e-6  # primer CG on template TT
