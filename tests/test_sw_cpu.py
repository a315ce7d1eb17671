"""CPU tier for K4: the product's Smith-Waterman core (pcramp_b200/csrc/sw.cuh), compiled for the host by
tests/native/host_sw_harness.cpp, against SO::SeqOverlap goldens of the UNMODIFIED reference (kat_background.npz)
and, in the dev container, against the live reference.  Integer work: bit-exact."""
import os

import numpy as np
import pytest

from tests import background_cases as bc
from tests.harness import REF_PATH, HostSw, RefLib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def host():
    return HostSw()


def compare(got, want):
    """scores always; coordinates where an alignment exists (the reference reports stale coordinates otherwise)"""
    assert np.array_equal(got[:, 0], want[:, 0])
    ok = got[:, 2] >= 0
    assert np.array_equal(got[ok], want[ok])
    assert np.all(want[~ok, 0] == 0)


def test_sw_matches_reference_golden(host):
    g = np.load(os.path.join(GOLD, "kat_background.npz"))
    got = host.sw_batch(g["sw_query"], g["sw_target"], with_start=True)
    compare(got, g["sw_out"])
    fast = host.sw_batch(g["sw_query"], g["sw_target"], with_start=False)   # the variant the background kernels use
    assert np.array_equal(fast[:, [0, 2, 4, 5]], got[:, [0, 2, 4, 5]])


def test_sw_known_answer(host, oracle):
    """SW(CAGCCACTGCACCTCTTCAT vs TTCAGCCACTGAACCTCTTCATGG) = 35, target range (2, 21)"""
    q = np.array([oracle.word_from_string("CAGCCACTGCACCTCTTCAT", True)], dtype=np.uint64)
    t = np.array([oracle.word_from_string("TTCAGCCACTGAACCTCTTCATGG", True)], dtype=np.uint64)
    out = host.sw_batch(q, t)
    assert out[0, 0] == 35 and (out[0, 3], out[0, 4]) == (2, 21)


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference not present")
def test_sw_matches_live_reference(host):
    ref = RefLib()
    q, t = bc.sw_problems(5, 30000, ref.word_from_string)
    compare(host.sw_batch(q, t), ref.sw_batch(q, t))
