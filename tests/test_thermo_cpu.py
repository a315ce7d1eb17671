"""CPU tier for K3 (NucCruc): the product's device functions (pcramp_b200/csrc/nuccruc.cuh), compiled for the host by
tests/native/host_thermo_harness.cpp, against the golden vectors the UNMODIFIED reference produced
(tests/golden/make_golden.py -> kat_thermo.npz, kat_thermo_batch.npz) and, in the dev container, against the live
reference on larger random batches.  Bit-exact: every float is compared by its bit pattern (the contract is
0.01 C / 0.01 kcal/mol; the arithmetic is reproduced exactly)."""
import os

import numpy as np
import pytest

from tests import thermo_cases as tc
from tests.harness import REF_PATH, HostThermo, RefLib, hetero_strand

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def host():
    return HostThermo()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("op", tc.OPS)
def test_host_build_matches_reference_golden(host, op):
    g = np.load(os.path.join(GOLD, "kat_thermo_batch.npz"))
    n = g["op0_salt0"].shape[0]
    for si, salt in enumerate(tc.SALTS):
        A, B, sa, sb = tc.problems(100 + si, n, op)
        want = g["op%d_salt%d" % (op, si)]
        got, cells = host.thermo_batch(op, A, B, salt, hetero_strand(sa, sb) if op in tc.TWO_SEQ else sa)
        assert np.array_equal(bits(got), bits(want[:, [0, 1, 2, 4]]))  # tm, dH, dS, dG_dp
        assert (cells > 0) == (op != 0)


def test_appendix_b_values(host):
    """SURVEY.md Appendix B (read from the compiled reference; salt 0.05 M, strand 9e-7 M)"""
    out, _ = host.thermo_batch(0, ["CAGCCACTGCACCTCTTCAT", "ACATAGCCTGATACGAGT", "GGGTGTGCATCGAGCGGGCG", "A" * 20, "GC" * 10])
    assert np.allclose(out[:, 0], [60.2054, 52.4262, 68.8933, 41.0242, 82.8735], atol=1e-3)
    assert np.allclose(out[0, 1:3], [-155.2, -0.437906], atol=1e-4)
    out, _ = host.thermo_batch(1, ["CAGCCACTGCACCTCTTCAT", "ACATAGCCTGATACGAGT", "GGGTGTGCATCGAGCGGGCG", "A" * 20, "GC" * 10, "ACGTACGTACGTACGTACGTACGTA"])
    assert np.allclose(out[:, 0], [12.8256, 11.3447, 13.3228, 0.0, 89.1549, 70.9397], atol=1e-3)
    out, _ = host.thermo_batch(2, ["ACGTACGTACGTACGTACGTACGTA", "ACAATCATTTCAGGCGCGAG"])
    assert np.allclose(out[:, 0], [62.3203, 6.1082], atol=1e-3)
    st = hetero_strand(9e-7, 9e-7)
    for op in (3, 4):
        out, _ = host.thermo_batch(op, ["CAGCCACTGCACCTCTTCAT"], ["ATGAAGAGGTGCAGTGGCTG"], strand=st)
        assert abs(out[0, 0] - 59.2220) < 1e-3


def test_old_known_answer_fixture(host):
    g = np.load(os.path.join(GOLD, "kat_thermo.npz"))
    seqs = [str(s) for s in g["thermo_seqs"]]
    k = 0
    for s in seqs:
        for op in (0, 1, 2):
            got, _ = host.thermo_batch(op, [s])
            assert np.array_equal(bits(got[0]), bits(g["thermo_self"][k][[0, 1, 2, 4]]))
            k += 1
    k = 0
    for i in range(0, len(seqs) - 1, 2):
        for op in (3, 4):
            got, _ = host.thermo_batch(op, [seqs[i]], [seqs[i + 1]], strand=hetero_strand(9e-7, 9e-7))
            assert np.array_equal(bits(got[0]), bits(g["thermo_het"][k][[0, 1, 2, 4]]))
            k += 1


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference not present")
@pytest.mark.parametrize("op", tc.OPS)
def test_host_build_matches_live_reference(host, op):
    ref = RefLib()
    A, B, sa, sb = tc.problems(7, 3000, op)
    want = ref.thermo_batch(op, A, B, 0.05, sa, sb)
    got, _ = host.thermo_batch(op, A, B, 0.05, hetero_strand(sa, sb) if op in tc.TWO_SEQ else sa)
    assert np.array_equal(bits(got), bits(want[:, [0, 1, 2, 4]]))
