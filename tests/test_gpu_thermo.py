"""GPU (-m gpu): K3 (SantaLucia nearest-neighbour thermodynamics) through the C ABI against the golden vectors of the
UNMODIFIED reference and, where the compiled reference travelled with the snapshot, against the live reference on
larger seeded batches.  Contract: Tm within 0.01 C, dG within 0.01 kcal/mol; the implementation reproduces the float
arithmetic, so the tests demand identical bit patterns and report the worst deviation if that ever fails."""
import os

import numpy as np
import pytest

from pcramp_b200 import api
from tests import thermo_cases as tc
from tests.golden.make_golden import thermo_filter_inputs
from tests.harness import REF_PATH, RefLib

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TM_TOL = 0.01


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def check(got, want, what):
    tm, dH, dS, dGdp = got
    g = np.stack([tm, dH, dS, dGdp], 1)
    w = want[:, [0, 1, 2, 4]]
    assert np.abs(g[:, 0] - w[:, 0]).max() <= TM_TOL, what
    dg_g = g[:, 1] - np.float32(310.15) * g[:, 2]
    dg_w = w[:, 1] - np.float32(310.15) * w[:, 2]
    assert np.abs(dg_g - dg_w).max() <= 0.01, what
    assert np.array_equal(bits(g), bits(w)), "%s: within tolerance but not bit-identical" % what


@pytest.mark.parametrize("op", tc.OPS)
def test_thermo_batch_matches_reference_golden(gpu, op):
    g = np.load(os.path.join(GOLD, "kat_thermo_batch.npz"))
    n = g["op0_salt0"].shape[0]
    for si, salt in enumerate(tc.SALTS):
        A, B, sa, sb = tc.problems(100 + si, n, op)
        got = gpu.thermo_batch(op, A, B, salt, sa, sb if op in tc.TWO_SEQ else None)
        check(got, g["op%d_salt%d" % (op, si)], "op %d salt %g" % (op, salt))
        st = gpu.thermo_stats()
        assert st["kernel_launches"] == 1 and st["n_problems"] == n


def test_thermo_filters_match_reference_golden(gpu):
    g = np.load(os.path.join(GOLD, "kat_thermo_batch.npz"))
    lib = api.load_library()

    def wfs(s, centre):
        import ctypes
        out = (ctypes.c_uint64 * 2)()
        lib.pcramp_word_from_string(s.encode(), int(centre), out)
        return (int(out[0]), int(out[1]))

    words, f, r, pool_f, pool_r = thermo_filter_inputs(wfs)
    assert np.array_equal(words, g["filter_words"]) and np.array_equal(f, g["filter_f"]) and np.array_equal(pool_r, g["pool_r"])
    for fast in (0, 1):
        for homo in (0, 1):
            got = gpu.is_valid(words, check_homo_dimer=bool(homo), fast_alignment=bool(fast))
            assert np.array_equal(got, g["is_valid_fast%d_homo%d" % (fast, homo)])
        got = gpu.is_valid(words, tm_range=(40.0, 90.0), max_hairpin=60.0, max_dimer=60.0, check_homo_dimer=True, fast_alignment=bool(fast))
        assert np.array_equal(got, g["is_valid_wide_fast%d" % fast])
        got = gpu.is_valid(words, tm_range=(30.0, 95.0), max_hairpin=95.0, max_dimer=30.0, check_homo_dimer=True, fast_alignment=bool(fast))
        assert np.array_equal(got, g["is_valid_dimer_fast%d" % fast])
        got = gpu.max_dimer_tm(f, r, fast_alignment=bool(fast))
        assert np.array_equal(bits(got), bits(g["max_dimer_fast%d" % fast]))
        got = gpu.multiplex_compatible(f, r, pool_f, pool_r, max_dimer=10.0, fast_alignment=bool(fast))
        assert np.array_equal(got, g["multiplex_fast%d" % fast])


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
@pytest.mark.parametrize("op", tc.OPS)
def test_thermo_batch_matches_live_reference(gpu, op):
    ref = RefLib()
    n = 20000
    A, B, sa, sb = tc.problems(21, n, op)
    want = ref.thermo_batch(op, A, B, 0.05, sa, sb)
    got = gpu.thermo_batch(op, A, B, 0.05, sa, sb if op in tc.TWO_SEQ else None)
    check(got, want, "live op %d" % op)


def test_resident_variant_and_properties(gpu):
    """staged run == host-pointer run; heterodimer of (a, revcomp(a)) at equal strands has the duplex enthalpy; hairpin of
    a homopolymer is 0; results do not depend on batch composition (one problem per thread)"""
    A, _, sa, _ = tc.problems(5, 4096, 2)
    ref_out = gpu.thermo_batch(2, A, None, 0.05, sa)
    gpu.thermo_stage(2, A, None, 0.05, sa)
    gpu.thermo_run_staged()
    gpu.thermo_run_staged()
    out = gpu.thermo_fetch()
    for a, b in zip(ref_out, out):
        assert np.array_equal(bits(a), bits(b))
    sub = gpu.thermo_batch(2, A[100:164], None, 0.05, sa[100:164])
    assert np.array_equal(bits(sub[0]), bits(ref_out[0][100:164]))
    plain = [s for s in A if "I" not in s][:512]
    rc = [tc.revcomp(s) for s in plain]
    het = gpu.thermo_batch(3, plain, rc, 0.05, 9e-7, 9e-7)
    pm = gpu.thermo_batch(0, plain, None, 0.05, np.float32(9e-7) - np.float32(0.5) * np.float32(9e-7))
    same = np.isclose(het[1], pm[1], atol=1e-3)
    assert same.mean() > 0.9  # the optimal local alignment is the full duplex unless a stronger partial one exists
    hp = gpu.thermo_batch(1, ["A" * 20, "C" * 25, "T" * 18])
    assert np.all(hp[0] == 0.0)


def test_thermo_errors_mirror_reference(gpu):
    with pytest.raises(api.GpuError, match="Unknown base"):
        gpu.thermo_batch(0, ["ACGTNACGT"])
    with pytest.raises(api.GpuError, match="Illegal base"):
        gpu.thermo_batch(1, ["ACGTNACGT"])
    with pytest.raises(api.GpuError, match="Empty query"):
        gpu.thermo_batch(1, [""])
    with pytest.raises(api.GpuError, match="Na"):
        gpu.thermo_batch(1, ["ACGTACGTAC"], salt=2.0)
    with pytest.raises(api.GpuError, match="strand"):
        gpu.thermo_batch(2, ["ACGTACGTAC"], strand_a=0.0)


def _cells(op, a, b):
    la, lb = len(a), len(b)
    if op == 1:
        return max(0, la - 4) * (max(0, la - 4) + 1) // 2
    return {0: 0, 2: la * la, 3: la * lb, 4: min(la, lb), 5: la}[op]


@pytest.mark.parametrize("op", tc.OPS)
def test_large_batches_equal_small_ones(gpu, op):
    """the three ways a string batch reaches the kernel -- host encoding (small), device encoding (>= 4096), chunks alternating
    between two streams (>= 65536 through thermo_batch) -- give the same bits"""
    n = 70001
    A, B, sa, sb = tc.problems(33, 5000, op)
    two = op in tc.TWO_SEQ
    idx = np.random.default_rng(3).integers(0, 5000, size=n)
    A2 = [A[i] for i in idx]
    B2 = [B[i] for i in idx] if two else None
    sa2, sb2 = sa[idx], (sb[idx] if two else None)
    small = gpu.thermo_batch(op, A[:4000], B[:4000] if two else None, 0.05, sa[:4000], sb[:4000] if two else None)
    piped = gpu.thermo_batch(op, A2, B2, 0.05, sa2, sb2)
    st = gpu.thermo_stats()
    assert st["kernel_launches"] == 4 and st["n_problems"] == n
    want_cells = sum(_cells(op, a, b) for a, b in zip(A2, B2 if two else A2))
    assert st["dp_cells"] == want_cells
    gpu.thermo_stage(op, A2, B2, 0.05, sa2, sb2)
    gpu.thermo_run_staged()
    staged = gpu.thermo_fetch()
    assert gpu.thermo_stats()["dp_cells"] == want_cells
    keep = idx < 4000
    for a, b, c in zip(piped, staged, small):
        assert np.array_equal(bits(a), bits(b))
        assert np.array_equal(bits(a[keep]), bits(c[idx[keep]]))


def test_large_batch_errors_keep_their_order(gpu):
    """the first failing problem decides the message, whether the device (text) or the host (strand concentrations) finds it"""
    n = 70000
    A = ["ACGTACGTACGTACGTAC"] * n
    for m in (5000, n):                      # device encoding; chunks on two streams
        a = list(A[:m])
        a[m - 7] = "ACGTNACGT"
        with pytest.raises(api.GpuError, match="Illegal base"):
            gpu.thermo_batch(2, a)
        with pytest.raises(api.GpuError, match="Unknown base"):
            gpu.thermo_batch(0, a)
        strands = np.full(m, 9e-7, np.float32)
        strands[m - 9] = -1.0
        with pytest.raises(api.GpuError, match="strand"):      # the bad concentration comes first
            gpu.thermo_batch(2, a, strand_a=strands)
        strands[m - 9], strands[m - 5] = 9e-7, -1.0
        with pytest.raises(api.GpuError, match="Illegal base"):  # the bad base comes first
            gpu.thermo_batch(2, a, strand_a=strands)
        a[m - 7] = "ACGT" * 9
        with pytest.raises(api.GpuError, match="longer than 32"):
            gpu.thermo_batch(2, api.pack_strings(a, 40))
        a[m - 7] = ""
        with pytest.raises(api.GpuError, match="Empty query"):
            gpu.thermo_batch(1, a)
        good = gpu.thermo_batch(2, A[:m])    # and the context still works afterwards
        assert np.all(bits(good[0]) == bits(good[0][:1]))


@pytest.mark.parametrize("op", (0, 1, 2, 3))
def test_words_in_equal_text_in(gpu, op):
    """pcramp_gpu_thermo_words: the oligos as words (what PCR::is_valid / max_dimer_tm hold) == their text through thermo_batch, for a
    small batch (converted on the host) and a chunked one (encoded on the device); a degenerate base fails like its text"""
    from pcramp_b200 import synth
    A, B, sa, sb = tc.problems(77, 3000, op)
    keep = [i for i in range(len(A)) if "I" not in A[i] and (B is None or "I" not in B[i])]
    A = [A[i] for i in keep]
    B = None if B is None else [B[i] for i in keep]
    sa, sb = sa[keep], sb[keep]
    two = op in tc.TWO_SEQ
    for n in (len(A), 70000):
        idx = np.arange(n) % len(A)
        a = [A[i] for i in idx]
        b = [B[i] for i in idx] if two else None
        wa = np.array([synth.word_from_string(x, bool(i & 1)) for i, x in enumerate(a)], dtype=np.uint64)
        wb = np.array([synth.word_from_string(x) for x in b], dtype=np.uint64) if two else None
        want = gpu.thermo_batch(op, a, b, 0.05, sa[idx], sb[idx] if two else None)
        got = gpu.thermo_words(op, wa, wb, 0.05, sa[idx], sb[idx] if two else None)
        for x, y in zip(got, want):
            assert np.array_equal(bits(x), bits(y))
        bad = wa.copy()
        bad[n - 3] = synth.word_from_string("ACGTRACGTACGTACGT")
        with pytest.raises(api.GpuError, match="Unknown base" if op == 0 else "Illegal base"):
            gpu.thermo_words(op, bad, wb, 0.05, sa[idx], sb[idx] if two else None)
