"""GPU (-m gpu): candidate generation through the C ABI -- PCR::random_assay with one GPU thread per seed
stream -- against goldens of the UNMODIFIED reference and, when the compiled reference travelled with the snapshot, the live
reference.  The selected oligos and the seed each stream ends on are bit-exact: a single rand_r draw more or less, or one
thermodynamic filter deciding differently, changes every later trial of the stream."""
import os

import numpy as np
import pytest

from pcramp_b200 import TARGET, GpuError
from pcramp_b200.api import RandomAssayOptions
from tests import random_assay_cases as rc
from tests.harness import REF_PATH, RefLib

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_random_assay.npz")


def load(gpu, case):
    gpu.upload_sequences(TARGET, case.coll.nibbles, case.coll.byte_off, case.coll.length, case.coll.weight)
    gpu.set_active(TARGET, case.active)
    for seq, pos in case.splits:
        gpu.split_sequence(TARGET, seq, pos)


@pytest.mark.parametrize("case", rc.ra_cases(), ids=lambda c: c.name)
def test_random_assays_match_reference_golden(gpu, case):
    g = np.load(GOLD)
    load(gpu, case)
    f, r, after, attempts = gpu.random_assays(TARGET, case.seeds, case.per, case.opt)
    assert np.array_equal(after, g["%s_seed_after" % case.name])
    assert np.array_equal(f, g["%s_f" % case.name]) and np.array_equal(r, g["%s_r" % case.name])
    assert attempts.min() >= 1 and attempts.max() > 1          # the filters rejected candidates along the way
    assert gpu.thermo_stats()["kernel_launches"] == 1
    # a stream is self-contained: any subset of the streams gives the same assays
    pick = np.arange(len(case.seeds))[::3]
    off = np.concatenate([[0], np.cumsum(case.per.astype(np.int64))]).astype(np.int64)
    f2, r2, after2, _ = gpu.random_assays(TARGET, case.seeds[pick], case.per[pick], case.opt)
    want = np.concatenate([np.arange(off[i], off[i + 1]) for i in pick])
    assert np.array_equal(f2, f[want]) and np.array_equal(r2, r[want]) and np.array_equal(after2, after[pick])


def test_selected_assays_pass_the_filters(gpu):
    """what random_assay returns is valid by its own rules: is_valid for both oligos, heterodimer Tm, amplicon geometry"""
    case = rc.ra_cases()[0]
    load(gpu, case)
    f, r, _, _ = gpu.random_assays(TARGET, case.seeds, case.per, case.opt)
    o = case.opt
    kw = dict(salt=o.salt, primer_strand=o.primer_strand, tm_range=(o.primer_tm_min, o.primer_tm_max), max_hairpin=o.max_hairpin,
              max_dimer=o.max_dimer)
    # (a fresh NucCruc object per oligo here; inside a stream the object carries the slots past the end of earlier, longer
    # oligos, nuccruc.cuh header, so a handful of borderline hairpins may differ)
    assert gpu.is_valid(f, **kw).mean() > 0.97 and gpu.is_valid(r, **kw).mean() > 0.97
    assert (gpu.max_dimer_tm(f, r, salt=o.salt, primer_strand=o.primer_strand) <= o.max_dimer).mean() > 0.97
    gpu.select_words(TARGET, f, r, 0.9)
    cov, _ = gpu.score_pairs(TARGET, f, r, 1.0, 1.0, o.amplicon_min, o.amplicon_max)
    assert (cov >= 1.0).all()                                   # every assay amplifies at least the target it was cut from


def test_errors_like_the_reference(gpu):
    case = rc.ra_cases()[0]
    load(gpu, case)
    gpu.set_active(TARGET, np.zeros(case.coll.n, np.uint8))
    with pytest.raises(GpuError, match="No active sequences found"):
        gpu.random_assays(TARGET, [1], [1])
    gpu.set_active(TARGET, case.active)
    with pytest.raises(GpuError, match="Unable to generate a valid initial assay"):
        gpu.random_assays(TARGET, [1], [1], RandomAssayOptions(primer_tm_range=(90.0, 95.0)))
    with pytest.raises(GpuError, match="sequence length is too small"):
        gpu.random_assays(TARGET, [1], [1], RandomAssayOptions(amplicon_range=(2500, 3000)))


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_against_live_reference(gpu):
    rng = np.random.default_rng(7)
    for case in rc.ra_cases()[:2]:
        ref = RefLib()
        ref.set_sequences(case.coll, case.active)
        for seq, pos in case.splits:
            ref.split_sequence(seq, pos)
        load(gpu, case)
        seeds = rng.integers(0, 2**32, size=64, dtype=np.uint64).astype(np.uint32)
        per = rng.integers(1, 10, size=64).astype(np.uint32)
        f, r, after, _ = gpu.random_assays(TARGET, seeds, per, case.opt)
        off = np.concatenate([[0], np.cumsum(per.astype(np.int64))]).astype(np.int64)
        for i, (seed, n) in enumerate(zip(seeds, per)):
            wf, wr, wa = ref.random_assay_stream(int(n), int(seed), case.opt)
            assert wa == int(after[i]), (case.name, i)
            assert np.array_equal(f[off[i]:off[i + 1]], wf) and np.array_equal(r[off[i]:off[i + 1]], wr), (case.name, i)
