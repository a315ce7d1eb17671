"""GPU (-m gpu): parity at the sizes BASELINE.json names and on every scale-dependent code path, against the LIVE reference
(oracle/_ref/libpcramp_ref.so, the unmodified sources compiled by oracle/Makefile; it travels with the snapshot).

  C1 full size          100 x 10 kb x 1000 pairs, --optimize.5 / --optimize.3 shift families: whole database, coverage, bitsets
  C3 shape              200 x 30 kb, primers of degeneracy up to 16 (-d 16): database, coverage, bitsets
  C5 / bench scale      20 000 x 30 kb resident (text index of 6e8 positions, byte tier table, 33-bit entry ids over five radix
                        passes): the columns of sampled targets against the reference on the sub-collection (the path is
                        independent per sequence: select_words.cpp:131-138, pcr_assay.cpp:348-360), one context and four
                        worker contexts sharing the index
  scale-dependent paths buffers that overflow and re-run (tiny_buffers), the sorted tier variant, thresholds with > 7 tiers
"""
import os
import threading

import numpy as np
import pytest

from pcramp_b200 import TARGET, synth
from pcramp_b200.api import unpack_bits
from tests.harness import REF_PATH, RefLib, canonical

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_PATH), reason="the compiled reference did not travel")]

THR = float(np.float32(1.0) * np.float32(0.9))


@pytest.fixture(scope="module")
def ref():
    r = RefLib()
    r.set_threads(0)
    return r


def widen(f, rng, max_degeneracy=16):
    """primers as `-d 16` leaves them (optimize.cpp:356-398 grows degeneracy one base at a time): up to four positions widened to
    two-letter codes, total degeneracy <= max_degeneracy"""
    out = f.copy()
    for w in out:
        k = int(rng.integers(0, 5))
        nib = [(int(w[i // 16]) >> ((15 - i % 16) * 4)) & 15 for i in range(32)]
        pos = [i for i in range(32) if nib[i]]
        deg = 1
        for i in rng.choice(pos, size=min(k, len(pos)), replace=False):
            add = int(synth.CODE[int(rng.integers(0, 4))])
            if nib[i] | add != nib[i] and deg * 2 <= max_degeneracy:
                nib[i] |= add
                deg *= 2
        hi = lo = 0
        for i in range(32):
            if i < 16:
                hi |= nib[i] << ((15 - i) * 4)
            else:
                lo |= nib[i] << ((31 - i) * 4)
        w[0], w[1] = hi, lo
    return out


def compare_all(gpu, ref, coll, f, r, thr=THR, **kw):
    """whole database + keys + coverage (search thr, detect 1.0) + find_target_match bitsets, GPU vs live reference"""
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length, coll.weight)
    ref.set_sequences(coll)
    ne, nk = gpu.select_words(TARGET, f, r, thr, **kw)
    no, nko = ref.select_words(f, r, thr, **kw)
    assert (ne, nk) == (no, nko), ((ne, nk), (no, nko))
    got = canonical(*gpu.db_copy(TARGET)[:4])
    for a, b in zip(got, ref.db()):
        assert np.array_equal(a, b)
    assert np.array_equal(gpu.keys_copy(TARGET), ref.keys())
    cov_r, bits_r = ref.score_pairs(f, r, 1.0, 0.9)
    cov_g, _ = gpu.score_pairs(TARGET, f, r, THR, 1.0)
    _, bits_g = gpu.score_pairs(TARGET, f, r, 1.0, 1.0)
    assert np.array_equal(cov_g, cov_r) and np.array_equal(unpack_bits(bits_g, coll.n), bits_r)
    assert bits_r.any()
    return gpu.stats(), ne


def test_c1_full_size_with_shift_families(gpu, ref):
    """BASELINE config 1 at full size (seed 1, 100 x 10 000 nt at 3 %), 1000 trial pairs, --optimize.5 --optimize.3"""
    coll = synth.make_targets(1, 100, 10000, n_clades=1, between=0.0, within=0.03)
    f, r = synth.make_pairs(11, coll, 1000)
    st, ne = compare_all(gpu, ref, coll, f, r, optimize_5=True, optimize_3=True)
    assert ne > 1000000 and st["n_patterns"] > 20000
    st, ne = compare_all(gpu, ref, coll, f, r)
    assert st["n_indexed"] > 0


def test_c3_shape_with_degenerate_primers(gpu, ref):
    """BASELINE config 3's shape: 30 kb targets in clades (15 % / 5 %), primers of degeneracy up to 16 (-d 16)"""
    coll = synth.make_targets(3, 200, 30000, n_clades=20, between=0.15, within=0.05)
    f, r = synth.make_pairs(31, coll, 400)
    rng = np.random.default_rng(32)
    f, r = widen(f, rng), widen(r, rng)
    st, ne = compare_all(gpu, ref, coll, f, r)
    assert st["n_seeded"] > 0 and ne > 0


def test_buffers_that_overflow_and_rerun(ref):
    """a fresh context whose hit / index-query / index-candidate / neighbour / work-list buffers all start far too small"""
    from pcramp_b200 import PcrampGpu
    coll = synth.make_targets(41, 120, 20000, n_clades=6, between=0.12, within=0.04)
    f, r = synth.make_pairs(42, coll, 300)
    g = PcrampGpu(0)
    try:
        g.set_option("tiny_buffers", 1)
        st, ne = compare_all(g, ref, coll, f, r)
        assert st["n_indexed"] > 0 and st["n_hits"] > 4096 and ne > 4096
    finally:
        g.close()


def test_sorted_tier_variant_and_wide_tier_span(gpu, ref):
    """the best-tier rule without the tier table (the form collections above 2^27 cells take), and a threshold that leaves more than
    7 tiers per candidate (25-mers at 0.65: counts 16..25), which takes the 4-byte table; both against the reference"""
    coll = synth.make_targets(51, 60, 8000, n_clades=3, between=0.12, within=0.05)
    f, r = synth.make_pairs(52, coll, 150)
    try:
        gpu.set_option("use_tier_table", 0)
        compare_all(gpu, ref, coll, f, r)
    finally:
        gpu.set_option("use_tier_table", 1)
    compare_all(gpu, ref, coll, f[:40], r[:40], thr=0.65)


# ---- bench scale -------------------------------------------------------------------------------------------------------------------
N_FULL, L_FULL = 20000, 30000


@pytest.fixture(scope="module")
def full():
    """the bench's own collection (bench.py make_factory: seed 3, 20 000 x 30 000 nt, 20 clades, 15 % / 5 %) resident on one context"""
    from pcramp_b200 import PcrampGpu
    factory = synth.TargetFactory(3, N_FULL, L_FULL, n_clades=20, between=0.15, within=0.05)
    coll = factory.collection()
    g = PcrampGpu(0)
    g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
    yield g, factory, coll
    g.close()


def reference_columns(ref, factory, idx, f, r):
    sample = factory.collection(idx)
    ref.set_sequences(sample)
    ref.select_words(f, r, THR)
    return ref.score_pairs(f, r, 1.0, 0.9)


def test_bench_scale_columns_vs_reference(full, ref):
    g, factory, coll = full
    f, r = synth.make_pairs(5, factory, 1000)                    # the bench's first batch
    idx = list(range(7, N_FULL, 311))                            # 65 targets across all clades
    cov_r, bits_r = reference_columns(ref, factory, idx, f, r)
    ne, _ = g.select_words(TARGET, f, r, THR, want_keys=False)
    st = g.stats()
    assert st["n_indexed"] > 0.9 * st["n_seeded"] > 0 and st["n_positions"] == N_FULL * L_FULL
    # entry ids: position bits of the longest sequence (+64) + 3 + sequence bits = 33 bits -> five 8-bit radix passes
    assert int(np.ceil(np.log2(L_FULL + 64))) + 3 + int(np.ceil(np.log2(N_FULL))) >= 33
    cov, bits = g.score_pairs(TARGET, f, r, THR, 1.0)
    _, bits_tm = g.score_pairs(TARGET, f, r, 1.0, 1.0)
    cols = np.asarray(idx)
    assert np.array_equal(unpack_bits(bits, N_FULL)[:, cols].sum(axis=1).astype(np.float32), cov_r)
    assert np.array_equal(unpack_bits(bits_tm, N_FULL)[:, cols], bits_r)
    assert np.array_equal(cov, unpack_bits(bits, N_FULL).sum(axis=1).astype(np.float32)) and bits_r.any()
    # the word database itself, for the sampled targets: entries of those sequences == the reference's database of the sub-collection
    w, seq, loc, strand, _ = g.db_copy(TARGET)
    remap = -np.ones(N_FULL, np.int64)
    remap[cols] = np.arange(len(cols))
    keep = remap[seq] >= 0
    got = canonical(w[keep], remap[seq[keep]].astype(np.uint32), loc[keep], strand[keep])
    for a, b in zip(got, ref.db()):
        assert np.array_equal(a, b)


def test_bench_scale_four_workers_vs_reference(full, ref):
    """four worker contexts share the parent's resident targets and text index; each scores its own batch from its own host thread
    (what bench.py times); every batch's sampled columns against the reference"""
    g, factory, coll = full
    f, r = synth.make_pairs(6, factory, 4 * 250)
    g.select_words(TARGET, f[:8], r[:8], THR, want_keys=False)   # the index exists before the workers start
    ctxs = [g] + [g.worker() for _ in range(3)]
    out, err = [None] * 4, []

    def run(k):
        try:
            fb, rb = f[250 * k:250 * (k + 1)], r[250 * k:250 * (k + 1)]
            for _ in range(3):                                    # several rounds: the contexts overlap on the device
                ctxs[k].select_words(TARGET, fb, rb, THR, want_keys=False)
                cov, bits = ctxs[k].score_pairs(TARGET, fb, rb, THR, 1.0)
            out[k] = (cov, bits)
        except Exception as e:                                    # noqa: BLE001
            err.append(e)
    th = [threading.Thread(target=run, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    try:
        assert not err, err
        idx = list(range(3, N_FULL, 409))
        cols = np.asarray(idx)
        for k in range(4):
            cov_r, _ = reference_columns(ref, factory, idx, f[250 * k:250 * (k + 1)], r[250 * k:250 * (k + 1)])
            cov, bits = out[k]
            assert np.array_equal(unpack_bits(bits, N_FULL)[:, cols].sum(axis=1).astype(np.float32), cov_r), k
            assert cov_r.sum() > 0
    finally:
        for c in ctxs[1:]:
            c.close()


# ---- index lifecycle ----------------------------------------------------------------------------------------------------------------
def test_index_in_several_parts(ref):
    """a collection cut into several index parts (what a collection above 2^31 bases is): every part queried, same database"""
    from pcramp_b200 import PcrampGpu
    coll = synth.make_targets(61, 50, 12000, n_clades=4, between=0.12, within=0.05)
    f, r = synth.make_pairs(62, coll, 300)
    g = PcrampGpu(0)
    try:
        g.set_option("index_part_positions", 100000)              # 8 sequences per part -> 7 parts
        st, ne = compare_all(g, ref, coll, f, r)
        assert st["n_indexed"] > 0 and st["index_bytes"] > 7 * (1 << 26) and st["n_index_builds"] == 1
    finally:
        g.close()


def test_splits_do_not_rebuild_the_index(ref):
    """main.cpp:1008-1017 splits the targets an assay amplifies; the index is kept: the split sequences' entries are ignored and,
    while such a sequence is still active, the table scan covers it.  Databases / coverage / bitsets == the reference with the same
    splits, before and after the stale share grows past the rebuild threshold."""
    from pcramp_b200 import PcrampGpu
    coll = synth.make_targets(71, 80, 9000, n_clades=4, between=0.12, within=0.05)
    f, r = synth.make_pairs(72, coll, 250)
    rng = np.random.default_rng(73)
    g = PcrampGpu(0)
    try:
        g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
        ref.set_sequences(coll)
        g.select_words(TARGET, f, r, THR)                         # builds the index
        assert g.stats()["n_index_builds"] == 1

        def check(expect_builds, expect_stale):
            ne, nk = g.select_words(TARGET, f, r, THR)
            st = g.stats()
            assert (ne, nk) == ref.select_words(f, r, THR)
            for a, b in zip(canonical(*g.db_copy(TARGET)[:4]), ref.db()):
                assert np.array_equal(a, b)
            cov_r, bits_r = ref.score_pairs(f, r, 1.0, 0.9)
            cov_g, _ = g.score_pairs(TARGET, f, r, THR, 1.0)
            _, bits_g = g.score_pairs(TARGET, f, r, 1.0, 1.0)
            assert np.array_equal(cov_g, cov_r) and np.array_equal(unpack_bits(bits_g, coll.n), bits_r)
            assert st["n_index_builds"] == expect_builds and st["n_index_stale"] == expect_stale, st
        # three sequences split (inside primer sites of some pairs, so that results change), all still active
        for seq in (3, 17, 40):
            for pos in rng.integers(100, 8900, size=6):
                g.split_sequence(TARGET, seq, int(pos))
                ref.split_sequence(seq, int(pos))
        check(1, 3)
        # one of them retired: its stale entries are simply never looked at
        active = np.ones(coll.n, np.uint8)
        active[17] = 0
        g.set_active(TARGET, active)
        ref.set_active(active)
        check(1, 2)
        # many more split and active: past 1/16 of the active text the index is rebuilt
        for seq in range(50, 62):
            g.split_sequence(TARGET, seq, 4000)
            ref.split_sequence(seq, 4000)
        check(2, 0)
    finally:
        g.close()
