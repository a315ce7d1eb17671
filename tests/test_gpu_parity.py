"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs and
against the golden vectors produced by the unmodified reference.  Bit-exact everywhere: binding-site
databases, keys, amplified-target bitsets and float coverages."""
import os

import numpy as np
import pytest

from pcramp_b200 import TARGET, BACKGROUND, synth
from pcramp_b200.api import unpack_bits
from tests import scenarios
from tests.harness import canonical

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
THR_09 = float(np.float32(1.0) * np.float32(0.9))
SCENARIOS = {fn.__name__[2:]: fn for fn in scenarios.ALL}


class GpuChecker:
    """adapts PcrampGpu to the checker surface scenarios.run_checker drives"""

    def __init__(self, gpu, kind=TARGET):
        self.g = gpu
        self.kind = kind
        self.n_seq = 0

    def set_sequences(self, coll, active=None):
        self.n_seq = coll.n
        self.g.upload_sequences(self.kind, coll.nibbles, coll.byte_off, coll.length, coll.weight)
        if active is not None:
            self.g.set_active(self.kind, active)

    def split_sequence(self, seq, pos):
        self.g.split_sequence(self.kind, seq, pos)

    def pack(self, seq, pack_max_degen, min_gc, max_gc, min_len):
        w, loc, st = self.g.pack(self.kind, seq, pack_max_degen, min_gc, max_gc, min_len)
        return canonical(w, np.full(len(loc), seq, np.uint32), loc, st)

    def select_words(self, f, r, threshold, **kw):
        return self.g.select_words(self.kind, f, r, threshold, **kw)

    def db(self):
        w, idx, loc, st, key = self.g.db_copy(self.kind)
        self.last_key_index = key
        return w, idx, loc, st

    def keys(self):
        return self.g.keys_copy(self.kind)

    def score_pairs(self, f, r, search, detect, amp_min, amp_max, taq):
        cov, bits = self.g.score_pairs(self.kind, f, r, search, detect, amp_min, amp_max, taq)
        return cov, unpack_bits(bits, self.n_seq)


def compare(got, want, what):
    assert sorted(got) == sorted(want), what
    for k in want:
        a, b = got[k], want[k]
        assert a.shape == b.shape, "%s: %s shape %s vs %s" % (what, k, a.shape, b.shape)
        assert np.array_equal(a, b), "%s: %s differs" % (what, k)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_scenario_vs_oracle_and_golden(gpu, oracle, name):
    sc = SCENARIOS[name]()
    got = scenarios.run_checker(GpuChecker(gpu), sc, "gpu")
    want = scenarios.run_checker(oracle, sc, "oracle")
    compare(got, want, name + " vs oracle")
    gold = np.load(os.path.join(GOLD, "scenario_%s.npz" % name))
    compare(got, {k: gold[k] for k in gold.files}, name + " vs reference golden")


@pytest.mark.parametrize("name", ["basic", "repeats", "shift"])
def test_db_is_sorted_and_keys_are_consistent(gpu, name):
    sc = SCENARIOS[name]()
    chk = GpuChecker(gpu)
    chk.set_sequences(sc.coll, sc.active)
    chk.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    w, idx, loc, st, key = gpu.db_copy(TARGET)
    keys = gpu.keys_copy(TARGET)
    order = np.lexsort((st, loc, idx, w[:, 1], w[:, 0]))
    assert np.array_equal(order, np.arange(len(order)))          # delivered in canonical order
    assert np.array_equal(keys[key], w)                          # key_index points at the entry's word
    uniq = np.unique(w, axis=0)
    assert len(uniq) == len(keys) and np.array_equal(np.unique(keys, axis=0), uniq)
    assert np.all((keys[1:, 0] > keys[:-1, 0]) | ((keys[1:, 0] == keys[:-1, 0]) & (keys[1:, 1] > keys[:-1, 1])))  # keys() order


def test_background_kind_and_independent_collections(gpu, oracle):
    """targets and backgrounds live side by side in one context (main.cpp keeps three deques)"""
    t = SCENARIOS["basic"]()
    b = SCENARIOS["lowthr"]()
    gt, gb = GpuChecker(gpu, TARGET), GpuChecker(gpu, BACKGROUND)
    gt.set_sequences(t.coll)
    gb.set_sequences(b.coll)
    gb.select_words(b.f, b.r, float(b.threshold), **b.select_kwargs())
    gt.select_words(t.f, t.r, float(t.threshold), **t.select_kwargs())
    for sc, g in ((t, gt), (b, gb)):
        oracle.set_sequences(sc.coll)
        oracle.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
        for a, c in zip(g.db(), oracle.db()):
            assert np.array_equal(a, c)
        cov_o, bits_o = oracle.score_pairs(sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
        cov_g, bits_g = g.score_pairs(sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
        assert np.array_equal(cov_g, cov_o) and np.array_equal(bits_g, bits_o)


def test_medium_random_vs_oracle(gpu, oracle):
    """a few hundred kb with several tiles per sequence, more pairs than one pattern chunk holds"""
    coll = synth.make_targets(501, 40, 5000, n_clades=4, between=0.15, within=0.05)
    f, r = synth.make_pairs(502, coll, 300)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    oracle.set_sequences(coll)
    ne, nk = g.select_words(f, r, thr)
    no, nko = oracle.select_words(f, r, thr)
    assert (ne, nk) == (no, nko) and ne > 1000
    for a, c in zip(g.db(), oracle.db()):
        assert np.array_equal(a, c)
    assert np.array_equal(g.keys(), oracle.keys())
    for search, detect in ((thr, 1.0), (1.0, 1.0), (thr, 0.9)):
        cov_o, bits_o = oracle.score_pairs(f, r, search, detect, 80, 200, False)
        cov_g, bits_g = g.score_pairs(f, r, search, detect, 80, 200, False)
        assert np.array_equal(bits_g, bits_o) and np.array_equal(cov_g, cov_o)
    st = gpu.stats()
    assert st["n_patterns"] == 4 * 300 and st["n_positions"] == 40 * 5000 and st["kernel_launches"] > 0


def test_staged_path_equals_host_pointer_path(gpu):
    sc = SCENARIOS["basic"]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll)
    g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    cov_a, bits_a = gpu.score_pairs(TARGET, sc.f, sc.r, sc.search_threshold, 1.0)
    gpu.stage_pairs(sc.f, sc.r)
    gpu.select_words_staged(TARGET, float(sc.threshold), **sc.select_kwargs())
    gpu.score_pairs_staged(TARGET, sc.search_threshold, 1.0)
    cov_b, bits_b = gpu.fetch_results(TARGET)
    assert np.array_equal(cov_a, cov_b) and np.array_equal(bits_a, bits_b)


def test_properties_at_scale(gpu):
    """size-independent properties on an input too large for the CPU oracle: every pair was cut from one of
    the targets, so (i) its source sequence must be amplified at threshold 0.9 search / 1.0 detect exactly when
    the primers match it perfectly -- they do, by construction; (ii) scoring is independent of batch
    composition for the bitset of find_target_match when the DB is rebuilt per batch from the same pairs;
    (iii) coverage == popcount(bitset) for unit weights."""
    n, L, P = 400, 30000, 256
    coll = synth.make_targets(601, n, L, n_clades=8, between=0.15, within=0.05)
    f, r = synth.make_pairs(602, coll, P)
    thr = float(np.float32(1.0) * np.float32(0.9))
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
    gpu.select_words(TARGET, f, r, thr)
    cov, bits = gpu.score_pairs(TARGET, f, r, 1.0, 1.0)
    b = unpack_bits(bits, n)
    assert np.array_equal(cov, b.sum(axis=1).astype(np.float32))
    assert (b.sum(axis=1) >= 1).all()            # each pair amplifies at least the sequence it was cut from
    # the same pairs in reverse order give the same rows
    gpu.select_words(TARGET, f[::-1].copy(), r[::-1].copy(), thr)
    cov2, bits2 = gpu.score_pairs(TARGET, f[::-1].copy(), r[::-1].copy(), 1.0, 1.0)
    assert np.array_equal(bits2[::-1], bits) and np.array_equal(cov2[::-1], cov)
    # deactivating the sequences a pair amplifies removes exactly those bits
    active = np.ones(n, np.uint8)
    active[b[0].astype(bool)] = 0
    gpu.set_active(TARGET, active)
    gpu.select_words(TARGET, f, r, thr)
    cov3, bits3 = gpu.score_pairs(TARGET, f, r, 1.0, 1.0)
    b3 = unpack_bits(bits3, n)
    assert b3[0].sum() == 0 and np.array_equal(b3, b * active[None, :])


def _db_tuple(gpu, kind=TARGET):
    w, idx, loc, st, key = gpu.db_copy(kind)
    return w, idx, loc, st, key, gpu.keys_copy(kind)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_seed_filter_equals_brute_force(gpu, name):
    """the pigeonhole seed path and the brute-force path are two CUDA implementations of the same scan"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    try:
        gpu.set_option("force_brute_scan", 1)
        g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
        assert gpu.stats()["n_seeded"] == 0
        brute = _db_tuple(gpu)
    finally:
        gpu.set_option("force_brute_scan", 0)
    g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    st = gpu.stats()
    fast = _db_tuple(gpu)
    for a, b in zip(fast, brute):
        assert np.array_equal(a, b)
    if name in ("lowthr", "taq_weights"):
        assert st["n_seeded"] == 0          # 0.72 n / 0.81 n leave pieces shorter than 5 bases: brute force
    elif name not in ("repeats",):
        assert st["n_seeded"] > 0


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_indexed_scan_equals_table_scan(gpu, name):
    """the indexed seed scan (text index + 1-mismatch k-mer neighbours) and the table-based seed scan are two CUDA
    implementations of the same filter: identical databases, and the index really is used where it applies"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    try:
        gpu.set_option("use_index", 0)
        g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
        assert gpu.stats()["n_indexed"] == 0
        table = _db_tuple(gpu)
    finally:
        gpu.set_option("use_index", 1)
    g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    st = gpu.stats()
    indexed = _db_tuple(gpu)
    for a, b in zip(indexed, table):
        assert np.array_equal(a, b)
    if st["n_seeded"] > 0 and name not in ("degenerate",):
        assert st["n_indexed"] > 0 and st["n_index_queries"] > 0


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_seed_table_filter_equals_compare_all(gpu, name):
    """pair scoring and the partial-word scan through the frame-aligned seed table (fst.cuh) == comparing every word with
    every oligo: databases, coverages and bitsets identical"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    out = []
    # (seed table, neighbour filter): compare-all; table walk (fst.cuh); neighbour lists of the generating candidate (score.cuh, the default)
    for use, neigh in ((0, 0), (1, 0), (1, 1)):
        gpu.set_option("use_seed_table", use)
        gpu.set_option("use_neighbours", neigh)
        try:
            g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
            db = _db_tuple(gpu)
            cov, bits = gpu.score_pairs(TARGET, sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            # other oligos than the ones the database was built for (what a move does): shifted / swapped pairs
            cov2, bits2 = gpu.score_pairs(TARGET, sc.r, np.roll(sc.f, 1, axis=0), sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            out.append((db, cov, bits, cov2, bits2))
        finally:
            gpu.set_option("use_seed_table", 1)
            gpu.set_option("use_neighbours", 1)
    for k in (1, 2):
        for a, b in zip(out[0][0], out[k][0]):
            assert np.array_equal(a, b)
        assert np.array_equal(out[0][1].view(np.uint32), out[k][1].view(np.uint32)) and np.array_equal(out[0][2], out[k][2])
        assert np.array_equal(out[0][3].view(np.uint32), out[k][3].view(np.uint32)) and np.array_equal(out[0][4], out[k][4])


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_tier_table_equals_sorted_tiers(gpu, name):
    """select_words' best-tier rule through the (sequence, candidate) table of maxima == through the sorted hit list"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    out = []
    for use in (0, 1):
        gpu.set_option("use_tier_table", use)
        try:
            g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
            out.append(_db_tuple(gpu))
        finally:
            gpu.set_option("use_tier_table", 1)
    assert len(out[0][0]) > 0
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_indexed_scan_large_vs_table_scan(gpu):
    """a collection of several tiles per sequence with degenerate text, EOS and degenerate primers: index path == table path
    (both also equal the oracle in the next test), and most patterns take the index"""
    rng = np.random.default_rng(79)
    base = synth.make_targets(711, 40, 12000, n_clades=4, between=0.12, within=0.05)
    codes = [base.codes(i).copy() for i in range(base.n)]
    for c in codes[:20]:
        k = rng.integers(0, len(c), size=10)
        c[k] |= synth.CODE[rng.integers(0, 4, size=10)]
    codes[7][5000] = 0
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(712, base, 500, degenerate_fraction=0.3)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    out = []
    for use in (0, 1):
        gpu.set_option("use_index", use)
        try:
            g.select_words(f, r, thr, optimize_5=True, optimize_3=True)
            out.append((_db_tuple(gpu), gpu.stats()))
        finally:
            gpu.set_option("use_index", 1)
    for a, b in zip(out[0][0], out[1][0]):
        assert np.array_equal(a, b)
    assert out[1][1]["n_indexed"] > 0.5 * out[1][1]["n_seeded"]
    assert out[1][1]["n_hits"] == out[0][1]["n_hits"]


def test_seed_filter_with_degenerate_text_vs_oracle(gpu, oracle):
    """IUPAC codes and N runs in the targets (dirty groups), degenerate primers (seed expansion), several tiles"""
    rng = np.random.default_rng(77)
    base = synth.make_targets(701, 24, 9000, n_clades=3, between=0.12, within=0.05)
    codes = [base.codes(i).copy() for i in range(base.n)]
    for c in codes:
        k = rng.integers(0, len(c), size=25)
        c[k] |= synth.CODE[rng.integers(0, 4, size=25)]
        s = int(rng.integers(0, len(c) - 40))
        c[s:s + int(rng.integers(1, 12))] = 15
    codes[3][2040:2056] = 15      # an N run across a tile boundary
    codes[5][4000] = 0            # and one EOS
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(702, base, 200, degenerate_fraction=0.5)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    oracle.set_sequences(coll)
    for kw in (dict(), dict(optimize_5=True, optimize_3=True)):
        ne, nk = g.select_words(f[:60] if kw else f, r[:60] if kw else r, thr, **kw)
        no, nko = oracle.select_words(f[:60] if kw else f, r[:60] if kw else r, thr, **kw)
        assert (ne, nk) == (no, nko)
        for a, c in zip(g.db(), oracle.db()):
            assert np.array_equal(a, c)
        assert gpu.stats()["n_seeded"] > 0
    cov_o, bits_o = oracle.score_pairs(f[:60], r[:60], thr, 0.9, 80, 200, False)
    cov_g, bits_g = g.score_pairs(f[:60], r[:60], thr, 0.9, 80, 200, False)
    assert np.array_equal(bits_g, bits_o) and np.array_equal(cov_g, cov_o)


def test_exact_threshold_and_mixed_classes(gpu, oracle):
    """threshold 1.0 (one piece = the whole primer) and 0.8 (pieces of 4-5: mixed seeded / brute-force batch)"""
    coll = synth.make_targets(801, 16, 6000, n_clades=2, between=0.1, within=0.04)
    f, r = synth.make_pairs(802, coll, 120)
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    oracle.set_sequences(coll)
    for thr in (1.0, 0.95, 0.8, 0.75):
        ne, nk = g.select_words(f, r, thr)
        assert (ne, nk) == oracle.select_words(f, r, thr)
        for a, c in zip(g.db(), oracle.db()):
            assert np.array_equal(a, c)


def test_merge_shards_equals_unsharded(gpu):
    """the N>1 exchange on one GPU: score three shards one after the other, concatenate their (any, pass-1) bitsets as an
    all-gather would, splice with pcramp_gpu_merge_shards -> identical bitsets and (weighted) coverage as the unsharded call"""
    import torch
    from pcramp_b200.sharding import shard_bounds, shard_sizes, shard_words
    coll = synth.make_targets(951, 70, 3000, n_clades=3, between=0.12, within=0.05)
    coll.weight = np.random.default_rng(5).uniform(0.1, 3.0, size=coll.n).astype(np.float32)
    f, r = synth.make_pairs(952, coll, 50)
    thr = float(np.float32(1.0) * np.float32(0.9))
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length, coll.weight)
    gpu.select_words(TARGET, f, r, thr)
    cov_all, bits_all = gpu.score_pairs(TARGET, f, r, thr, 0.9)
    world = 3
    b = shard_bounds(coll.n, world)
    words = shard_words(coll.n, world)
    any_parts, p1_parts = [], []
    for k in range(world):
        sh = coll.subset(range(b[k], b[k + 1]))
        gpu.upload_sequences(TARGET, sh.nibbles, sh.byte_off, sh.length, sh.weight)
        gpu.stage_pairs(f, r)
        gpu.select_words_staged(TARGET, thr)
        gpu.score_pairs_staged(TARGET, thr, 0.9)
        d_cov, d_any, d_p1 = gpu.device_pointers()
        n = len(f) * words[k]

        class Dev:
            def __init__(self, ptr):
                self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (n,), "typestr": "<i4", "version": 2}
        any_parts.append(torch.as_tensor(Dev(d_any), device="cuda").clone())
        p1_parts.append(torch.as_tensor(Dev(d_p1), device="cuda").clone())
    g_any, g_p1 = torch.cat(any_parts), torch.cat(p1_parts)
    out_bits = torch.zeros((len(f), (coll.n + 31) // 32), dtype=torch.int32, device="cuda")
    out_cov = torch.zeros(len(f), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    gpu.merge_shards(g_any.data_ptr(), g_p1.data_ptr(), shard_sizes(coll.n, world), len(f), out_bits.data_ptr(), out_cov.data_ptr(), coll.weight)
    assert np.array_equal(out_bits.cpu().numpy().view(np.uint32), bits_all)
    assert np.array_equal(out_cov.cpu().numpy(), cov_all)
    assert bits_all.any()


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_fused_sequence_scoring_equals_item_list(gpu, name):
    """the per-sequence scoring kernel (entries staged in shared memory, pairs read off the bit rows) == the item list +
    one warp per item it replaces: coverages and bitsets identical, for the oligos of the database and for others"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    out = []
    for fused in (0, 1):
        gpu.set_option("use_fused_score", fused)
        gpu.set_option("use_entry_score", 0)
        try:
            cov, bits = gpu.score_pairs(TARGET, sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            cov2, bits2 = gpu.score_pairs(TARGET, sc.r, np.roll(sc.f, 1, axis=0), sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            out.append((cov, bits, cov2, bits2))
        finally:
            gpu.set_option("use_fused_score", 0)
            gpu.set_option("use_entry_score", 1)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_entry_driven_scoring_equals_item_list(gpu, name):
    """pair scoring driven by the plus-strand entries (binary search for the partner among the minus-strand entries in amplicon
    range; score.cuh, the default) == bit rows + item list + one warp per item: coverages and bitsets identical, for the oligos of
    the database, for others, and in variant mode (the scoring of an optimisation move)"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    out = []
    rolled = np.roll(sc.f, 1, axis=0)
    for entry in (0, 1):
        gpu.set_option("use_entry_score", entry)
        try:
            a = gpu.score_pairs(TARGET, sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            b = gpu.score_pairs(TARGET, sc.r, rolled, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            c = gpu.score_variants(TARGET, sc.f, sc.r, sc.f, np.roll(sc.r, 1, axis=0), sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1],
                                   sc.taq)
            d = gpu.score_pairs(TARGET, sc.f, sc.r, sc.search_threshold, sc.target_threshold, 0, 2000, sc.taq)
            out.append(a + b + c + d)
        finally:
            gpu.set_option("use_entry_score", 1)
    for x, y in zip(out[0], out[1]):
        assert np.array_equal(np.asarray(x).view(np.uint32), np.asarray(y).view(np.uint32))


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_segmented_database_equals_sorted_ids(gpu, name):
    """the database built from per-(sequence, strand) segments sorted in shared memory (db.cuh, the default) == the radix sort +
    unique-by-key of the entry ids: same entries, keys, coverages and bitsets (the words are materialised on demand in the first)"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    out = []
    for seg in (0, 1):
        gpu.set_option("use_segmented_db", seg)
        try:
            g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
            cov, bits = gpu.score_pairs(TARGET, sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
            out.append(_db_tuple(gpu) + (cov, bits))
        finally:
            gpu.set_option("use_segmented_db", 1)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_segments_longer_than_shared_memory(gpu, oracle):
    """a tandem repeat: thousands of hits in ONE (sequence, strand) segment -- the in-place global-memory sort of db.cuh -- next to
    ordinary sequences; against the oracle and the sorted-id build"""
    rng = np.random.default_rng(91)
    unit = synth.CODE[rng.integers(0, 4, size=120)]
    rep = np.tile(unit, 450)                                    # 54 kb of one 120-base unit
    base = synth.make_targets(92, 6, 4000, n_clades=2, between=0.1, within=0.03)
    coll = synth.Collection([rep] + [base.codes(i) for i in range(base.n)])
    f, r = synth.make_pairs(93, synth.Collection([np.tile(unit, 4)]), 12)
    f2, r2 = synth.make_pairs(94, base, 20)
    f, r = np.concatenate([f, f2]), np.concatenate([r, r2])
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    oracle.set_sequences(coll)
    ne, nk = g.select_words(f, r, THR_09)
    assert (ne, nk) == oracle.select_words(f, r, THR_09) and ne > 5000
    for a, c in zip(g.db(), oracle.db()):
        assert np.array_equal(a, c)
    cov_o, bits_o = oracle.score_pairs(f, r, THR_09, 1.0, 80, 200, False)
    cov_g, bits_g = g.score_pairs(f, r, THR_09, 1.0, 80, 200, False)
    assert np.array_equal(bits_g, bits_o) and np.array_equal(cov_g, cov_o)
    seg = _db_tuple(gpu)
    gpu.set_option("use_segmented_db", 0)
    try:
        g.select_words(f, r, THR_09)
        for a, b in zip(seg, _db_tuple(gpu)):
            assert np.array_equal(a, b)
    finally:
        gpu.set_option("use_segmented_db", 1)


def test_fast_form_equals_general_form(oracle):
    """batches of one shape after the first run without the host in the loop (select_words_fast): databases, coverages and bitsets
    equal the oracle's; a batch that breaks an assumption -- more hits than the buffers were sized for (tiny_buffers keeps them
    tight), a primer too degenerate for the index (more than 16 letter combinations in a segment prefix) -- is detected from the read-back and re-run in the general form"""
    from pcramp_b200 import PcrampGpu
    base = synth.make_targets(131, 60, 7000, n_clades=3, between=0.12, within=0.04)
    loner = synth.make_targets(135, 1, 7000)                       # related to nothing: its primers hit one sequence
    coll = synth.Collection([base.codes(i) for i in range(base.n)] + [loner.codes(0)])
    f, r = synth.make_pairs(132, base, 180)
    fd, rd = synth.make_pairs(133, base, 60, degenerate_fraction=1.0)
    for w in fd:                                                   # five two-letter positions at the 5' end: 32 letter combinations in
        nib = [(int(w[i // 16]) >> ((15 - i % 16) * 4)) & 15 for i in range(32)]   # one segment prefix, more than the index enumerates
        pos = [i for i in range(32) if nib[i]][:5]
        for i in pos:
            nib[i] |= int(synth.CODE[(int(np.log2(nib[i] & -nib[i])) + 1) % 4]) if bin(nib[i]).count("1") == 1 else 0
        w[0] = sum(nib[i] << ((15 - i) * 4) for i in range(16))
        w[1] = sum(nib[i] << ((31 - i) * 4) for i in range(16, 32))
    oracle.set_sequences(coll)
    g = PcrampGpu(0)
    try:
        g.set_option("tiny_buffers", 1)
        g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
        chk = GpuChecker(g)
        chk.n_seq = coll.n

        def batch(fb, rb, staged):
            if staged:
                g.stage_pairs(fb, rb)
                g.select_words_staged(TARGET, THR_09, want_keys=False, want_entries=False)
                g.score_pairs_staged(TARGET, THR_09, 1.0)
                cov, bits = g.fetch_results(TARGET)
            else:
                g.select_words(TARGET, fb, rb, THR_09, want_keys=False, want_entries=False)
                cov, bits = g.score_pairs(TARGET, fb, rb, THR_09, 1.0)
            no, nko = oracle.select_words(fb, rb, THR_09)
            cov_o, bits_o = oracle.score_pairs(fb, rb, THR_09, 1.0, 80, 200, False)
            assert np.array_equal(cov, cov_o) and np.array_equal(unpack_bits(bits, coll.n), bits_o)
            assert g.db_size(TARGET) == (no, nko)
            for a, c in zip(chk.db(), oracle.db()):
                assert np.array_equal(a, c)
            return g.stats()
        # 60 pairs cut from the loner first: a few hundred hits, tight buffers
        f1, r1 = synth.make_pairs(134, loner, 60)
        st = batch(f1, r1, False)
        assert st["n_fast"] == 0                                   # the first batch of a shape: general form
        st = batch(f1[::-1].copy(), r1[::-1].copy(), True)
        assert st["n_fast"] == 1 and st["n_fast_redo"] == 0        # the second: fast, verified
        st = batch(f[:60], r[:60], True)                           # pairs from everywhere: several times the hits -> overflow -> re-run
        assert st["n_fast"] == 2 and st["n_fast_redo"] == 1
        st = batch(f[60:120], r[60:120], False)                    # fast again (the buffers stay tight: it may overflow once more)
        assert st["n_fast"] == 3 and st["n_fast_redo"] in (1, 2)
        if st["n_fast_redo"] == 2:                                 # the re-run is a general-form batch: the next one is fast again
            st = batch(f[60:120], r[60:120], True)
        redo, fast = st["n_fast_redo"], st["n_fast"]
        st = batch(fd, rd, True)                                   # very degenerate primers: the index cannot take them all
        assert st["n_fast"] == fast + 1 and st["n_fast_redo"] == redo + 1
        st = batch(f[120:180], r[120:180], True)                   # and the general form of that batch withdrew the hint
        assert st["n_fast"] == fast + 1
    finally:
        g.close()


def _end_primers(rng, coll, seqs, n_each):
    """primers cut from the first and last ~35 bases of sequences: the ones partial words (FILL / TAIL / EOS events) can match"""
    out = []
    for s_ in seqs:
        c = coll.codes(int(s_))
        L = len(c)
        for _ in range(n_each):
            n = int(rng.integers(18, 26))
            if L < n:
                continue
            a = int(rng.integers(0, min(10, L - n + 1)))
            out.append(synth.word_from_codes(c[a:a + n]))
            b = L - n - int(rng.integers(0, min(10, L - n + 1)))
            out.append(synth.word_from_codes(synth.revcomp_codes(c[b:b + n])))
    return np.array(out, np.uint64)


def test_partial_word_table_equals_scan_and_oracle(oracle):
    """edge.cuh: in the fast form the candidates look themselves up in the collection's table of partial words.  Primers cut from
    sequence ends (so that FILL / TAIL / EOS-event words really match), short sequences, IUPAC codes inside the first and last 32
    bases, splits and switched-off sequences: table == scan kernel == oracle, and the table follows splits.  (Every batch the fast
    form accepts has pieces of five slots or more -- shorter pieces are not seedable, select_words takes the general form -- so the
    table's runs of five serve all of them; the flag for a candidate it cannot serve is a guard.)"""
    from pcramp_b200 import PcrampGpu
    from bench_legs import widen
    rng = np.random.default_rng(91)
    base = synth.make_targets(141, 36, 420, n_clades=2, between=0.10, within=0.03)
    seqs = [base.codes(i).copy() for i in range(base.n)]
    for i in (3, 9, 20):                                            # degenerate bases near both ends: words without a seed code
        seqs[i][int(rng.integers(2, 30))] = 15
        seqs[i][len(seqs[i]) - int(rng.integers(2, 30))] = 5
    for n in (17, 18, 24, 31, 32, 33):                              # shorter than / about one word
        seqs.append(seqs[1][:n].copy())
        seqs.append(seqs[2][-n:].copy())
    coll = synth.Collection(seqs)
    ends = _end_primers(rng, coll, range(0, 36, 3), 3)
    inner_f, inner_r = synth.make_pairs(142, base, 40)
    P = 48
    f = np.concatenate([ends[0::2][:P - 16], inner_f[:16]])
    r = np.concatenate([ends[1::2][:P - 16], inner_r[:16]])
    f2, r2 = widen(f[::-1].copy(), rng), widen(r[::-1].copy(), rng)
    assert len(f) == P and len(r) == P
    active = np.ones(coll.n, np.uint8)
    active[[5, 11, 40]] = 0
    oracle.set_sequences(coll, active)
    g = PcrampGpu(0)
    try:
        g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
        g.set_active(TARGET, active)
        chk = GpuChecker(g)
        chk.n_seq = coll.n

        def batch(fb, rb, thr, table):
            g.set_option("use_edge_table", table)
            g.select_words(TARGET, fb, rb, thr, want_keys=False, want_entries=False)
            cov, bits = g.score_pairs(TARGET, fb, rb, thr, 1.0)
            no, nko = oracle.select_words(fb, rb, thr)
            cov_o, bits_o = oracle.score_pairs(fb, rb, thr, 1.0, 80, 200, False)
            assert g.db_size(TARGET) == (no, nko)
            db_o = oracle.db()
            for a, c in zip(chk.db(), db_o):
                assert np.array_equal(a, c)
            assert np.array_equal(cov, cov_o) and np.array_equal(unpack_bits(bits, coll.n), bits_o)
            w = db_o[0]                                               # entries that are partial words: an empty first or last slot
            partial = int((((w[:, 0] >> np.uint64(60)) == 0) | ((w[:, 1] & np.uint64(15)) == 0)).sum())
            return g.stats(), partial
        def fast_batch(fb, rb, thr, table):
            """a batch that must run in the fast form; if its buffers (sized by the batch before) overflowed it was re-run in the
            general form, which grew them: the same batch again is then fast and verified"""
            before = g.stats()
            st, _ = batch(fb, rb, thr, table)
            assert st["n_fast"] == before["n_fast"] + 1
            if st["n_fast_redo"] != before["n_fast_redo"]:
                st2, _ = batch(fb, rb, thr, table)
                assert st2["n_fast"] == st["n_fast"] + 1 and st2["n_fast_redo"] == st["n_fast_redo"]
                st = st2
            assert st["edge_table_used"] == table
            return st
        st, partial = batch(f, r, THR_09, 1)
        assert st["n_fast"] == 0 and partial >= 20                  # general form; the end primers do find partial words
        st = fast_batch(f2, r2, THR_09, 1)
        assert st["n_edge_words"] > 1000
        hits_table = st["n_hits"]
        st = fast_batch(f2, r2, THR_09, 0)
        assert st["n_hits"] == hits_table
        # splits make new partial words (EOS events) and retire old ones: the table is rebuilt
        for s_, p_ in ((0, 25), (0, 200), (6, 31), (6, 32), (12, 400), (12, 401), (37, 9), (21, 100)):
            g.split_sequence(TARGET, s_, p_)
            oracle.split_sequence(s_, p_)
        active[[5, 40]] = 1
        active[[0, 15]] = 0
        g.set_active(TARGET, active)
        oracle.set_active(active)
        near = np.array([synth.word_from_codes(coll.codes(12)[401 + 1:401 + 21]), synth.word_from_codes(coll.codes(0)[26:26 + 22])], np.uint64)
        f3, r3 = f.copy(), r.copy()
        f3[:2] = near                                                # primers that start right behind a split
        fast = g.stats()["n_fast"]
        st, _ = batch(f3, r3, THR_09, 1)                             # (the split withdrew the hint: general form)
        assert st["n_fast"] == fast
        st = fast_batch(f3[::-1].copy(), r3[::-1].copy(), THR_09, 1)
        hits_table = st["n_hits"]
        st = fast_batch(f3[::-1].copy(), r3[::-1].copy(), THR_09, 0)
        assert st["n_hits"] == hits_table
        # another threshold, same table (it belongs to the collection and the pack() parameters, not to the batch)
        thr95 = float(np.float32(1.0) * np.float32(0.95))
        batch(f, r, thr95, 1)
        st = fast_batch(f2, r2, thr95, 1)
        assert st["n_edge_words"] > 1000
    finally:
        g.close()


def test_degenerate_primers_through_the_index(gpu, oracle):
    """primers as `-d 16` leaves them (up to four two-letter positions): the degenerate positions of a segment prefix are enumerated
    letter by letter in the index queries -- same database, keys, coverage and bits as the oracle and as the table scan, the same
    number of hits (no alignment reported twice), and the patterns really take the index"""
    from bench_legs import widen
    coll = synth.make_targets(801, 30, 6000, n_clades=3, between=0.10, within=0.04)
    f, r = synth.make_pairs(802, coll, 300)
    rng = np.random.default_rng(5)
    f, r = widen(f, rng), widen(r, rng)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    oracle.set_sequences(coll)
    out = []
    for use in (0, 1):
        gpu.set_option("use_index", use)
        try:
            ne, nk = g.select_words(f, r, thr)
            out.append((g.db(), g.keys(), gpu.stats(), (ne, nk)))
        finally:
            gpu.set_option("use_index", 1)
    no, nko = oracle.select_words(f, r, thr)
    assert out[1][3] == (no, nko) == out[0][3] and no > 1000
    for a, b, c in zip(out[1][0], out[0][0], oracle.db()):
        assert np.array_equal(a, c) and np.array_equal(b, c)
    assert np.array_equal(out[1][1], oracle.keys())
    st = out[1][2]
    assert st["n_indexed"] > 0.8 * st["n_seeded"] and st["n_hits"] == out[0][2]["n_hits"]
    for search, detect in ((thr, 1.0), (1.0, 1.0)):
        cov_o, bits_o = oracle.score_pairs(f, r, search, detect, 80, 200, False)
        cov_g, bits_g = g.score_pairs(f, r, search, detect, 80, 200, False)
        assert np.array_equal(bits_g, bits_o) and np.array_equal(cov_g, cov_o)
    assert int(bits_o.sum()) > 0


def test_seed_table_reports_a_site_once_when_the_bucket_partner_matches(gpu, oracle):
    """A 7-base seed bucket holds two codes (base 0 and base 6 off by one code bit each).  Primer CTGAAT MGCGCTC ATTTGGG on the site
    CTGAAA AGCGCTT ATTTGGG: piece 1's text is the bucket partner of the M = C expansion, M admits the A, and the alignment passes with
    two mismatches -- the table scan used to report it through piece 1 AND through the exact piece 2 (found by scripts/stress_parity.py:
    hits 747684 by the table against 747558 through the index, databases identical).  One site, one hit, by both scans."""
    rng = np.random.default_rng(77)
    site = [int(synth.CODE["ACGT".index(c)]) for c in "CTGAAAAGCGCTTATTTGGG"]
    seqs = []
    for k in range(4):
        c = synth.CODE[rng.integers(0, 4, 700)].astype(np.uint8)
        at = 100 + 37 * k
        c[at:at + 20] = site if k % 2 == 0 else synth.revcomp_codes(site)
        seqs.append(c)
    coll = synth.Collection(seqs)
    f = np.array([synth.word_from_string("CTGAATMGCGCTCATTTGGG")], np.uint64)
    r = np.array([synth.word_from_codes(seqs[0][400:420])], np.uint64)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = GpuChecker(gpu)
    g.set_sequences(coll)
    oracle.set_sequences(coll)
    no, nko = oracle.select_words(f, r, thr)
    hits = []
    for use in (0, 1):
        gpu.set_option("use_index", use)
        try:
            assert g.select_words(f, r, thr) == (no, nko)
            for a, c in zip(g.db(), oracle.db()):
                assert np.array_equal(a, c)
            hits.append(gpu.stats()["n_hits"])
        finally:
            gpu.set_option("use_index", 1)
    assert no >= 5 and hits[0] == hits[1] == no  # four embedded sites + the plain primer's own, each seen by one pattern only


@pytest.mark.parametrize("name", ["basic", "degenerate", "splits", "shift"])
def test_async_index_scan_equals_register_scan(gpu, name):
    """scan_index_async_kernel (a cp.async ring of chunks in shared memory; option use_async_scan) == scan_index_kernel"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    out = []
    for use in (0, 1):
        gpu.set_option("use_async_scan", use)
        try:
            g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
            out.append((_db_tuple(gpu), gpu.stats()["n_hits"]))
        finally:
            gpu.set_option("use_async_scan", 0)
    for a, b in zip(out[0][0], out[1][0]):
        assert np.array_equal(a, b)
    assert out[0][1] == out[1][1]


@pytest.mark.parametrize("name", ["lowthr", "taq_weights", "repeats"])
def test_unit_scoring_equals_key_matrix(gpu, name):
    """pair scoring by (sequence, pair) units (option use_unit_score; thresholds at which the filters exclude nobody) == the key-matrix path"""
    sc = SCENARIOS[name]()
    g = GpuChecker(gpu)
    g.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        g.split_sequence(s, p)
    g.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    out = []
    for use in (1, 0):
        gpu.set_option("use_unit_score", use)
        try:
            res = []
            for search, detect in ((sc.search_threshold, sc.target_threshold), (0.72, 0.8), (0.6, 0.7)):
                res.append(g.score_pairs(sc.f, sc.r, float(search), float(detect), sc.amp[0], sc.amp[1], sc.taq))
            out.append(res)
        finally:
            gpu.set_option("use_unit_score", 1)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert any(a[1].sum() > 0 for a in out[0])
