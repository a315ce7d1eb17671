"""Seeded parity scenarios shared by the oracle-vs-reference tests, the golden generator and the GPU parity tests.

Each scenario is small enough for the CPU oracle to finish in well under a second and pokes at one
family of edge cases of the reference: EOS absorption in Sequence::pack, short
and empty sequences, degenerate bases on either side, the GC and degeneracy filters, the 5'/3' shift
families of select_words, TaqMAMA, weights and inactive sequences, repeats and palindromes.
"""
import numpy as np

from pcramp_b200 import synth


class Scenario:
    def __init__(self, name, coll, f, r, **kw):
        self.name = name
        self.coll = coll
        self.f = f
        self.r = r
        self.active = kw.pop("active", None)
        # seed-scan parameters (main.cpp:585-601 / :657-669)
        self.threshold = np.float32(kw.pop("target_threshold", 1.0)) * np.float32(kw.pop("search_multiplier", 0.9))
        self.target_threshold = float(kw.pop("target_threshold_raw", 1.0))
        self.search_multiplier = float(kw.pop("search_multiplier_raw", 0.9))
        self.optimize_5 = kw.pop("optimize_5", False)
        self.optimize_3 = kw.pop("optimize_3", False)
        self.pack_max_degen = kw.pop("pack_max_degen", 256)
        self.pack_min_gc = kw.pop("pack_min_gc", 0.0)
        self.pack_max_gc = kw.pop("pack_max_gc", 1.0)
        self.min_oligo_length = kw.pop("min_oligo_length", 18)
        self.amp = kw.pop("amp", (80, 200))
        self.taq = kw.pop("taq", False)
        self.splits = kw.pop("splits", [])  # (seq, pos) applied with split_sequence after upload
        assert not kw, kw

    def select_kwargs(self):
        return dict(optimize_5=self.optimize_5, optimize_3=self.optimize_3, pack_max_degen=self.pack_max_degen,
                    pack_min_gc=self.pack_min_gc, pack_max_gc=self.pack_max_gc, min_oligo_length=self.min_oligo_length)

    @property
    def search_threshold(self):
        """float product exactly as assay.h:407 forms it"""
        return float(np.float32(self.target_threshold) * np.float32(self.search_multiplier))


def _thr(target, mult):
    return dict(target_threshold=target, search_multiplier=mult, target_threshold_raw=target, search_multiplier_raw=mult)


def _codes(coll):
    return [coll.codes(i).copy() for i in range(coll.n)]


def s_basic():
    coll = synth.make_targets(11, 12, 700, n_clades=2, between=0.15, within=0.04)
    f, r = synth.make_pairs(12, coll, 30)
    return Scenario("basic", coll, f, r, **_thr(1.0, 0.9))


def s_lowthr():
    """background-style parameters: threshold 0.8*0.9, shorter minimum word, long amplicon window (main.cpp:592-601)"""
    coll = synth.make_targets(21, 10, 900, n_clades=3, between=0.2, within=0.08)
    f, r = synth.make_pairs(22, coll, 24)
    return Scenario("lowthr", coll, f, r, **_thr(0.8, 0.9), min_oligo_length=int(18 * 0.9), amp=(0, 2000))


def s_eos_short():
    rng = np.random.default_rng(31)
    base = synth.make_targets(31, 10, 400, n_clades=1, within=0.03)
    codes = _codes(base)
    codes[0][150] = 0                       # one split in the steady state
    codes[1][10] = 0                        # EOS while the first word is still filling
    codes[1][11] = 0
    codes[2][31] = 0                        # exactly where the first full word would complete
    codes[2][32] = 0
    codes[3][0] = 0                         # leading EOS
    codes[3][399] = 0                       # trailing EOS
    codes[4][200:203] = 0                   # a run of EOS
    codes[4][260] = 0
    codes[5] = codes[5][:33]                # barely longer than a word
    codes[6] = codes[6][:32]                # exactly one word
    codes[7] = codes[7][:20]                # shorter than a word, longer than min_len
    codes[8] = codes[8][:17]                # shorter than min_len: emits nothing
    codes[9] = np.concatenate([codes[9][:120], np.zeros(1, np.uint8), codes[9][120:]])  # multi-record style pad
    codes.append(np.zeros(0, np.uint8))     # empty sequence
    codes.append(np.zeros(5, np.uint8))     # only EOS
    extra = codes[0].copy()
    extra[395:] = 0                         # several trailing EOS
    codes.append(extra)
    coll = synth.Collection(codes)
    clean = synth.Collection([c[c != 0] for c in codes[:5]])
    f, r = synth.make_pairs(32, clean, 40)
    # primers cut right at sequence starts / ends so the partial (centred) words matter
    c0 = codes[5]
    f[0] = synth.word_from_codes(c0[:20])
    r[0] = synth.word_from_codes(synth.revcomp_codes(c0[-19:]))
    c7 = codes[7]
    f[1] = synth.word_from_codes(c7[:18])
    r[1] = synth.word_from_codes(synth.revcomp_codes(c7[-18:]))
    c3 = codes[3][codes[3] != 0]
    f[2] = synth.word_from_codes(c3[:22])
    r[2] = synth.word_from_codes(synth.revcomp_codes(c3[-25:]))
    return Scenario("eos_short", coll, f, r, **_thr(1.0, 0.9), amp=(10, 400))


def s_degenerate():
    rng = np.random.default_rng(41)
    base = synth.make_targets(41, 10, 600, n_clades=2, between=0.12, within=0.04)
    codes = _codes(base)
    for i in range(10):
        k = rng.integers(0, 600, size=12)
        codes[i][k] |= synth.CODE[rng.integers(0, 4, size=12)]   # scattered two-letter codes
    codes[0][100:106] = 15                   # 6 N: degeneracy 4096 > 256 -> windows dropped
    codes[1][300:304] = 15                   # 4 N: 256, not > 256 -> kept
    codes[2][0:5] = 15                       # N run at the very start
    codes[3][595:600] = 15                   # ... and at the very end
    codes[4][200:203] = 7                    # V V V
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(42, base, 40, degenerate_fraction=0.7)
    return Scenario("degenerate", coll, f, r, **_thr(1.0, 0.9))


def s_gc():
    coll = synth.make_targets(51, 8, 800, n_clades=2, between=0.1, within=0.03)
    codes = _codes(coll)
    codes[0][300:360] = 1                    # an AT desert
    codes[1][100:170] = 4                    # a G island
    codes[2][50] = 0                         # GC window with an EOS inside
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(52, coll.subset(range(3, 8)), 30)
    return Scenario("gc", coll, f, r, **_thr(1.0, 0.9), pack_min_gc=0.35, pack_max_gc=0.62)


def s_shift():
    coll = synth.make_targets(61, 8, 500, n_clades=2, between=0.1, within=0.05)
    f, r = synth.make_pairs(62, coll, 12)
    return Scenario("shift", coll, f, r, **_thr(1.0, 0.9), optimize_5=True, optimize_3=True)


def s_taq_weights():
    coll = synth.make_targets(71, 14, 600, n_clades=2, between=0.1, within=0.06)
    w = np.random.default_rng(71).uniform(0.1, 3.0, size=14).astype(np.float32)
    coll.weight = w
    active = np.ones(14, np.uint8)
    active[[3, 8]] = 0
    f, r = synth.make_pairs(72, coll, 40)
    return Scenario("taq_weights", coll, f, r, **_thr(0.9, 0.9), taq=True, active=active)


def s_repeats():
    rng = np.random.default_rng(81)
    unit = synth.CODE[rng.integers(0, 4, size=37)]
    codes = []
    codes.append(np.tile(unit, 12))                                   # tandem repeat: many identical words
    codes.append(np.full(300, 1, np.uint8))                           # poly-A
    pal = synth.CODE[rng.integers(0, 4, size=60)]
    codes.append(np.concatenate([pal, synth.revcomp_codes(pal), pal, synth.revcomp_codes(pal)]))  # palindromes: (+) word == (-) word
    codes.append(np.concatenate([np.tile(unit, 3), synth.CODE[rng.integers(0, 4, size=150)], np.tile(unit, 3)]))
    coll = synth.Collection(codes)
    f = np.zeros((6, 2), np.uint64)
    r = np.zeros((6, 2), np.uint64)
    f[0] = synth.word_from_codes(unit[:20]); r[0] = synth.word_from_codes(synth.revcomp_codes(np.tile(unit, 4)[100:120]))
    f[1] = synth.word_from_codes(np.full(20, 1, np.uint8)); r[1] = synth.word_from_codes(np.full(19, 8, np.uint8))
    f[2] = synth.word_from_codes(pal[5:27]); r[2] = synth.word_from_codes(pal[30:50])
    f[3] = synth.word_from_codes(unit[10:30]); r[3] = synth.word_from_codes(synth.revcomp_codes(unit[3:24]))
    f[4] = synth.word_from_codes(pal[0:18]); r[4] = synth.word_from_codes(synth.revcomp_codes(pal[40:60]))
    f[5] = synth.word_from_codes(unit[:25]); r[5] = synth.word_from_codes(synth.revcomp_codes(unit[5:30]))
    return Scenario("repeats", coll, f, r, **_thr(1.0, 0.9), amp=(30, 200))


def s_splits():
    """split_sequence applied after upload (main.cpp:1010-1016) -- exercises the on-device re-compaction"""
    coll = synth.make_targets(91, 8, 500, n_clades=1, within=0.04)
    f, r = synth.make_pairs(92, coll, 30)
    splits = [(0, 250), (0, 251), (1, 40), (2, 499), (3, 0), (4, 100), (4, 300), (4, 301)]
    return Scenario("splits", coll, f, r, **_thr(1.0, 0.9), splits=splits)


def s_odd():
    """odd lengths: pack() also pushes the pad nibble of the last byte (sequence.cpp:111)"""
    base = synth.make_targets(101, 9, 420, n_clades=1, within=0.04)
    codes = _codes(base)
    lens = [419, 401, 333, 35, 33, 31, 19, 417, 1]
    codes = [c[:n] for c, n in zip(codes, lens)]
    codes[7][416] = 0                        # a real EOS right before the pad nibble
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(102, synth.Collection(codes[:3]), 30)
    c = codes[0]
    f[0] = synth.word_from_codes(c[300:320]); r[0] = synth.word_from_codes(synth.revcomp_codes(c[-21:]))
    c = codes[3]
    f[1] = synth.word_from_codes(c[:18]); r[1] = synth.word_from_codes(synth.revcomp_codes(c[-18:]))
    return Scenario("odd", coll, f, r, **_thr(1.0, 0.9), amp=(20, 300))


ALL = [s_odd, s_basic, s_lowthr, s_eos_short, s_degenerate, s_gc, s_shift, s_taq_weights, s_repeats, s_splits]


def all_scenarios():
    return [fn() for fn in ALL]


def pack_record(w, idx, loc, st):
    """canonically ordered (word_hi, word_lo, loc, strand) rows of one sequence's pack()"""
    if len(loc) == 0:
        return np.zeros((0, 4), np.int64)
    return np.concatenate([np.ascontiguousarray(w).view(np.int64), loc[:, None].astype(np.int64), st[:, None].astype(np.int64)], axis=1)


def digest(a):
    import hashlib
    return int.from_bytes(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], "little")


def run_checker(chk, sc, kind="oracle"):
    """Run one scenario through OracleLib or RefLib; returns a dict of arrays (the golden record)."""
    chk.set_sequences(sc.coll, sc.active)
    for (s, p) in sc.splits:
        chk.split_sequence(s, p)
    out = {}
    packs, digests, counts = [], [], []
    for i in range(sc.coll.n):
        rec = pack_record(*chk.pack(i, sc.pack_max_degen, sc.pack_min_gc, sc.pack_max_gc, sc.min_oligo_length))
        counts.append(len(rec))
        digests.append(digest(rec))
        if i < 2 or sc.coll.length[i] <= 64:
            packs.append(rec)  # full Sequence::pack output for a few sequences, a digest for all of them
    out["pack_count"] = np.array(counts, np.int64)
    out["pack_digest"] = np.array(digests, np.uint64)
    out["pack_full"] = np.concatenate(packs, axis=0) if packs else np.zeros((0, 4), np.int64)
    ne, nk = chk.select_words(sc.f, sc.r, float(sc.threshold), **sc.select_kwargs())
    w, idx, loc, st = chk.db()
    out["db_words"], out["db_index"], out["db_loc"], out["db_strand"] = w, idx, loc, st
    out["keys"] = chk.keys()
    if kind == "ref":
        cov, bits = chk.score_pairs(sc.f, sc.r, sc.target_threshold, sc.search_multiplier, sc.amp[0], sc.amp[1], sc.taq)
    else:
        cov, _ = chk.score_pairs(sc.f, sc.r, sc.search_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
        _, bits = chk.score_pairs(sc.f, sc.r, sc.target_threshold, sc.target_threshold, sc.amp[0], sc.amp[1], sc.taq)
    out["coverage"] = cov
    out["bits"] = bits
    return out
