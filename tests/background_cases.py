"""Seeded cases for K4 (Smith-Waterman + find_background_match + find_multiplex_background_match), shared by the
golden generator, the CPU tier and the GPU tier."""
import numpy as np

from pcramp_b200 import synth

BG_THRESHOLD = np.float32(0.8)      # DEFAULT_BACKGROUND_THRESHOLD (pcramp.h:40)
BG_MULT = np.float32(0.9)           # DEFAULT_SEARCH_THRESHOLD_MULTIPLIER (pcramp.h:51)
BG_AMP = (0, 2000)                  # DEFAULT_MIN/MAX_BACKGROUND_AMPLICON (pcramp.h:17-18)
BG_MIN_LEN = int(18 * 0.9)          # main.cpp:592-595


def sw_problems(seed, n, word_from_string):
    """(query words, target words): primer-like queries against related / unrelated 32-mers, IUPAC codes, short operands"""
    rng = np.random.default_rng(seed)
    sym = list("ACGTMRSVWYHKDBN")

    def rnd(lo, hi, deg=0.05):
        k = int(rng.integers(lo, hi + 1))
        p = [(1 - deg) / 4] * 4 + [deg / 11] * 11
        return "".join(rng.choice(sym, size=k, p=p))

    def mutate(s, k):
        s = list(s)
        for _ in range(k):
            i = int(rng.integers(0, len(s)))
            u = rng.random()
            if u < 0.5:
                s[i] = str(rng.choice(list("ACGT")))
            elif u < 0.75 and len(s) > 4:
                del s[i]
            elif len(s) < 32:
                s.insert(i, str(rng.choice(list("ACGT"))))
        return "".join(s)

    Q, T = [], []
    for _ in range(n):
        q = rnd(1, 32) if rng.random() < 0.2 else rnd(15, 27)
        if rng.random() < 0.5:
            t = (rnd(0, 8) + mutate(q, int(rng.integers(0, 6))) + rnd(0, 8))[:32]
        else:
            t = rnd(1, 32)
        Q.append(word_from_string(q, bool(rng.integers(0, 2))))
        T.append(word_from_string(t or "A", bool(rng.integers(0, 2))))
    return np.array(Q, dtype=np.uint64), np.array(T, dtype=np.uint64)


class BgCase:
    def __init__(self, name, coll, f, r, taq=False, splits=()):
        self.name, self.coll, self.f, self.r, self.taq, self.splits = name, coll, f, r, taq, list(splits)

    @property
    def search_threshold(self):
        return float(BG_THRESHOLD * BG_MULT)   # assay.h:417: float product


def bg_cases():
    """background collections near the primers' source (so that the permissive 0.8*0.9 seed threshold finds many binding
    sites), one with few sequences and many candidate amplicons per pair (the odd-index guard of
    background_match.cpp:122 then drops candidates), one with splits and TaqMAMA"""
    out = []
    src = synth.make_targets(51, 6, 1500, n_clades=2, between=0.10, within=0.03)
    near = synth.make_targets(52, 14, 1500, n_clades=3, between=0.12, within=0.06)
    codes = [near.codes(i).copy() for i in range(near.n)]
    for i in range(0, near.n, 2):              # graft stretches of the primers' source into the background
        j = i % src.n
        codes[i][200:900] = src.codes(j)[200:900]
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(53, src, 40, amplicon_range=(80, 600))
    out.append(BgCase("near", coll, f, r))
    # few sequences, tandem repeats of the amplified region: dozens of candidate amplicons per pair, more than sequences
    unit = src.codes(0)[100:400]
    rep = [np.concatenate([src.codes(1)[:50], unit, unit, unit, unit, src.codes(1)[50:80]]),
           np.concatenate([unit, src.codes(2)[:40], unit, unit]),
           np.concatenate([src.codes(3)[:300], unit])]
    coll2 = synth.Collection(rep)
    base = synth.Collection([unit])
    f2, r2 = synth.make_pairs(54, base, 24, amplicon_range=(80, 250))
    out.append(BgCase("repeats", coll2, f2, r2))
    codes3 = [c.copy() for c in codes[:8]]
    out.append(BgCase("taq_splits", synth.Collection(codes3), f[:24], r[:24], taq=True, splits=[(0, 500), (2, 350), (2, 351), (5, 10)]))
    # The reference scores F(+)&R(-) as SW(F, f-key) * SW(rc(R), r-key) (background_match.cpp:83: slots 0 and 3), and the r-key is a
    # word R matches in its own orientation -- so a sequence is only flagged when the reverse primer also resembles its own
    # reverse complement.  Palindromic primers make the product large: the only way to exercise the detected branch.
    rng = np.random.default_rng(55)
    pal_f, pal_r, pal_seqs = [], [], []
    sym = synth.CODE
    n_pal = 12
    sites = []
    for k in range(n_pal):
        hf = sym[rng.integers(0, 4, size=int(rng.integers(9, 13)))]
        hr = sym[rng.integers(0, 4, size=int(rng.integers(9, 13)))]
        F = np.concatenate([hf, synth.revcomp_codes(hf)])
        R = np.concatenate([hr, synth.revcomp_codes(hr)])
        if k % 3 == 0:                          # only the reverse primer palindromic
            F = sym[rng.integers(0, 4, size=20)]
        sites.append((F, R))
        pal_f.append(synth.word_from_codes(F))
        pal_r.append(synth.word_from_codes(synth.revcomp_codes(R)))
    for i in range(9):
        body = sym[rng.integers(0, 4, size=2600)]
        pos = 60
        for k, (F, R) in enumerate(sites):
            if (k + i) % 3 == 0:
                continue                        # this sequence lacks this assay
            Fm, Rm = F.copy(), R.copy()
            for _ in range((k + i) % 4):        # 0..3 substitutions per site
                Fm[int(rng.integers(0, len(Fm)))] = sym[int(rng.integers(0, 4))]
                Rm[int(rng.integers(0, len(Rm)))] = sym[int(rng.integers(0, 4))]
            body[pos:pos + len(Fm)] = Fm
            body[pos + 90:pos + 90 + len(Rm)] = Rm
            pos += 200
        pal_seqs.append(body)
    out.append(BgCase("palindromes", synth.Collection(pal_seqs), np.array(pal_f, dtype=np.uint64), np.array(pal_r, dtype=np.uint64)))
    # palindromic assays in tandem copies on only three sequences: far more candidates than sequences, the good copies placed
    # late, so whether a sequence is flagged depends on the odd-index guard of background_match.cpp:122
    rep_seqs = []
    for i in range(3):
        body = sym[rng.integers(0, 4, size=3000)]
        pos = 40
        for copy in range(6):
            for k, (F, R) in enumerate(sites[:4]):
                Fm, Rm = F.copy(), R.copy()
                n_mut = 5 if copy < 4 - (k % 2) else (copy + i) % 2   # early copies fall below the 0.8 threshold
                for _ in range(n_mut):
                    Fm[int(rng.integers(0, len(Fm)))] = sym[int(rng.integers(0, 4))]
                    Rm[int(rng.integers(0, len(Rm)))] = sym[int(rng.integers(0, 4))]
                body[pos:pos + len(Fm)] = Fm
                body[pos + 60:pos + 60 + len(Rm)] = Rm
                pos += 110
        rep_seqs.append(body)
    out.append(BgCase("palindrome_repeats", synth.Collection(rep_seqs), np.array(pal_f[:4], dtype=np.uint64), np.array(pal_r[:4], dtype=np.uint64)))
    out.append(BgCase("palindromes_taq", synth.Collection(pal_seqs[:5]), np.array(pal_f, dtype=np.uint64), np.array(pal_r, dtype=np.uint64), taq=True))
    return out


def multiplex_case():
    """multiplex background = amplicon-sized sequences (main.cpp:989-1008); some contain a primer site, some a split"""
    src = synth.make_targets(61, 4, 1200, n_clades=1, within=0.03)
    f, r = synth.make_pairs(62, src, 32, amplicon_range=(80, 300))
    rng = np.random.default_rng(63)
    seqs = []
    for k in range(21):
        i = int(rng.integers(0, src.n))
        a = int(rng.integers(0, 900))
        L = int(rng.integers(60, 300))
        c = src.codes(i)[a:a + L].copy()
        if k % 5 == 0:
            c[len(c) // 2] = 0                  # EOS inside
        if k % 7 == 0:
            c[5] = 15                           # N
        seqs.append(c)
    seqs.append(np.zeros(0, np.uint8))          # an empty sequence
    seqs.append(src.codes(0)[:17].copy())       # shorter than a primer
    return synth.Collection(seqs), f, r
