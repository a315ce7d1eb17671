"""Seeded NucCruc test problems shared by the golden generator, the CPU tier and the GPU tier."""
import random

import numpy as np

OPS = (0, 1, 2, 3, 4, 5)  # pm duplex, hairpin, homodimer, heterodimer, heterodimer (diagonal), homodimer (diagonal)
TWO_SEQ = (3, 4)
SALTS = (0.05, 0.2)


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def _mutate(rng, s, k):
    s = list(s)
    for _ in range(k):
        i = rng.randrange(len(s))
        r = rng.random()
        if r < 0.6:
            s[i] = rng.choice("ACGT")
        elif r < 0.8 and len(s) > 8:
            del s[i]
        elif len(s) < 32:
            s.insert(i, rng.choice("ACGT"))
    return "".join(s)


def _rnd(rng, lo, hi, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(rng.randint(lo, hi)))


def problems(seed, n, op):
    """-> (seq_a, seq_b, strand_a, strand_b): primer-like random oligos, low-complexity ones, built-in hairpins / self-dimers,
    and for the two-sequence ops near-complements with substitutions and indels"""
    rng = random.Random(seed * 1000 + op)
    A, B = [], []
    for _ in range(n):
        kind = rng.random()
        if kind < 0.4:
            a = _rnd(rng, 15, 32)
        elif kind < 0.55:
            a = _rnd(rng, 6, 32, rng.choice(["ACGT", "AT", "GC", "ACGTGC"]))
        elif kind < 0.8:
            h = _rnd(rng, 4, 12)
            a = (_rnd(rng, 0, 4) + h + _rnd(rng, 3, 8) + _mutate(rng, revcomp(h), rng.randint(0, 2)) + _rnd(rng, 0, 4))[:32]
        else:
            h = _rnd(rng, 5, 15)
            a = (h + _mutate(rng, revcomp(h), rng.randint(0, 3)))[:32]
        if len(a) < 5:
            a += "ACGTA"
        if op in TWO_SEQ:
            b = _mutate(rng, revcomp(a), rng.randint(0, 5))[:32] if rng.random() < 0.6 else _rnd(rng, 15, 32)
            if op != 0 and rng.random() < 0.05:
                b = b[:3] + "I" + b[4:]
        else:
            b = a
        if op != 0 and rng.random() < 0.03:
            a = a[:2] + "I" + a[3:]
        A.append(a)
        B.append(b)
    sa = np.array([rng.choice([9e-7, 9e-7 / 4, 2e-7, 1e-6]) for _ in range(n)], dtype=np.float32)
    sb = np.array([rng.choice([9e-7, 9e-7 / 2, 1e-6]) for _ in range(n)], dtype=np.float32)
    return A, (B if op in TWO_SEQ else None), sa, sb


def primer_words(seed, n, word_from_string, degenerate=True, lo=18, hi=25):
    """n centred primer words, some with IUPAC degenerate positions (degeneracy <= 16)"""
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        s = list(_rnd(rng, lo, hi))
        kind = rng.random()
        if kind < 0.12:  # self-complementary: a strong homodimer
            h = _rnd(rng, (lo + 1) // 2, hi // 2)
            s = list(h + revcomp(h))
        elif kind < 0.24:  # stem-loop: a strong hairpin
            h = _rnd(rng, 6, 8)
            s = list((h + _rnd(rng, 4, 6) + revcomp(h) + _rnd(rng, 0, 4))[:hi])
            while len(s) < lo:
                s.append(rng.choice("ACGT"))
        if degenerate and rng.random() < 0.5:
            for _ in range(rng.randint(1, 3)):
                s[rng.randrange(len(s))] = rng.choice("RYKMSW" if rng.random() < 0.8 else "BDHVN")
            d = 1
            for ch in s:
                d *= {"R": 2, "Y": 2, "K": 2, "M": 2, "S": 2, "W": 2, "B": 3, "D": 3, "H": 3, "V": 3, "N": 4}.get(ch, 1)
            if d > 16:
                s = [c if c in "ACGT" else "A" for c in s]
        out.append(word_from_string("".join(s), True))
    return np.array(out, dtype=np.uint64)
