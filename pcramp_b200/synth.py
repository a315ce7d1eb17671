"""Synthetic inputs of the shapes BASELINE.json names and oligo packing.

All generators are pure functions of (seed, index) so that any rank, the CPU baseline and the tests
can regenerate exactly the same sequence without shipping data.

Nibble codes: A=1 C=2 G=4 T=8, 0 = EOS (base_table.h:9-28 of the reference).
"""
import numpy as np

CODE = np.array([1, 2, 4, 8], dtype=np.uint8)  # A C G T
_COMP = np.zeros(16, dtype=np.uint8)
for _b in range(16):
    _COMP[_b] = ((_b & 1) << 3) | ((_b & 8) >> 3) | ((_b & 2) << 1) | ((_b & 4) >> 1)
_SYM = "-ACMGRSVTWYHKDBN"


def _mutate(rng, letters, rate):
    """i.i.d. substitutions: each position is replaced, with probability `rate`, by one of the 3 other letters."""
    if rate <= 0.0:
        return letters.copy()
    hit = rng.integers(0, 1 << 16, size=letters.shape, dtype=np.uint16) < np.uint16(min(65535, int(rate * 65536.0)))
    shift = rng.integers(1, 4, size=letters.shape, dtype=np.uint8)
    return np.where(hit, (letters + shift) & 3, letters).astype(np.uint8)


class Collection:
    """A set of sequences as the reference stores them: packed nibbles + offsets + lengths (+ weights)."""

    def __init__(self, codes_list, weights=None):
        self.n = len(codes_list)
        self.length = np.array([len(c) for c in codes_list], dtype=np.uint32)
        nbytes = (self.length.astype(np.uint64) + 1) // 2
        # 16-byte aligned starts (not required by the ABI, friendlier to vector loads)
        padded = (nbytes + 15) // 16 * 16
        self.byte_off = np.zeros(self.n, dtype=np.uint64)
        if self.n:
            self.byte_off[1:] = np.cumsum(padded)[:-1]
        total = int(padded.sum()) if self.n else 0
        self.nibbles = np.zeros(max(total, 16), dtype=np.uint8)
        for i, c in enumerate(codes_list):
            c = np.asarray(c, dtype=np.uint8)
            if len(c) % 2:
                c = np.concatenate([c, np.zeros(1, np.uint8)])
            o = int(self.byte_off[i])
            self.nibbles[o:o + len(c) // 2] = (c[0::2] << 4) | c[1::2]
        self.weight = None if weights is None else np.asarray(weights, dtype=np.float32)

    def codes(self, i):
        o, L = int(self.byte_off[i]), int(self.length[i])
        b = self.nibbles[o:o + (L + 1) // 2]
        out = np.empty(2 * len(b), dtype=np.uint8)
        out[0::2] = b >> 4
        out[1::2] = b & 15
        return out[:L]

    def text(self, i):
        return "".join(_SYM[c] for c in self.codes(i))

    def subset(self, idx):
        w = None if self.weight is None else self.weight[idx]
        return Collection([self.codes(int(i)) for i in idx], w)


class TargetFactory:
    """n sequences of `length` nt: a random root, `n_clades` clade ancestors at `between` divergence from it,
    every sequence at `within` divergence from its clade ancestor (sequence i belongs to clade i % n_clades).
    Sequence i is a pure function of (seed, i), so shards and samples can be generated independently."""

    def __init__(self, seed, n, length, n_clades=1, between=0.0, within=0.03):
        self.seed, self.n, self.seq_len, self.n_clades, self.within = seed, n, length, n_clades, within
        root = np.random.default_rng([seed, 0]).integers(0, 4, size=length, dtype=np.uint8)
        self.clades = [_mutate(np.random.default_rng([seed, 1, c]), root, between) for c in range(n_clades)]
        self.length = np.full(n, length, dtype=np.uint32)

    def codes(self, i):
        return CODE[_mutate(np.random.default_rng([self.seed, 2, int(i)]), self.clades[int(i) % self.n_clades], self.within)]

    def collection(self, indices=None):
        idx = range(self.n) if indices is None else indices
        return Collection([self.codes(i) for i in idx])


def make_targets(seed, n, length, n_clades=1, between=0.0, within=0.03):
    return TargetFactory(seed, n, length, n_clades, between, within).collection()


def word_from_codes(codes, centre=True):
    """Pack <= 32 nibble codes into {buffer[0], buffer[1]} (word.h:290-317), optionally centred (word.h:392-418)."""
    n = len(codes)
    assert n <= 32
    shift = (33 - n) // 2 if (centre and n > 0) else 0
    hi = lo = 0
    for k, c in enumerate(codes):
        i = k + shift
        if i < 16:
            hi |= int(c) << ((15 - i) * 4)
        else:
            lo |= int(c) << ((31 - i) * 4)
    return hi, lo


def word_from_string(s, centre=True):
    from .api import load_library  # host helper of the C ABI (same packing as word_from_codes)
    import ctypes
    out = (ctypes.c_uint64 * 2)()
    load_library().pcramp_word_from_string(s.encode(), int(centre), out)
    return int(out[0]), int(out[1])


def revcomp_codes(codes):
    return _COMP[np.asarray(codes, dtype=np.uint8)[::-1]]


def make_pairs(seed, coll, n_pairs, primer_range=(18, 25), amplicon_range=(80, 200), degenerate_fraction=0.0):
    """Primer pairs with the geometry of PCR::random_assay (pcr_assay.cpp:636-688), without the thermodynamic
    filter: F = an 18-25-mer cut from a random sequence, R = reverse complement of an 18-25-mer downstream of it
    such that the amplicon is 80-200 nt.  `degenerate_fraction` of the primers get one position widened to a
    two-letter IUPAC code (config C3's "degenerate primers")."""
    rng = np.random.default_rng([seed, 3])
    if not isinstance(coll.length, np.ndarray):
        raise TypeError("coll must be a Collection or a TargetFactory")
    f = np.zeros((n_pairs, 2), dtype=np.uint64)
    r = np.zeros((n_pairs, 2), dtype=np.uint64)
    t = 0
    while t < n_pairs:
        i = int(rng.integers(0, coll.n))
        L = int(coll.length[i])
        f_len = int(rng.integers(primer_range[0], primer_range[1] + 1))
        r_len = int(rng.integers(primer_range[0], primer_range[1] + 1))
        amp = int(rng.integers(amplicon_range[0], amplicon_range[1] + 1))
        if amp < f_len + r_len or L < amp:
            continue
        f_start = int(rng.integers(0, L - amp + 1))
        r_start = f_start + amp - r_len
        c = coll.codes(i)[f_start:f_start + amp]
        fc = c[:f_len].copy()
        rc = revcomp_codes(c[amp - r_len:]).copy()
        if (fc == 0).any() or (rc == 0).any():
            continue
        if degenerate_fraction > 0.0:
            for oc in (fc, rc):
                if rng.random() < degenerate_fraction:
                    k = int(rng.integers(0, len(oc)))
                    oc[k] |= CODE[int(rng.integers(0, 4))]
        f[t] = word_from_codes(fc)
        r[t] = word_from_codes(rc)
        t += 1
    return f, r


def words_to_strings(w):
    out = []
    for hi, lo in np.asarray(w, dtype=np.uint64):
        s = []
        for i in range(32):
            limb = int(hi) if i < 16 else int(lo)
            c = (limb >> ((15 - (i % 16)) * 4)) & 15
            if c:
                s.append(_SYM[c])
        out.append("".join(s))
    return out
