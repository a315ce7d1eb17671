"""ctypes wrappers around the two CPU checkers (test infrastructure only):

  OracleLib  oracle/liboracle.so          -- my restatement of the reference algorithm
  RefLib     oracle/_ref/libpcramp_ref.so -- the unmodified reference sources + oracle/ref_driver.cpp

Both expose the same Python surface so tests can run one scenario through either.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libpcramp_ref.so")

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f32p = ctypes.POINTER(ctypes.c_float)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _w(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 2
    return a


def build_oracle():
    if not os.path.exists(ORACLE_PATH) or os.path.getmtime(ORACLE_PATH) < os.path.getmtime(os.path.join(ROOT, "oracle", "pcramp_oracle.cpp")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True, stdout=subprocess.DEVNULL)


def canonical(words, index, loc, strand):
    """Sort DB entries by (word, index, loc, strand) -- the reference's order among equal keys is unspecified."""
    words = np.asarray(words, dtype=np.uint64).reshape(-1, 2)
    order = np.lexsort((strand, loc, index, words[:, 1], words[:, 0]))
    return words[order], np.asarray(index)[order], np.asarray(loc)[order], np.asarray(strand)[order]


class _Base:
    prefix = ""

    def _fn(self, name, res, args):
        f = getattr(self.lib, self.prefix + name)
        f.restype = res
        f.argtypes = args
        return f

    def _words_api(self):
        P = self.prefix
        self.f_word_from_string = self._fn("word_from_string", None, [ctypes.c_char_p, ctypes.c_int, _u64p])
        self.f_word_and = self._fn("word_and", ctypes.c_uint32, [_u64p, _u64p])
        self.f_word_size = self._fn("word_size", ctypes.c_uint32, [_u64p])
        self.f_word_start = self._fn("word_start", ctypes.c_int, [_u64p])
        self.f_word_stop = self._fn("word_stop", ctypes.c_int, [_u64p])
        self.f_word_degeneracy = self._fn("word_degeneracy", ctypes.c_double, [_u64p])
        self.f_word_complement = self._fn("word_complement", None, [_u64p, _u64p])
        self.f_word_center = self._fn("word_center", None, [_u64p, _u64p])
        self.f_word_shift = self._fn("word_shift", None, [_u64p, ctypes.c_int, _u64p])
        self.f_word_push_back = self._fn("word_push_back", None, [_u64p, ctypes.c_uint8, _u64p])
        self.f_taq = self._fn("taq_mama", ctypes.c_float, [ctypes.c_uint8] * 4)
        self.f_word_expand = self._fn("word_expand", ctypes.c_long, [_u64p, ctypes.c_long, _u64p])

    # --- word helpers, value semantics ---
    @staticmethod
    def _mk(w):
        return (ctypes.c_uint64 * 2)(int(w[0]), int(w[1]))

    def word_from_string(self, s, centre=True):
        out = (ctypes.c_uint64 * 2)()
        self.f_word_from_string(s.encode(), int(centre), out)
        return (int(out[0]), int(out[1]))

    def word_and(self, a, b):
        return int(self.f_word_and(self._mk(a), self._mk(b)))

    def word_size(self, a):
        return int(self.f_word_size(self._mk(a)))

    def word_start(self, a):
        return int(self.f_word_start(self._mk(a)))

    def word_stop(self, a):
        return int(self.f_word_stop(self._mk(a)))

    def word_degeneracy(self, a):
        return float(self.f_word_degeneracy(self._mk(a)))

    def _unary(self, fn, a, *extra):
        out = (ctypes.c_uint64 * 2)()
        fn(self._mk(a), *extra, out)
        return (int(out[0]), int(out[1]))

    def word_complement(self, a):
        return self._unary(self.f_word_complement, a)

    def word_center(self, a):
        return self._unary(self.f_word_center, a)

    def word_shift(self, a, left):
        return self._unary(self.f_word_shift, a, int(left))

    def word_push_back(self, a, b):
        return self._unary(self.f_word_push_back, a, int(b))

    def taq_mama(self, p0, p1, t0, t1):
        return float(self.f_taq(p0, p1, t0, t1))

    def word_expand(self, a, cap=4096):
        out = np.zeros((cap, 2), np.uint64)
        n = self.f_word_expand(self._mk(a), cap, _p(out, _u64p))
        return n, out[:min(n, cap)]


class OracleLib(_Base):
    prefix = "oracle_"

    def __init__(self):
        build_oracle()
        self.lib = ctypes.CDLL(ORACLE_PATH)
        self._words_api()
        self.lib.oracle_create.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(self.lib.oracle_create())
        vp = ctypes.c_void_p
        self.f_set = self._fn("set_sequences", ctypes.c_int, [vp, ctypes.c_uint32, _u8p, _u64p, _u32p, _f32p, _u8p])
        self.f_active = self._fn("set_active", ctypes.c_int, [vp, _u8p])
        self.f_split = self._fn("split_sequence", ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32])
        self.f_pack = self._fn("pack", ctypes.c_long, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint32,
                                                      _u64p, _u32p, _i32p, _u32p])
        self.f_select = self._fn("select_words", ctypes.c_long, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                                ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint32])
        self.f_db_size = self._fn("db_size", ctypes.c_long, [vp])
        self.f_num_keys = self._fn("num_keys", ctypes.c_long, [vp])
        self.f_db_copy = self._fn("db_copy", None, [vp, _u64p, _u32p, _i32p, _u32p])
        self.f_keys_copy = self._fn("keys_copy", None, [vp, _u64p])
        self.f_db_set = self._fn("db_set", ctypes.c_int, [vp, ctypes.c_long, _u64p, _u32p, _i32p, _u32p])
        self.f_score = self._fn("score_pairs", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                             ctypes.c_int, ctypes.c_int, _f32p, _u8p])
        self.f_has_split = self._fn("has_split", ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int])
        self.n_seq = 0

    def set_sequences(self, coll, active=None):
        self.n_seq = coll.n
        a = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        assert self.f_set(self.h, coll.n, _p(coll.nibbles, _u8p), _p(coll.byte_off, _u64p), _p(coll.length, _u32p), _p(coll.weight, _f32p),
                          _p(a, _u8p)) == 0

    def set_active(self, active):
        a = np.ascontiguousarray(active, dtype=np.uint8)
        self.f_active(self.h, _p(a, _u8p))

    def split_sequence(self, seq, pos):
        self.f_split(self.h, seq, pos)

    def pack(self, seq, pack_max_degen=256, min_gc=0.0, max_gc=1.0, min_len=18):
        n = self.f_pack(self.h, seq, pack_max_degen, min_gc, max_gc, min_len, None, None, None, None)
        words = np.zeros((n, 2), np.uint64)
        index = np.zeros(n, np.uint32)
        loc = np.zeros(n, np.int32)
        strand = np.zeros(n, np.uint32)
        self.f_pack(self.h, seq, pack_max_degen, min_gc, max_gc, min_len, _p(words, _u64p), _p(index, _u32p), _p(loc, _i32p), _p(strand, _u32p))
        return canonical(words, index, loc, strand)

    def select_words(self, f, r, threshold, optimize_5=False, optimize_3=False, pack_max_degen=256, pack_min_gc=0.0, pack_max_gc=1.0,
                     min_oligo_length=18):
        f, r = _w(f), _w(r)
        n = self.f_select(self.h, len(f), _p(f, _u64p), _p(r, _u64p), int(optimize_5), int(optimize_3), threshold, pack_max_degen, pack_min_gc,
                          pack_max_gc, min_oligo_length)
        assert n >= 0
        return n, self.f_num_keys(self.h)

    def db(self):
        n = self.f_db_size(self.h)
        words = np.zeros((n, 2), np.uint64)
        index = np.zeros(n, np.uint32)
        loc = np.zeros(n, np.int32)
        strand = np.zeros(n, np.uint32)
        if n:
            self.f_db_copy(self.h, _p(words, _u64p), _p(index, _u32p), _p(loc, _i32p), _p(strand, _u32p))
        return canonical(words, index, loc, strand)

    def keys(self):
        n = self.f_num_keys(self.h)
        k = np.zeros((n, 2), np.uint64)
        if n:
            self.f_keys_copy(self.h, _p(k, _u64p))
        return k

    def db_set(self, words, index, loc, strand):
        words = _w(words)
        index = np.ascontiguousarray(index, np.uint32)
        loc = np.ascontiguousarray(loc, np.int32)
        strand = np.ascontiguousarray(strand, np.uint32)
        assert self.f_db_set(self.h, len(index), _p(words, _u64p), _p(index, _u32p), _p(loc, _i32p), _p(strand, _u32p)) == 0

    def score_pairs(self, f, r, search_threshold, detect_threshold, amp_min=80, amp_max=200, taq=False):
        f, r = _w(f), _w(r)
        cov = np.zeros(len(f), np.float32)
        bits = np.zeros((len(f), self.n_seq), np.uint8)
        assert self.f_score(self.h, len(f), _p(f, _u64p), _p(r, _u64p), search_threshold, detect_threshold, amp_min, amp_max, int(taq),
                            _p(cov, _f32p), _p(bits, _u8p)) == 0
        return cov, bits

    def has_split(self, seq, loc, length):
        return self.f_has_split(self.h, seq, loc, length)


class RefLib(_Base):
    prefix = "ref_"

    def __init__(self):
        self.lib = ctypes.CDLL(REF_PATH)
        self._words_api()
        self.lib.ref_create.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(self.lib.ref_create())
        vp = ctypes.c_void_p
        self.f_err = self._fn("last_error", ctypes.c_char_p, [vp])
        self.f_threads = self._fn("set_threads", None, [ctypes.c_int])
        self.f_max_threads = self._fn("max_threads", ctypes.c_int, [])
        self.f_set = self._fn("set_sequences", ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_char_p, _u64p, _u32p, _f32p, _u8p])
        self.f_active = self._fn("set_active", ctypes.c_int, [vp, _u8p])
        self.f_split = self._fn("split_sequence", ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32])
        self.f_pack = self._fn("pack", ctypes.c_long, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint32,
                                                      _u64p, _u32p, _i32p, _u32p])
        self.f_select = self._fn("select_words", ctypes.c_long, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                                ctypes.c_uint32, ctypes.c_float, ctypes.c_float, ctypes.c_uint32])
        self.f_db_size = self._fn("db_size", ctypes.c_long, [vp])
        self.f_num_keys = self._fn("num_keys", ctypes.c_long, [vp])
        self.f_db_copy = self._fn("db_copy", None, [vp, _u64p, _u32p, _i32p, _u32p])
        self.f_keys_copy = self._fn("keys_copy", None, [vp, _u64p])
        self.f_db_set = self._fn("db_set", ctypes.c_int, [vp, ctypes.c_long, _u64p, _u32p, _i32p, _u32p])
        self.f_score = self._fn("score_pairs", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                             ctypes.c_int, ctypes.c_int, _f32p, _u8p])
        self.f_random = self._fn("random_assays", ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                ctypes.c_int, ctypes.c_uint32, ctypes.c_float, _u64p, _u64p])
        self.f_has_split = self._fn("has_split", ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int])
        self.f_thermo = self._fn("thermo", ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_float, ctypes.c_float,
                                                          ctypes.c_float, _f32p])
        cp = ctypes.c_char_p
        self.f_thermo_batch = self._fn("thermo_batch", ctypes.c_int, [ctypes.c_int, ctypes.c_int, cp, cp, ctypes.c_float, _f32p, _f32p])
        self.f_is_valid = self._fn("is_valid", ctypes.c_int, [ctypes.c_uint32, _u64p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                               ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, _u8p])
        self.f_max_dimer = self._fn("max_dimer_tm", ctypes.c_int, [ctypes.c_uint32, _u64p, _u64p, ctypes.c_float, ctypes.c_float, ctypes.c_int, _f32p])
        self.f_multiplex = self._fn("multiplex_compatible", ctypes.c_int, [ctypes.c_uint32, _u64p, _u64p, ctypes.c_uint32, _u64p, _u64p,
                                                                             ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, _u8p])
        self.f_sw = self._fn("sw_batch", ctypes.c_int, [ctypes.c_uint32, _u64p, _u64p, _i32p])
        self.f_bg = self._fn("background_match", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                                ctypes.c_int, ctypes.c_int, _u8p, _u32p])
        self.f_mbg = self._fn("multiplex_background_match", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_float, ctypes.c_int, _u8p])
        self.f_variants = self._fn("score_variants", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, _u64p, _u64p, ctypes.c_float, ctypes.c_float,
                                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p])
        self.f_optimize = self._fn("optimize", ctypes.c_int, [vp, vp, ctypes.c_uint32, _u64p, _u64p, _i32p, ctypes.c_uint32, ctypes.c_void_p, _f32p])
        self.f_pack_all = self._fn("pack_all", ctypes.c_long, [vp, ctypes.c_uint32, ctypes.c_uint32])
        self.f_optimize_mpx = self._fn("optimize_multiplex", ctypes.c_int, [vp, vp, vp, ctypes.c_uint32, _u64p, _u64p, _i32p, ctypes.c_uint32,
                                                                            ctypes.c_void_p, ctypes.c_uint32, _u64p, _u64p, _f32p])
        self.f_mpx_cov = self._fn("multiplex_coverage", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, _u64p, _u64p, ctypes.c_float, ctypes.c_int,
                                                                       _f32p])
        self.f_overlap = self._fn("oligo_overlap", ctypes.c_int, [ctypes.c_uint32, _u64p, _u64p, ctypes.c_uint32, _u64p, _u64p, _f32p])
        self.f_uamp = self._fn("unique_amplicons", ctypes.c_int, [vp, _u64p, _u64p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u64p,
                                                                  _u64p, ctypes.c_char_p, _u32p])
        self.f_pool_amp = self._fn("pool_amplicon_coverage", ctypes.c_int, [vp, ctypes.c_uint32, _u64p, _u64p, ctypes.c_uint32, _u64p, _u64p,
                                                                            ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                                                            _f32p])
        self.f_accept = self._fn("accept_assay", ctypes.c_long, [vp, vp, _u64p, _u64p, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                                                 ctypes.c_uint32])
        self.n_seq = 0

    def set_threads(self, n):
        self.f_threads(int(n))

    def max_threads(self):
        return int(self.f_max_threads())

    def set_sequences(self, coll, active=None):
        self.n_seq = coll.n
        texts = [coll.text(i) for i in range(coll.n)]
        off = np.zeros(coll.n, np.uint64)
        ln = np.array([len(t) for t in texts], np.uint32)
        if coll.n:
            off[1:] = np.cumsum(ln.astype(np.uint64))[:-1]
        blob = "".join(texts).encode()
        a = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        rc = self.f_set(self.h, coll.n, blob, _p(off, _u64p), _p(ln, _u32p), _p(coll.weight, _f32p), _p(a, _u8p))
        assert rc == 0, self.f_err(self.h)

    def set_active(self, active):
        a = np.ascontiguousarray(active, dtype=np.uint8)
        self.f_active(self.h, _p(a, _u8p))

    def split_sequence(self, seq, pos):
        self.f_split(self.h, seq, pos)

    def pack(self, seq, pack_max_degen=256, min_gc=0.0, max_gc=1.0, min_len=18):
        n = self.f_pack(self.h, seq, pack_max_degen, min_gc, max_gc, min_len, None, None, None, None)
        assert n >= 0, self.f_err(self.h)
        words = np.zeros((n, 2), np.uint64)
        index = np.zeros(n, np.uint32)
        loc = np.zeros(n, np.int32)
        strand = np.zeros(n, np.uint32)
        self.f_pack(self.h, seq, pack_max_degen, min_gc, max_gc, min_len, _p(words, _u64p), _p(index, _u32p), _p(loc, _i32p), _p(strand, _u32p))
        return canonical(words, index, loc, strand)

    def select_words(self, f, r, threshold, optimize_5=False, optimize_3=False, pack_max_degen=256, pack_min_gc=0.0, pack_max_gc=1.0,
                     min_oligo_length=18):
        f, r = _w(f), _w(r)
        n = self.f_select(self.h, len(f), _p(f, _u64p), _p(r, _u64p), int(optimize_5), int(optimize_3), threshold, pack_max_degen, pack_min_gc,
                          pack_max_gc, min_oligo_length)
        assert n >= 0, self.f_err(self.h)
        return n, self.f_num_keys(self.h)

    def db(self):
        n = self.f_db_size(self.h)
        words = np.zeros((n, 2), np.uint64)
        index = np.zeros(n, np.uint32)
        loc = np.zeros(n, np.int32)
        strand = np.zeros(n, np.uint32)
        if n:
            self.f_db_copy(self.h, _p(words, _u64p), _p(index, _u32p), _p(loc, _i32p), _p(strand, _u32p))
        return canonical(words, index, loc, strand)

    def keys(self):
        n = self.f_num_keys(self.h)
        k = np.zeros((n, 2), np.uint64)
        if n:
            self.f_keys_copy(self.h, _p(k, _u64p))
        return k

    def db_set(self, words, index, loc, strand):
        words = _w(words)
        index = np.ascontiguousarray(index, np.uint32)
        loc = np.ascontiguousarray(loc, np.int32)
        strand = np.ascontiguousarray(strand, np.uint32)
        assert self.f_db_set(self.h, len(index), _p(words, _u64p), _p(index, _u32p), _p(loc, _i32p), _p(strand, _u32p)) == 0

    def score_pairs(self, f, r, target_threshold, search_multiplier, amp_min=80, amp_max=200, taq=False, want_cov=True, want_bits=True):
        """coverage: optimize()'s first score (search = thr*mult, detect = thr); bits: find_target_match (search = detect = thr)."""
        f, r = _w(f), _w(r)
        cov = np.zeros(len(f), np.float32)
        bits = np.zeros((len(f), self.n_seq), np.uint8)
        rc = self.f_score(self.h, len(f), _p(f, _u64p), _p(r, _u64p), target_threshold, search_multiplier, amp_min, amp_max, int(taq),
                          _p(cov, _f32p) if want_cov else None, _p(bits, _u8p) if want_bits else None)
        assert rc == 0, self.f_err(self.h)
        return cov, bits

    def random_assays(self, n_pairs, seed, primer_range=(18, 25), amp_range=(80, 200), degen=1, salt=0.05):
        f = np.zeros((n_pairs, 2), np.uint64)
        r = np.zeros((n_pairs, 2), np.uint64)
        rc = self.f_random(self.h, n_pairs, seed, primer_range[0], primer_range[1], amp_range[0], amp_range[1], degen, salt, _p(f, _u64p),
                           _p(r, _u64p))
        assert rc == 0, self.f_err(self.h)
        return f, r

    def random_assay_stream(self, n_trials, seed, opt):
        """one OpenMP thread of main.cpp:527-548: fresh NucCruc object, local seed -> (f, r, seed after); opt = RandomAssayOptions"""
        fn = self._fn("random_assay_stream", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint32, _u32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_int, ctypes.c_uint32] + [ctypes.c_float] * 6 + [_u64p, _u64p])
        f, r = np.zeros((n_trials, 2), np.uint64), np.zeros((n_trials, 2), np.uint64)
        sd = np.array([seed], np.uint32)
        rc = fn(self.h, n_trials, _p(sd, _u32p), opt.primer_min, opt.primer_max, opt.amplicon_min, opt.amplicon_max, opt.degen, opt.salt,
                opt.primer_strand, opt.primer_tm_min, opt.primer_tm_max, opt.max_hairpin, opt.max_dimer, _p(f, _u64p), _p(r, _u64p))
        if rc:
            return None
        return f, r, int(sd[0])

    def has_split(self, seq, loc, length):
        return self.f_has_split(self.h, seq, loc, length)

    def thermo(self, op, a, b="", salt=0.05, strand_a=9e-7, strand_b=9e-7):
        out = np.zeros(5, np.float32)
        rc = self.f_thermo(op, a.encode(), b.encode(), salt, strand_a, strand_b, _p(out, _f32p))
        assert rc == 0
        return out

    def thermo_batch(self, op, seq_a, seq_b=None, salt=0.05, strand_a=9e-7, strand_b=9e-7):
        """-> (n, 5) float32 {tm, dH, dS, dG, dG_dp}; one NucCruc per OpenMP thread, stale ring-buffer slots pinned to 'A'"""
        a = pack_strings(seq_a)
        b = pack_strings(seq_b) if seq_b is not None else None
        n = len(a)
        strand = np.stack([np.broadcast_to(np.asarray(strand_a, np.float32), (n,)), np.broadcast_to(np.asarray(strand_b, np.float32), (n,))], 1)
        strand = np.ascontiguousarray(strand, dtype=np.float32)
        out = np.zeros((n, 5), np.float32)
        rc = self.f_thermo_batch(int(op), n, a.ctypes.data_as(ctypes.c_char_p), None if b is None else b.ctypes.data_as(ctypes.c_char_p),
                                 float(salt), _p(strand, _f32p), _p(out, _f32p))
        assert rc == 0
        return out

    def is_valid(self, words, salt=0.05, primer_strand=9e-7, tm_range=(50.0, 75.0), max_hairpin=40.0, max_dimer=40.0, check_homo_dimer=True,
                 fast_alignment=False):
        w = _w(words)
        out = np.zeros(len(w), np.uint8)
        rc = self.f_is_valid(len(w), _p(w, _u64p), salt, primer_strand, tm_range[0], tm_range[1], max_hairpin, max_dimer, int(check_homo_dimer),
                             int(fast_alignment), _p(out, _u8p))
        assert rc == 0
        return out

    def max_dimer_tm(self, f, r, salt=0.05, primer_strand=9e-7, fast_alignment=False):
        f, r = _w(f), _w(r)
        out = np.zeros(len(f), np.float32)
        rc = self.f_max_dimer(len(f), _p(f, _u64p), _p(r, _u64p), salt, primer_strand, int(fast_alignment), _p(out, _f32p))
        assert rc == 0
        return out

    def multiplex_compatible(self, f, r, pool_f, pool_r, salt=0.05, primer_strand=9e-7, max_dimer=40.0, fast_alignment=False):
        f, r, pf, pr = _w(f), _w(r), _w(pool_f), _w(pool_r)
        out = np.zeros(len(f), np.uint8)
        rc = self.f_multiplex(len(f), _p(f, _u64p), _p(r, _u64p), len(pf), _p(pf, _u64p), _p(pr, _u64p), salt, primer_strand, max_dimer,
                              int(fast_alignment), _p(out, _u8p))
        assert rc == 0
        return out


    def score_variants(self, base_f, base_r, var_f, var_r, target_threshold, search_multiplier, amp_min=80, amp_max=200, taq=False):
        bf, br, vf, vr = _w(base_f), _w(base_r), _w(var_f), _w(var_r)
        cov = np.zeros(len(bf), np.float32)
        rc = self.f_variants(self.h, len(bf), _p(bf, _u64p), _p(br, _u64p), _p(vf, _u64p), _p(vr, _u64p), target_threshold, search_multiplier,
                             amp_min, amp_max, int(taq), _p(cov, _f32p))
        assert rc == 0, self.f_err(self.h)
        return cov

    def optimize(self, f, r, moves, options, background=None):
        """the reference's optimize() per trial (serial); options = pcramp_b200.api.OptimizeOptions; background = another RefLib"""
        f, r = _w(f).copy(), _w(r).copy()
        mv = np.ascontiguousarray(moves, dtype=np.int32)
        score = np.zeros((len(f), 3), np.float32)
        rc = self.f_optimize(self.h, background.h if background is not None else None, len(f), _p(f, _u64p), _p(r, _u64p), _p(mv, _i32p), len(mv),
                             ctypes.byref(options), _p(score, _f32p))
        assert rc == 0, self.f_err(self.h)
        return f, r, score

    def word_max_overlap(self, a, b):
        fn = self._fn("word_max_overlap", ctypes.c_float, [_u64p, _u64p])
        a, b = _w(a), _w(b)
        return np.array([fn(_p(a[i], _u64p), _p(b[i], _u64p)) for i in range(len(a))], np.float32)

    def parse_fasta(self, paths, min_len=0, max_len=1 << 40, ignore=()):
        """the reference's parse_fasta over the files, in order -> [(length, weight, nibbles uint8[length])]"""
        fn = self._fn("parse_fasta", ctypes.c_long, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint64, ctypes.c_uint64,
                                                     ctypes.c_int, ctypes.c_char_p])
        get = self._fn("sequence_get", ctypes.c_long, [ctypes.c_void_p, ctypes.c_uint32, _f32p, _u8p])
        arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
        ig = b"".join(x.encode() + b"\0" for x in ignore)
        n = fn(self.h, len(paths), arr, int(min_len), int(max_len), len(ignore), ig)
        assert n >= 0, self.f_err(self.h)
        out = []
        for i in range(n):
            w = np.zeros(1, np.float32)
            ln = get(self.h, i, _p(w, _f32p), None)
            nib = np.zeros(max(ln, 1), np.uint8)
            get(self.h, i, _p(w, _f32p), _p(nib, _u8p))
            out.append((int(ln), float(w[0]), nib[:ln].copy()))
        return out

    def append_fasta_groups(self, paths, file_group, min_len=0, max_len=1 << 40, num_pad=1, ignore=()):
        """append_fasta_group per group as main.cpp:296-341 -> [(length, weight, nibbles)] of the groups that kept something"""
        fn = self._fn("append_fasta_groups", ctypes.c_long, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), _u32p, ctypes.c_uint64,
                                                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_char_p])
        arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
        fg = np.ascontiguousarray(file_group, dtype=np.uint32)
        ig = b"".join(x.encode() + b"\0" for x in ignore)
        n = fn(self.h, len(paths), arr, _p(fg, _u32p), int(min_len), int(max_len), int(num_pad), len(ignore), ig)
        assert n >= 0, self.f_err(self.h)
        return self.sequences()

    def pack_all(self, pack_max_degen=256, min_oligo_length=18):
        """the multiplex background database of main.cpp:989-1003 (every sequence packed whole) -> its keys"""
        n = self.f_pack_all(self.h, pack_max_degen, min_oligo_length)
        assert n >= 0, self.f_err(self.h)
        return self.keys()

    def optimize_multiplex(self, f, r, moves, options, background=None, multiplex=None, pool_f=None, pool_r=None):
        """optimize() with a multiplex background context (pack_all) and an assay pool"""
        f, r = _w(f).copy(), _w(r).copy()
        mv = np.ascontiguousarray(moves, dtype=np.int32)
        score = np.zeros((len(f), 3), np.float32)
        pf = _w(pool_f) if pool_f is not None else np.zeros((0, 2), np.uint64)
        pr = _w(pool_r) if pool_r is not None else np.zeros((0, 2), np.uint64)
        rc = self.f_optimize_mpx(self.h, background.h if background is not None else None, multiplex.h if multiplex is not None else None,
                                 len(f), _p(f, _u64p), _p(r, _u64p), _p(mv, _i32p), len(mv), ctypes.byref(options), len(pf), _p(pf, _u64p),
                                 _p(pr, _u64p), _p(score, _f32p))
        assert rc == 0, self.f_err(self.h)
        return f, r, score

    def multiplex_coverage(self, base_f, base_r, var_f, var_r, threshold, taq=False):
        bf, br, vf, vr = _w(base_f), _w(base_r), _w(var_f), _w(var_r)
        cov = np.zeros(len(bf), np.float32)
        rc = self.f_mpx_cov(self.h, len(bf), _p(bf, _u64p), _p(br, _u64p), _p(vf, _u64p), _p(vr, _u64p), threshold, int(taq), _p(cov, _f32p))
        assert rc == 0, self.f_err(self.h)
        return cov

    def oligo_overlap(self, f, r, pool_f, pool_r):
        f, r, pf, pr = _w(f), _w(r), _w(pool_f), _w(pool_r)
        ov = np.zeros(len(f), np.float32)
        rc = self.f_overlap(len(f), _p(f, _u64p), _p(r, _u64p), len(pf), _p(pf, _u64p), _p(pr, _u64p), _p(ov, _f32p))
        assert rc == 0
        return ov

    def sw_batch(self, query, target):
        """SO::SeqOverlap, 8 problems per align(): -> (n, 6) int32 {score, q_start, q_stop, t_start, t_stop, last_two}"""
        q, t = _w(query), _w(target)
        out = np.zeros((len(q), 6), np.int32)
        rc = self.f_sw(len(q), _p(q, _u64p), _p(t, _u64p), _p(out, _i32p))
        assert rc == 0
        return out

    def background_match(self, f, r, background_threshold, search_multiplier, amp_min=0, amp_max=2000, taq=False):
        """PCR::find_background_match per pair -> (bits (n_pairs, n_seq) uint8, candidate amplicon counts)"""
        f, r = _w(f), _w(r)
        bits = np.zeros((len(f), self.n_seq), np.uint8)
        cnt = np.zeros(len(f), np.uint32)
        rc = self.f_bg(self.h, len(f), _p(f, _u64p), _p(r, _u64p), background_threshold, search_multiplier, amp_min, amp_max, int(taq),
                       _p(bits, _u8p), _p(cnt, _u32p))
        assert rc == 0, self.f_err(self.h)
        return bits, cnt

    def multiplex_background_match(self, f, r, background_threshold, taq=False):
        f, r = _w(f), _w(r)
        bits = np.zeros((len(f), self.n_seq), np.uint8)
        rc = self.f_mbg(self.h, len(f), _p(f, _u64p), _p(r, _u64p), background_threshold, int(taq), _p(bits, _u8p))
        assert rc == 0, self.f_err(self.h)
        return bits


    def unique_amplicons(self, f, r, threshold, amp_min=80, amp_max=200, want_bounds=True):
        """PCR::collect_unique_amplicons of ONE assay -> (amplicon strings in the returned order, bounds (n, 3) uint32), or None when the
        reference threw (last_error() tells what)"""
        f, r = _w(np.asarray(f).reshape(1, 2)).reshape(-1), _w(np.asarray(r).reshape(1, 2)).reshape(-1)
        cnt = np.zeros(3, np.uint64)
        if self.f_uamp(self.h, _p(f, _u64p), _p(r, _u64p), threshold, amp_min, amp_max, int(want_bounds), _p(cnt, _u64p), None, None, None):
            return None
        off = np.zeros(int(cnt[0]) + 1, np.uint64)
        text = ctypes.create_string_buffer(max(1, int(cnt[1])))
        bounds = np.zeros((max(1, int(cnt[2])), 3), np.uint32)
        assert self.f_uamp(self.h, _p(f, _u64p), _p(r, _u64p), threshold, amp_min, amp_max, int(want_bounds), _p(cnt, _u64p), _p(off, _u64p), text,
                           _p(bounds, _u32p)) == 0
        raw = text.raw[:int(cnt[1])]
        return [raw[int(off[i]):int(off[i + 1])].decode("ascii") for i in range(int(cnt[0]))], bounds[:int(cnt[2])]

    def last_error(self):
        return self.f_err(self.h).decode()

    def pool_amplicon_coverage(self, f, r, pool_f, pool_r, target_threshold, amp_min, amp_max, background_threshold, taq=False):
        f, r, pf, pr = _w(f), _w(r), _w(pool_f), _w(pool_r)
        cov = np.zeros(len(f), np.float32)
        rc = self.f_pool_amp(self.h, len(f), _p(f, _u64p), _p(r, _u64p), len(pf), _p(pf, _u64p), _p(pr, _u64p), target_threshold, amp_min, amp_max,
                             background_threshold, int(taq), _p(cov, _f32p))
        assert rc == 0, self.f_err(self.h)
        return cov

    def accept_assay(self, multiplex, f, r, threshold, amp_min=80, amp_max=200, pack_max_degen=256, min_oligo_length=18):
        """main.cpp:989-1017: amplicons appended to the RefLib `multiplex` (database + keys rebuilt), this context's sequences split"""
        f, r = _w(np.asarray(f).reshape(1, 2)).reshape(-1), _w(np.asarray(r).reshape(1, 2)).reshape(-1)
        n = self.f_accept(self.h, multiplex.h, _p(f, _u64p), _p(r, _u64p), threshold, amp_min, amp_max, pack_max_degen, min_oligo_length)
        assert n >= 0, self.f_err(self.h)
        multiplex.n_seq += n
        return n

    def best_assay(self, target, background, overlap, f, r, max_background_cover):
        """main.cpp:829-858 folded over the trials with the reference's Score / PCR -> (index or -1, accuracy, overlap, degeneracy)"""
        fn = self._fn("best_assay", ctypes.c_int, [ctypes.c_uint32, _f32p, _f32p, _f32p, _u64p, _u64p, ctypes.c_float, ctypes.POINTER(ctypes.c_int64),
                                                   _f32p, _f32p, ctypes.POINTER(ctypes.c_double)])
        f, r = _w(f), _w(r)
        t, b = np.ascontiguousarray(target, np.float32), np.ascontiguousarray(background, np.float32)
        o = None if overlap is None else np.ascontiguousarray(overlap, np.float32)
        bi, acc, ov, dg = ctypes.c_int64(-1), np.zeros(1, np.float32), np.zeros(1, np.float32), ctypes.c_double(0.0)
        assert fn(len(f), _p(t, _f32p), _p(b, _f32p), _p(o, _f32p), _p(f, _u64p), _p(r, _u64p), float(max_background_cover), ctypes.byref(bi),
                  _p(acc, _f32p), _p(ov, _f32p), ctypes.byref(dg)) == 0
        return int(bi.value), float(acc[0]), float(ov[0]), float(dg.value)

    def reduce_best(self, score, degeneracy):
        """the root's receive loop of reduce_best_assay (main.cpp:1455-1480) in rank order -> owning rank"""
        fn = self._fn("reduce_best", ctypes.c_int, [ctypes.c_uint32, _f32p, ctypes.POINTER(ctypes.c_double)])
        sc = np.ascontiguousarray(score, np.float32)
        dg = np.ascontiguousarray(degeneracy, np.float64)
        return int(fn(len(dg), _p(sc, _f32p), dg.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))

    def sequences(self):
        """[(length, weight, nibbles uint8[length])] of the context's sequences"""
        get = self._fn("sequence_get", ctypes.c_long, [ctypes.c_void_p, ctypes.c_uint32, _f32p, _u8p])
        out = []
        i = 0
        while True:
            w = np.zeros(1, np.float32)
            ln = get(self.h, i, _p(w, _f32p), None)
            if ln < 0:
                return out
            nib = np.zeros(max(ln, 1), np.uint8)
            get(self.h, i, _p(w, _f32p), _p(nib, _u8p))
            out.append((int(ln), float(w[0]), nib[:ln].copy()))
            i += 1


def pack_strings(strs, stride=33):
    buf = np.zeros((len(strs), stride), dtype=np.uint8)
    for i, s in enumerate(strs):
        b = str(s).encode()
        assert len(b) < stride
        buf[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    return buf


HOST_THERMO_SRC = os.path.join(ROOT, "tests", "native", "host_thermo_harness.cpp")
HOST_THERMO_LIB = os.path.join(ROOT, "build", "libhost_thermo.so")


class HostThermo:
    """The product's NucCruc device functions (pcramp_b200/csrc/nuccruc.cuh, all __host__ __device__) compiled for the
    host by tests/native/host_thermo_harness.cpp -- lets the CPU tier check the K3 arithmetic without a GPU."""

    def __init__(self):
        csrc = os.path.join(ROOT, "pcramp_b200", "csrc")
        deps = [HOST_THERMO_SRC, os.path.join(csrc, "nuccruc.cuh"), os.path.join(csrc, "word128.cuh"), os.path.join(csrc, "santalucia_tables.inc")]
        if not os.path.exists(HOST_THERMO_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOST_THERMO_LIB) for d in deps):
            os.makedirs(os.path.dirname(HOST_THERMO_LIB), exist_ok=True)
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-o", HOST_THERMO_LIB,
                            HOST_THERMO_SRC], check=True)
        self.lib = ctypes.CDLL(HOST_THERMO_LIB)
        self.lib.host_thermo_batch.restype = ctypes.c_int
        self.lib.host_thermo_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_float, _f32p, _f32p,
                                               ctypes.POINTER(ctypes.c_longlong)]

    def thermo_batch(self, op, seq_a, seq_b=None, salt=0.05, strand=9e-7):
        """strand = the effective total strand concentration (what NucCruc::strand() holds); -> (n, 4) {tm, dH, dS, dG_dp}, cells"""
        a = pack_strings(seq_a)
        b = pack_strings(seq_b) if seq_b is not None else None
        n = len(a)
        st = np.ascontiguousarray(np.broadcast_to(np.asarray(strand, np.float32), (n,)))
        out = np.zeros((n, 4), np.float32)
        cells = ctypes.c_longlong(0)
        rc = self.lib.host_thermo_batch(int(op), n, a.ctypes.data_as(ctypes.c_char_p), None if b is None else b.ctypes.data_as(ctypes.c_char_p),
                                        float(salt), _p(st, _f32p), _p(out, _f32p), ctypes.byref(cells))
        assert rc == 0
        return out, cells.value


def hetero_strand(c_a, c_b):
    """NucCruc::strand(c_a, c_b) (nuc_cruc.h:818-838) in float32"""
    c_a = np.asarray(c_a, np.float32)
    c_b = np.asarray(c_b, np.float32)
    return np.where(c_a > c_b, c_a - np.float32(0.5) * c_b, c_b - np.float32(0.5) * c_a).astype(np.float32)


HOST_SW_SRC = os.path.join(ROOT, "tests", "native", "host_sw_harness.cpp")
HOST_SW_LIB = os.path.join(ROOT, "build", "libhost_sw.so")


class HostSw:
    """pcramp_b200/csrc/sw.cuh (the K4 Smith-Waterman core, __host__ __device__) compiled for the host"""

    def __init__(self):
        csrc = os.path.join(ROOT, "pcramp_b200", "csrc")
        deps = [HOST_SW_SRC, os.path.join(csrc, "sw.cuh"), os.path.join(csrc, "word128.cuh")]
        if not os.path.exists(HOST_SW_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOST_SW_LIB) for d in deps):
            os.makedirs(os.path.dirname(HOST_SW_LIB), exist_ok=True)
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", HOST_SW_LIB, HOST_SW_SRC], check=True)
        self.lib = ctypes.CDLL(HOST_SW_LIB)
        self.lib.host_sw_batch.restype = ctypes.c_int
        self.lib.host_sw_batch.argtypes = [ctypes.c_uint32, _u64p, _u64p, ctypes.c_int, _i32p]

    def sw_batch(self, query, target, with_start=True):
        q, t = _w(query), _w(target)
        out = np.zeros((len(q), 6), np.int32)
        self.lib.host_sw_batch(len(q), _p(q, _u64p), _p(t, _u64p), int(with_start), _p(out, _i32p))
        return out
