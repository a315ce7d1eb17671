"""CPU: the C-ABI library loads, exports every symbol include/pcramp_gpu.h declares, its host-side word helpers
agree with the oracle, and it refuses to run without a GPU (no CPU fallback).  No kernels are launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pcramp_b200 import build
    from pcramp_b200.api import load_library
    build.build_cuda()
    return load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcramp_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcramp_(?:gpu|word|fasta)_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_python_mirror_covers_header():
    from pcramp_b200 import api
    assert sorted(api.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pcramp_b200 import PcrampGpu, GpuError
    with pytest.raises(GpuError):
        PcrampGpu(0)


def test_product_does_not_touch_oracle():
    """nothing under pcramp_b200/ may import, link or dlopen oracle/ (the oracle is test infrastructure)"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pcramp_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "liboracle" not in src and "pcramp_oracle" not in src and "libpcramp_ref" not in src, fn


def test_host_word_helpers_match_oracle(lib, oracle):
    rng = np.random.default_rng(5)
    sym = "ACGTMRSVWYHKDBN"
    for _ in range(200):
        n = int(rng.integers(1, 33))
        s = "".join(rng.choice(list(sym), size=n))
        for centre in (0, 1):
            out = (ctypes.c_uint64 * 2)()
            lib.pcramp_word_from_string(s.encode(), centre, out)
            w = (int(out[0]), int(out[1]))
            assert w == oracle.word_from_string(s, bool(centre))
            assert lib.pcramp_word_size(out) == oracle.word_size(w)
            assert lib.pcramp_word_start(out) == oracle.word_start(w)
            assert lib.pcramp_word_stop(out) == oracle.word_stop(w)
            c = (ctypes.c_uint64 * 2)()
            lib.pcramp_word_complement(out, c)
            assert (int(c[0]), int(c[1])) == oracle.word_complement(w)
            lib.pcramp_word_center(out, c)
            assert (int(c[0]), int(c[1])) == oracle.word_center(w)
            buf = ctypes.create_string_buffer(33)
            assert lib.pcramp_word_to_string(out, buf) == n and buf.value.decode() == s
            other = (ctypes.c_uint64 * 2)()
            t = "".join(rng.choice(list(sym), size=int(rng.integers(1, 33))))
            lib.pcramp_word_from_string(t.encode(), 1, other)
            assert lib.pcramp_word_and(out, other) == oracle.word_and(w, (int(other[0]), int(other[1])))


def test_synth_word_packing_matches_library(lib):
    from pcramp_b200 import synth
    rng = np.random.default_rng(9)
    for _ in range(50):
        n = int(rng.integers(1, 33))
        codes = synth.CODE[rng.integers(0, 4, size=n)]
        s = "".join("ACGT"[int(np.log2(c))] for c in codes)
        assert synth.word_from_codes(codes) == synth.word_from_string(s)


def test_word_max_overlap_matches_reference_golden(lib):
    """Word::max_overlap (word.h:38-92) of the host helper (the same function the pool-overlap kernel runs) vs the reference"""
    import os
    from tests import optimize_cases as oc
    from tests.harness import REF_PATH, RefLib
    a, b = oc.overlap_words()
    got = np.array([lib.pcramp_word_max_overlap(a[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                b[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) for i in range(len(a))], np.float32)
    want = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_optimize_multiplex.npz"))["maxov"]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if os.path.exists(REF_PATH):
        a2, b2 = oc.overlap_words(seed=23, n=3000)
        got = np.array([lib.pcramp_word_max_overlap(a2[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                    b2[i].ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) for i in range(len(a2))], np.float32)
        assert np.array_equal(got.view(np.uint32), RefLib().word_max_overlap(a2, b2).view(np.uint32))
