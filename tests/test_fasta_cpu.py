"""CPU tier of the FASTA ingest (pcramp_b200/csrc/fasta.cuh): the HOST half -- the reference reader's record split (gzgets
chunks, '>' anywhere in a chunk, CR/LF), weights and the length window -- against goldens of the unmodified reference's
parse_fasta (tests/golden/kat_fasta.npz).  The residue -> nibble mapping is restated here in numpy for the check; the device
kernels that do it in the product are compared with the same goldens in tests/test_gpu_fasta.py."""
import ctypes
import os

import numpy as np
import pytest

from tests import fasta_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_fasta.npz")
LUT = np.full(256, 255, np.uint8)
for ch, v in {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "S": 6, "V": 7, "W": 9, "Y": 10, "H": 11, "K": 12, "D": 13, "B": 14,
              "N": 15, "I": 15, "X": 15}.items():
    LUT[ord(ch)] = LUT[ord(ch.lower())] = v
LUT[ord("-")] = 0
SPACE = np.zeros(256, bool)
SPACE[[32, 9, 10, 11, 12, 13]] = True


def host_records(lib, blob):
    u64p, u32p, f32p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_float)
    n = lib.pcramp_fasta_scan(blob, len(blob), 0, None, None, None, None, None)
    off, dl, b, e, w = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.float32)
    lib.pcramp_fasta_scan(blob, len(blob), n, off.ctypes.data_as(u64p), dl.ctypes.data_as(u32p), b.ctypes.data_as(u64p), e.ctypes.data_as(u64p),
                          w.ctypes.data_as(f32p))
    raw = np.frombuffer(blob, np.uint8)
    out = []
    for i in range(n):
        span = raw[int(b[i]):int(e[i])]
        res = span[~SPACE[span]]
        out.append((blob[int(off[i]):int(off[i]) + int(dl[i])].decode("latin1"), LUT[res], float(w[i])))
    return out


@pytest.fixture(scope="module")
def lib():
    from pcramp_b200.api import load_library
    return load_library()


@pytest.mark.parametrize("case", fasta_cases.cases(), ids=lambda c: c.name)
def test_host_split_matches_reference_golden(lib, case):
    g = np.load(GOLD)
    kept = []
    for blob in case.files:
        for defline, nib, w in host_records(lib, blob):
            if case.min_len <= len(nib) <= case.max_len and not any(s in defline.lower() for s in case.ignore):
                kept.append((nib, w))
    assert [len(k[0]) for k in kept] == list(g["%s_len" % case.name])
    assert np.array_equal(np.array([k[1] for k in kept], np.float32).view(np.uint32), g["%s_weight" % case.name].view(np.uint32))
    got = np.concatenate([k[0] for k in kept]) if kept else np.zeros(0, np.uint8)
    assert not (got == 255).any()
    assert np.array_equal(got, g["%s_nibbles" % case.name])
