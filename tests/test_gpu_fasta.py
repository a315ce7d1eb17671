"""GPU (-m gpu): FASTA text -> device-resident collection (pcramp_gpu_upload_fasta, fasta.cuh) against goldens of the unmodified
reference's parse_fasta + Sequence packing (and the live reference when it travelled): record count, lengths, weights, deflines
and every nibble bit-exact; the resulting collection behaves like one uploaded through pcramp_gpu_upload_sequences."""
import os
import tempfile

import numpy as np
import pytest

from pcramp_b200 import TARGET, BACKGROUND, synth
from tests import fasta_cases
from tests.harness import REF_PATH, RefLib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_fasta.npz")


def device_nibbles(gpu, kind):
    off, ln, raw = gpu.sequences_copy(kind)
    out = []
    for o, L in zip(off, ln):
        b = raw[int(o):int(o) + (int(L) + 1) // 2]
        nib = np.empty(2 * len(b), np.uint8)
        nib[0::2], nib[1::2] = b >> 4, b & 15
        if int(L) % 2:
            assert nib[int(L)] == 0                       # the pad nibble
        out.append(nib[:int(L)])
    return ln, out


@pytest.mark.parametrize("case", fasta_cases.cases(), ids=lambda c: c.name)
def test_upload_fasta_matches_reference_golden(gpu, case):
    g = np.load(GOLD)
    recs = gpu.upload_fasta(TARGET, case.files, case.min_len, case.max_len, case.ignore)
    ln, nibs = device_nibbles(gpu, TARGET)
    assert list(ln) == list(g["%s_len" % case.name]) == [r[2] for r in recs]
    assert np.array_equal(np.array([r[3] for r in recs], np.float32).view(np.uint32), g["%s_weight" % case.name].view(np.uint32))
    got = np.concatenate(nibs) if nibs else np.zeros(0, np.uint8)
    assert np.array_equal(got, g["%s_nibbles" % case.name])
    for r in recs:
        assert b"\n" not in r[1] and b"\r" not in r[1]


def test_illegal_symbol_is_the_references_error(gpu):
    with pytest.raises(RuntimeError, match="Illegal base"):
        gpu.upload_fasta(TARGET, [b">a\nACGTACGT\n>b\nACGT*ACGT\n"])
    # ... but not in a record the length window drops (the reference never converts it)
    recs = gpu.upload_fasta(TARGET, [b">a\nACGTACGTAC\n>b\nAC*T\n"], min_length=8)
    assert [r[2] for r in recs] == [10]


def test_fasta_collection_equals_uploaded_collection(gpu):
    """same sequences through both doors -> same word database and scores (EOS, IUPAC and odd lengths included)"""
    coll = synth.make_targets(971, 40, 2500, n_clades=3, between=0.1, within=0.04)
    codes = [coll.codes(i).copy() for i in range(coll.n)]
    codes[3][700] = 0
    codes[7][100:103] = [5, 15, 0]
    codes[9] = codes[9][:2499]
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(972, coll, 60)
    thr = float(np.float32(1.0) * np.float32(0.9))
    text = "".join(">s%d\n%s\n" % (i, "\n".join(coll.text(i)[k:k + 70] for k in range(0, int(coll.length[i]), 70))) for i in range(coll.n)).encode()
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
    ne_a = gpu.select_words(TARGET, f, r, thr)
    db_a = gpu.db_copy(TARGET)
    cov_a, bits_a = gpu.score_pairs(TARGET, f, r, thr, 0.9)
    recs = gpu.upload_fasta(BACKGROUND, [text])
    assert [x[2] for x in recs] == list(coll.length)
    ne_b = gpu.select_words(BACKGROUND, f, r, thr)
    db_b = gpu.db_copy(BACKGROUND)
    cov_b, bits_b = gpu.score_pairs(BACKGROUND, f, r, thr, 0.9)
    assert ne_a == ne_b and all(np.array_equal(a, b) for a, b in zip(db_a, db_b))
    assert np.array_equal(cov_a, cov_b) and np.array_equal(bits_a, bits_b) and bits_a.any()
    # weights set after the fact (per-file normalisation) == weights given at upload
    w = np.random.default_rng(4).uniform(0.1, 2.0, size=coll.n).astype(np.float32)
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length, w)
    gpu.select_words(TARGET, f, r, thr)
    cov_w, _ = gpu.score_pairs(TARGET, f, r, thr, 0.9)
    gpu.set_weights(BACKGROUND, w)
    cov_w2, _ = gpu.score_pairs(BACKGROUND, f, r, thr, 0.9)
    assert np.array_equal(cov_w.view(np.uint32), cov_w2.view(np.uint32)) and not np.array_equal(cov_w, cov_a)


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_large_random_fasta_vs_live_reference(gpu):
    rng = np.random.default_rng(5)
    parts = []
    for i in range(300):
        n = int(rng.integers(0, 20000))
        body = "".join(rng.choice(list("ACGTacgtNRYKM-"), size=n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .005, .005, .005, .005]))
        w = int(rng.integers(20, 200))
        parts.append(">r%d [w=%g]\n%s" % (i, float(rng.integers(1, 9)) / 4, "".join(body[k:k + w] + "\n" for k in range(0, n, w))))
    blob = "".join(parts).encode()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "x.fa")
        open(path, "wb").write(blob)
        want = RefLib().parse_fasta([path], 100, 15000)
    recs = gpu.upload_fasta(TARGET, [blob], 100, 15000)
    ln, nibs = device_nibbles(gpu, TARGET)
    assert [w[0] for w in want] == list(ln)
    assert [np.float32(w[1]) for w in want] == [np.float32(x[3]) for x in recs]
    assert all(np.array_equal(a[2], b) for a, b in zip(want, nibs))


@pytest.mark.parametrize("case", fasta_cases.group_cases(), ids=lambda c: c.name)
def test_upload_fasta_groups_matches_reference_golden(gpu, case):
    """append_fasta_group (parse_fasta.cpp:91-169): the files of a group as ONE sequence, pads between the kept records"""
    g = np.load(GOLD)
    recs = gpu.upload_fasta_groups(TARGET, case.files, case.file_group, case.min_len, case.max_len, case.num_pad, case.ignore)
    ln, nibs = device_nibbles(gpu, TARGET)
    assert list(ln) == list(g["group_%s_len" % case.name]) == [r[2] for r in recs]
    got = np.concatenate(nibs) if nibs else np.zeros(0, np.uint8)
    assert np.array_equal(got, g["group_%s_nibbles" % case.name])
    # the grouped collection scans like the same sequences uploaded directly (pads and '-' are EOS the device text knows about)
    off, _, raw = gpu.sequences_copy(TARGET)
    rng = np.random.default_rng(3)
    words = []
    for _ in range(16):                                         # 20-mers cut from the grouped sequences
        i = int(rng.integers(0, len(nibs)))
        a = int(rng.integers(0, max(1, len(nibs[i]) - 20)))
        w = nibs[i][a:a + 20]
        if len(w) == 20 and (w != 0).all():
            words.append(synth.word_from_codes(w))
    if len(words) >= 2:
        f = np.array(words[:len(words) // 2], dtype=np.uint64)
        r = np.array(words[len(words) // 2:2 * (len(words) // 2)], dtype=np.uint64)
        ne_a = gpu.select_words(TARGET, f, r, 0.9)
        db_a = gpu.db_copy(TARGET)
        gpu.upload_sequences(BACKGROUND, raw, off, ln)
        ne_b = gpu.select_words(BACKGROUND, f, r, 0.9)
        db_b = gpu.db_copy(BACKGROUND)
        assert ne_a == ne_b and ne_a[0] > 0 and all(np.array_equal(x, y) for x, y in zip(db_a, db_b))


@pytest.mark.skipif(not os.path.exists(REF_PATH), reason="compiled reference did not travel with the snapshot")
def test_fasta_groups_vs_live_reference(gpu):
    rng = np.random.default_rng(15)
    files, groups = [], []
    for gidx in range(12):
        for _ in range(int(rng.integers(1, 4))):
            parts = []
            for i in range(int(rng.integers(1, 6))):
                n = int(rng.integers(0, 3000))
                body = "".join(rng.choice(list("ACGTacgtN-"), size=n, p=[.23, .23, .23, .23, .02, .02, .01, .01, .01, .01]))
                w = int(rng.integers(30, 120))
                parts.append(">g%d_%d%s\n%s" % (gidx, i, " phage" if rng.random() < 0.15 else "", "".join(body[k:k + w] + "\n" for k in range(0, n, w))))
            files.append("".join(parts).encode())
            groups.append(gidx)
    for num_pad in (1, 3):
        with tempfile.TemporaryDirectory() as d:
            paths = []
            for k, blob in enumerate(files):
                paths.append(os.path.join(d, "x%d.fa" % k))
                open(paths[-1], "wb").write(blob)
            want = RefLib().append_fasta_groups(paths, groups, 80, 2500, num_pad, ["phage"])
        gpu.upload_fasta_groups(TARGET, files, groups, 80, 2500, num_pad, ["phage"])
        ln, nibs = device_nibbles(gpu, TARGET)
        assert [w[0] for w in want] == list(ln) and len(ln) > 3
        assert all(np.array_equal(a[2], b) for a, b in zip(want, nibs))
