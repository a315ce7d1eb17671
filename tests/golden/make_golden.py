"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libpcramp_ref.so, built from
/root/reference by oracle/Makefile) on the seeded scenarios of tests/scenarios.py, plus a table of
known-answer values of the reference's word and NucCruc primitives.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
The fixtures travel with the repository; the GPU box never reads /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import scenarios  # noqa: E402
from tests import thermo_cases  # noqa: E402
from tests import background_cases  # noqa: E402
from tests import optimize_cases  # noqa: E402
from tests import amplicon_cases  # noqa: E402
from tests import random_assay_cases  # noqa: E402
from tests.harness import RefLib  # noqa: E402


def word_kats(ref):
    """Known answers of Word primitives on a fixed list of oligos (word.h / word.cpp)."""
    rng = np.random.default_rng(2024)
    sym = "ACGTMRSVWYHKDBN"
    oligos = ["CAGCCACTGCACCTCTTCAT", "ACATAGCCTGATACGAGT", "GGGTGTGCATCGAGCGGGCG", "A" * 32, "N", "ACGTN", "RYKMSWBDHVN",
              "CAGCCTCTGCACCTNTTCAT", "RAGCCACTGCACCTCTTCAT"]
    for n in (1, 2, 17, 18, 25, 31, 32):
        for _ in range(4):
            p = [0.22, 0.22, 0.22, 0.22] + [0.12 / 11] * 11
            oligos.append("".join(rng.choice(list(sym), size=n, p=p)))
    rows = []
    words = []
    for s in oligos:
        for centre in (0, 1):
            w = ref.word_from_string(s, centre)
            words.append(w)
    words = np.array(words, dtype=np.uint64)
    for w in words:
        w = (int(w[0]), int(w[1]))
        comp = ref.word_complement(w)
        cen = ref.word_center(w)
        sl = ref.word_shift(w, 1)
        sr = ref.word_shift(w, 0)
        rows.append([w[0], w[1], ref.word_size(w), ref.word_start(w) & 0xFFFFFFFF, ref.word_stop(w) & 0xFFFFFFFF,
                     int(min(ref.word_degeneracy(w), 2.0 ** 62)), comp[0], comp[1], cen[0], cen[1], sl[0], sl[1], sr[0], sr[1]])
    pair = []
    for i in range(0, len(words), 3):
        for j in range(1, len(words), 5):
            a, b = (int(words[i][0]), int(words[i][1])), (int(words[j][0]), int(words[j][1]))
            pair.append([a[0], a[1], b[0], b[1], ref.word_and(a, b)])
    taq = np.array([[ref.taq_mama(a, b, c, d) for c in (1, 2, 4, 8, 5) for d in (1, 2, 4, 8, 0)] for a in (1, 2, 4, 8, 15) for b in (1, 2, 4, 8, 3)],
                   dtype=np.float32)
    expand = []
    for s in ("ACGT", "RAGY", "NAN", "ACRTNGB", "SWKM"):
        w = ref.word_from_string(s, 1)
        n, ws = ref.word_expand(w)
        expand.append(np.concatenate([[n], ws.reshape(-1)]).astype(np.uint64))
    return dict(oligos=np.array(oligos), word_rows=np.array(rows, dtype=np.uint64), and_rows=np.array(pair, dtype=np.uint64), taq=taq,
                expand=np.concatenate(expand))


def thermo_kats(ref):
    """NucCruc front-ends on fixed oligos (SURVEY.md Appendix B lists a few of these)."""
    rng = np.random.default_rng(7)
    seqs = ["CAGCCACTGCACCTCTTCAT", "ACATAGCCTGATACGAGT", "GGGTGTGCATCGAGCGGGCG", "AAAAAAAAAAAAAAAAAAAA", "GCGCGCGCGCGCGCGCGCGC",
            "ACGTACGTACGTACGTACGTACGTA", "ACAATCATTTCAGGCGCGAG", "ATGAAGAGGTGCAGTGGCTG"]
    for n in (18, 19, 20, 21, 22, 23, 24, 25):
        for _ in range(3):
            seqs.append("".join(rng.choice(list("ACGT"), size=n)))
    rows = []
    for s in seqs:
        for op in (0, 1, 2):
            rows.append(ref.thermo(op, s))
    het = []
    for i in range(0, len(seqs) - 1, 2):
        for op in (3, 4):
            het.append(ref.thermo(op, seqs[i], seqs[i + 1]))
    return dict(thermo_seqs=np.array(seqs), thermo_self=np.array(rows, np.float32), thermo_het=np.array(het, np.float32))


THERMO_N = 160          # problems per (op, salt) in the committed fixture
FILTER_N = 96           # trial oligos / pairs in the is_valid, max_dimer_tm and multiplex_compatible fixtures


def thermo_filter_inputs(word_from_string):
    """the words used by the is_valid / max_dimer_tm / multiplex_compatible fixtures (shared with the tests)"""
    words = thermo_cases.primer_words(11, FILTER_N, word_from_string)
    f = thermo_cases.primer_words(12, FILTER_N, word_from_string)
    r = thermo_cases.primer_words(13, FILTER_N, word_from_string)
    # pool assays keep len(F) <= len(R): see the note above ref_multiplex_compatible in oracle/ref_driver.cpp
    pool_f = thermo_cases.primer_words(14, 3, word_from_string, lo=18, hi=20)
    pool_r = thermo_cases.primer_words(15, 3, word_from_string, lo=21, hi=25)
    # a few trials that do dimerise with the pool: reverse complements of pool oligos with a mismatch
    return words, f, r, pool_f, pool_r


def thermo_batch_kats(ref):
    """Batches of NucCruc problems per op and salt (tests/thermo_cases.py) + the three PCR-level thermodynamic filters."""
    rec = {}
    for op in thermo_cases.OPS:
        for si, salt in enumerate(thermo_cases.SALTS):
            A, B, sa, sb = thermo_cases.problems(100 + si, THERMO_N, op)
            rec["op%d_salt%d" % (op, si)] = ref.thermo_batch(op, A, B, salt, sa, sb)
    words, f, r, pool_f, pool_r = thermo_filter_inputs(ref.word_from_string)
    rec["filter_words"] = words
    rec["filter_f"], rec["filter_r"], rec["pool_f"], rec["pool_r"] = f, r, pool_f, pool_r
    for fast in (0, 1):
        for homo in (0, 1):
            rec["is_valid_fast%d_homo%d" % (fast, homo)] = ref.is_valid(words, check_homo_dimer=bool(homo), fast_alignment=bool(fast))
        rec["is_valid_wide_fast%d" % fast] = ref.is_valid(words, tm_range=(40.0, 90.0), max_hairpin=60.0, max_dimer=60.0, check_homo_dimer=True,
                                                           fast_alignment=bool(fast))
        rec["is_valid_dimer_fast%d" % fast] = ref.is_valid(words, tm_range=(30.0, 95.0), max_hairpin=95.0, max_dimer=30.0, check_homo_dimer=True,
                                                            fast_alignment=bool(fast))
        rec["max_dimer_fast%d" % fast] = ref.max_dimer_tm(f, r, fast_alignment=bool(fast))
        rec["multiplex_fast%d" % fast] = ref.multiplex_compatible(f, r, pool_f, pool_r, max_dimer=10.0, fast_alignment=bool(fast))
    return rec


SW_N = 4000


def background_kats(ref):
    """K4: raw SeqOverlap alignments, find_background_match per case, find_multiplex_background_match"""
    rec = {}
    q, t = background_cases.sw_problems(71, SW_N, ref.word_from_string)
    rec["sw_query"], rec["sw_target"], rec["sw_out"] = q, t, ref.sw_batch(q, t)
    for case in background_cases.bg_cases():
        ref.set_sequences(case.coll)
        for seq, pos in case.splits:
            ref.split_sequence(seq, pos)
        ref.select_words(case.f, case.r, case.search_threshold, min_oligo_length=background_cases.BG_MIN_LEN)
        bits, cnt = ref.background_match(case.f, case.r, float(background_cases.BG_THRESHOLD), float(background_cases.BG_MULT),
                                         background_cases.BG_AMP[0], background_cases.BG_AMP[1], case.taq)
        rec["bg_%s_bits" % case.name], rec["bg_%s_count" % case.name] = bits, cnt
        print("background %-10s candidates %6d  max/pair %4d  detected %4d  sequences %d" % (case.name, int(cnt.sum()), int(cnt.max()),
                                                                                            int(bits.sum()), case.coll.n))
    coll, f, r = background_cases.multiplex_case()
    ref.set_sequences(coll)
    for taq in (0, 1):
        rec["multiplex_bits_taq%d" % taq] = ref.multiplex_background_match(f, r, float(background_cases.BG_THRESHOLD), bool(taq))
    print("multiplex detected", int(rec["multiplex_bits_taq0"].sum()), int(rec["multiplex_bits_taq1"].sum()))
    return rec


def optimize_kats(ref_factory):
    """the local search: score_variants (one move evaluation) and optimize() end to end, per case"""
    rec = {}
    for case in optimize_cases.cases():
        ref = ref_factory()
        ref.set_sequences(case.targets)
        ref.select_words(case.f, case.r, case.target_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3)
        bg = None
        if case.background is not None:
            bg = ref_factory()
            bg.set_sequences(case.background)
            bg.select_words(case.f, case.r, case.background_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3,
                            min_oligo_length=background_cases.BG_MIN_LEN)
        f, r, score = ref.optimize(case.f, case.r, case.moves, case.options, bg)
        rec["opt_%s_f" % case.name], rec["opt_%s_r" % case.name], rec["opt_%s_score" % case.name] = f, r, score
        # one move evaluation: the optimised oligos scored against the candidate lists of the starting assays
        rec["var_%s_cov" % case.name] = ref.score_variants(case.f, case.r, f, r, float(case.options.target_threshold),
                                                          float(case.options.target_search_multiplier), case.options.target_amplicon_min,
                                                          case.options.target_amplicon_max, bool(case.options.use_taq_mama))
        changed = int(((f != case.f).any(1) | (r != case.r).any(1)).sum())
        print("optimize %-18s changed %3d of %3d  mean score %s" % (case.name, changed, len(f), score.mean(0)))
    return rec


def multiplex_optimize_kats(ref_factory):
    """optimize() with the multiplex terms (multiplex background keys + assay pool), and the three pieces on their own:
    keys of the whole-sequence pack, compute_multiplex_background_coverage of trial oligos, compute_oligo_overlap"""
    rec = {}
    wa, wb = optimize_cases.overlap_words()
    rec["maxov"] = ref_factory().word_max_overlap(wa, wb)
    print("max_overlap: %d pairs, %d exact re-uses" % (len(wa), int((rec["maxov"] == 1.0).sum())))
    for case in optimize_cases.multiplex_cases():
        ref = ref_factory()
        ref.set_sequences(case.targets)
        ref.select_words(case.f, case.r, case.target_search, optimize_5=case.optimize_5, optimize_3=case.optimize_3)
        mp = None
        if case.multiplex is not None:
            mp = ref_factory()
            mp.set_sequences(case.multiplex)
            rec["mpx_%s_keys" % case.name] = mp.pack_all()
        pf, pr = case.pool
        f, r, score = ref.optimize_multiplex(case.f, case.r, case.moves, case.options, None, mp, pf, pr)
        rec["opt_%s_f" % case.name], rec["opt_%s_r" % case.name], rec["opt_%s_score" % case.name] = f, r, score
        if mp is not None:
            for taq in (0, 1):
                rec["mpx_%s_cov_taq%d" % (case.name, taq)] = mp.multiplex_coverage(case.f, case.r, f, r, float(case.options.background_threshold),
                                                                                 bool(taq))
        rec["mpx_%s_overlap" % case.name] = ref.oligo_overlap(f, r, pf, pr)
        changed = int(((f != case.f).any(1) | (r != case.r).any(1)).sum())
        print("multiplex optimize %-18s changed %3d of %3d  mean score %s  keys %d  cov>0: %d" % (
            case.name, changed, len(f), score.mean(0), len(rec.get("mpx_%s_keys" % case.name, [])),
            int((rec.get("mpx_%s_cov_taq0" % case.name, np.zeros(1)) > 0).sum())))
    return rec


def fasta_kats(ref):
    """parse_fasta + Sequence packing of the reference on the texts of tests/fasta_cases.py"""
    import tempfile
    from tests import fasta_cases
    rec = {}
    for case in fasta_cases.cases():
        with tempfile.TemporaryDirectory() as d:
            paths = []
            for k, blob in enumerate(case.files):
                paths.append(os.path.join(d, "f%d.fa" % k))
                open(paths[-1], "wb").write(blob)
            seqs = ref.parse_fasta(paths, case.min_len, case.max_len, case.ignore)
        rec["%s_len" % case.name] = np.array([s[0] for s in seqs], np.uint32)
        rec["%s_weight" % case.name] = np.array([s[1] for s in seqs], np.float32)
        rec["%s_nibbles" % case.name] = np.concatenate([s[2] for s in seqs]) if seqs else np.zeros(0, np.uint8)
        print("fasta %-16s records %3d  bases %7d  weights %s" % (case.name, len(seqs), int(sum(s[0] for s in seqs)),
                                                                 sorted(set(s[1] for s in seqs))))
    for case in fasta_cases.group_cases():                       # append_fasta_group (parse_fasta.cpp:91-169)
        with tempfile.TemporaryDirectory() as d:
            paths = []
            for k, blob in enumerate(case.files):
                paths.append(os.path.join(d, "g%d.fa" % k))
                open(paths[-1], "wb").write(blob)
            seqs = ref.append_fasta_groups(paths, case.file_group, case.min_len, case.max_len, case.num_pad, case.ignore)
        rec["group_%s_len" % case.name] = np.array([s[0] for s in seqs], np.uint32)
        rec["group_%s_nibbles" % case.name] = np.concatenate([s[2] for s in seqs]) if seqs else np.zeros(0, np.uint8)
        print("fasta groups %-20s sequences %2d  lengths %s  EOS %d" % (case.name, len(seqs), [s[0] for s in seqs],
                                                                        int(sum((s[2] == 0).sum() for s in seqs))))
    return rec


def amplicon_kats(ref_factory):
    """multiplex bookkeeping: collect_unique_amplicons per pair, pool x amplicon coverage, two accept steps"""
    rec = {}
    for case in amplicon_cases.amp_cases():
        ref, mref = ref_factory(), ref_factory()
        ref.set_sequences(case.coll, case.active)
        for seq, pos in case.splits:
            ref.split_sequence(seq, pos)
        ref.select_words(case.f, case.r, case.search_threshold)
        thr = float(case.threshold)
        rec.update(amplicon_cases.flatten(case.name, [ref.unique_amplicons(case.f[p], case.r[p], thr, *case.amp) for p in range(len(case.f))]))
        counts = rec["%s_n_amp" % case.name]
        order = np.argsort(-counts, kind="stable")
        pool = order[:case.pool]
        rec["%s_pool" % case.name] = pool.astype(np.uint32)
        rec["%s_pool_cov" % case.name] = ref.pool_amplicon_coverage(case.f, case.r, case.f[pool], case.r[pool], thr, case.amp[0], case.amp[1],
                                                                    float(amplicon_cases.BG_THRESHOLD), case.taq)
        # two accept steps (main.cpp:989-1017): the second one sees the split targets and appends to a non-empty multiplex background
        for step, p in enumerate(order[:2]):
            if step:
                ref.select_words(case.f, case.r, case.search_threshold)
            n = ref.accept_assay(mref, case.f[p], case.r[p], thr, *case.amp)
            rec["%s_accept%d_n" % (case.name, step)] = np.array([n, len(mref.keys())], np.uint64)
            rec["%s_accept%d_keys" % (case.name, step)] = mref.keys()
            for tag, ctx in (("mpx", mref), ("tgt", ref)):
                seqs = ctx.sequences()
                rec["%s_accept%d_%s_len" % (case.name, step, tag)] = np.array([q[0] for q in seqs], np.uint32)
                rec["%s_accept%d_%s_nib" % (case.name, step, tag)] = np.concatenate([q[2] for q in seqs] + [np.zeros(0, np.uint8)])
        print("amplicons %-16s pairs %3d unique amplicons %4d bounds %5d pool coverage sum %4d accepted %s" % (
            case.name, len(case.f), int(counts.sum()), int(rec["%s_n_bounds" % case.name].sum()), int(rec["%s_pool_cov" % case.name].sum()),
            [int(rec["%s_accept%d_n" % (case.name, k)][0]) for k in (0, 1)]))
    return rec


def random_assay_kats(ref_factory):
    """candidate generation: PCR::random_assay per seed stream (one OpenMP thread of main.cpp:527-548 each)"""
    rec = {}
    for case in random_assay_cases.ra_cases():
        ref = ref_factory()
        ref.set_sequences(case.coll, case.active)
        for seq, pos in case.splits:
            ref.split_sequence(seq, pos)
        F, R, S = [], [], []
        for seed, n in zip(case.seeds, case.per):
            f, r, after = ref.random_assay_stream(int(n), int(seed), case.opt)
            F.append(f); R.append(r); S.append(after)
        rec["%s_f" % case.name], rec["%s_r" % case.name] = np.concatenate(F), np.concatenate(R)
        rec["%s_seed_after" % case.name] = np.array(S, np.uint32)
        print("random_assay %-14s streams %3d trials %4d" % (case.name, len(case.seeds), int(case.per.sum())))
    return rec


def best_assay_kats(ref):
    """a15: the best-assay update rule and the cross-rank fold, by the reference's own Score / PCR objects (ref_driver.cpp)"""
    from tests import best_assay_cases
    rec = {}
    for name, tgt, bg, ov, f, r, max_bg in best_assay_cases.trial_cases():
        rec["trial_" + name] = np.array(ref.best_assay(tgt, bg, ov, f, r, max_bg), np.float64)
    for name, score, deg, valid in best_assay_cases.rank_cases():
        rec["rank_" + name] = np.array([ref.reduce_best(score, deg)], np.int64)
    print("best_assay: %d fixtures" % len(rec))
    return rec


def main():
    if "--best-assay-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "kat_best_assay.npz"), **best_assay_kats(RefLib()))
        return
    if "--random-assay-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "kat_random_assay.npz"), **random_assay_kats(RefLib))
        return
    if "--amplicons-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "kat_amplicons.npz"), **amplicon_kats(RefLib))
        return
    if "--fasta-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "kat_fasta.npz"), **fasta_kats(RefLib()))
        return
    if "--multiplex-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "kat_optimize_multiplex.npz"), **multiplex_optimize_kats(RefLib))
        return
    ref = RefLib()
    ref.set_threads(1)
    for sc in scenarios.all_scenarios():
        rec = scenarios.run_checker(ref, sc, "ref")
        np.savez_compressed(os.path.join(HERE, "scenario_%s.npz" % sc.name), **rec)
        print("%-12s entries %5d keys %5d detected %4d" % (sc.name, len(rec["db_loc"]), len(rec["keys"]), int(rec["bits"].sum())))
    np.savez_compressed(os.path.join(HERE, "kat_words.npz"), **word_kats(ref))
    np.savez_compressed(os.path.join(HERE, "kat_thermo.npz"), **thermo_kats(ref))
    np.savez_compressed(os.path.join(HERE, "kat_thermo_batch.npz"), **thermo_batch_kats(ref))
    np.savez_compressed(os.path.join(HERE, "kat_background.npz"), **background_kats(ref))
    np.savez_compressed(os.path.join(HERE, "kat_optimize.npz"), **optimize_kats(RefLib))
    np.savez_compressed(os.path.join(HERE, "kat_optimize_multiplex.npz"), **multiplex_optimize_kats(RefLib))
    np.savez_compressed(os.path.join(HERE, "kat_fasta.npz"), **fasta_kats(RefLib()))
    np.savez_compressed(os.path.join(HERE, "kat_amplicons.npz"), **amplicon_kats(RefLib))
    np.savez_compressed(os.path.join(HERE, "kat_random_assay.npz"), **random_assay_kats(RefLib))
    np.savez_compressed(os.path.join(HERE, "kat_best_assay.npz"), **best_assay_kats(RefLib()))
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
