"""bench_legs.py -- the other BASELINE configurations as legs of bench.py's JSON line (rank 0, N = 1).

The headline of bench.py is config 5 (the fixed-pair sweep at seed threshold 0.9).  The legs here time what the other
configurations add to the hot path, each through the C ABI, each next to the unmodified reference (oracle/_ref) on a bounded
sample with a bit-for-bit comparison of that sample:

  background_scan      C2: the background screen -- select_words at the background thresholds (0.8 x 0.9: cannot be seeded, the
                       brute-force scan) + find_background_match (Smith-Waterman on every candidate amplicon)
  degenerate_primers   C3: the sweep's own collection scored with primers as `-d 16` leaves them (the degenerate positions of a segment
                       prefix are enumerated letter by letter in the index queries)
  optimize_moves       C1: optimize() with all six moves on the trials of one design iteration
  design_iteration     C1: whole iterations of pcramp's main loop (candidates, index maintenance, optimize, screens, accept,
                       splits) through pcramp_gpu_design_iteration, next to the stock program
  design_c2            C2: whole design iterations with backgrounds at config 2's size, next to the stock program on a sample of the trials
  design_c3            C3: whole design iterations with -d 16 on the headline's collection
  large_genomes        C4: 1000 x 5 Mb genomes (5 x 10^9 positions: the text index in parts of < 2^31 positions) -- index build time /
                       bytes, one batch per step on the indexed scan

Only bench.py imports this module; the reference is loaded as the checker / CPU arm only (tests.harness.RefLib)."""
import os
import subprocess
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
BG_THRESHOLD, BG_MULT = np.float32(0.8), np.float32(0.9)   # pcramp.h:40,51
BG_AMP = (0, 2000)                                          # pcramp.h:17-18
BG_MIN_LEN = int(18 * 0.9)                                  # main.cpp:592-595
EVAL_UNIT = "evaluations/s"


def _ref():
    from tests.harness import REF_PATH, RefLib
    if not os.path.exists(REF_PATH):
        return None
    r = RefLib()
    r.set_threads(0)
    return r


def _wall(fn, reps):
    best, total = 1e30, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best, total = min(best, dt), total + dt
    return total / reps, best


def widen(words, rng, max_degeneracy=16):
    """primers as `-d 16` leaves them (optimize.cpp:356-398 grows degeneracy one base at a time): up to four positions widened to
    two-letter codes, total degeneracy <= max_degeneracy"""
    from pcramp_b200 import synth
    out = words.copy()
    for w in out:
        nib = [(int(w[i // 16]) >> ((15 - i % 16) * 4)) & 15 for i in range(32)]
        pos = [i for i in range(32) if nib[i]]
        deg = 1
        for i in rng.choice(pos, size=min(int(rng.integers(1, 5)), len(pos)), replace=False):
            add = int(synth.CODE[int(rng.integers(0, 4))])
            if nib[i] | add != nib[i] and deg * 2 <= max_degeneracy:
                nib[i] |= add
                deg *= 2
        hi = lo = 0
        for i in range(32):
            if i < 16:
                hi |= nib[i] << ((15 - i) * 4)
            else:
                lo |= nib[i] << ((31 - i) * 4)
        w[0], w[1] = hi, lo
    return out


# ---------------------------------------------------------------------------------------------------------------------
def background_leg(a, device):
    """C2: 1000 trial pairs against 1000 x 5000 nt backgrounds (a sister clade 10 % away, 2 % within)"""
    from pcramp_b200 import BACKGROUND, PcrampGpu, synth
    from pcramp_b200.api import unpack_bits
    P, n_bg, L = a.pairs, 1000, 5000
    tf = synth.TargetFactory(2, 10000, L, n_clades=1, between=0.0, within=0.02)
    bf = synth.TargetFactory(2, n_bg, L, n_clades=1, between=0.10, within=0.02)
    f, r = synth.make_pairs(2, tf, P)
    bg = bf.collection()
    thr = float(BG_THRESHOLD * BG_MULT)
    g = PcrampGpu(device)
    try:
        g.upload_sequences(BACKGROUND, bg.nibbles, bg.byte_off, bg.length)
        acc = {"ms_scan": 0.0, "ms_seed": 0.0, "ms_edge": 0.0, "ms_db": 0.0, "launches": 0, "n": 0}
        last = {}

        def step():
            g.select_words(BACKGROUND, f, r, thr, min_oligo_length=BG_MIN_LEN, want_keys=False)
            st = g.stats()
            for k in ("ms_scan", "ms_seed", "ms_edge", "ms_db"):
                acc[k] += st[k]
            acc["launches"] += st["kernel_launches"]
            acc["n"] += 1
            last.update(st)
            bits, n_amp = g.background_match(BACKGROUND, f, r, thr, float(BG_THRESHOLD), BG_AMP[0], BG_AMP[1], False)
            acc["launches"] += g.stats()["kernel_launches"]
            last["n_amplicons"] = n_amp
            last["bits"] = bits

        step()
        acc.update({k: 0 for k in acc})
        mean_s, best_s = _wall(step, 3)
        int_peak = g.measure_int_peak()
        n = acc["n"]
        alignments = float(last["n_patterns"]) * float(last["n_positions"])
        scan_ms = acc["ms_scan"] / n
        out = {
            "config": "C2 background screen: %d pairs x %d backgrounds of %d nt (sister clade 10 %% away), thresholds %.1f x %.1f; step = "
                      "select_words + find_background_match through host pointers" % (P, n_bg, L, BG_THRESHOLD, BG_MULT),
            "metric": "background_pair_x_sequence_evaluations_per_s", "value": P * n_bg / mean_s, "unit": EVAL_UNIT,
            "ms_per_step": mean_s * 1e3, "ms_per_step_best": best_s * 1e3, "e2e": {"value": P * n_bg / mean_s, "unit": EVAL_UNIT,
                                                                                  "h2d_bytes_per_step": 2 * 2 * P * 16, "d2h_bytes_per_step": P * ((n_bg + 31) // 32) * 4},
            "db_entries": int(last["n_entries"]), "candidate_amplicons": int(last["n_amplicons"]), "gpu_launches": acc["launches"],
            "breakdown_ms": {k: acc[k] / n for k in ("ms_scan", "ms_seed", "ms_edge", "ms_db")},
            "roofline": {"kernel": "scan_full_kernel", "bound": "integer issue", "achieved": alignments / (scan_ms * 1e-3) if scan_ms > 0 else None,
                         "peak": int_peak, "unit": "alignments/s", "frac": (alignments / (scan_ms * 1e-3) / int_peak) if scan_ms > 0 and int_peak else None,
                         "traffic": None, "avg_launch_ms": scan_ms,
                         "peak_source": "measured live (pcramp_gpu_measure_int_peak: the brute-force scan's own mix, 4 LOP3 + POPC + compare per alignment)",
                         "note": "background thresholds leave pieces of < 5 bases: no seed filter applies, every (pattern, position) alignment is counted"},
            "cpu_baseline": None, "parity": None}
        ref = _ref()
        if ref is not None and not a.no_cpu_baseline:
            idx = list(range(0, n_bg, max(1, n_bg // 24)))[:24]
            sample = bg.subset(idx)
            ref.set_sequences(sample)
            t0 = time.perf_counter()
            ref.select_words(f, r, thr, min_oligo_length=BG_MIN_LEN)
            want, cnt = ref.background_match(f, r, float(BG_THRESHOLD), float(BG_MULT), BG_AMP[0], BG_AMP[1], False)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": P * len(idx) / dt, "unit": EVAL_UNIT, "cores": ref.max_threads(), "kind": "reference", "seconds": dt,
                                   "sample": "the step's %d pairs x %d of the %d backgrounds" % (P, len(idx), n_bg)}
            # find_background_match depends on the number of sequences (background_match.cpp:122): the comparison is sample against sample
            g.upload_sequences(BACKGROUND, sample.nibbles, sample.byte_off, sample.length)
            g.select_words(BACKGROUND, f, r, thr, min_oligo_length=BG_MIN_LEN, want_keys=False)
            bits, n_amp = g.background_match(BACKGROUND, f, r, thr, float(BG_THRESHOLD), BG_AMP[0], BG_AMP[1], False)
            got = unpack_bits(bits, sample.n)
            defined = ~(want == 255).any(1)           # pairs where the reference indexes past its list (undefined there) are left out
            out["parity"] = {"ok": bool(n_amp == int(cnt.sum()) and np.array_equal(got[defined], want[defined])), "pairs_compared": int(defined.sum()),
                             "candidate_amplicons": int(n_amp), "matched_bits_reference": int(want[defined].sum()), "checker": "reference"}
        return out
    finally:
        g.close()


# ---------------------------------------------------------------------------------------------------------------------
def degenerate_leg(a, g, factory, coll, device, parity_at_bench):
    """C3: the resident collection of the headline, primers of degeneracy up to 16"""
    from pcramp_b200 import TARGET, synth
    P = a.pairs
    f, r = synth.make_pairs(31, factory, P * 4)
    rng = np.random.default_rng(32)
    f, r = widen(f, rng), widen(r, rng)
    thr = float(np.float32(1.0) * np.float32(0.9))
    acc = {"ms_seed": 0.0, "ms_scan": 0.0, "ms_edge": 0.0, "ms_db": 0.0, "ms_score": 0.0, "launches": 0, "n": 0}
    last = {}
    g.stage_pairs(f, r)
    batch = [0]

    def step():
        g.set_batch((batch[0] % 4) * P, P)
        batch[0] += 1
        g.select_words_staged(TARGET, thr, want_keys=False, want_entries=False)
        g.score_pairs_staged(TARGET, thr, 1.0)
        g.synchronize()
        st = g.stats()
        for k in ("ms_seed", "ms_scan", "ms_edge", "ms_db", "ms_score"):
            acc[k] += st[k]
        acc["launches"] += st["kernel_launches"]
        acc["n"] += 1
        last.update(st)

    step()
    step()
    acc.update({k: 0 for k in acc})
    mean_s, best_s = _wall(step, 4)
    n = acc["n"]
    int_peak = g.measure_int_peak()
    alignments = float(last["n_patterns"]) * float(last["n_positions"])
    scan_ms = (acc["ms_seed"] + acc["ms_scan"]) / n
    out = {
        "config": "C3 degenerate primers: %d pairs of degeneracy <= 16 (-d 16) x %d x %d nt targets (the headline's resident collection), "
                  "thresholds 1.0 x 0.9; step = seed scan + pair scoring, pairs and results resident" % (P, a.targets, a.length),
        "metric": "primer_pair_x_target_evaluations_per_s", "value": P * a.targets / mean_s, "unit": EVAL_UNIT, "ms_per_step": mean_s * 1e3,
        "ms_per_step_best": best_s * 1e3, "gpu_launches": acc["launches"],
        "patterns_indexed": int(last.get("n_indexed", 0)), "patterns_seeded": int(last.get("n_seeded", 0)), "patterns": int(last["n_patterns"]),
        "breakdown_ms": {k: acc[k] / n for k in ("ms_seed", "ms_scan", "ms_edge", "ms_db", "ms_score")},
        "roofline": {"kernel": "index_query_kernel + scan_index_kernel + index_hits_kernel (+ scan_seed_kernel for patterns the index cannot take)", "bound": "integer issue",
                     "achieved": alignments / (scan_ms * 1e-3) if scan_ms > 0 else None, "peak": int_peak, "unit": "alignments/s (brute-force equivalent)",
                     "frac": (alignments / (scan_ms * 1e-3) / int_peak) if scan_ms > 0 and int_peak else None, "traffic": None, "avg_stage_ms": scan_ms,
                     "peak_source": "measured live (pcramp_gpu_measure_int_peak)",
                     "note": "a degenerate base inside a segment prefix multiplies the index queries (one per letter combination, at most 16); "
                             "patterns beyond that take the table-based seed filter.  Above 1.0 = alignments the exact filters skip"},
        "cpu_baseline": None, "parity": None}
    if not a.no_cpu_baseline:
        cpu, parity = parity_at_bench(a, g, factory, coll, f[:P], r[:P], thr, device)
        out["cpu_baseline"], out["parity"] = cpu, parity
    return out


# ---------------------------------------------------------------------------------------------------------------------
def _c1_targets():
    from pcramp_b200 import synth
    return synth.make_targets(1, 100, 10000, within=0.03)


def _empty_other_collections(g):
    from pcramp_b200 import BACKGROUND, MULTIPLEX
    none = np.zeros((0, 2), np.uint64)
    for kind in (BACKGROUND, MULTIPLEX):
        g.upload_sequences(kind, np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
    g.multiplex_keys()
    g.set_pool(none, none)


def optimize_leg(a, device):
    """C1: optimize() with the six moves (-d 16, --optimize.5, --optimize.3) on 1000 trials against 100 x 10 kb targets"""
    from pcramp_b200 import BACKGROUND, TARGET, PcrampGpu, synth
    from pcramp_b200.api import MOVES, OptimizeOptions
    tg = _c1_targets()
    T = 1000
    f, r = synth.make_pairs(42, tg, T)
    moves = [MOVES[m] for m in ("IncreaseDegeneracy", "DecreaseDegeneracy", "Trim5", "Grow5", "Trim3", "Grow3")]   # main.cpp:77-96
    o = OptimizeOptions(degen=16)
    thr = float(np.float32(o.target_threshold) * np.float32(o.target_search_multiplier))
    bthr = float(np.float32(o.background_threshold) * np.float32(o.background_search_multiplier))
    g = PcrampGpu(device)
    try:
        _empty_other_collections(g)
        g.upload_sequences(TARGET, tg.nibbles, tg.byte_off, tg.length)
        t0 = time.perf_counter()
        g.select_words(TARGET, f, r, thr, optimize_5=True, optimize_3=True)
        g.select_words(BACKGROUND, f, r, bthr)
        ms_db = (time.perf_counter() - t0) * 1e3
        res = {}

        def step():
            res["out"] = g.optimize(f, r, moves, o)

        step()
        mean_s, best_s = _wall(step, 3)
        of, orr, tc, bcov, ov, it = res["out"]
        out = {
            "config": "C1 local search: optimize() with all six moves (-d 16, --optimize.5, --optimize.3) on %d trial assays x %d x %d nt targets; "
                      "step = one pcramp_gpu_optimize call, host pointers" % (T, tg.n, int(tg.length[0])),
            "metric": "optimized_trial_assays_per_s", "value": T / mean_s, "unit": "trial assays/s", "ms_per_step": mean_s * 1e3,
            "ms_per_step_best": best_s * 1e3, "ms_word_database_first_call": ms_db, "move_rounds_max": int(it.max()), "move_rounds_mean": float(it.mean()),
            "trials_changed": int(((of != f).any(1) | (orr != r).any(1)).sum()),
            "roofline": {"kernel": "score_kernel<true> + thermo_kernel per move round", "bound": "latency", "achieved": None, "peak": None, "unit": None,
                         "frac": None, "traffic": None,
                         "note": "every round of the search is two small batched launches (variant scoring, thermodynamic filter) and a host replay of "
                                 "the reference's accept rule: the call is bound by the number of rounds, not by a kernel"},
            "cpu_baseline": None, "parity": None}
        ref = _ref()
        if ref is not None and not a.no_cpu_baseline:
            m = 48
            ref.set_sequences(tg)
            ref.select_words(f, r, thr, optimize_5=True, optimize_3=True)
            t0 = time.perf_counter()
            wf, wr, wscore = ref.optimize(f[:m], r[:m], moves, o, None)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": m / dt, "unit": "trial assays/s", "cores": 1, "kind": "reference", "seconds": dt,
                                   "sample": "the first %d trials through the reference's optimize(), one thread (the stock program runs one trial per "
                                             "OpenMP thread)" % m}
            got = np.stack([tc, bcov, ov], 1)[:m]
            out["parity"] = {"ok": bool(np.array_equal(of[:m], wf) and np.array_equal(orr[:m], wr) and
                                        np.array_equal(got.view(np.uint32), wscore.view(np.uint32))), "trials_compared": m, "checker": "reference"}
        return out
    finally:
        g.close()


# ---------------------------------------------------------------------------------------------------------------------
def _write_fasta(path, coll, prefix):
    with open(path, "w") as fh:
        for i in range(coll.n):
            s = coll.text(i)
            fh.write(">%s%d\n" % (prefix, i))
            for k in range(0, len(s), 70):
                fh.write(s[k:k + 70] + "\n")


def design_leg(a, device):
    """C1: whole design iterations (--seed 42 --count 3 --trial 1000), index maintenance included"""
    from pcramp_b200 import TARGET, PcrampGpu, synth
    from pcramp_b200.api import DesignLoop
    tg = _c1_targets()
    assays = []
    streams = max(1, min(64, os.cpu_count() or 1))
    runs = {}
    for label, n_streams in (("thread_1", 1), ("threads_%d" % streams, streams), ("threads_1000", 1000)):
        g = PcrampGpu(device)
        try:
            _empty_other_collections(g)
            g.upload_sequences(TARGET, tg.nibbles, tg.byte_off, tg.length)
            loop = DesignLoop(g, 42, num_trial=1000, n_streams=n_streams)
            its = []
            try:
                t0 = time.perf_counter()
                for _ in range(3):
                    res = loop.iteration()
                    if label == "thread_1":
                        assays.append((synth.words_to_strings(np.array([[res.f[0], res.f[1]]], np.uint64))[0],
                                       synth.words_to_strings(np.array([[res.r[0], res.r[1]]], np.uint64))[0], float(res.target_coverage)))
                    its.append({"ms_total": res.ms_total, "ms_candidates": res.ms_candidates, "ms_index_and_database": res.ms_select_target + res.ms_select_background,
                                "ms_optimize": res.ms_optimize, "ms_screen": res.ms_screen, "ms_accept_and_splits": res.ms_accept,
                                "found": int(res.found), "target_coverage": float(res.target_coverage), "targets_remaining": int(res.targets_remaining),
                                "splits": int(res.n_splits)})
                    if not res.found:
                        break
                wall = time.perf_counter() - t0
            finally:
                loop.close()
            st = g.stats()
            runs[label] = {"n_streams": n_streams, "iterations": its, "ms_per_iteration": wall * 1e3 / max(1, len(its)),
                           "ms_fastest_iteration": min(i["ms_total"] for i in its),   # without the first iterations' buffer growth
                           "index_builds": int(st.get("n_index_builds", 0)), "ms_index_build": float(st.get("ms_index_build", 0.0))}
        finally:
            g.close()
    key = "threads_1000"
    out = {
        "config": "C1 design run: 100 x 10 kb targets at 3 %, --seed 42 --count 3 --trial 1000; step = one pcramp_gpu_design_iteration "
                  "(candidates, word database incl. index maintenance after the previous assay's splits, optimize, screens, accept + splits)",
        "metric": "design_iterations_per_s", "value": 1e3 / runs[key]["ms_per_iteration"], "unit": "iterations/s",
        "ms_per_iteration": runs[key]["ms_per_iteration"], "ms_fastest_iteration": runs[key]["ms_fastest_iteration"], "runs": runs,
        "note": "value: one seed stream per trial (the static schedule of --thread 1000: candidate generation is a serial rand_r chain per seed "
                "stream, so the device wants as many streams as trials); thread_1 = the stock program at --thread 1 (one stream draws all trials: "
                "its reports are what tests/test_gpu_design_loop.py compares line by line); threads_%d = the schedule of the stock program's best "
                "run on this host (cpu_baseline)" % streams,
        "roofline": {"kernel": "random_assay_kernel (candidates) + the headline's kernels", "bound": "latency", "achieved": None, "peak": None, "unit": None,
                     "frac": None, "traffic": None, "note": "C1 is 10^6 bases: every stage is launch- / latency-bound at this size"},
        "cpu_baseline": None, "parity": None}
    # the --thread 1 run is BASELINE config 1 as the stock program ran it for tests/golden/design_c1_seed42_count3.txt: same assays?
    gold = os.path.join(ROOT, "tests", "golden", "design_c1_seed42_count3.txt")
    if os.path.exists(gold):
        want = []
        cover = []
        with open(gold) as fh:
            for line in fh:
                if line.startswith("ASSAY"):
                    want.append(tuple(line.split("\t")[1:3]))
                elif line.startswith("# Assay") and "target coverage score = " in line:
                    cover.append(float(line.split("target coverage score = ")[1].split()[0]))
        got = [(a.upper(), b.upper()) for a, b, _ in assays]      # a re-used oligo is written in lower case in the report
        out["parity"] = {"ok": bool(got == [(a.upper(), b.upper()) for a, b in want] and [c for _, _, c in assays] == cover[:len(assays)]),
                         "assays_compared": len(want), "checker": "golden report of the stock program (tests/golden/make_design_golden.py)",
                         "assays": ["%s / %s" % (a, b) for a, b in got]}
    stock = os.path.join(ROOT, "oracle", "_ref", "pcramp")
    if os.path.exists(stock) and not a.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "c1.fa")
            _write_fasta(fa, tg, "t")
            per = {}
            for threads in ("1", str(streams)):
                t0 = time.perf_counter()
                p = subprocess.run([stock, "-t", fa, "--thread", threads, "--seed", "42", "--count", "3", "-o", os.path.join(d, "o.txt"), "-v", "silent"],
                                   stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL, timeout=600)
                per[threads] = (time.perf_counter() - t0, p.returncode)
            out["cpu_baseline"] = {"value": 3.0 / per[str(streams)][0], "unit": "iterations/s", "cores": streams, "kind": "reference",
                                   "seconds": per[str(streams)][0], "seconds_thread_1": per["1"][0],
                                   "sample": "the stock program (oracle/_ref/pcramp) on the same FASTA, --count 3, wall clock of the whole run "
                                             "(reading 1 MB of FASTA included) / 3"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def design_c2_leg(a, device):
    """C2: whole design iterations with backgrounds -- 10 000 x 5 kb targets at 2 %, 1 000 x 5 kb backgrounds from a sister ancestor 10 % away"""
    from pcramp_b200 import BACKGROUND, MULTIPLEX, TARGET, PcrampGpu, synth
    from pcramp_b200.api import DesignLoop
    tg = synth.TargetFactory(2, 10000, 5000, n_clades=1, between=0.0, within=0.02).collection()
    bg = synth.TargetFactory(2, 1000, 5000, n_clades=1, between=0.10, within=0.02).collection()
    def run():
        g = PcrampGpu(device)
        its = []
        try:
            g.upload_sequences(TARGET, tg.nibbles, tg.byte_off, tg.length)
            g.upload_sequences(BACKGROUND, bg.nibbles, bg.byte_off, bg.length)
            g.upload_sequences(MULTIPLEX, np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
            g.multiplex_keys()
            g.set_pool(np.zeros((0, 2), np.uint64), np.zeros((0, 2), np.uint64))
            loop = DesignLoop(g, 42, num_trial=1000, n_streams=1000)
            try:
                t0 = time.perf_counter()
                for _ in range(4):
                    res = loop.iteration()
                    its.append({"ms_total": res.ms_total, "ms_candidates": res.ms_candidates, "ms_background_database": res.ms_select_background,
                                "ms_target_database": res.ms_select_target, "ms_optimize": res.ms_optimize, "ms_screen": res.ms_screen,
                                "ms_accept_and_splits": res.ms_accept, "found": int(res.found), "target_coverage": float(res.target_coverage),
                                "background_coverage": float(res.background_coverage), "targets_remaining": int(res.targets_remaining),
                                "splits": int(res.n_splits), "target_entries": int(res.n_target_entries), "background_entries": int(res.n_background_entries)})
                    if not res.found:
                        break
                wall = time.perf_counter() - t0
            finally:
                loop.close()
        finally:
            g.close()
        return wall, its
    # the run is made twice, each from the upload on (new context, new index, the same four iterations): the first in a cold process
    # (every device buffer is a first cudaMalloc: 100-900 ms of allocator time, box to box), the second with the process-wide block
    # cache warm -- what the second and later design runs of a process, or a run after the first few iterations, see
    wall_cold, its_cold = run()
    wall, its = run()
    ms = wall * 1e3 / max(1, len(its))
    out = {
        "config": "C2 design run: 10 000 x 5 000 nt targets at 2 %, 1 000 x 5 000 nt backgrounds (sister clade 10 % away), --seed 42 --trial 1000, "
                  "one seed stream per trial; step = one pcramp_gpu_design_iteration (candidates, background and target word databases, optimize, "
                  "screens incl. find_background_match, accept + splits)",
        "metric": "design_iterations_per_s", "value": 1e3 / ms, "unit": "iterations/s", "ms_per_iteration": ms,
        "ms_fastest_iteration": min(i["ms_total"] for i in its), "iterations": its,
        "ms_per_iteration_cold_process": wall_cold * 1e3 / max(1, len(its_cold)), "iterations_cold_process": [i["ms_total"] for i in its_cold],
        "note": "value / ms_per_iteration: the second of two identical runs in this process (each from the upload on: index build and splits "
                "inside), device blocks served by the process-wide cache; *_cold_process: the first run, every buffer a first cudaMalloc",
        "roofline": {"kernel": "scan_full_kernel (background database) + score kernels on 1.8 x 10^7 database entries", "bound": "integer issue / latency",
                     "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                     "note": "see configs.background_scan for the brute-force scan's roofline"},
        "cpu_baseline": None, "parity": None}
    stock = os.path.join(ROOT, "oracle", "_ref", "pcramp")
    if os.path.exists(stock) and not a.no_cpu_baseline:
        threads = max(1, os.cpu_count() or 1)
        trials = 50
        with tempfile.TemporaryDirectory() as d:
            ft, fb = os.path.join(d, "t.fa"), os.path.join(d, "b.fa")
            _write_fasta(ft, tg, "t")
            _write_fasta(fb, bg, "b")
            t0 = time.perf_counter()
            try:
                p = subprocess.run([stock, "-t", ft, "-b", fb, "--thread", str(threads), "--seed", "42", "--count", "1", "--trial", str(trials),
                                    "-o", os.path.join(d, "o.txt"), "-v", "silent"], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL, timeout=240)
                dt = time.perf_counter() - t0
                if p.returncode == 0:
                    scale = 1000.0 / trials
                    out["cpu_baseline"] = {"value": 1.0 / (dt * scale), "unit": "iterations/s", "cores": threads, "kind": "reference", "seconds": dt,
                                           "sample": "the stock program (oracle/_ref/pcramp) on the same FASTA files, ONE iteration of %d trials instead "
                                                     "of 1000 (%.1f s incl. reading 55 MB of FASTA), scaled x %.0f: its word-database loops are candidates x "
                                                     "words (select_words.cpp:93-117)" % (trials, dt, scale)}
            except subprocess.TimeoutExpired:
                out["cpu_baseline"] = {"value": None, "unit": "iterations/s", "cores": threads, "kind": "reference",
                                       "sample": "the stock program did not finish one iteration of %d trials in 240 s" % trials}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def design_c3_leg(a, coll, device, cpu_evals_per_s):
    """C3: whole design iterations on the headline's own collection (20 000 x 30 kb in 20 clades) with degenerate primers (-d 16)"""
    from pcramp_b200 import BACKGROUND, MULTIPLEX, TARGET, PcrampGpu
    from pcramp_b200.api import DesignLoop
    def run():
        g = PcrampGpu(device)
        its = []
        try:
            g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
            for kind in (BACKGROUND, MULTIPLEX):
                g.upload_sequences(kind, np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
            g.multiplex_keys()
            g.set_pool(np.zeros((0, 2), np.uint64), np.zeros((0, 2), np.uint64))
            loop = DesignLoop(g, 42, num_trial=1000, n_streams=1000, degen=16)
            try:
                t0 = time.perf_counter()
                for _ in range(8):
                    res = loop.iteration()
                    its.append({"ms_total": res.ms_total, "ms_candidates": res.ms_candidates, "ms_target_database": res.ms_select_target,
                                "ms_optimize": res.ms_optimize, "ms_screen": res.ms_screen, "ms_accept_and_splits": res.ms_accept, "found": int(res.found),
                                "target_coverage": float(res.target_coverage), "targets_remaining": int(res.targets_remaining), "splits": int(res.n_splits),
                                "target_entries": int(res.n_target_entries)})
                    if not res.found:
                        break
                wall = time.perf_counter() - t0
            finally:
                loop.close()
            st = g.stats()
        finally:
            g.close()
        return wall, its, st
    # ONE run of eight iterations, cold (design_c2_leg reports the second of two runs): this collection's buffers -- 9.6 GB of index
    # entries, 20 GB while it is sorted -- are larger than the blocks the process-wide cache keeps (1 GB), so a second run pays the same
    # allocations again (measured: a second run's first iteration took 1.2 s).  The first iteration carries the one-time costs (index
    # build, first-use device and page-locked allocations, kernel modules: 0.4-2.0 s from box to box), so it is reported on its own too
    wall, its, st = run()
    ms = wall * 1e3 / max(1, len(its))
    out = {
        "config": "C3 design run: %d x %d nt targets in clades (the headline's collection), -d 16, --seed 42 --trial 1000, one seed stream per trial; "
                  "step = one pcramp_gpu_design_iteration (candidates, target word database incl. the text index and its upkeep across splits, "
                  "optimize() with the degeneracy moves, screens, accept + splits)" % (coll.n, int(coll.length[0])),
        "metric": "design_iterations_per_s", "value": 1e3 / ms, "unit": "iterations/s", "ms_per_iteration": ms,
        "ms_fastest_iteration": min(i["ms_total"] for i in its), "iterations": its,
        "ms_first_iteration": its[0]["ms_total"], "ms_later_iterations": float(np.mean([i["ms_total"] for i in its[1:]])) if len(its) > 1 else None,
        "note": "value / ms_per_iteration: wall clock of the whole run over its iterations, the first included -- it builds the text index and "
                "makes the context's first-use allocations (device and page-locked buffers, kernel modules: 0.4-2.0 s from box to box); "
                "ms_later_iterations: what every further iteration of a design run costs",
        "index": {"ms_build": float(st["ms_index_build"]), "builds": int(st["n_index_builds"]), "stale_sequences_last_call": int(st["n_index_stale"])},
        "roofline": {"kernel": "thermo_kernel / score_entries_groups_kernel per move round + host replay of the accept rule", "bound": "latency",
                     "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                     "note": "optimize() with -d 16 is 12-18 lock-step rounds over ~9 x 10^5 trial oligos: see configs.optimize_moves"},
        "cpu_baseline": None, "parity": None}
    if cpu_evals_per_s:
        evals = 1000.0 * coll.n
        out["cpu_baseline"] = {"value": cpu_evals_per_s / evals, "unit": "iterations/s", "cores": os.cpu_count(), "kind": "reference",
                               "seconds": evals / cpu_evals_per_s,
                               "sample": "LOWER BOUND from this run's reference sample (cpu_baseline of the main line: %.3g (pair, target) evaluations/s "
                                         "through the reference's select_words + scoring): one iteration's 1000 trials x %d targets need %.0f s for "
                                         "the word database and first scores alone, before optimize()'s moves" % (cpu_evals_per_s, coll.n,
                                                                                                                  evals / cpu_evals_per_s)}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def large_genome_leg(a, device):
    """C4: a.c4_targets x 5 Mb genomes at 1 % (the default 1000 is BASELINE config 4's target set)"""
    from pcramp_b200 import TARGET, PcrampGpu, synth
    from pcramp_b200.api import unpack_bits
    n, L, P = a.c4_targets, a.c4_length, a.pairs
    fac = synth.TargetFactory(4, n, L, n_clades=1, between=0.0, within=0.01)
    t0 = time.perf_counter()
    coll = fac.collection()
    gen_s = time.perf_counter() - t0
    f, r = synth.make_pairs(4, coll, P * 3)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = PcrampGpu(device)
    try:
        t0 = time.perf_counter()
        g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
        upload_s = time.perf_counter() - t0
        g.stage_pairs(f, r)
        g.set_batch(0, P)
        t0 = time.perf_counter()
        g.select_words_staged(TARGET, thr, want_keys=False)     # builds the text index
        first_s = time.perf_counter() - t0
        st0 = g.stats()
        acc = {"ms_seed": 0.0, "ms_scan": 0.0, "ms_edge": 0.0, "ms_db": 0.0, "ms_score": 0.0, "ms_index_kernel": 0.0, "launches": 0, "n": 0}
        last = {}
        batch = [0]

        def step():
            g.set_batch((batch[0] % 3) * P, P)
            batch[0] += 1
            g.select_words_staged(TARGET, thr, want_keys=False, want_entries=False)
            g.score_pairs_staged(TARGET, thr, 1.0)
            g.synchronize()
            st = g.stats()
            for k in ("ms_seed", "ms_scan", "ms_edge", "ms_db", "ms_score", "ms_index_kernel"):
                acc[k] += st[k]
            acc["launches"] += st["kernel_launches"]
            acc["n"] += 1
            last.update(st)

        step()
        step()
        acc.update({k: 0 for k in acc})
        mean_s, best_s = _wall(step, 6)
        k = acc["n"]
        hbm_peak = a.hbm_peak
        stream_bytes = 16.0 * last["n_index_entries"]
        kern_ms = acc["ms_index_kernel"] / k
        out = {
            "config": "C4 targets: %d x %d nt genomes at 1 %%, %d pairs per step, thresholds 1.0 x 0.9; step = seed scan + pair scoring, pairs and "
                      "results resident" % (n, L, P),
            "metric": "primer_pair_x_target_evaluations_per_s", "value": P * n / mean_s, "unit": EVAL_UNIT, "ms_per_step": mean_s * 1e3,
            "ms_per_step_best": best_s * 1e3, "positions": int(last["n_positions"]), "positions_x_patterns_per_s": float(last["n_positions"]) * last["n_patterns"] / mean_s,
            "index": {"ms_build": float(st0["ms_index_build"]), "bytes": int(st0["index_bytes"]), "builds": int(st0["n_index_builds"]),
                      "first_call_s": first_s, "upload_s": upload_s, "patterns_indexed": int(last["n_indexed"]), "patterns": int(last["n_patterns"])},
            "gpu_launches": acc["launches"], "breakdown_ms": {x: acc[x] / k for x in ("ms_seed", "ms_scan", "ms_edge", "ms_db", "ms_score")},
            "roofline": {"kernel": "scan_index_kernel", "bound": "hbm", "achieved": stream_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None,
                         "peak": hbm_peak, "unit": "GB/s", "frac": (stream_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak) if kern_ms > 0 else None,
                         "traffic": None, "avg_launch_ms": kern_ms,
                         "note": "bytes = 16-byte index entries in the queried ranges (the kernel's stream); the SURVEY 8d figure (one pass over the "
                                 "nibbles = %.0f MB) over the same time is %.3f of peak" % (
                                     sum((int(x) + 1) // 2 for x in coll.length) / 1e6,
                                     (sum((int(x) + 1) // 2 for x in coll.length) / (kern_ms * 1e-3) / 1e9 / hbm_peak) if kern_ms > 0 else 0.0)},
            "whole_step_on_8d_bytes": {
                "algorithmic_bytes": float(sum((int(x) + 1) // 2 for x in coll.length)) + 16.0 * 2 * P + 28.0 * float(last["n_entries"]),
                "GB_per_s": (float(sum((int(x) + 1) // 2 for x in coll.length)) + 16.0 * 2 * P + 28.0 * float(last["n_entries"])) / mean_s / 1e9,
                "frac_of_hbm_peak": (float(sum((int(x) + 1) // 2 for x in coll.length)) + 16.0 * 2 * P + 28.0 * float(last["n_entries"])) / mean_s / 1e9 / hbm_peak,
                "note": "SURVEY.md 8d: nibbles of the active targets + 16 B per candidate + 28 B per database entry, over the WHOLE step (seed scan, "
                        "database, pair scoring)"},
            "host_generation_s": gen_s, "cpu_baseline": None, "parity": None}
        ref = _ref()
        if ref is not None and not a.no_cpu_baseline:
            m = 100                                               # the reference's loop is candidates x words: 100 pairs x one 5 Mb genome
            one = coll.subset([0])
            ref.set_sequences(one)
            t0 = time.perf_counter()
            ref.select_words(f[:m], r[:m], thr)
            cov, bits = ref.score_pairs(f[:m], r[:m], 1.0, 0.9)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": m * 1 / dt, "unit": EVAL_UNIT, "cores": ref.max_threads(), "kind": "reference", "seconds": dt,
                                   "sample": "%d of the step's pairs x 1 of the %d genomes" % (m, n)}
            g.select_words(TARGET, f[:m], r[:m], thr, want_keys=False)
            cov_g, bits_g = g.score_pairs(TARGET, f[:m], r[:m], thr, 1.0)
            _, bits_tm = g.score_pairs(TARGET, f[:m], r[:m], 1.0, 1.0)
            col = unpack_bits(bits_g, coll.n)[:, 0]
            out["parity"] = {"ok": bool(np.array_equal(col.astype(np.float32), cov) and np.array_equal(unpack_bits(bits_tm, coll.n)[:, :1], bits)),
                             "pairs_compared": m, "targets_compared": 1, "checker": "reference"}
        return out
    finally:
        g.close()
