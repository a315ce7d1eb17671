#!/usr/bin/env python
"""bench.py -- primer-pair x target evaluations/s of the B200 scoring path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU code (oracle/_ref) on host cores

Workload (BASELINE.json config 5, the one the headline target is quoted on): fixed candidate primer pairs in
batches of --pairs (the reference's num_trial, default 1000) against --targets x --length synthetic viral
targets (20 clades, 15 % between / 5 % within).  One step = one batch through the
whole hot path: seed scan (Sequence::pack + select_words over every active target, sort, keys) and pair scoring
(collect_candidates, identity, coverage, amplified-target bitsets).  One evaluation = one (pair, target)
decision.  With N > 1 (one rank per GPU under torchrun) the sweep is sharded by PAIRS -- every GPU holds every target and scores its own
batches, nothing is exchanged on the data path ("scaling": "weak") -- and the target-sharded arrangement (every GPU scores the same batch
on its shard of the targets, the shards' bitsets merged by peer stores over NVLink, xchg.cuh) is measured beside it (`target_sharded`).
At N = 1 the line also carries one leg per other BASELINE configuration (bench_legs.py) and the DP / Smith-Waterman / ingest legs.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "primer_pair_x_target_evaluations_per_s"
UNIT = "evaluations/s"
SEARCH_MULT = np.float32(0.9)   # DEFAULT_SEARCH_THRESHOLD_MULTIPLIER (pcramp.h:51)
TARGET_THR = np.float32(1.0)    # DEFAULT_TARGET_THRESHOLD (pcramp.h:36)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--targets", type=int, default=20000)
    ap.add_argument("--length", type=int, default=30000)
    ap.add_argument("--clades", type=int, default=20)
    ap.add_argument("--pairs", type=int, default=1000)
    ap.add_argument("--cpu-targets", type=int, default=200, help="targets in the bounded CPU sample of the b200 arm (also the parity pin)")
    ap.add_argument("--ref-targets", type=int, default=200, help="targets per step of --impl reference (every k-th target of the collection)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=1.0,
                    help="the timed region of --steps steps is repeated (same steps, same brackets) until this much device time was measured")
    ap.add_argument("--shard", default="pairs", choices=["pairs", "targets"],
                    help="N>1: pairs = every GPU holds all targets and scores its own batches of the sweep (no data-path exchange, the "
                         "headline); targets = the sequences are split across the GPUs, every GPU scores the same batch and the shards' "
                         "bitsets are exchanged (one design iteration at minimum latency; also measured, as `target_sharded`, in pairs mode)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: how the shards' bitsets are merged -- p2p = peer stores over NVLink fused into the tail of pair scoring "
                         "(xchg.cuh, the product path); nccl = all-gather + pcramp_gpu_merge_shards (the baseline it replaces)")
    ap.add_argument("--workers", type=int, default=0,
                    help="batches in flight per GPU (pairs sharding / one GPU): worker contexts (pcramp_gpu_create_worker) that share the resident "
                         "targets + text index, one host thread each; 1 = one batch at a time; 0 = 4, 5 or 6, whichever divides --steps with "
                         "the fewest idle slots in the last round")
    ap.add_argument("--fasta-targets", type=int, default=8000, help="sequences in the FASTA-ingest leg (0 = skip; rank 0 only)")
    ap.add_argument("--dp-problems", type=int, default=262144, help="NucCruc problems per step of the DP GCUPS leg (0 = skip the leg)")
    ap.add_argument("--dp-cpu-problems", type=int, default=60000, help="problems in the bounded CPU sample of the DP leg")
    ap.add_argument("--config-legs", default="background,degenerate,optimize,design,design_c2,design_c3,large",
                    help="legs for the other BASELINE configurations (bench_legs.py; rank 0, N = 1): any of background (C2), degenerate (C3), "
                         "optimize (C1 moves), design (C1 iterations), design_c2 (C2 iterations with backgrounds), design_c3 (C3 iterations with -d 16), large (C4); 'none' skips them")
    ap.add_argument("--c4-targets", type=int, default=1000, help="genomes in the large-genome leg (1000 x 5 Mb = BASELINE config 4's targets)")
    ap.add_argument("--c4-length", type=int, default=5000000, help="bases per genome in the large-genome leg")
    return ap.parse_args()


def workload_name(a):
    return ("C5 sweep: batches of %d fixed primer pairs (18-25 nt, amplicon 80-200 nt) vs %d x %d nt synthetic viral targets "
            "(%d clades, 15%%/5%% divergence); step = seed scan + pair scoring of one batch" % (a.pairs, a.targets, a.length, a.clades))


def bench_config(a):
    """the workload, identical in both arms (the driver compares the two lines' `config`)"""
    return {"workload": workload_name(a), "targets": a.targets, "target_len": a.length, "pairs_per_step": a.pairs,
            "seed_threshold": float(TARGET_THR * SEARCH_MULT), "detect_threshold": float(TARGET_THR),
            "l2": "inputs larger than L2 (%.0f MB of packed bases)" % (a.targets * a.length / 2e6)}


def make_factory(a):
    from pcramp_b200 import synth
    return synth.TargetFactory(3, a.targets, a.length, n_clades=a.clades, between=0.15, within=0.05)


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  The timed region is a few
    tens of milliseconds, so the samples come from NVML in a thread (one every ~2 ms); `nvidia-smi -lms` is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []          # nvidia-smi rows
        self.samples = []       # NVML: (sm MHz, power W, reasons bitmask)
        self.proc = None
        self.device = device
        self.nvml = None
        self.handle = None
        self.sm_max = None
        self.running = False
        self.source = None

    def _nvml_open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except (ValueError, IndexError):
                    pass
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self._nvml_sample()
            return True
        except Exception:
            self.nvml = None
            return False

    def _nvml_sample(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            pw = None
        try:
            rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            try:
                rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                rs = 0
        return sm, pw, rs

    def _nvml_loop(self):
        while self.running:
            try:
                self.samples.append(self._nvml_sample())
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self._nvml_open():
            self.source = "nvml"
            self.running = True
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.source == "nvml":
            self.running = False
            self.thread.join(timeout=2)
            n = self.nvml
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}   # nvml.h reasons
            sm = [x[0] for x in self.samples]
            pw = [x[1] for x in self.samples if x[1] is not None]
            allr = 0
            for x in self.samples:
                allr |= x[2]
            try:
                n.nvmlShutdown()
            except Exception:
                pass
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max, "power_w_max": max(pw) if pw else None,
                    "samples": len(sm), "reasons": sorted(k for k, b in bits.items() if allr & b), "source": "nvml, one sample / ~2 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


# --------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref when the compiled reference travelled, else the port)
# --------------------------------------------------------------------------------------------------
def cpu_checker():
    from tests.harness import RefLib, OracleLib, REF_PATH
    if os.path.exists(REF_PATH):
        chk = RefLib()
        chk.set_threads(0)  # all host threads (the reference's --thread default)
        return chk, "reference", chk.max_threads()
    return OracleLib(), "port", 1


def cpu_step(chk, kind, sample, f, r, want_bits=False):
    """one step on the CPU; -> (coverage at search 0.9 / detect 1.0, find_target_match bits or None)"""
    thr = float(TARGET_THR * SEARCH_MULT)
    chk.select_words(f, r, thr)
    if kind == "reference":
        cov, bits = chk.score_pairs(f, r, float(TARGET_THR), float(SEARCH_MULT), 80, 200, False, want_cov=True, want_bits=want_bits)
        return cov, (bits if want_bits else None)
    cov, _ = chk.score_pairs(f, r, thr, float(TARGET_THR), 80, 200, False)
    bits = chk.score_pairs(f, r, float(TARGET_THR), float(TARGET_THR), 80, 200, False)[1] if want_bits else None
    return cov, bits


def cpu_sample(a, factory, n_targets):
    stride = max(1, a.targets // n_targets)
    idx = [i * stride for i in range(n_targets)]
    return factory.collection(idx), idx


def run_reference(a):
    from pcramp_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    factory = make_factory(a)
    chk, kind, cores = cpu_checker()
    n_t = min(a.targets, a.ref_targets)       # a stated subsample of >= 200 targets, scaled linearly
    sample, _ = cpu_sample(a, factory, n_t)
    chk.set_sequences(sample)
    total = a.steps + a.warmup
    f, r = synth.make_pairs(5, factory, a.pairs * total)
    for s in range(a.warmup):
        cpu_step(chk, kind, sample, f[s * a.pairs:(s + 1) * a.pairs], r[s * a.pairs:(s + 1) * a.pairs])
    t0 = time.perf_counter()
    for s in range(a.warmup, total):
        cpu_step(chk, kind, sample, f[s * a.pairs:(s + 1) * a.pairs], r[s * a.pairs:(s + 1) * a.pairs])
    dt = time.perf_counter() - t0
    value = a.pairs * n_t * a.steps / dt
    sample_desc = "%d pairs x %d of the %d targets per step (every %d-th target); throughput is per (pair, target), the loop is linear in targets" % (
        a.pairs, n_t, a.targets, max(1, a.targets // n_t))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": bench_config(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# --------------------------------------------------------------------------------------------------
# DP GCUPS leg (the second half of BASELINE.json's metric): batches of gapped heterodimer problems
# (NucCruc::approximate_tm_heterodimer on primer pairs, what PCR::max_dimer_tm / multiplex_compatible run)
# --------------------------------------------------------------------------------------------------
DP_INT_OPS_PER_CELL = 45.0   # 3 states, 7 delta_g look-ups per gapped nearest-neighbour cell


def dp_problems(seed, n):
    """primer-like oligo pairs: 18-25-mers, uniform ACGT (random_assay geometry, pcr_assay.cpp:636-688)"""
    rng = np.random.default_rng(seed)
    la = rng.integers(18, 26, size=n)
    lb = rng.integers(18, 26, size=n)
    sym = np.frombuffer(b"ACGT", dtype=np.uint8)
    a = np.zeros((n, 33), np.uint8)
    b = np.zeros((n, 33), np.uint8)
    ra = sym[rng.integers(0, 4, size=(n, 32))]
    rb = sym[rng.integers(0, 4, size=(n, 32))]
    col = np.arange(32)[None, :]
    a[:, :32] = np.where(col < la[:, None], ra, 0)
    b[:, :32] = np.where(col < lb[:, None], rb, 0)
    return a, b, int((la * lb).sum())


def ctypes_long():
    import ctypes
    return ctypes.c_long


def ctypes_char_p():
    import ctypes
    return ctypes.c_char_p


def parse_fasta_args():
    import ctypes
    return [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_char_p]


def _ncu_traffic(kernel, **match):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/r02_traffic.json) when it was taken on this workload"""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(kernel)
        if tr and all(tr.get(k) == v for k, v in match.items()):
            return tr["dram_bytes_per_launch"]
    except (OSError, ValueError):
        pass
    return None


def sw_leg(a, g, ext, torch):
    """the Smith-Waterman half of the DP metric (cells = q x t per slot): SO::SeqOverlap alignments of primer-like queries
    (18-25 nt) against 32-base database words, what find_background_match runs four times per candidate amplicon.  Timed through
    pcramp_gpu_sw_batch with host pointers (H2D of the words, kernel, D2H of score + coordinates inside the region); CPU: the
    reference's SSE SeqOverlap, 8 problems per align(), one host thread."""
    from pcramp_b200 import synth
    from tests.harness import RefLib, REF_PATH
    rng = np.random.default_rng(29)
    sym = synth.CODE
    qs = np.array([synth.word_from_codes(sym[rng.integers(0, 4, size=int(rng.integers(18, 26)))]) for _ in range(4096)], dtype=np.uint64)
    ts = np.array([synth.word_from_codes(sym[rng.integers(0, 4, size=32)]) for _ in range(4096)], dtype=np.uint64)
    n = 1 << 20
    qi, ti = rng.integers(0, 4096, size=n), rng.integers(0, 4096, size=n)
    qlen = np.array([bin(int(w[0])).count("1") + bin(int(w[1])).count("1") for w in qs])[qi]   # single letters: one bit per base
    cells = float((qlen * 32).sum())
    # the caller's host buffers, page-locked (what a host that calls this per design iteration keeps): words in, one array per field out
    q_pin, t_pin = torch.from_numpy(qs[qi].view(np.int64)).pin_memory(), torch.from_numpy(ts[ti].view(np.int64)).pin_memory()
    q, t = q_pin.numpy().view(np.uint64), t_pin.numpy().view(np.uint64)
    out_pin = [torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(5)], torch.zeros((n, 2), dtype=torch.uint8).pin_memory()
    out = [c.numpy() for c in out_pin[0]], out_pin[1].numpy()
    g.sw_batch(q, t, out=out)                                     # warm-up at full size
    torch.cuda.synchronize()
    ms, ms_kernel = 1e30, 1e30
    for _ in range(5):                                            # wall clock around the whole call: copies in, kernel, copies out
        t0 = time.perf_counter()
        g.sw_batch(q, t, out=out)
        ms = min(ms, (time.perf_counter() - t0) * 1e3)
        ms_kernel = min(ms_kernel, g.sw_timing())
    scores = out[0][0].copy()
    int32_peak = g.measure_int32_peak()
    ops_per_cell = 12.0                                           # 12 INT32 operations per SW cell
    achieved = cells * ops_per_cell / (ms_kernel * 1e-3)
    res = {"metric": "sw_gcups", "value": cells / (ms_kernel * 1e-3) / 1e9, "unit": "GCUPS (sw_words_kernel, cells = q x t, with start coordinates)",
           "problems": n, "ms_kernel": ms_kernel,
           "e2e": {"value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms": ms, "h2d_bytes": 32 * n, "d2h_bytes": 22 * n,
                   "note": "wall clock of pcramp_gpu_sw_batch with page-locked host arrays: two copies in, kernel, six copies out (one per field)"},
           "roofline": {"kernel": "sw_words_kernel", "bound": "int32 issue", "achieved": achieved / 1e12, "peak": int32_peak / 1e12, "unit": "Tops/s",
                        "frac": achieved / int32_peak, "traffic": _ncu_traffic("sw_words_kernel", problems=n),
                        "peak_source": "measured live (pcramp_gpu_measure_int32_peak: VIMNMX3 / VIADDMNMX / LOP3 / IADD chains)",
                        "note": "algorithmic cost 12 INT32 operations per cell (SURVEY.md 8d) x cells / kernel time; the kernel spends ~19 "
                                "instructions per cell of the rows it computes (packed score + start words) and computes 20 / 24 / 28 / 32 rows "
                                "for the longest query of a warp"},
           "gpu_launches": 1, "cpu_baseline": None}
    if os.path.exists(REF_PATH) and not a.no_cpu_baseline:
        ref = RefLib()
        m = 1 << 16
        t0 = time.perf_counter()
        want = ref.sw_batch(q[:m], t[:m])
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": float((qlen[:m] * 32).sum()) / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "reference", "seconds": dt,
                               "sample": "the first %d problems, SSE int16 x 8 slots" % m, "scores_identical_to_gpu": bool(np.array_equal(want[:, 0], scores[:m]))}
    return res


def candidate_leg(a, g, factory):
    """candidate generation on the device (random_assay.cuh): opt.num_trial = 1000 trial assays drawn by
    PCR::random_assay, one seed stream per trial (the reference at --thread 1000), against 256 of the targets in the BACKGROUND slot;
    kernel time from the library's CUDA events, e2e = the whole call (seeds in, oligos out).  CPU: the same 1000 streams through the
    unmodified reference, one after the other on one host thread (each stream owns its NucCruc object, as an OpenMP thread would)."""
    from pcramp_b200 import BACKGROUND
    from tests.harness import RefLib, REF_PATH
    n_t, n_streams = min(a.targets, 256), 1000
    sample = factory.collection(range(n_t))
    g.upload_sequences(BACKGROUND, sample.nibbles, sample.byte_off, sample.length)
    rng = np.random.default_rng(17)
    seeds = rng.integers(0, 2**32, size=n_streams, dtype=np.uint64).astype(np.uint32)
    per = np.ones(n_streams, np.uint32)
    g.random_assays(BACKGROUND, seeds, per)                       # warm-up (tables, buffers)
    t0 = time.perf_counter()
    f, r, _, attempts = g.random_assays(BACKGROUND, seeds, per)
    dt = time.perf_counter() - t0
    ms_kernel = g.thermo_stats()["ms_kernel"]
    out = {"metric": "random_assay_trials_per_s", "value": n_streams / (ms_kernel * 1e-3), "unit": "trial assays/s (kernel, one GPU thread per seed stream)",
           "streams": n_streams, "ms_kernel": ms_kernel, "candidates_tried_per_trial": float(attempts.mean()),
           "e2e": {"value": n_streams / dt, "unit": "trial assays/s", "ms": dt * 1e3}, "gpu_launches": 1, "cpu_baseline": None}
    if os.path.exists(REF_PATH) and not a.no_cpu_baseline:
        from pcramp_b200.api import RandomAssayOptions
        ref = RefLib()
        ref.set_sequences(sample)
        opt = RandomAssayOptions()
        t0 = time.perf_counter()
        same = True
        for k in range(n_streams):
            wf, wr, _ = ref.random_assay_stream(1, int(seeds[k]), opt)
            same = same and np.array_equal(wf[0], f[k]) and np.array_equal(wr[0], r[k])
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_streams / dt, "unit": "trial assays/s", "cores": 1, "kind": "reference", "seconds": dt,
                               "sample": "the same 1000 seed streams, one host thread", "identical_to_gpu": bool(same)}
    return out


def fasta_leg(a, g, coll, hbm_peak):
    """FASTA ingest on the device (fasta.cuh): the text of the first --fasta-targets sequences (line width 70) through
    pcramp_gpu_upload_fasta into the BACKGROUND slot of the context.  Kernel figures from the library's CUDA events; e2e = the whole
    call from host text (pageable) to a scan-ready collection (H2D of the text, both kernels, bit-planes, tiles)."""
    from pcramp_b200 import BACKGROUND
    n = min(coll.n, a.fasta_targets)
    if n <= 0:
        return None
    sym = np.frombuffer(b"-ACMGRSVTWYHKDBN", np.uint8)
    parts = []
    for i in range(n):
        letters = sym[coll.codes(i)]
        L = len(letters)
        rows = (L + 69) // 70
        pad = np.full(rows * 70, 10, np.uint8)
        pad[:L] = letters
        body = np.concatenate([pad.reshape(rows, 70), np.full((rows, 1), 10, np.uint8)], axis=1).reshape(-1)
        body = body[:L + (L // 70) + (1 if L % 70 else 0)] if L % 70 else body
        parts.append(b">t%d\n" % i)
        parts.append(body.tobytes())
    blob = b"".join(parts)
    g.upload_fasta(BACKGROUND, [blob])            # warm-up (allocations)
    reps, ms_c, ms_p, t_e2e = 3, 0.0, 0.0, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        recs = g.upload_fasta(BACKGROUND, [blob])
        t_e2e += time.perf_counter() - t0
        tm = g.fasta_timing(BACKGROUND)
        ms_c += tm["ms_count"]
        ms_p += tm["ms_pack"]
    assert len(recs) == n and tm["n_bases"] == int(coll.length[:n].sum())
    ms_c, ms_p, t_e2e = ms_c / reps, ms_p / reps, t_e2e / reps
    chars, bases = float(tm["text_bytes"]), float(tm["n_bases"])
    alg = chars + 0.5 * bases                                      # one read of the text + the packed nibbles
    moved = 2.0 * chars + 0.5 * bases                              # this implementation reads the text twice (count, then pack)
    kern_s = (ms_c + ms_p) * 1e-3
    g.upload_sequences(BACKGROUND, np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
    cpu = None
    if not a.no_cpu_baseline:
        from tests.harness import RefLib, REF_PATH
        if os.path.exists(REF_PATH):       # the reference's own reader (single-threaded by construction) on a slice of the same text
            import tempfile
            m = max(1, n // 10)
            cut = blob.find(b">t%d\n" % m) if m < n else len(blob)
            with tempfile.TemporaryDirectory() as d:
                path = os.path.join(d, "slice.fa")
                with open(path, "wb") as fh:
                    fh.write(blob[:cut])
                ref = RefLib()
                ref.f_parse = ref._fn("parse_fasta", ctypes_long(), parse_fasta_args())
                t0 = time.perf_counter()
                k = ref.f_parse(ref.h, 1, (ctypes_char_p() * 1)(path.encode()), 0, 1 << 40, 0, b"")
                dt = time.perf_counter() - t0
            cpu = {"value": float(coll.length[:m].sum()) / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": "reference", "seconds": dt,
                   "sample": "parse_fasta of the first %d sequences (%d records read) from a file in the page cache" % (m, k)}
    return {"metric": "fasta_ingest", "cpu_baseline": cpu, "value": bases / kern_s / 1e9, "unit": "Gbases/s (both kernels, text resident in HBM)",
            "sequences": n, "text_bytes": int(chars), "bases": int(bases), "ms_count_kernel": ms_c, "ms_pack_kernel": ms_p,
            "e2e": {"value": bases / t_e2e / 1e9, "unit": "Gbases/s", "ms": t_e2e * 1e3, "h2d_bytes": int(chars),
                    "note": "host text (pageable memory) -> scan-ready collection: host record split, H2D, count + pack, bit-planes, tiles"},
            "roofline": {"kernel": "fasta_count_kernel + fasta_pack_kernel", "bound": "hbm", "achieved": alg / kern_s / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg / kern_s / 1e9 / hbm_peak, "algorithmic_bytes": alg, "bytes_moved_by_design": moved,
                         "achieved_on_moved_bytes": moved / kern_s / 1e9}}


def dp_leg(a, g, torch, ext, rank, world, dist):
    """-> dict for the JSON line (rank 0) or None"""
    import ctypes
    n = a.dp_problems
    sa, sb, cells = dp_problems(1000 + rank, n)
    # the caller's host buffers, page-locked: sequence text in, one float array per field out
    pins = [torch.from_numpy(sa).pin_memory(), torch.from_numpy(sb).pin_memory(), torch.full((n,), 9e-7, dtype=torch.float32).pin_memory()]
    sa, sb, strand = pins[0].numpy(), pins[1].numpy(), pins[2].numpy()
    out_pin = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(4)]
    out_np = [o.numpy() for o in out_pin]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(k):
            fn()
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    g.thermo_stage(3, sa, sb, 0.05, strand, strand)
    for _ in range(max(3, a.warmup)):
        g.thermo_run_staged()
    kernel_ms = []

    def resident():
        g.thermo_run_staged()

    ms_res = timed(resident, a.steps)
    g.synchronize()
    st = g.thermo_stats()
    assert st["dp_cells"] == cells and st["n_problems"] == n
    g.thermo_batch(3, sa, sb, 0.05, strand, strand, out=out_np)
    ms_e2e = timed(lambda: g.thermo_batch(3, sa, sb, 0.05, strand, strand, out=out_np), a.steps)
    # the same batch handed over as words (what the reference's own callers hold): 36 instead of 70 bytes per problem across the host link
    lut = np.zeros(256, np.uint8)
    lut[[65, 67, 71, 84]] = [1, 2, 4, 8]

    def to_words(text):
        nib = lut[text[:, :32]].astype(np.uint64)                        # codes, left-justified (Word::str() reads start()..stop())
        sh = (np.uint64(60) - np.uint64(4) * np.arange(16, dtype=np.uint64))
        return np.stack([(nib[:, :16] << sh).sum(axis=1, dtype=np.uint64), (nib[:, 16:] << sh).sum(axis=1, dtype=np.uint64)], 1)

    wpins = [torch.from_numpy(to_words(sa).view(np.int64)).pin_memory(), torch.from_numpy(to_words(sb).view(np.int64)).pin_memory()]
    wa, wb = wpins[0].numpy().view(np.uint64), wpins[1].numpy().view(np.uint64)
    ref_out = [o.copy() for o in out_np]
    g.thermo_words(3, wa, wb, 0.05, strand, strand, out=out_np)
    words_identical = all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(out_np, ref_out))
    ms_words = timed(lambda: g.thermo_words(3, wa, wb, 0.05, strand, strand, out=out_np), a.steps)
    g.thermo_stage(3, sa, sb, 0.05, strand, strand)
    g.thermo_run_staged()
    g.thermo_fetch()                                              # reads the kernel's own event time of the staged run
    kernel_ms = g.thermo_stats()["ms_kernel"]
    if rank != 0:
        return None
    total_cells = float(cells) * world   # every rank runs its own batch of the same shape (replicas)
    gcups = total_cells * a.steps / (ms_res * 1e-3) / 1e9
    gcups_e2e = total_cells * a.steps / (ms_e2e * 1e-3) / 1e9
    gcups_words = total_cells * a.steps / (ms_words * 1e-3) / 1e9
    int_peak = g.measure_int32_peak()   # issue-bound INT32 operations/s of the DP instruction mix, measured live on this GPU
    out = {
        "metric": "dp_gcups", "value": gcups, "unit": "GCUPS (1e9 DP cells/s, cells = q x t)", "ms_per_step": ms_res / a.steps,
        "problems_per_step_per_gpu": n, "cells_per_step_per_gpu": cells, "scaling": "weak (replicas: every rank runs its own batch)",
        "op": "approximate_tm_heterodimer, gapped (align_dimer + enumeration + evaluation), 18-25-mer pairs",
        "e2e": {"value": gcups_e2e, "unit": "GCUPS", "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": n * (33 + 33 + 4),
                "d2h_bytes_per_step": n * 16,
                "note": "pcramp_gpu_thermo_batch with page-locked host arrays: sequence text in (encoded on the device while the host takes "
                        "logf of the strand concentrations), size-ordered launch, one copy per result field out"},
        "e2e_words": {"value": gcups_words, "unit": "GCUPS", "ms_per_step": ms_words / a.steps, "h2d_bytes_per_step": n * (16 + 16 + 4),
                      "d2h_bytes_per_step": n * 16, "identical_to_text": bool(words_identical),
                      "note": "pcramp_gpu_thermo_words: the same problems as 16-byte words instead of text"},
        "gpu_launches": a.steps,
        "roofline": {"kernel": "thermo_kernel", "bound": "int32 issue", "achieved": total_cells / world * DP_INT_OPS_PER_CELL / (kernel_ms * 1e-3) / 1e12,
                     "peak": int_peak / 1e12, "unit": "Tops/s (INT32)", "frac": (total_cells / world * DP_INT_OPS_PER_CELL / (kernel_ms * 1e-3)) / int_peak,
                     "avg_launch_ms": kernel_ms,
                     "peak_source": "measured live (pcramp_gpu_measure_int32_peak: VIMNMX3 / VIADDMNMX / LOP3 / IADD chains)",
                     "traffic": _ncu_traffic("thermo_kernel", problems=n),
                     "note": "algorithmic cost 45 INT32 ops per gapped cell (SURVEY.md 8d) against the measured issue peak of the DP instruction "
                             "mix; the kernel also runs the traceback / enumeration / float evaluation epilogue, which this model does not credit"},
    }
    if world == 1 and not a.no_cpu_baseline:
        from tests.harness import RefLib, REF_PATH
        if os.path.exists(REF_PATH):
            ref = RefLib()
            ref.set_threads(0)
            m = min(n, a.dp_cpu_problems)
            A = [bytes(r[:int(np.argmax(r == 0))]).decode() for r in sa[:m]]
            B = [bytes(r[:int(np.argmax(r == 0))]).decode() for r in sb[:m]]
            ref.thermo_batch(3, A[:2000], B[:2000], 0.05, np.float32(9e-7), np.float32(9e-7))
            c1 = float(sum(len(x) * len(y) for x, y in zip(A, B)))
            reps, c = 0, 0.0
            t0 = time.perf_counter()
            while reps < 64 and (reps == 0 or time.perf_counter() - t0 < 8.0):
                ref.thermo_batch(3, A, B, 0.05, np.float32(9e-7), np.float32(9e-7))
                reps += 1
                c += c1
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": c / dt / 1e9, "unit": "GCUPS", "cores": ref.max_threads(), "kind": "reference", "seconds": dt,
                                   "sample": "the first %d problems of the step x %d passes, one NucCruc per OpenMP thread" % (m, reps)}
    return out


# --------------------------------------------------------------------------------------------------
# this repository's CUDA path
# --------------------------------------------------------------------------------------------------
class DevArray:
    """zero-copy torch view of a device pointer owned by the C library"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}


def target_sharded_leg(a, factory, rank, world, local, dist, torch, thr):
    """N>1, second arrangement: the targets are split across the GPUs (contiguous shards on bitset-word boundaries), every GPU scores
    the SAME batch on its shard and the shards' bitsets are merged over peer memory (xchg.cuh) -- what one design iteration of pcramp
    needs (its 1000 trials against all targets at minimum latency).  value = pairs x ALL targets / max-over-ranks time."""
    from pcramp_b200 import PcrampGpu, TARGET, synth
    from pcramp_b200.sharding import shard_bounds
    P = a.pairs
    bounds = shard_bounds(a.targets, world, align=32)
    sizes = np.array([bounds[k + 1] - bounds[k] for k in range(world)], dtype=np.uint32)
    coll = factory.collection(range(bounds[rank], bounds[rank + 1]))
    steps, warm = a.steps, 2
    f, r = synth.make_pairs(5, factory, P * (steps + warm))
    g = PcrampGpu(local)
    try:
        g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
        g.exchange_create(rank, world, sizes, P, None)
        handles = [None] * world
        dist.all_gather_object(handles, g.exchange_ipc_handle())
        g.exchange_connect_ipc(handles)
        g.stage_pairs(f, r)
        ext = torch.cuda.ExternalStream(g.stream, device=local)

        def step(b):
            g.set_batch(b * P, P)
            g.select_words_staged(TARGET, thr, want_keys=False, want_entries=False)
            g.score_pairs_staged(TARGET, thr, float(TARGET_THR))
            g.exchange_step(TARGET)

        for b in range(warm):
            step(b)
        g.exchange_fetch(P)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for b in range(warm, warm + steps):
            step(b)
        e1.record(ext)
        cov, bits = g.exchange_fetch(P)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        chk = torch.tensor([float(cov.sum()), float(np.unpackbits(bits.view(np.uint8)).sum())], device="cuda", dtype=torch.float64)
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        same = bool((lo_ == hi_).all().item())      # every rank must hold the same merged result
        dist.barrier()
        del ext
    finally:
        g.close()
    return {"value": float(P) * a.targets * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "scaling": "strong",
            "shards": [int(x) for x in sizes], "exchange": "peer stores over NVLink + flag wait (xchg.cuh)", "ranks_agree": same,
            "detected_bits_last_step": int(chk[1].item())}


def parity_at_bench(a, g, factory, coll, fb, rb, thr, device):
    """The bounded CPU sample of the bench's own workload doubles as the parity pin at bench scale: the reference scores batch 0
    (the bench's own 1000 pairs) against every k-th target of the bench's own collection, and its coverage (optimize()'s first
    score: search 0.9, detect 1.0) and find_target_match bitsets (search = detect = 1.0) must equal, bit for bit,
      (i)  the same sub-collection through a second context of the CUDA library, and
      (ii) those targets' columns of the FULL collection's result for the same pairs (per-sequence independence of the path,
           select_words.cpp:131-138, pcr_assay.cpp:348-360) -- which is the resident index / tier table / id sort at full size.
    -> (cpu_baseline dict, parity dict)"""
    from pcramp_b200 import PcrampGpu, TARGET
    from pcramp_b200.api import unpack_bits
    chk, kind, cores = cpu_checker()
    sample, idx = cpu_sample(a, factory, a.cpu_targets)
    chk.set_sequences(sample)
    t0 = time.perf_counter()
    cov_ref, bits_ref = cpu_step(chk, kind, sample, fb, rb, want_bits=True)
    dt = time.perf_counter() - t0
    P = len(fb)
    cpu_baseline = {"value": P * len(idx) / dt, "unit": UNIT, "cores": cores, "kind": kind, "seconds": dt,
                    "sample": "one step of %d pairs x %d of the %d targets (every %d-th target), coverage + find_target_match bitsets" % (
                        P, len(idx), a.targets, max(1, a.targets // len(idx)))}
    one = float(TARGET_THR)
    # (ii) the full collection, through the resident context the timed region used
    g.select_words(TARGET, fb, rb, thr, want_keys=False)
    cov_full, bits_full = g.score_pairs(TARGET, fb, rb, thr, one)
    _, bits_full_tm = g.score_pairs(TARGET, fb, rb, one, one)
    col = np.asarray(idx, dtype=np.int64)
    full_cols = unpack_bits(bits_full, coll.n)[:, col]
    full_cols_tm = unpack_bits(bits_full_tm, coll.n)[:, col]
    # (i) the sub-collection on its own context
    g2 = PcrampGpu(device)
    try:
        g2.upload_sequences(TARGET, sample.nibbles, sample.byte_off, sample.length)
        g2.select_words(TARGET, fb, rb, thr, want_keys=False)
        cov_sub, bits_sub = g2.score_pairs(TARGET, fb, rb, thr, one)
        _, bits_sub_tm = g2.score_pairs(TARGET, fb, rb, one, one)
    finally:
        g2.close()
    sub = unpack_bits(bits_sub, sample.n)
    sub_tm = unpack_bits(bits_sub_tm, sample.n)
    checks = {
        "coverage_subsample_vs_reference": bool(np.array_equal(cov_sub, cov_ref)),
        "bitsets_subsample_vs_reference": bool(np.array_equal(sub_tm, bits_ref)),
        "coverage_full_columns_vs_reference": bool(np.array_equal(full_cols.sum(axis=1).astype(np.float32), cov_ref)),
        "bitsets_full_columns_vs_reference": bool(np.array_equal(full_cols_tm, bits_ref)),
        "full_columns_vs_subsample_search": bool(np.array_equal(full_cols, sub)),
        "full_coverage_is_popcount": bool(np.array_equal(cov_full, unpack_bits(bits_full, coll.n).sum(axis=1).astype(np.float32))),
    }
    parity = {"ok": all(checks.values()), "checks": checks, "pairs": P, "targets_compared": len(idx), "of_targets": coll.n,
              "detected_bits_reference": int(bits_ref.sum()), "detected_bits_full_collection": int(unpack_bits(bits_full_tm, coll.n).sum()),
              "checker": kind}
    return cpu_baseline, parity


def run_b200(a):
    import torch
    import torch.distributed as dist
    from pcramp_b200 import PcrampGpu, TARGET, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (a.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    factory = make_factory(a)
    from pcramp_b200.sharding import shard_bounds
    by_targets = world > 1 and a.shard == "targets"
    p2p = by_targets and a.exchange == "p2p"
    if by_targets:
        bounds = shard_bounds(a.targets, world, align=32 if p2p else 1)   # contiguous shards; p2p: on bitset words
        shard_nseq = np.array([bounds[k + 1] - bounds[k] for k in range(world)], dtype=np.uint32)
        lo, hi = bounds[rank], bounds[rank + 1]
    else:                                                                 # one GPU, or the pair-sharded sweep: every rank holds every target
        shard_nseq = np.array([a.targets] * world, dtype=np.uint32)
        lo, hi = 0, a.targets
    coll = factory.collection(range(lo, hi))
    total = a.steps + a.warmup
    n_batches = total + a.steps + 1                                # resident (warm-up + timed) then e2e (1 warm-up + timed)
    # target-sharded: every rank scores the SAME batches; pair-sharded: every rank has its own slice of the sweep
    f_all, r_all = synth.make_pairs(5 if by_targets else 5 + 1000 * rank, factory, a.pairs * n_batches)
    f_pin = torch.from_numpy(f_all.view(np.int64)).pin_memory()
    r_pin = torch.from_numpy(r_all.view(np.int64)).pin_memory()
    f_host, r_host = f_pin.numpy().view(np.uint64), r_pin.numpy().view(np.uint64)

    g = PcrampGpu(local)
    g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
    ext = torch.cuda.ExternalStream(g.stream, device=local)
    thr = float(TARGET_THR * SEARCH_MULT)
    if a.workers > 0:
        W = a.workers
    elif a.steps < 4:
        W = max(1, a.steps)
    else:                                                             # rounds x W >= steps: the fewest idle slots, then the fewest contexts
        W = min((4, 5, 6), key=lambda w: (-(-a.steps // w) * w - a.steps, w))
    if by_targets:
        W = 1
    g.stage_pairs(f_all[:a.pairs], r_all[:a.pairs])
    t_first = time.perf_counter()
    g.select_words_staged(TARGET, thr, want_keys=False)   # builds the text index the workers share
    first_call_s = time.perf_counter() - t_first
    st_index = g.stats()
    index_info = {"ms_build": float(st_index["ms_index_build"]), "bytes": int(st_index["index_bytes"]), "builds": int(st_index["n_index_builds"]),
                  "first_select_words_s": first_call_s,
                  "note": "one-time per upload, outside the timed region (the sweep's targets never change); a design run pays it once and "
                          "keeps the index across splits (configs.design_iteration times whole iterations, index maintenance included)"}
    ctxs = [g] + [g.worker() for _ in range(W - 1)]
    import threading
    lock = threading.Lock()
    n_words_local = (int(shard_nseq[rank]) + 31) // 32
    n_words_global = (a.targets + 31) // 32
    P = a.pairs
    if p2p:
        g.exchange_create(rank, world, shard_nseq, P, None)
        handles = [None] * world
        dist.all_gather_object(handles, g.exchange_ipc_handle())
        g.exchange_connect_ipc(handles)
        dist.barrier()
    elif by_targets:
        max_words = int(max((int(n) + 31) // 32 for n in shard_nseq))
        gat_any = torch.zeros((world, P * max_words), dtype=torch.int32, device="cuda")
        gat_p1 = torch.zeros((world, P * max_words), dtype=torch.int32, device="cuda")
        packed_any = torch.zeros(int(sum(P * ((int(n) + 31) // 32) for n in shard_nseq)), dtype=torch.int32, device="cuda")
        packed_p1 = torch.zeros_like(packed_any)
        out_bits = torch.zeros((P, n_words_global), dtype=torch.int32, device="cuda")
        out_cov = torch.zeros(P, dtype=torch.float32, device="cuda")
    host_cov = [torch.zeros(P, dtype=torch.float32).pin_memory() for _ in ctxs]
    host_bits = [torch.zeros((P, n_words_global), dtype=torch.int32).pin_memory() for _ in ctxs]
    launches = [0]
    stats_acc = {"ms_seed": 0.0, "ms_scan": 0.0, "ms_edge": 0.0, "ms_db": 0.0, "ms_score": 0.0, "ms_index_kernel": 0.0, "n_entries": 0, "n_hits": 0,
                 "scan_launches": 0,
                 "n_index_entries": 0, "n_index_queries": 0, "n_indexed": 0, "n_seeded": 0}
    last = {}

    def exchange():
        """NCCL: all-gather the shard bitsets, then splice + re-sum on every rank"""
        d_cov, d_any, d_p1 = g.device_pointers()
        my_any = torch.as_tensor(DevArray(d_any, (P * n_words_local,), "<i4"), device="cuda")
        my_p1 = torch.as_tensor(DevArray(d_p1, (P * n_words_local,), "<i4"), device="cuda")
        gat_any[rank, :P * n_words_local].copy_(my_any)
        gat_p1[rank, :P * n_words_local].copy_(my_p1)
        dist.all_gather_into_tensor(gat_any.view(-1), gat_any[rank].clone())
        dist.all_gather_into_tensor(gat_p1.view(-1), gat_p1[rank].clone())
        o = 0
        for s in range(world):   # drop the padding so shard s holds exactly P x words_s words
            n = P * ((int(shard_nseq[s]) + 31) // 32)
            packed_any[o:o + n].copy_(gat_any[s, :n])
            packed_p1[o:o + n].copy_(gat_p1[s, :n])
            o += n
        torch.cuda.current_stream().synchronize()
        g.merge_shards(packed_any.data_ptr(), packed_p1.data_ptr(), shard_nseq, P, out_bits.data_ptr(), out_cov.data_ptr())
        launches[0] += 1

    def account(c, timed):
        """timed: True = count launches; "stats" = also accumulate the library's stage / kernel event times (sequential pass only:
        under several batches in flight a kernel's own time is not its time alone)"""
        st = c.stats()
        with lock:
            last.update(st)
            if timed:
                launches[0] += st["kernel_launches"]
            if timed == "stats":
                for k in ("ms_seed", "ms_scan", "ms_edge", "ms_db", "ms_score", "ms_index_kernel"):
                    stats_acc[k] += st[k]
                for k in ("n_entries", "n_hits", "n_index_entries", "n_index_queries", "n_indexed", "n_seeded"):
                    stats_acc[k] += st[k]
                stats_acc["scan_launches"] += 1

    def step_resident(b, timed, k=0):
        c = ctxs[k]
        c.set_batch(b * P, P)
        c.select_words_staged(TARGET, thr, want_keys=False, want_entries=False)   # sizes / keys() only for hosts that walk the database
        c.score_pairs_staged(TARGET, thr, float(TARGET_THR))
        account(c, timed)
        if p2p:
            g.exchange_step(TARGET)
            launches[0] += 3 if timed else 0
        elif by_targets:
            exchange()

    def step_e2e(b, timed, k=0):
        fb, rb = f_host[b * P:(b + 1) * P], r_host[b * P:(b + 1) * P]
        host_cov_k, host_bits_k = host_cov[k], host_bits[k]
        if not by_targets:
            c = ctxs[k]
            c.select_words(TARGET, fb, rb, thr, want_keys=False, want_entries=False)
            c.score_pairs(TARGET, fb, rb, thr, float(TARGET_THR), out=(host_cov_k.numpy(), host_bits_k.numpy().view(np.uint32)))
            if timed:
                with lock:
                    launches[0] += c.stats()["kernel_launches"]
        elif p2p:
            g.stage_pairs(fb, rb)
            g.select_words_staged(TARGET, thr, want_keys=False, want_entries=False)
            g.score_pairs_staged(TARGET, thr, float(TARGET_THR))
            g.exchange_step(TARGET)
            cov, bits = g.exchange_fetch(P)
            host_cov_k.numpy()[:] = cov
            host_bits_k.numpy().view(np.uint32)[:] = bits
        else:
            g.stage_pairs(fb, rb)
            g.select_words_staged(TARGET, thr, want_keys=False)
            g.score_pairs_staged(TARGET, thr, float(TARGET_THR))
            exchange()
            host_cov_k.copy_(out_cov, non_blocking=True)
            host_bits_k.copy_(out_bits, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(fn, first_batch, n_steps, timed, workers):
        """steps first_batch .. first_batch + n_steps - 1; with several workers step s runs on context s % workers, one host thread each"""
        if workers == 1:
            for s in range(n_steps):
                fn(first_batch + s, timed)
            return
        errors = []

        def work(k):
            try:
                for s in range(k, n_steps, workers):
                    fn(first_batch + s, timed, k)
            except Exception as e:                                 # noqa: BLE001
                errors.append(e)
        th = [threading.Thread(target=work, args=(k,)) for k in range(workers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errors:
            raise errors[0]

    def timed_region(fn, first_batch, n_steps, timed=True, workers=1):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)                      # the device is idle here (barrier): the event marks the start for every stream
        run_steps(fn, first_batch, n_steps, timed, workers)
        if workers > 1:
            torch.cuda.synchronize()        # every context's stream has drained before the closing event
        e1.record(ext)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.cuda.stream(ext):
        for k, c in enumerate(ctxs):
            c.stage_pairs(f_all, r_all)                  # every batch resident in HBM before the timed region
            for s in range(a.warmup):                    # every context warms up (its scratch buffers grow to their working size)
                step_resident(s, False, k)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        def repeated(fn, first_batch):
            """the K-step region, repeated until --min-seconds of device time have been measured (K steps at ~1.4 ms are tens of
            milliseconds): every repeat is the same K steps bracketed by barrier + synchronize; -> (mean ms per K steps, repeats)"""
            ms, reps = timed_region(fn, first_batch, a.steps, True, W), 1
            while ms < a.min_seconds * 1e3 and reps < 4096:
                ms += timed_region(fn, first_batch, a.steps, False, W)
                reps += 1
            return ms / reps, reps

        ms_resident, reps_resident = repeated(step_resident, a.warmup)
        launches_resident = launches[0]
        for k in range(W):
            step_e2e(total, False, k)
        launches[0] = 0
        ms_e2e, reps_e2e = repeated(step_e2e, total + 1)
        clocks = sampler.stop() if rank == 0 else None
        # one batch at a time on one context: the library's per-kernel event times (roofline, stage breakdown) are taken here,
        # where a kernel's duration is its own
        g.stage_pairs(f_all, r_all)                      # (the e2e steps staged their own batches)
        ms_sequential = timed_region(step_resident, a.warmup, a.steps, "stats", 1)
        # In the fast form the partial-word scan runs on a second stream beside the indexed scan, so scan_index_kernel's event time in the
        # pass above is its time while it shares the SMs.  The roofline wants the kernel's own duration: a second pass in the general form
        # (one stream, every kernel alone on the GPU; same kernels, same launch parameters) gives it.
        fast_acc = dict(stats_acc)
        for k in stats_acc:
            stats_acc[k] = 0.0 if isinstance(stats_acc[k], float) else 0
        g.set_option("use_fast_path", 0)
        timed_region(step_resident, a.warmup, a.steps, "stats", 1)
        g.set_option("use_fast_path", 1)
        alone_acc = dict(stats_acc)
        stats_acc.update(fast_acc)
        for _ in range(2):                                   # back in the fast form for the legs that follow
            step_resident(a.warmup, False, 0)
        st_edge = g.stats()                                  # (settles the batch; the partial-word table is per context)
        edge_info = {"used": bool(st_edge["edge_table_used"]), "words": int(st_edge["n_edge_words"]), "bytes": int(st_edge["edge_table_bytes"]),
                     "ms_build": float(st_edge["ms_edge_table_build"]),
                     "note": "the FILL / EOS-event / TAIL words of the collection, both strands, indexed by their runs of five slots (edge.cuh): "
                             "one-time per upload / split and context, outside the timed region"}
        int_peak = g.measure_int_peak() if rank == 0 else 0.0
        dp = dp_leg(a, g, torch, ext, rank, world, dist if world > 1 else None) if a.dp_problems > 0 else None

    evals_per_step = float(P) * a.targets * (1 if by_targets else world)   # pair-sharded: every rank completes its own batch per step
    tsh = None
    if world > 1 and not by_targets:
        tsh = target_sharded_leg(a, factory, rank, world, local, dist, torch, thr)
    value = evals_per_step * a.steps / (ms_resident * 1e-3)      # ms_*: mean over the repeats of the --steps-step region
    e2e_value = evals_per_step * a.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        n_scan = max(1, stats_acc["scan_launches"])
        scan_ms = (stats_acc["ms_seed"] + stats_acc["ms_scan"]) / n_scan
        n_cand = last["n_patterns"] // 2
        # nibbles of the active sequences + 16 B per candidate + 28 B per emitted entry
        alg_bytes = float(sum((int(L) + 1) // 2 for L in coll.length)) + 16.0 * n_cand + 28.0 * stats_acc["n_entries"] / n_scan
        align_per_launch = float(last["n_patterns"]) * float(last["n_positions"])
        align_rate = align_per_launch / (scan_ms * 1e-3) if scan_ms > 0 else 0.0
        indexed = stats_acc["n_indexed"] > 0
        seeded = stats_acc["ms_seed"] >= stats_acc["ms_scan"]
        kernel = ("scan_index_kernel" if indexed else "scan_seed_kernel") if seeded else "scan_full_kernel"
        # the dominant kernel's own launch time: scan_index_kernel is timed alone by the library (CUDA events on its stream);
        # the table / brute-force scans are the whole stage
        kernel_ms_beside = (stats_acc["ms_index_kernel"] / n_scan) if indexed else scan_ms
        kernel_ms = (alone_acc["ms_index_kernel"] / max(1, alone_acc["scan_launches"])) if indexed else scan_ms
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        stream_bytes = 16.0 * stats_acc["n_index_entries"] / n_scan   # 16-byte index entries in the queried ranges
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(kernel)
            if tr and tr.get("targets") == a.targets and tr.get("target_len") == a.length and tr.get("pairs_per_step") == a.pairs and world == 1:
                traffic = tr["dram_bytes_per_launch"]
        except (OSError, ValueError):
            pass
        roofline = {
            "kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic,
            "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": kernel_ms,
            "avg_launch_ms_beside_partial_word_scan": kernel_ms_beside, "share_of_step": kernel_ms_beside * n_scan / ms_sequential,
            "note": "avg_launch_ms: CUDA events around the kernel, one batch at a time in the general form (every kernel alone on the GPU); in the "
                    "fast form the partial-word scan runs beside it on a second stream (avg_launch_ms_beside_partial_word_scan, share_of_step).  "
                    "algorithmic bytes = SURVEY.md 8d (nibbles of the active targets + 16 B/candidate + 28 B/entry): what ONE pass over the text "
                    "would move.  The seeded scan does not stream the text: %d patterns are resolved through a text index (index.cuh) whose "
                    "16-byte entries are the kernel's real HBM stream, see `index_stream` and `traffic` (ncu dram bytes, profiles/)" % last["n_patterns"],
            "index_stream": {
                "unit": "GB/s", "bytes_per_launch": stream_bytes, "achieved": stream_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0,
                "frac_of_hbm_peak": (stream_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak) if kernel_ms > 0 else 0.0,
                "queries_per_launch": stats_acc["n_index_queries"] / n_scan, "entries_per_launch": stats_acc["n_index_entries"] / n_scan,
                "patterns_indexed": stats_acc["n_indexed"] / n_scan, "patterns_seeded": stats_acc["n_seeded"] / n_scan,
                "seeded_scan_stage_ms": scan_ms},
            "brute_force_equivalent": {
                "unit": "alignments/s", "achieved": align_rate, "issue_peak": int_peak,
                "ratio": (align_rate / int_peak) if int_peak else None,
                "alignments_per_launch": align_per_launch,
                "note": "alignments the reference's select_words loop would count (patterns x positions) per second of the seeded-scan stage, "
                        "against the measured issue-bound ceiling of the brute-force instruction mix (4 LOP3 + POPC + ISETP, "
                        "pcramp_gpu_measure_int_peak).  The brute-force kernel sits at ~1.0 of it; the exact filters exceed 1.0 because they "
                        "skip alignments."},
        }
        cpu_baseline, parity = None, None
        if world == 1 and not a.no_cpu_baseline:
            cpu_baseline, parity = parity_at_bench(a, g, factory, coll, f_all[:P], r_all[:P], thr, local)
        jobs = 1 if by_targets else world                  # pair-sharded: every rank moves its own batch
        h2d = 2 * P * 16 * jobs
        d2h = (P * 4 + P * n_words_global * 4) * jobs
        import bench_legs
        want = set() if (a.config_legs == "none" or world > 1) else set(a.config_legs.split(","))
        a.hbm_peak = hbm_peak
        cfg = {}
        if "degenerate" in want:
            cfg["degenerate_primers"] = bench_legs.degenerate_leg(a, g, factory, coll, local, parity_at_bench)
        if "background" in want:
            cfg["background_scan"] = bench_legs.background_leg(a, local)
        if "optimize" in want:
            cfg["optimize_moves"] = bench_legs.optimize_leg(a, local)
        if "design" in want:
            cfg["design_iteration"] = bench_legs.design_leg(a, local)
        if "design_c2" in want:
            cfg["design_iteration_c2"] = bench_legs.design_c2_leg(a, local)
        if "design_c3" in want:
            cfg["design_iteration_c3"] = bench_legs.design_c3_leg(a, coll, local, None if cpu_baseline is None else cpu_baseline["value"])
        if "large" in want:
            cfg["large_genomes"] = bench_legs.large_genome_leg(a, local)
        legs = {
            "dp_gcups": dp, "target_sharded": tsh, "configs": cfg,
            "sw_gcups": sw_leg(a, g, ext, torch) if a.dp_problems > 0 else None,
            "candidate_generation": candidate_leg(a, g, factory) if a.fasta_targets > 0 else None,
            "fasta_ingest": fasta_leg(a, g, coll, hbm_peak) if a.fasta_targets > 0 else None}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_resident / a.steps, "higher_is_better": True, "scaling": "strong" if (by_targets or world == 1) else "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": bench_config(a),
            "layout": {"sharding": ("targets, contiguous, %d shard(s); every GPU scores the same batch" % world) if (by_targets or world == 1) else
                       ("pairs: %d GPUs x all %d targets, each GPU scores its own batch of %d pairs per step" % (world, a.targets, P)),
                       "exchange": ("none" if not by_targets else "peer stores over NVLink from the scoring stream + flag wait (xchg.cuh), no NCCL "
                                    "on the data path" if p2p else "NCCL all-gather + pcramp_gpu_merge_shards"),
                       "db_entries_per_step": stats_acc["n_entries"] / n_scan, "hits_per_step": stats_acc["n_hits"] / n_scan,
                       "text_index": index_info, "partial_word_table": edge_info},
            "timed_region": {"repeats_resident": reps_resident, "repeats_e2e": reps_e2e, "seconds_resident": ms_resident * reps_resident * 1e-3,
                             "seconds_e2e": ms_e2e * reps_e2e * 1e-3,
                             "note": "the region of --steps steps (barrier + synchronize on both sides, CUDA events) is repeated until "
                                     "--min-seconds of device time; ms_per_step is the mean over all repeats"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_resident,
            "breakdown_ms_per_step": {k: stats_acc[k] / n_scan for k in ("ms_seed", "ms_scan", "ms_edge", "ms_db", "ms_score")},
            "pipeline": {"workers": W, "ms_per_step_one_batch_at_a_time": ms_sequential / a.steps,
                         "launches_per_step": launches_resident / max(1, a.steps),
                         "note": "value / e2e: %d batch(es) in flight per GPU (worker contexts sharing the resident targets and text index, one "
                                 "host thread each); roofline / breakdown: from a pass with one batch at a time" % W},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_at_bench": parity}
        line.update(legs)
        # the driver keeps the last ~1500 characters of the line: the facts a reader needs first go last
        def sig(x, n=4):
            """n significant digits: the tail of the line has to hold the whole summary"""
            if x is None or isinstance(x, (bool, str)):
                return x
            return float("%.*g" % (n, float(x)))

        # the legs first, the headline LAST: what the driver keeps of a long line is its tail
        line["summary"] = {
            "configs": {k: {"value": sig(v["value"]), "unit": v["unit"], "ms": sig(v.get("ms_per_step", v.get("ms_per_iteration"))),
                            "cpu": None if not v.get("cpu_baseline") else sig(v["cpu_baseline"]["value"]),
                            "parity": None if not v.get("parity") else v["parity"]["ok"],
                            "frac": sig((v.get("roofline") or {}).get("frac")),
                            **({"ms_first": sig(v["ms_first_iteration"]), "ms_later": sig(v["ms_later_iterations"])} if v.get("ms_later_iterations") else {})}
                        for k, v in cfg.items()},
            "index_build_ms": sig(index_info["ms_build"]), "index_gb": sig(index_info["bytes"] / 1e9),
            "dp_gcups": None if dp is None else sig(dp["value"]), "dp_gcups_e2e": None if dp is None else sig(dp["e2e"]["value"]),
            "dp_gcups_e2e_words": None if dp is None else sig(dp["e2e_words"]["value"]),
            "sw_gcups": None if legs["sw_gcups"] is None else sig(legs["sw_gcups"]["value"]),
            "sw_gcups_e2e": None if legs["sw_gcups"] is None else sig(legs["sw_gcups"]["e2e"]["value"]),
            "target_sharded_evals_per_s": None if tsh is None else sig(tsh["value"]),
            "target_sharded_ms_per_step": None if tsh is None else sig(tsh["ms_per_step"]),
            "cpu_reference_evals_per_s": None if cpu_baseline is None else sig(cpu_baseline["value"]),
            "parity_at_bench": None if parity is None else parity["ok"],
            "launches_per_step": launches_resident / max(1, a.steps), "ms_one_batch_at_a_time": sig(ms_sequential / a.steps),
            "roofline_kernel": roofline["kernel"], "roofline_frac": sig(roofline["frac"]),
            "ms_per_step": sig(ms_resident / a.steps), "e2e": sig(e2e_value, 5), "value": sig(value, 5)}
        print(json.dumps(line))
    # teardown order matters: torch tensors that were used on the library's stream must die before the stream does
    sys.stdout.flush()
    torch.cuda.synchronize()
    if world > 1:
        if by_targets and not p2p:
            del gat_any, gat_p1, packed_any, packed_p1, out_bits, out_cov
        dist.barrier()                      # p2p: nobody frees its exchange buffer while a peer could still store into it
    for c in ctxs[1:]:
        c.close()                           # workers before their parent (include/pcramp_gpu.h)
    del ext, host_cov, host_bits, f_pin, r_pin
    torch.cuda.empty_cache()
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # a normal interpreter exit from here on (exit hooks run): nothing of torch's still refers to the library's stream


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
