"""GPU (-m gpu): whole design runs through the command-line host (pcramp_b200/host/pcramp_b200 = pcramp's main loop, main.cpp:471-1130,
over the C ABI) against the STOCK program's reports: goldens written by tests/golden/make_design_golden.py from oracle/_ref/pcramp
(the unmodified reference built by oracle/Makefile) and, when that binary travelled with the snapshot, the live program.
Every line of the report must be identical: ASSAY lines (selected pairs, degeneracies, lower-cased re-used oligos), the coverage
scores, the T- / B- lists of detected sequences and the closing summary."""
import os
import subprocess

import pytest

from tests import design_cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pcramp_b200", "host", "pcramp_b200")
STOCK = os.path.join(ROOT, "oracle", "_ref", "pcramp")
CASES = {c.name: c for c in design_cases.cases()}


def run_host(case, tmp_path, extra=()):
    argv = design_cases.materialise(case, str(tmp_path))
    out = str(tmp_path / (case.name + ".out"))
    p = subprocess.run([HOST] + argv + list(extra) + ["-o", out, "--timing"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return design_cases.report_lines(out), argv, p.stderr


@pytest.mark.parametrize("name", sorted(CASES))
def test_report_equals_stock_program_golden(name, tmp_path):
    case = CASES[name]
    got, _, log = run_host(case, tmp_path)
    with open(os.path.join(ROOT, "tests", "golden", "design_%s.txt" % name)) as fh:
        want = fh.read().splitlines()
    assert [x for x in got if x.startswith("ASSAY")] == [x for x in want if x.startswith("ASSAY")], log[-1500:]
    assert got == want
    assert sum(x.startswith("ASSAY") for x in got) >= 3


@pytest.mark.skipif(not os.path.exists(STOCK), reason="the stock program did not travel")
def test_report_equals_live_stock_program(tmp_path):
    """another seed than the goldens', both programs run here"""
    case = CASES["exhaust"]
    case = design_cases.DesignCase("exhaust_live", case.targets, ["--seed", "1234", "--count", "6", "--trial", "120", "-d", "2"])
    got, argv, _ = run_host(case, tmp_path)
    out = str(tmp_path / "stock.out")
    subprocess.run([STOCK] + argv + ["-o", out, "-v", "silent"], check=True, stderr=subprocess.DEVNULL, timeout=900)
    assert got == design_cases.report_lines(out)


def test_design_iteration_through_the_c_abi(gpu):
    """the same iteration called in-process (what bench.py times): a found assay detects what it claims"""
    import numpy as np
    from pcramp_b200 import TARGET
    from pcramp_b200.api import DesignLoop
    case = CASES["exhaust"]
    coll = case.targets
    gpu.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
    gpu.upload_sequences(1, np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
    gpu.upload_sequences(2, np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32))
    gpu.set_pool(np.zeros((0, 2), np.uint64), np.zeros((0, 2), np.uint64))
    loop = DesignLoop(gpu, 5, num_trial=100)
    try:
        res = loop.iteration()
        assert res.found == 1 and res.target_coverage == 6.0 and res.targets_remaining == 12 and res.n_amplicons_added > 0
        t, _ = loop.matches(coll.n)
        assert int(t.sum()) == 6
        res2 = loop.iteration()
        assert res2.found == 1 and res2.targets_remaining == 6 and res2.iteration == 2
    finally:
        loop.close()
