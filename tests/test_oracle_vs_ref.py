"""CPU, dev container only: the oracle restatement against the compiled reference itself (oracle/_ref), on the
seeded scenarios and on randomly drawn ones beyond the committed golden set.  Skipped where the reference
library is absent."""
import numpy as np
import pytest

from pcramp_b200 import synth
from tests import scenarios

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("fn", scenarios.ALL, ids=lambda f: f.__name__[2:])
def test_scenario(oracle, ref, fn):
    sc = fn()
    a = scenarios.run_checker(oracle, sc, "oracle")
    b = scenarios.run_checker(ref, sc, "ref")
    for k in b:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("seed", range(6))
def test_random_collections(oracle, ref, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 9))
    base = synth.make_targets(2000 + seed, n, int(rng.integers(120, 500)), n_clades=2, between=0.1, within=0.05)
    codes = [base.codes(i).copy() for i in range(n)]
    for c in codes:  # sprinkle EOS, degenerate codes, ragged lengths
        if rng.random() < 0.6:
            c[rng.integers(0, len(c), size=int(rng.integers(1, 4)))] = 0
        if rng.random() < 0.6:
            k = rng.integers(0, len(c), size=5)
            c[k] |= synth.CODE[rng.integers(0, 4, size=5)]
    codes = [c[:int(rng.integers(max(1, len(c) - 70), len(c) + 1))] for c in codes]
    coll = synth.Collection(codes)
    f, r = synth.make_pairs(3000 + seed, base, 20, degenerate_fraction=0.3)
    sc = scenarios.Scenario("rand%d" % seed, coll, f, r, target_threshold=0.9, search_multiplier=0.9, target_threshold_raw=0.9,
                            search_multiplier_raw=0.9, optimize_5=bool(seed & 1), optimize_3=bool(seed & 2), taq=bool(seed & 1),
                            min_oligo_length=int(rng.integers(12, 20)), amp=(40, 300))
    a = scenarios.run_checker(oracle, sc, "oracle")
    b = scenarios.run_checker(ref, sc, "ref")
    for k in b:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
