"""GPU (-m gpu): worker contexts (pcramp_gpu_create_worker) -- independent batches driven from several host threads against
ONE resident copy of the targets and the text index give exactly what the parent computes batch after batch."""
import threading

import numpy as np
import pytest

from pcramp_b200 import TARGET, GpuError, PcrampGpu, synth

pytestmark = pytest.mark.gpu


def test_workers_equal_the_parent_batch_by_batch():
    coll = synth.make_targets(31, 300, 6000, n_clades=5, between=0.15, within=0.04)
    n_batches, P = 12, 96
    f, r = synth.make_pairs(32, coll, n_batches * P)
    thr = float(np.float32(1.0) * np.float32(0.9))
    g = PcrampGpu(0)
    try:
        g.upload_sequences(TARGET, coll.nibbles, coll.byte_off, coll.length)
        want = []
        for b in range(n_batches):
            sl = slice(b * P, (b + 1) * P)
            ne, _ = g.select_words(TARGET, f[sl], r[sl], thr, want_keys=False)
            cov, bits = g.score_pairs(TARGET, f[sl], r[sl], thr, 1.0)
            want.append((ne, cov.copy(), bits.copy()))
        assert g.stats()["n_indexed"] > 0                       # the shared text index is what the workers will read
        ctxs = [g, g.worker(), g.worker()]
        got = [None] * n_batches
        err = []

        def run(k):
            try:
                c = ctxs[k]
                for b in range(k, n_batches, len(ctxs)):
                    sl = slice(b * P, (b + 1) * P)
                    ne, _ = c.select_words(TARGET, f[sl], r[sl], thr, want_keys=False)
                    cov, bits = c.score_pairs(TARGET, f[sl], r[sl], thr, 1.0)
                    got[b] = (ne, cov.copy(), bits.copy())
            except Exception as e:                                # noqa: BLE001
                err.append(e)

        for rep in range(3):
            th = [threading.Thread(target=run, args=(k,)) for k in range(len(ctxs))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not err, err
            for b in range(n_batches):
                assert got[b][0] == want[b][0] and np.array_equal(got[b][1], want[b][1]) and np.array_equal(got[b][2], want[b][2]), (rep, b)
        assert any(w[1].sum() > 0 for w in want)
        # a worker cannot change the shared sequences, and goes stale when the parent does
        w = ctxs[1]
        with pytest.raises(GpuError, match="worker"):
            w.split_sequence(TARGET, 0, 100)
        g.split_sequence(TARGET, 0, 100)
        with pytest.raises(GpuError, match="changed"):
            w.select_words(TARGET, f[:P], r[:P], thr, want_keys=False)
    finally:
        g.close()
