"""CPU: bench.py's parity checker must accept the oracle's own answers and reject doctored ones
(missing expected row, wrong score, broken order, bad padding) on all three synthetic distributions."""
import importlib.util
import os

import numpy as np
import pytest

from multimodal_audio_search_b200 import synth
from oracle import numpy_oracle as no

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def _answers(mode, n, nq, plants, k):
    a, b, f, _ = synth.library(bench.SEED, n, nq, plants, False, mode=mode)
    q = synth.bench_queries(bench.SEED, mode, 0, nq)
    wa = np.array([bench.W_CLASSES[i % len(bench.W_CLASSES)] for i in range(nq)]); wb = 1.0 - wa
    ind = np.full((nq, k), -1, np.int64); fus = np.zeros((nq, k)); sa = np.zeros((nq, k), np.float32)
    sb = np.zeros((nq, k), np.float32); cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        o = no.search(q[i], a, b, f, wa[i], wb[i], k=k)
        c = len(o.indices)
        ind[i, :c], fus[i, :c], cnt[i] = o.indices, o.fusion, c
        sa[i, :c], sb[i, :c] = no.cosine_rows(q[i], a[o.indices]), no.cosine_rows(q[i], b[o.indices])
    return q, wa, wb, (list(range(nq)), ind, fus, sa, sb, cnt)


@pytest.mark.parametrize("mode,n", [("planted", 30_000), ("clustered", 60_000), ("ascending", 30_000)])
def test_parity_checker_accepts_the_oracle_and_rejects_doctored_results(mode, n):
    nq, plants, k = 3, 30, 10
    q, wa, wb, res = _answers(mode, n, nq, plants, k)
    args = (q, wa, wb, k, "fp32", n, nq, plants, 0.1, mode)
    good = bench.parity_check([res], *args)
    assert good["ok"] and good["queries"] == nq and good["max_abs_err"] <= 1e-6, good

    def doctored(fn):
        qids, ind, fus, sa, sb, cnt = (x.copy() if hasattr(x, "copy") else list(x) for x in res)
        fn(ind, fus, sa, sb, cnt)
        return bench.parity_check([(qids, ind, fus, sa, sb, cnt)], *args)

    def wrong_score(ind, fus, sa, sb, cnt):
        fus[1, 2] += 5e-5
    assert not doctored(wrong_score)["ok"]

    def swapped(ind, fus, sa, sb, cnt):
        for arr in (ind, fus, sa, sb):
            arr[0, [0, 1]] = arr[0, [1, 0]]
    assert not doctored(swapped)["ok"]

    def bad_padding(ind, fus, sa, sb, cnt):
        cnt[2] -= 1                                  # the last result is now "padding" that is not -1
    assert not doctored(bad_padding)["ok"]

    def lost_best(ind, fus, sa, sb, cnt):           # the best row dropped, everything shifted up, padded correctly
        for arr, pad in ((ind, -1), (fus, 0.0), (sa, 0.0), (sb, 0.0)):
            arr[0, :-1] = arr[0, 1:]
            arr[0, -1] = pad
        cnt[0] -= 1
    r = doctored(lost_best)
    if mode in ("planted", "clustered"):            # expected rows are known: completeness is checked
        assert not r["ok"] and "missing" in r["problems"][0]
    else:
        assert not r["ok"]                           # ascending: count != k


def test_reference_arm_and_our_arm_describe_the_same_config():
    assert bench.workload_config("1m_fp32_q1_top10", 2, 0.1)["global_segments"] == 2_000_000
    for name, (n, dtype, nq, k, path, mode) in bench.WORKLOADS.items():
        assert mode in synth.MODES and path in ("gemv", "gemm") and k <= 128 and nq <= 4096
