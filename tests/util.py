"""Shared helpers for the parity tests."""
import numpy as np

FP32_TOL = 1e-5     # BASELINE.json north_star: fp32 scores within 1e-5 absolute
BF16_TOL = 2e-3     # bf16 mode: 2e-3 absolute


def assert_topk_matches(got_idx, got_fusion, ref, tol, k=None, threshold=0.1):
    """`ref` is an oracle OracleResult (with `all_fusion` for every row).

    Index lists must be identical except where oracle scores are closer than `tol` (adjacent
    gaps, or the gap to the k-th / to the threshold): a differing index is accepted only if the
    oracle's own score for it is within tol of the oracle's score at that rank."""
    got_idx = [int(i) for i in got_idx]
    ref_idx = [int(i) for i in ref.indices]
    allf = ref.all_fusion
    n_common = min(len(got_idx), len(ref_idx))
    for r in range(n_common):
        assert abs(got_fusion[r] - ref.fusion[r]) <= tol, (r, got_fusion[r], ref.fusion[r])
        if got_idx[r] != ref_idx[r]:
            assert abs(allf[got_idx[r]] - ref.fusion[r]) <= tol, \
                f"rank {r}: got segment {got_idx[r]} (oracle score {allf[got_idx[r]]}) vs {ref_idx[r]} ({ref.fusion[r]})"
    # length differences only through threshold / k-boundary straddlers
    for extra in got_idx[n_common:]:
        assert abs(allf[extra] - threshold) <= tol, f"extra result {extra} score {allf[extra]}"
    for missing in ref_idx[n_common:]:
        assert abs(allf[missing] - threshold) <= tol, f"missing result {missing} score {allf[missing]}"
    assert len(set(got_idx)) == len(got_idx)


def result_row(res, q=0):
    c = int(res.count[q])
    assert (res.indices[q, c:] == -1).all()
    return res.indices[q, :c], res.fusion[q, :c], res.asr_sim[q, :c], res.audio_sim[q, :c], res.flags[q, :c]


def python_fusion(sa, sb, flags, wa, wb):
    """The reference's float64 arithmetic (audio_search.py:656-670) on one result."""
    ea = wa if flags & 1 else 0
    eb = wb if flags & 2 else 0
    tot = ea + eb
    ea /= tot
    eb /= tot
    return ea * float(sa) + eb * float(sb), ea, eb
