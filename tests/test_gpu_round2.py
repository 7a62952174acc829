"""GPU tests added in round 2: stream/launch ordering of the search kernels (programmatic dependent
launch), non-finite queries with device outputs, indices on several devices in one process, and
the NVLink peer-memory exchange protocol (fused push + flag + merge in the finalize kernel)
exercised by TWO PROCESSES sharing one GPU, so it runs on the driver's single-GPU box too.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from multimodal_audio_search_b200 import SegmentIndex, synth          # noqa: E402
from oracle import numpy_oracle as no                                 # noqa: E402  (checker only)
from tests.util import BF16_TOL, FP32_TOL, assert_topk_matches, result_row   # noqa: E402

pytestmark = pytest.mark.gpu


def _lib(seed, n, nq, plants, dtype="fp32", partial=False):
    idx = SegmentIndex(dtype, capacity=n)
    idx.append_synth(seed, n, 0, n, n_queries=nq, plants=plants, partial=partial)
    return idx


def test_query_written_by_the_kernel_in_front_of_a_single_query_search():
    """the scan is launched with programmatic stream serialization; a 1-query search
    with a DEVICE query has no staging copy in front of it, so the kernel in front of the scan is
    the caller's own producer (MiniLM's last layer in the app).  The scan must not read the query
    before that kernel has completed."""
    torch = pytest.importorskip("torch")
    seed, n, nq = 91, 400_000, 24
    idx = _lib(seed, n, nq, 30)
    q = synth.raw_queries(seed, 0, nq)
    want = [idx.search(q[i:i + 1], 0.5, 0.5, k=10) for i in range(nq)]
    qd = torch.from_numpy(q).cuda()
    buf = torch.zeros(384, device="cuda")
    big = torch.randn(64 << 20, device="cuda")              # producer: a long kernel, then the query write
    got = []
    for i in range(nq):
        # a slow elementwise kernel followed by the (tiny) kernel that writes the query; the search
        # goes out right behind them on the same stream
        big.mul_(1.0000001)
        buf.copy_(qd[i] + big[:384] * 0.0)
        got.append(idx.search(buf.unsqueeze(0), 0.5, 0.5, k=10))
    torch.cuda.synchronize()
    for i in range(nq):
        np.testing.assert_array_equal(got[i].indices.cpu().numpy(), want[i].indices)
        np.testing.assert_array_equal(got[i].fusion.cpu().numpy(), want[i].fusion)
    idx.close()


@pytest.mark.parametrize("dtype,path,nq", [("fp32", "gemv", 1), ("fp32", "gemv", 5), ("bf16", "gemm", 70)])
def test_nonfinite_query_with_device_outputs_is_reported_per_query_and_leaves_no_residue(dtype, path, nq):
    """with out_loc = CAB_DEVICE nobody read the non-finite flag; it stayed set and
    broke the next append / host search.  Now: out_count = -1 for exactly the bad queries, nothing
    sticky on the handle."""
    torch = pytest.importorskip("torch")
    seed, n = 17, 30_000
    idx = _lib(seed, n, nq, 24, dtype)
    q = synth.raw_queries(seed, 0, nq)
    bad = q.copy()
    bad_rows = sorted({0, nq // 2})
    for r in bad_rows:
        bad[r, 7] = np.nan if r == 0 else np.inf
    good = idx.search(q, 0.5, 0.5, k=10, path=path)
    dev = idx.search(torch.from_numpy(bad).cuda(), 0.5, 0.5, k=10, path=path)
    cnt = dev.count.cpu().numpy()
    ind = dev.indices.cpu().numpy()
    for r in range(nq):
        if r in bad_rows:
            assert cnt[r] == -1 and (ind[r] == -1).all()
        else:
            assert cnt[r] == good.count[r]
            np.testing.assert_array_equal(ind[r], good.indices[r])
    # no residue: a valid append and a valid host search right after
    a, b, f, _ = synth.library(seed + 1, 64, 1, 4)
    idx.append(a, b, f)
    again = idx.search(q, 0.5, 0.5, k=10, path=path)
    assert (again.count >= good.count).all()
    # host outputs: the reference's ValueError (sklearn's message)
    with pytest.raises(ValueError, match="NaN|infinity"):
        idx.search(bad, 0.5, 0.5, k=10, path=path)
    ok = idx.search(q, 0.5, 0.5, k=10, path=path)
    np.testing.assert_array_equal(ok.indices, again.indices)
    # the same through candidates + merge (the merge has no query: the marker rides in the records)
    cands = idx.search_candidates(bad, 0.5, 0.5, k=10, path=path)
    merged = idx.merge_candidates(cands.unsqueeze(0), 0.5, 0.5, k=10, to_host=False)
    mc = merged.count.cpu().numpy()
    assert [int(mc[r]) == -1 for r in range(nq)] == [r in bad_rows for r in range(nq)]
    with pytest.raises(ValueError):
        idx.merge_candidates(cands.unsqueeze(0), 0.5, 0.5, k=10, to_host=True)
    idx.close()


def test_indices_on_two_devices_in_one_process():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: the 64 KB multi-query scan and
    the 214 KB tensor-core scan must launch on a second device of the same process."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    seed, n, nq = 5, 60_000, 70
    q = synth.raw_queries(seed, 0, nq)
    out = []
    for dev in (0, 1):
        idx = SegmentIndex("bf16", capacity=n, device=dev)
        idx.append_synth(seed, n, 0, n, n_queries=nq, plants=24)
        out.append((idx.search(q[:8], 0.4, 0.6, k=10, path="gemv"), idx.search(q, 0.4, 0.6, k=10, path="gemm")))
        idx.close()
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a.indices, b.indices)
        np.testing.assert_array_equal(a.fusion, b.fusion)


def test_failed_append_leaves_table_and_index_in_step():
    """the columnar table used to forget a batch's embeddings before the device append
    had succeeded."""
    from multimodal_audio_search_b200 import DualPipelineAudioSearch
    from multimodal_audio_search_b200.segment_table import SegmentTable

    class Emb:
        def encode(self, text):
            return synth.raw_queries(3, 0, 1)[0]

    a, b, f, _ = synth.library(3, 50, 1, 10)

    def seg(i, bad=False):
        e = a[i].copy()
        if bad:
            e[3] = np.nan
        return {"segment_id": f"seg_{i}", "start_time": 10.0 * i, "end_time": 10.0 * i + 10, "duration": 10.0,
                "asr_text": "t", "asr_embedding": e, "asr_success": True, "audio_description": "d",
                "audio_embedding": b[i], "audio_success": True, "audio_data": np.zeros(4, np.float32),
                "sample_rate": 16000}
    eng = DualPipelineAudioSearch(text_embedder=Emb())
    eng.audio_segments = SegmentTable.from_segments([seg(i) for i in range(20)])
    r0, _ = eng.search_with_fusion("anything")
    eng.audio_segments.extend([seg(i, bad=(i == 25)) for i in range(20, 30)])
    with pytest.raises(ValueError):
        eng.search_with_fusion("anything")
    with pytest.raises(ValueError):                     # still the reference's error, not "out of step"
        eng.search_with_fusion("anything")
    assert eng.audio_segments.n_pending == 10           # nothing was forgotten
    assert len(eng._cab_library.index) == 20


def test_pipelined_searches_match_one_at_a_time():
    """Option "queries_settled": search i+1's scan starts while search i's finalize is in flight.
    A long run of back-to-back device searches (GEMV single, GEMV batch, tensor-core) must equal
    the same searches issued one at a time with a synchronisation in between."""
    torch = pytest.importorskip("torch")
    seed, n, nq = 23, 300_000, 96
    for dtype in ("fp32", "bf16"):
        idx = _lib(seed, n, nq, 30, dtype, partial=True)
        q = synth.raw_queries(seed, 0, nq)
        qd = torch.from_numpy(q).cuda()
        wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
        plan = [(i, i + 1, "gemv") for i in range(40)] + [(40, 47, "gemv"), (0, 33, "gemv")]
        if dtype == "bf16":
            plan += [(0, 96, "gemm"), (3, 4, "gemv"), (10, 80, "gemm")]
        want = []
        for lo, hi, path in plan:
            want.append(idx.search(qd[lo:hi], wa[lo:hi], wb[lo:hi], k=10, path=path))
            torch.cuda.synchronize()
        idx.set_option("queries_settled", 1)
        for rep in range(3):
            got = [idx.search(qd[lo:hi], wa[lo:hi], wb[lo:hi], k=10, path=path) for lo, hi, path in plan]
            torch.cuda.synchronize()
            for g, w in zip(got, want):
                assert torch.equal(g.indices, w.indices) and torch.equal(g.fusion, w.fusion) and torch.equal(g.count, w.count)
        idx.close()


def test_fused_exchange_world_of_one_equals_plain_search():
    """cab_search_sharded with a world of one rank: the finalize kernel pushes into its own exchange
    buffer, raises and waits on its own flag and merges -- same answer as cab_search, over many
    epochs (both buffer parities), host and device outputs, k = 10 and 100."""
    torch = pytest.importorskip("torch")
    seed, n, nq = 29, 120_000, 40
    for dtype in ("fp32", "bf16"):
        idx = _lib(seed, n, nq, 220, dtype, partial=True)
        idx.row_base = 1000
        idx.peer_attach(idx.peer_init(0, 1, max_queries=128, max_k=128))
        idx.set_option("stamp_exchange", 1)
        q = synth.raw_queries(seed, 0, nq)
        qd = torch.from_numpy(q).cuda()
        for k in (10, 100):
            for lo, hi in ((0, 1), (1, 2), (2, 9), (0, 40)):
                wa = np.linspace(0.3, 0.7, hi - lo); wb = 1 - wa
                ref = idx.search(q[lo:hi], wa, wb, k=k)
                host = idx.search_sharded(q[lo:hi], wa, wb, k=k)
                dev = idx.search_sharded(qd[lo:hi], wa, wb, k=k, to_host=False)
                np.testing.assert_array_equal(host.indices, ref.indices)
                np.testing.assert_array_equal(host.fusion, ref.fusion)
                np.testing.assert_array_equal(host.count, ref.count)
                np.testing.assert_array_equal(dev.indices.cpu().numpy(), ref.indices)
                np.testing.assert_array_equal(dev.asr_sim.cpu().numpy(), ref.asr_sim)
        st = idx.exchange_stamps()
        assert len(st) > 0 and (np.diff(st.astype(np.int64), axis=1) >= 0).all()
        idx.close()


# ---- two processes, one GPU: the peer-memory exchange end to end -----------------------------------------
def _peer_worker(rank, world, port, seed, n, nq, out_path, dtype):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodal_audio_search_b200 import SegmentIndex, ShardedSearcher, synth
    from multimodal_audio_search_b200.sharded import shard_range
    torch.cuda.set_device(0)
    lo, hi = shard_range(n, rank, world)
    idx = SegmentIndex(dtype.split("+")[0], capacity=hi - lo, device=0)
    idx.append_synth(seed, n, lo, hi, n_queries=nq, plants=60, partial=True)
    idx.row_base = lo
    idx.set_option("gemm_min_queries", 4)          # bf16: batches of >= 4 queries take the tensor-core scan (fused merge too)
    if dtype.endswith("+shadow"):                  # fp32 shard: batches preselected from bf16 shadows, exact local top-k pushed
        idx.enable_tensor_core_batches()
    sh = ShardedSearcher(idx, rank, world, exchange="p2p", max_queries=128, max_k=128)
    q = synth.raw_queries(seed, 0, nq)
    qd = torch.from_numpy(q).cuda()
    wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
    res = {}
    # fused merge (few queries, one pass), separate merge launch (many queries), both buffer parities,
    # host and device outputs, k = 10 / 100
    cases = [(0, 1, 10, True), (1, 2, 10, False), (2, 3, 100, True), (3, 12, 10, False), (0, 48, 10, True),
             (0, 100, 10, False), (5, 6, 100, False)]
    for rep in range(2):
        for ci, (a, b, k, to_host) in enumerate(cases):
            r = sh.search(q[a:b] if to_host else qd[a:b], wa[a:b], wb[a:b], k=k, to_host=to_host)
            ind = r.indices if to_host else r.indices.cpu().numpy()
            fus = r.fusion if to_host else r.fusion.cpu().numpy()
            cnt = r.count if to_host else r.count.cpu().numpy()
            res[f"i{rep}_{ci}"], res[f"f{rep}_{ci}"], res[f"c{rep}_{ci}"] = ind, fus, cnt
    # a library so small that the second rank's shard is EMPTY (nothing to scan: it pushes empty lists)
    tiny = SegmentIndex("fp32", capacity=8, device=0)
    t_lo, t_hi = (0, 3) if rank == 0 else (3, 3)                   # all 3 rows on rank 0
    if t_hi > t_lo:
        tiny.append_synth(seed + 7, 3, t_lo, t_hi, n_queries=1, plants=3)
    tiny.row_base = t_lo
    sh2 = ShardedSearcher(tiny, rank, world, exchange="p2p", max_queries=4, max_k=16)
    q2 = synth.raw_queries(seed + 7, 0, 1)
    for rep in range(2):
        r = sh2.search(q2, 0.5, 0.5, k=10, threshold=-1.0, to_host=True)
        res[f"tiny_i{rep}"], res[f"tiny_f{rep}"], res[f"tiny_c{rep}"] = r.indices, r.fusion, r.count
    res["shadow_queries"] = np.array([idx.get_option("total_shadow_queries")])
    torch.cuda.synchronize()
    dist.barrier()
    np.savez(out_path + f".{rank}.npz", **res)
    idx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype", [(2, "fp32"), (3, "fp32"), (2, "bf16"), (2, "fp32+shadow")])
def test_peer_exchange_two_processes_one_gpu(tmp_path, world, dtype):
    """World of 2 (3) ranks = 2 (3) processes, all on cuda:0, exchange buffers shared through CUDA IPC: the
    finalize kernel's peer stores + epoch flags + merge must give every rank exactly the answer of
    one index over the whole library.  (The GPU time-slices the two contexts; the flag waits are
    bounded, so a protocol bug fails the test instead of hanging the box.)"""
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    seed, n, nq = 41, 90_000, 100                    # world = 3, k = 100: 300 candidates per query -> the merge SORTS
    whole = SegmentIndex(dtype.split("+")[0], capacity=n)
    whole.append_synth(seed, n, 0, n, n_queries=nq, plants=60, partial=True)
    q = synth.raw_queries(seed, 0, nq)
    wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
    cases = [(0, 1, 10), (1, 2, 10), (2, 3, 100), (3, 12, 10), (0, 48, 10), (0, 100, 10), (5, 6, 100)]
    # the reference answer comes from the exact scan; bf16 shards answer batches through the tensor cores
    want = [whole.search(q[a:b], wa[a:b], wb[a:b], k=k, path="gemv") for a, b, k in cases]
    whole.close()
    out = str(tmp_path / "peer")
    port = 29500 + (os.getpid() % 2000) + 7 * world + {"fp32": 0, "bf16": 3, "fp32+shadow": 5}[dtype]   # one rendezvous port per variant
    ctx = mp.spawn(_peer_worker, args=(world, port, seed, n, nq, out, dtype), nprocs=world, join=False)
    deadline = 240
    import time
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail("peer exchange workers did not finish")
    tiny = SegmentIndex("fp32", capacity=8)
    tiny.append_synth(seed + 7, 3, 0, 3, n_queries=1, plants=3)
    tiny_want = tiny.search(synth.raw_queries(seed + 7, 0, 1), 0.5, 0.5, k=10, threshold=-1.0)
    tiny.close()
    assert tiny_want.count[0] >= 1
    for rank in range(world):
        z = np.load(out + f".{rank}.npz")
        # the fp32 + shadow shards really answered their batches (9, 48 and 100 queries, twice) on the tensor cores
        assert int(z["shadow_queries"][0]) == (2 * (9 + 48 + 100) if dtype.endswith("+shadow") else 0)
        for rep in range(2):
            np.testing.assert_array_equal(z[f"tiny_i{rep}"], tiny_want.indices, err_msg=f"rank {rank}: empty second shard")
            np.testing.assert_array_equal(z[f"tiny_f{rep}"], tiny_want.fusion)
            for ci, w in enumerate(want):
                if dtype == "fp32":
                    np.testing.assert_array_equal(z[f"i{rep}_{ci}"], w.indices, err_msg=f"rank {rank} case {ci}")
                    np.testing.assert_array_equal(z[f"f{rep}_{ci}"], w.fusion)
                    np.testing.assert_array_equal(z[f"c{rep}_{ci}"], w.count)
                    continue
                # bf16: tensor-core and GEMV scans may order rows whose scores differ by rounding noise
                # differently at the k-th boundary; every row both hold carries the same score
                gi, gf = z[f"i{rep}_{ci}"], z[f"f{rep}_{ci}"]
                for j in range(len(gi)):
                    a_ = {int(r): float(f) for r, f in zip(w.indices[j], w.fusion[j]) if r >= 0}
                    b_ = {int(r): float(f) for r, f in zip(gi[j], gf[j]) if r >= 0}
                    assert all(a_[r] == b_[r] for r in set(a_) & set(b_))
                    kth = min(list(a_.values()) + list(b_.values()))
                    assert all(abs((a_.get(r) or b_.get(r)) - kth) <= 1e-5 for r in set(a_) ^ set(b_)), (rank, ci, j)


@pytest.mark.parametrize("dtype,rel", [("fp32", 2e-6), ("bf16", 4e-3)])
def test_raw_dot_product_scoring_for_non_unit_embeddings(dtype, rel, tmp_path):
    """Option "raw_dot": similarities are <query, row> of the vectors as appended (the ranking rule of
    previous_iterations/clean_audio_search.py:306-310), for rows and queries of any length; fusion,
    threshold and order unchanged.  The row lengths survive growth of the store and a save / load."""
    rng = np.random.default_rng(12)
    n, nq = 50_000, 5
    a = rng.standard_normal((n, 384)).astype(np.float32) * rng.uniform(0.01, 0.3, (n, 1)).astype(np.float32)
    b = rng.standard_normal((n, 384)).astype(np.float32) * rng.uniform(0.01, 0.3, (n, 1)).astype(np.float32)
    f = rng.choice(np.array([1, 2, 3, 3, 3], np.uint8), n)
    a[(f & 1) == 0] = 0; b[(f & 2) == 0] = 0
    q = rng.standard_normal((nq, 384)).astype(np.float32) * np.float32(0.7)
    idx = SegmentIndex(dtype, capacity=1000)                 # grows several times while appending
    for lo in range(0, n, 7000):
        idx.append(a[lo:lo + 7000], b[lo:lo + 7000], f[lo:lo + 7000])
    cos = idx.search(q, 0.4, 0.6, k=20)
    idx.set_option("raw_dot", 1)
    res = idx.search(q, 0.4, 0.6, k=20)

    def check(res):
        for qi in range(nq):
            sa, sb = a @ q[qi], b @ q[qi]
            fusion, _, _ = no.fuse(sa, sb, f, 0.4, 0.6)
            order = np.argsort(-fusion, kind="stable")
            want = [int(r) for r in order[:20] if fusion[r] > 0.1]
            gi, gf, ga, gb, _ = result_row(res, qi)
            tol = rel * max(1.0, float(np.abs(fusion[np.isfinite(fusion)]).max()))
            assert len(gi) == len(want) or abs(fusion[order[len(gi)]] - 0.1) <= tol
            for p, r in enumerate(gi.tolist()):
                assert abs(gf[p] - fusion[r]) <= tol and abs(ga[p] - sa[r]) <= tol and abs(gb[p] - sb[r]) <= tol
                if p < len(want) and r != want[p]:
                    assert abs(fusion[r] - fusion[want[p]]) <= tol
    check(res)
    assert res.indices.tolist() != cos.indices.tolist()          # not the cosine ranking
    with pytest.raises(Exception, match="GEMV|bf16"):
        idx.search(np.tile(q, (20, 1))[:70], 0.4, 0.6, k=10, path="gemm")
    path = str(tmp_path / "raw.idx")
    idx.save(path)
    again = SegmentIndex.load(path)
    again.set_option("raw_dot", 1)
    r2 = again.search(q, 0.4, 0.6, k=20)
    assert r2.indices.tolist() == res.indices.tolist() and r2.fusion.tolist() == res.fusion.tolist()
    idx.close(); again.close()


def test_alternating_device_and_host_calls_on_one_handle():
    """A device-tensor search runs on torch's stream and returns without waiting; a host search on
    the same handle runs on the handle's own stream.  Both use the handle's workspace (partial
    lists, chunk tickets): the second call must be ordered after the first (found on two B200s: the
    two multi-batch scans ran concurrently and each lost rows)."""
    torch = pytest.importorskip("torch")
    seed, n, nq = 31, 400_000, 80
    idx = _lib(seed, n, nq, 30, "fp32", partial=True)
    q = synth.raw_queries(seed, 0, nq)
    qd = torch.from_numpy(q).cuda()
    wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
    want_a = idx.search(q[:40], wa[:40], wb[:40], k=100)
    want_b = idx.search(q[40:], wa[40:], wb[40:], k=100)
    for _ in range(6):
        dev = idx.search(qd[:40], wa[:40], wb[:40], k=100)            # async, torch's stream
        host = idx.search(q[40:], wa[40:], wb[40:], k=100)            # handle's own stream, right behind it
        dev2 = idx.search(qd[:40], wa[:40], wb[:40], k=100)
        np.testing.assert_array_equal(host.indices, want_b.indices)
        np.testing.assert_array_equal(dev.indices.cpu().numpy(), want_a.indices)
        np.testing.assert_array_equal(dev2.indices.cpu().numpy(), want_a.indices)
    idx.close()


def test_fp32_batches_on_tensor_cores_with_exactness_certificates():
    """fp32 library + bf16 shadows: batches are preselected on the tensor cores, re-scored exactly and
    certified per query; uncertified queries are re-run on the exact scan.  The answer must be the
    fp32 GEMV path's (same re-score code, float64 fusion), for planted data (everything certified),
    for dense threshold-level noise at k = 100 (certificates fail, re-runs one by one and as a whole
    batch), host and device outputs, and after the store has grown."""
    torch = pytest.importorskip("torch")
    seed, n, nq = 61, 300_000, 96
    idx = SegmentIndex("fp32", capacity=150_000)
    idx.append_synth(seed, n, 0, 150_000, n_queries=nq, plants=30, partial=True)
    idx.enable_tensor_core_batches()
    idx.append_synth(seed, n, 150_000, n, n_queries=nq, plants=30, partial=True)      # grows the shadows too
    q = synth.raw_queries(seed, 0, nq)
    qd = torch.from_numpy(q).cuda()
    wa = np.linspace(0.2, 0.8, nq); wb = 1 - wa
    a, b, f, _ = synth.library(seed, n, nq, 30, True)

    def same(x, y, what):
        xi = x.indices.cpu().numpy() if hasattr(x.indices, "cpu") else x.indices
        xf = x.fusion.cpu().numpy() if hasattr(x.fusion, "cpu") else x.fusion
        bad = [i for i in range(len(xi)) if not (np.array_equal(xi[i], y.indices[i]) and np.array_equal(xf[i], y.fusion[i]))]
        for i in bad:                                  # only swaps between rows closer than the fp32 tolerance are allowed
            sx = dict(zip(xi[i].tolist(), xf[i].tolist())); sy = dict(zip(y.indices[i].tolist(), y.fusion[i].tolist()))
            kth = min(min(sx.values()), min(sy.values()))
            for r in set(sx) ^ set(sy):
                assert abs(sx.get(r, sy.get(r)) - kth) <= FP32_TOL, (what, i, r)
        return len(bad)

    exact10 = idx.search(q, wa, wb, k=10, path="gemv")
    l0 = idx.launch_count
    got = idx.search(q, wa, wb, k=10)                              # auto: >= 64 queries -> tensor-core preselection
    assert idx.get_option("total_shadow_queries") == nq and idx.get_option("last_uncertified") == 0
    assert same(got, exact10, "planted k=10") == 0
    assert idx.launch_count - l0 == 3                                # prologue + tensor-core scan + finalize
    for qi in (0, 17, 95):
        o = no.search(q[qi], a, b, f, wa[qi], wb[qi], k=10)
        gi, gf, _, _, _ = result_row(got, qi)
        assert_topk_matches(gi, gf, o, FP32_TOL)
    dev = idx.search(qd, wa, wb, k=10, path="gemm")
    assert same(dev, exact10, "planted k=10 device") == 0
    # k = 100 on a library with 30 plants per query: ranks 31..100 are threshold-level noise rows,
    # dense in score -> the certificate (100 candidates above the 128th scan score + 4e-3) fails
    exact100 = idx.search(q, wa, wb, k=100, path="gemv")
    got100 = idx.search(q, wa, wb, k=100)
    n_rerun = idx.get_option("last_uncertified")
    assert n_rerun > 0
    same(got100, exact100, "dense k=100")
    dev100 = idx.search(qd, wa, wb, k=100, path="gemm")
    same(dev100, exact100, "dense k=100 device")
    # a few uncertified queries in a certified batch: mix planted k=10 queries with ... threshold 0.05 makes more rows compete
    low = idx.search(q, wa, wb, k=40, threshold=0.12)
    exact_low = idx.search(q, wa, wb, k=40, threshold=0.12, path="gemv")
    same(low, exact_low, "k=40")
    assert idx.get_option("total_shadow_queries") >= 5 * nq
    # a batch in which only SOME queries fail their certificate (gathered re-run): 70 queries with 128
    # strong planted rows each (certified at k = 100), 26 queries whose whole top-100 is dense noise
    mix = SegmentIndex("fp32", capacity=n)
    mix.append_synth(seed + 1, n, 0, n, n_queries=70, plants=128)
    mix.enable_tensor_core_batches()
    qm = synth.raw_queries(seed + 1, 0, nq)
    exact_mix = mix.search(qm, wa, wb, k=100, path="gemv")
    for variant in (qm, torch.from_numpy(qm).cuda()):
        got_mix = mix.search(variant, wa, wb, k=100)
        bad = mix.get_option("last_uncertified")
        assert 0 < bad <= 0.75 * nq, bad
        same(got_mix, exact_mix, "mixed batch")
        gc = got_mix.count.cpu().numpy() if hasattr(got_mix.count, "cpu") else got_mix.count
        np.testing.assert_array_equal(gc, exact_mix.count)
    mix.close()
    with pytest.raises(Exception, match="fp32"):
        SegmentIndex("bf16").set_option("tensor_core_shadow", 1)
    idx.enable_tensor_core_batches(False)
    with pytest.raises(Exception, match="bf16 index"):
        idx.search(q, wa, wb, k=10, path="gemm")
    idx.close()
