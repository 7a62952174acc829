"""Persistent index file: header checks on the CPU, save / load /
sharded range-load round trips on the GPU."""
import struct

import numpy as np
import pytest

from multimodal_audio_search_b200 import SegmentIndex, synth
from multimodal_audio_search_b200 import _native as N


def _header(n_rows=0, dtype=0, magic=b"CABIDX01", dim=384, row_base=0):
    eb = 2 if dtype == 1 else 4
    off_asr = 4096
    pad = lambda x: (x + 4095) // 4096 * 4096  # noqa: E731
    off_audio = pad(off_asr + n_rows * dim * eb)
    off_flags = pad(off_audio + n_rows * dim * eb)
    total = pad(off_flags + n_rows)
    h = magic + struct.pack("<IIIIQQQQQQ", 1, dim, dtype, 0, n_rows, row_base, off_asr, off_audio, off_flags, total)
    return h + b"\0" * (4096 - len(h)), total


def test_file_info_and_header_validation(tmp_path):
    from multimodal_audio_search_b200 import build
    build.build()
    good, total = _header(n_rows=3, dtype=1, row_base=77)
    p = tmp_path / "ok.cab"
    p.write_bytes(good + b"\0" * (total - 4096))
    assert SegmentIndex.file_info(str(p)) == {"dim": 384, "dtype": "bf16", "n_rows": 3, "row_base": 77}
    for name, blob in (("magic", _header(magic=b"NOTCABXX")[0]), ("dim", _header(dim=128)[0]),
                       ("short", good[:100]), ("truncated", good)):          # last: sections missing
        q = tmp_path / f"{name}.cab"
        q.write_bytes(blob)
        with pytest.raises(N.CabError):
            SegmentIndex.file_info(str(q))
    with pytest.raises(N.CabError):
        SegmentIndex.file_info(str(tmp_path / "missing.cab"))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_save_load_round_trip_is_bit_identical(tmp_path, dtype):
    seed, n = 606, 7001
    idx = SegmentIndex(dtype)
    idx.append_synth(seed, n, 0, n, n_queries=3, plants=30, partial=True)
    idx.row_base = 1000
    q = synth.raw_queries(seed, 0, 3)
    want = idx.search(q, [0.5, 0.3, 0.68], [0.5, 0.7, 0.32], k=25)
    path = str(tmp_path / f"lib_{dtype}.cab")
    idx.save(path)
    assert SegmentIndex.file_info(path) == {"dim": 384, "dtype": dtype, "n_rows": n, "row_base": 1000}
    back = SegmentIndex.load(path)
    assert len(back) == n and back.row_base == 1000 and back.dtype == dtype
    got = back.search(q, [0.5, 0.3, 0.68], [0.5, 0.7, 0.32], k=25)
    for f in ("indices", "fusion", "asr_sim", "audio_sim", "flags", "count"):
        np.testing.assert_array_equal(getattr(got, f), getattr(want, f))
    np.testing.assert_array_equal(back.read_rows(0, 0, 50), idx.read_rows(0, 0, 50))
    # the loaded index keeps accepting appends (incremental ingest)
    a, b, f2, _ = synth.library(seed + 1, 10, 1, 0)
    back.append(a, b, f2)
    assert len(back) == n + 10


@pytest.mark.gpu
def test_range_loads_are_shards_of_the_same_library(tmp_path):
    torch = pytest.importorskip("torch")
    seed, n, k = 707, 12000, 40
    idx = SegmentIndex("fp32")
    idx.append_synth(seed, n, 0, n, n_queries=2, plants=50, partial=True)
    q = synth.raw_queries(seed, 0, 2)
    want = idx.search(q, [0.4, 0.6], [0.6, 0.4], k=k)
    path = str(tmp_path / "lib.cab")
    idx.save(path)
    shards = [SegmentIndex.load(path, rows=r) for r in ((0, 5000), (5000, 5001), (5001, n))]
    assert [s.row_base for s in shards] == [0, 5000, 5001]
    gathered = torch.stack([s.search_candidates(q, [0.4, 0.6], [0.6, 0.4], k=k) for s in shards]).contiguous()
    got = shards[0].merge_candidates(gathered, [0.4, 0.6], [0.6, 0.4], k=k)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.fusion, want.fusion)
    with pytest.raises(N.CabError):
        SegmentIndex.load(path, rows=(5, n + 1))
