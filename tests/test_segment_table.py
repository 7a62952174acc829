"""Columnar segment metadata: a SegmentTable must behave like the
reference's `audio_segments` list of dicts (audio_search.py:275-294, :797) on the search path,
round-trip through its file, and materialise results without per-library work.  CPU only."""
import os
import time

import numpy as np
import pytest

from multimodal_audio_search_b200.segment_table import (LAZY_FIELDS, RECORD_ORDER, SegmentRecord,
                                                         SegmentTable)


def _segments(n, seed=0, none_every=4):
    """Records with the reference's key set and key order (:275-294)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        asr_ok = i % none_every != 1
        audio_ok = i % none_every != 2
        out.append({
            "segment_id": f"seg_{i}",
            "start_time": 5.0 * i,
            "end_time": 5.0 * i + 10.0,
            "duration": 10.0,
            "asr_text": f"hello wörld {i}" if asr_ok else "",
            "asr_embedding": rng.standard_normal(384).astype(np.float32) if asr_ok else None,
            "asr_success": asr_ok,
            "audio_description": f"a dog barks {i} times" if audio_ok else "",
            "audio_embedding": rng.standard_normal(384).astype(np.float32) if audio_ok else None,
            "audio_success": audio_ok,
            "audio_data": rng.standard_normal(160 + i).astype(np.float32),
            "sample_rate": 16000,
        })
    return out


def _same(record, segment):
    assert list(record.keys()) == list(segment.keys())
    for k, v in segment.items():
        got = record[k]
        if isinstance(v, np.ndarray):
            assert np.array_equal(got, v), k
        else:
            assert got == v and type(got) is type(v), (k, got, v)


def test_list_protocol_matches_reference_records():
    segs = _segments(9)
    t = SegmentTable()
    assert not t and len(t) == 0
    t.extend(segs[:5])
    t.extend(segs[5:])
    assert t and len(t) == 9
    assert list(RECORD_ORDER) == list(segs[0].keys())
    for i, s in enumerate(segs):
        _same(t[i], s)
    _same(t[-1], segs[-1])
    assert [r["segment_id"] for r in t] == [s["segment_id"] for s in segs]
    assert [r["segment_id"] for r in t[2:5]] == ["seg_2", "seg_3", "seg_4"]
    with pytest.raises(IndexError):
        t[9]


def test_records_are_lazy_until_read():
    t = SegmentTable.from_segments(_segments(3))
    r = t[1]
    assert not any(dict.__contains__(r, k) for k in LAZY_FIELDS)      # nothing fetched yet
    assert all(k in r for k in LAZY_FIELDS)
    assert r.get("audio_data").shape == (161,)
    assert dict.__contains__(r, "audio_data")                          # cached after first read
    assert r.get("nope", 7) == 7
    with pytest.raises(KeyError):
        r["nope"]
    merged = r.with_fields(fusion_score=0.5)                            # {**segment, ...} of :673-682
    assert isinstance(merged, SegmentRecord) and merged["fusion_score"] == 0.5
    assert list(merged.keys())[-1] == "fusion_score"
    plain = {**t[2]}                                                    # generic mapping unpack
    assert set(LAZY_FIELDS) <= set(plain) and type(plain) is dict


def test_drain_pending_gives_index_rows():
    segs = _segments(8)
    t = SegmentTable.from_segments(segs)
    assert t.n_pending == 8
    row0, asr, audio, flags = t.drain_pending()
    assert row0 == 0 and asr.shape == (8, 384) and t.n_pending == 0
    for i, s in enumerate(segs):
        assert np.array_equal(asr[i], s["asr_embedding"] if s["asr_embedding"] is not None else np.zeros(384, "f4"))
        assert np.array_equal(audio[i], s["audio_embedding"] if s["audio_embedding"] is not None else np.zeros(384, "f4"))
        assert flags[i] == (1 if s["asr_success"] else 0) | (2 if s["audio_success"] else 0)
    t.extend(_segments(2, seed=1))
    row0, asr, _, flags = t.drain_pending()
    assert row0 == 8 and asr.shape == (2, 384) and len(flags) == 2
    with pytest.raises(KeyError):                 # drained rows need the device index for embeddings
        t[0]["asr_embedding"]
    bad = dict(segs[0]); bad["asr_embedding"] = np.zeros(100, "f4")
    t.append(bad)
    with pytest.raises(ValueError, match="Incompatible dimension"):
        t.drain_pending()


def test_file_round_trip(tmp_path):
    segs = _segments(7)
    segs[3]["speaker"] = "alice"                                       # a key the schema does not know
    t = SegmentTable.from_segments(segs, file="talk.wav")
    path = str(tmp_path / "lib.meta")
    t.save(path)
    info = SegmentTable.file_info(path)
    assert info["n_rows"] == 7
    u = SegmentTable.load(path)
    assert len(u) == 7 and u.n_pending == 0
    for i, s in enumerate(segs):
        r = u[i]
        for k in ("segment_id", "start_time", "end_time", "duration", "asr_text", "asr_success",
                  "audio_description", "audio_success", "sample_rate"):
            assert r[k] == s[k] and type(r[k]) is type(s[k]), k
        assert r["file"] == "talk.wav"
        assert np.array_equal(r["audio_data"], s["audio_data"])
    assert u[3]["speaker"] == "alice" and "speaker" not in u[2]
    # append after load, save again: loaded part + tail are merged
    u.extend(_segments(2, seed=5), file="second.wav")
    path2 = str(tmp_path / "lib2.meta")
    u.save(path2)
    v = SegmentTable.load(path2)
    assert len(v) == 9 and v[8]["file"] == "second.wav" and v[0]["file"] == "talk.wav"
    assert np.array_equal(v[8]["audio_data"], u[8]["audio_data"])
    assert np.array_equal(v[2]["audio_data"], segs[2]["audio_data"])
    # without the audio blob
    t.save(path, audio=False)
    assert SegmentTable.load(path)[0]["audio_data"] is None


def test_file_validation(tmp_path):
    t = SegmentTable.from_segments(_segments(4))
    path = str(tmp_path / "lib.meta")
    t.save(path)
    raw = open(path, "rb").read()
    bad = str(tmp_path / "bad.meta")
    open(bad, "wb").write(b"NOTMETA1" + raw[8:])
    with pytest.raises(ValueError, match="bad magic"):
        SegmentTable.load(bad)
    open(bad, "wb").write(raw[:40])
    with pytest.raises(ValueError, match="too short"):
        SegmentTable.load(bad)
    open(bad, "wb").write(raw[:len(raw) - 10])
    with pytest.raises(ValueError, match="truncated"):
        SegmentTable.load(bad)
    os.remove(path + ".audio")
    with pytest.raises(ValueError, match="missing or truncated"):
        SegmentTable.load(path)


def test_clear_bumps_generation():
    t = SegmentTable.from_segments(_segments(2))
    g = t.generation
    t.clear(); t.clear()
    assert len(t) == 0 and t.generation == g + 2


def test_bulk_table_materialises_in_constant_time(tmp_path):
    n = 2_000_000
    rng = np.random.default_rng(0)
    t = SegmentTable.from_columns(n, asr_success=rng.integers(0, 2, n), audio_success=np.ones(n))
    assert len(t) == n and t[n - 1]["segment_id"] == f"seg_{n - 1}"
    path = str(tmp_path / "big.meta")
    t.save(path, audio=False)
    t0 = time.perf_counter()
    u = SegmentTable.load(path)
    rows = [u[int(i)] for i in rng.integers(0, n, 10)]
    dt = time.perf_counter() - t0
    assert len(rows) == 10 and rows[0]["end_time"] - rows[0]["start_time"] == 10.0
    assert dt < 0.5, f"load + 10 records took {dt:.3f}s for {n} rows"


def test_round_trip_property(tmp_path_factory):
    """Any list of reference-shaped records (unicode texts, empty strings, missing audio, extra
    JSON-able keys) survives extend -> save -> load, field by field."""
    from hypothesis import given, settings, strategies as st
    text = st.text(max_size=40)
    record = st.fixed_dictionaries({
        "start_time": st.floats(0, 1e6, allow_nan=False), "duration": st.floats(0, 30, allow_nan=False),
        "asr_text": text, "audio_description": text, "asr_success": st.booleans(), "audio_success": st.booleans(),
        "sample_rate": st.sampled_from([8000, 16000, 44100]), "n_audio": st.integers(0, 50),
        "extra": st.one_of(st.none(), st.dictionaries(st.text(min_size=1, max_size=8).filter(lambda k: k not in RECORD_ORDER and k != "file"),
                                                      st.one_of(st.integers(-5, 5), text), max_size=2)),
    })
    base = tmp_path_factory.mktemp("prop")
    counter = [0]

    @settings(max_examples=40, deadline=None)
    @given(st.lists(record, min_size=1, max_size=12), st.sampled_from([None, "clip.wav", "ünï.flac"]))
    def run(specs, file):
        segs = []
        for i, s in enumerate(specs):
            seg = {"segment_id": f"seg_{i}", "start_time": s["start_time"], "end_time": s["start_time"] + s["duration"],
                   "duration": s["duration"], "asr_text": s["asr_text"], "asr_embedding": None, "asr_success": s["asr_success"],
                   "audio_description": s["audio_description"], "audio_embedding": None, "audio_success": s["audio_success"],
                   "audio_data": np.arange(s["n_audio"], dtype=np.float32) if s["n_audio"] else None,
                   "sample_rate": s["sample_rate"]}
            if s["extra"]:
                seg.update(s["extra"])
            segs.append(seg)
        t = SegmentTable.from_segments(segs, file=file)
        counter[0] += 1
        path = str(base / f"t{counter[0]}.meta")
        t.save(path)
        u = SegmentTable.load(path)
        assert len(u) == len(segs)
        for i, seg in enumerate(segs):
            r = u[i]
            for k, v in seg.items():
                if k in ("asr_embedding", "audio_embedding"):
                    continue
                if k == "audio_data":
                    assert (r[k] is None) if v is None else np.array_equal(r[k], v)
                else:
                    assert r[k] == v, (k, r[k], v)
            assert r.get("file") == file
    run()
