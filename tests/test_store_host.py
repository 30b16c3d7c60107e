"""EmbeddingStore host logic against the reference's own behaviour (tests/golden/store_golden.json, produced by the
reference's EmbeddingStore on the cinderella chunks).  A fake embedding model replays the reference's embeddings, so
no GPU is touched: ids, dedup, return values, lookups, parquet schema and reload are what is checked."""
import json
import os

import numpy as np
import pytest

from comorag_b200.embedding_store import EmbeddingStore, compute_mdhash_id

HERE = os.path.dirname(__file__)


@pytest.fixture(scope="module")
def gold():
    g = json.load(open(os.path.join(HERE, "golden", "store_golden.json")))
    g["embeddings"] = np.load(os.path.join(HERE, "golden", "store_golden_embeddings.npy"))
    return g


class ReplayModel:
    """batch_encode returns the row the reference stored for that text."""

    def __init__(self, gold):
        self.table = dict(zip(gold["texts"], gold["embeddings"]))
        self.embedding_dim = gold["embeddings"].shape[1]
        self.calls = []

    def batch_encode(self, texts, **kw):
        self.calls.append(list(texts))
        return np.stack([self.table[t] for t in texts]).astype(np.float32)


def test_insert_lookup_and_parquet_match_reference(gold, tmp_path):
    model = ReplayModel(gold)
    chunks = gold["texts"][:6]
    store = EmbeddingStore(model, str(tmp_path / "chunk_embeddings"), 4, "chunk")
    assert os.path.basename(store.filename) == gold["filename"]
    assert store.hash_ids == [] and store.hash_id_to_text == {} and store.text_to_hash_id == {}
    r1 = store.insert_strings(chunks)
    r2 = store.insert_strings(chunks[:2] + ["a brand new chunk"])
    r3 = store.insert_strings(chunks[:2])
    assert [repr(r1), repr(r2), repr(r3)] == gold["insert_returns"]
    assert model.calls == [chunks, ["a brand new chunk"]]          # only missing texts are encoded
    assert store.get_all_ids() == gold["hash_ids"]
    assert store.texts == gold["texts"]
    assert store.get_row(store.hash_ids[0]) == gold["row0"]
    assert store.get_hash_id_to_order() == gold["hash_id_to_order"]
    assert store.get_missing_string_hash_ids([chunks[0], "never seen"]) == gold["missing"]
    assert store.get_missing_string_hash_ids([]) == {} and store.insert_strings([]) is None
    e = store.get_embeddings(store.hash_ids[:3])
    assert list(e.shape) == gold["get_embeddings_shape"] and str(e.dtype) == gold["get_embeddings_dtype"]
    np.testing.assert_allclose(store.get_embedding(store.hash_ids[0])[:8], gold["embedding_row0_first8"], atol=1e-7)
    assert store.get_embeddings([]) == [] and store.get_rows([]) == {}
    assert len(store.embeddings) == 7 and store.embeddings[0].dtype == np.float32
    with pytest.raises(KeyError):
        store.get_row("chunk-doesnotexist")
    import pyarrow.parquet as pq
    schema = pq.read_schema(store.filename)
    assert {n: str(schema.field(n).type) for n in schema.names} == gold["parquet_schema"]
    # reload (checkpoint/resume path, embedding_store.py:92-107)
    again = EmbeddingStore(model, str(tmp_path / "chunk_embeddings"), 4, "chunk")
    assert again.get_all_ids() == gold["reloaded_hash_ids"]
    np.testing.assert_array_equal(again.get_embeddings(again.hash_ids), store.get_embeddings(store.hash_ids))
    assert again.text_to_hash_id[chunks[1]] == gold["hash_ids"][1]


def test_reference_written_parquet_is_readable(gold, tmp_path):
    """A parquet file in the reference's own layout (pandas: list<float> of per-row arrays) loads into the engine store."""
    import pandas as pd
    d = tmp_path / "s"
    d.mkdir()
    pd.DataFrame({"hash_id": gold["hash_ids"], "content": gold["texts"],
                  "embedding": [r for r in gold["embeddings"]]}).to_parquet(d / "vdb_chunk.parquet", index=False)
    store = EmbeddingStore(ReplayModel(gold), str(d), 4, "chunk")
    assert store.get_all_ids() == gold["hash_ids"]
    np.testing.assert_array_equal(store.get_embeddings(store.hash_ids), gold["embeddings"])


def test_id_scheme_and_in_call_dedup(tmp_path):
    assert compute_mdhash_id("abc", prefix="ns-") == "ns-900150983cd24fb0d6963f7d28e17f72"

    class M:
        embedding_dim = 4

        def batch_encode(self, texts, **kw):
            return np.arange(len(texts) * 4, dtype=np.float32).reshape(len(texts), 4)

    s = EmbeddingStore(M(), str(tmp_path / "x"), 2, "entity")
    s.insert_strings(["a", "b", "a"])          # duplicate inside one call collapses (dict keyed by id)
    assert s.texts == ["a", "b"] and len(s.embeddings) == 2


def test_append_only_persistence_roundtrip(gold, tmp_path):
    """Opt-in O(new rows) persistence: raw fp32/bf16 shards + jsonl; same lookups after reload; parquet on demand."""
    import types
    import torch
    model = ReplayModel(gold)
    model.global_config = types.SimpleNamespace(embedding_store_append_only=True)
    d = str(tmp_path / "chunk_embeddings")
    store = EmbeddingStore(model, d, 4, "chunk")
    store.insert_strings(gold["texts"][:4])
    store.insert_strings(gold["texts"][2:])            # appends only the 3 new rows
    assert not os.path.exists(store.filename)          # no parquet rewrite happened
    assert store.get_all_ids() == gold["hash_ids"]
    again = EmbeddingStore(model, d, 4, "chunk")
    assert again.get_all_ids() == gold["hash_ids"] and again.texts == gold["texts"]
    np.testing.assert_array_equal(again.get_embeddings(again.hash_ids), gold["embeddings"])
    assert again.text_to_hash_id[gold["texts"][3]] == gold["hash_ids"][3]
    shard = again.raw_shard_path()
    raw = torch.from_numpy(np.fromfile(shard, dtype=np.int16)).view(torch.bfloat16).view(7, 128)
    assert torch.equal(raw, torch.from_numpy(gold["embeddings"]).bfloat16())
    import pyarrow.parquet as pq
    schema = pq.read_schema(again.export_parquet())
    assert {n: str(schema.field(n).type) for n in schema.names} == gold["parquet_schema"]
    ref_style = EmbeddingStore(ReplayModel(gold), d, 4, "chunk")       # a parquet-mode store reads the export
    assert ref_style.get_all_ids() == gold["hash_ids"]


def test_interrupted_append_is_rolled_back(gold, tmp_path):
    """A crash between the data appends and the meta commit must not shift later rows: reload cuts the files back
    to what the meta file vouches for, and the next insert lands on the right offsets."""
    import types
    model = ReplayModel(gold)
    model.global_config = types.SimpleNamespace(embedding_store_append_only=True)
    d = str(tmp_path / "chunk_embeddings")
    store = EmbeddingStore(model, d, 4, "chunk")
    store.insert_strings(gold["texts"][:4])
    base = os.path.join(d, "vdb_chunk")
    with open(base + ".f32", "ab") as f:                 # half a row of an append that never committed
        f.write(b"\x01" * 100)
    with open(base + ".bf16", "ab") as f:
        f.write(b"\x02" * 300)
    with open(base + ".rows.jsonl", "a") as f:
        f.write('{"hash_id": "chunk-torn", "content": "torn')
    again = EmbeddingStore(model, d, 4, "chunk")
    assert again.get_all_ids() == gold["hash_ids"][:4]
    again.insert_strings(gold["texts"][4:])
    third = EmbeddingStore(model, d, 4, "chunk")
    assert third.get_all_ids() == gold["hash_ids"] and third.texts == gold["texts"]
    np.testing.assert_array_equal(third.get_embeddings(third.hash_ids), gold["embeddings"])
    assert third.raw_shard_path() is not None
    os.truncate(base + ".f32", 10)                        # data shorter than the commit record: refuse, do not guess
    with pytest.raises(ValueError):
        EmbeddingStore(model, d, 4, "chunk")


def test_parquet_is_written_in_row_batches(gold, tmp_path, monkeypatch):
    """Stores past 2^31 floats cannot be one list<float> array (int32 offsets); the writer batches rows.  Forced
    here with a tiny batch bound: same schema, same rows after reload."""
    monkeypatch.setattr(EmbeddingStore, "_PARQUET_MAX_VALUES", 2 * gold["embeddings"].shape[1] + 5)
    model = ReplayModel(gold)
    store = EmbeddingStore(model, str(tmp_path / "e"), 4, "chunk")
    store.insert_strings(gold["texts"])
    import pyarrow.parquet as pq
    f = pq.ParquetFile(store.filename)
    assert f.metadata.num_row_groups == 4 and f.metadata.num_rows == 7          # 2 + 2 + 2 + 1 rows
    schema = pq.read_schema(store.filename)
    assert {n: str(schema.field(n).type) for n in schema.names} == gold["parquet_schema"]
    again = EmbeddingStore(model, str(tmp_path / "e"), 4, "chunk")
    assert again.get_all_ids() == gold["hash_ids"]
    np.testing.assert_array_equal(again.get_embeddings(again.hash_ids), gold["embeddings"])


def test_reads_the_parquet_the_reference_wrote(gold, tmp_path):
    """tests/golden/vdb_chunk_reference.parquet was written by the REFERENCE's EmbeddingStore (pandas -> pyarrow,
    embedding_store.py:109-120; generator: tests/golden/make_golden_parquet.py).  Our store must load it as-is; the
    opposite direction (the reference loading our file) was checked where the reference can run and is recorded in
    tests/golden/parquet_interop.json."""
    import shutil
    d = tmp_path / "chunk_embeddings"
    d.mkdir()
    shutil.copy(os.path.join(HERE, "golden", "vdb_chunk_reference.parquet"), d / "vdb_chunk.parquet")
    model = ReplayModel(gold)
    store = EmbeddingStore(model, str(d), 4, "chunk")
    assert store.get_all_ids() == gold["hash_ids"] and store.texts == gold["texts"]
    np.testing.assert_array_equal(store.get_embeddings(store.hash_ids), gold["embeddings"])
    assert store.insert_strings(gold["texts"][:3]) == {} and model.calls == []        # nothing re-encoded
    store.insert_strings(["one more chunk"]) if "one more chunk" in model.table else None
    interop = json.load(open(os.path.join(HERE, "golden", "parquet_interop.json")))
    assert all(v is True for k, v in interop["reference_read_ours"].items() if isinstance(v, bool))
    import pyarrow.parquet as pq
    ref_schema = pq.read_schema(os.path.join(HERE, "golden", "vdb_chunk_reference.parquet"))
    assert {n: str(ref_schema.field(n).type) for n in ref_schema.names} == gold["parquet_schema"]


def test_switching_persistence_modes_keeps_every_row(tmp_path):
    """parquet -> append-only -> reload, and append-only -> parquet, over the same directory."""
    import types
    from comorag_b200.embedding_store import EmbeddingStore

    def model(append):
        m = types.SimpleNamespace(embedding_dim=8, global_config=types.SimpleNamespace(embedding_store_append_only=append))
        m.batch_encode = lambda texts, **kw: np.stack([np.full(8, float(len(t)), np.float32) for t in texts])
        return m

    d = str(tmp_path / "chunk_embeddings")
    s1 = EmbeddingStore(model(False), d, 4, "chunk")            # reference behaviour: whole-file parquet
    s1.insert_strings(["a", "bb", "ccc"])
    s2 = EmbeddingStore(model(True), d, 4, "chunk")             # append-only over the existing parquet
    assert s2.get_all_ids() == s1.get_all_ids()
    s2.insert_strings(["dddd", "eeeee"])
    s3 = EmbeddingStore(model(True), d, 4, "chunk")             # reload from the raw shards: all five rows
    assert len(s3.get_all_ids()) == 5 and s3.texts == ["a", "bb", "ccc", "dddd", "eeeee"]
    np.testing.assert_array_equal(s3.get_embeddings(s3.get_all_ids())[:, 0], [1, 2, 3, 4, 5])
    assert s3.raw_shard_path() is not None
    s4 = EmbeddingStore(model(False), d, 4, "chunk")            # back to parquet mode: must not lose the appended rows
    assert s4.texts == s3.texts
    s4.insert_strings(["ffffff"])
    s5 = EmbeddingStore(model(False), d, 4, "chunk")
    assert len(s5.get_all_ids()) == 6 and s5.get_missing_string_hash_ids(["a", "zz"]).keys() == {s5._nodes(["zz"]).popitem()[0]}
    s6 = EmbeddingStore(model(True), d, 4, "chunk")             # append mode again: the parquet now has one row more
    assert len(s6.get_all_ids()) == 6
    s6.insert_strings(["ggggggg"])
    s7 = EmbeddingStore(model(True), d, 4, "chunk")
    assert s7.texts[-2:] == ["ffffff", "ggggggg"] and len(s7.get_all_ids()) == 7 and s7.raw_shard_path() is not None


def test_sqlite_embedding_cache_round_trip(tmp_path):
    """make_cache_embed (reference base.py:112-187): rows keyed by (instruction, prompt, max_length); a second call is
    served from sqlite without touching the encoder, a changed instruction is a different key."""
    import torch
    from comorag_b200.embedding_model.base import make_cache_embed
    calls = []

    def encode(prompts, **kw):
        calls.append(list(prompts))
        return torch.tensor([[float(len(p)), float(len(kw.get("instruction", "")))] for p in prompts])

    cached = make_cache_embed(encode, str(tmp_path / "cache.db"), "cpu")
    a = cached(prompts=["x", "yy"], instruction="I:", max_length=16)
    b = cached(prompts=["yy", "zzz"], instruction="I:", max_length=16)
    assert calls == [["x", "yy"], ["zzz"]] and a.shape == (2, 2) and torch.equal(b[0], a[1])
    c = cached(prompts=["x"], instruction="other", max_length=16)
    assert calls[-1] == ["x"] and float(c[0, 1]) == 5.0
    again = make_cache_embed(encode, str(tmp_path / "cache.db"), "cpu")(prompts=["x", "yy", "zzz"], instruction="I:", max_length=16)
    assert len(calls) == 3 and again[:, 0].tolist() == [1.0, 2.0, 3.0]


def test_dspy_filter_can_drop_candidates(tmp_path):
    """the reference's LLM filter returns a subset, possibly empty (ComoRAG.py:486-488 then falls back to
    dense retrieval); the scorer-based filter does too once a threshold / keep-fraction is configured."""
    import types
    from comorag_b200.rerank import DSPyFilter
    emb = {"q": [1.0, 0.0], "a b c": [0.9, 0.1], "d e f": [0.2, 0.8], "g h i": [0.5, 0.5]}
    model = types.SimpleNamespace(batch_encode=lambda texts, **kw: np.array([emb[t] for t in texts], np.float32))
    items = [("a", "b", "c"), ("d", "e", "f"), ("g", "h", "i")]
    rag = lambda **cfg: types.SimpleNamespace(global_config=types.SimpleNamespace(**cfg), embedding_model=model)
    keep_all = DSPyFilter(rag())("q", items, [10, 11, 12], len_after_rerank=5)
    assert keep_all[0] == [10, 12, 11]
    thr = DSPyFilter(rag(rerank_score_threshold=0.4))("q", items, [10, 11, 12], len_after_rerank=5)
    assert thr[0] == [10, 12] and thr[1] == [items[0], items[2]]
    none = DSPyFilter(rag(rerank_score_threshold=2.0))("q", items, [10, 11, 12], len_after_rerank=5)
    assert none[0] == [] and none[1] == []
    half = DSPyFilter(rag(rerank_keep_fraction=0.5))("q", items, [10, 11, 12], len_after_rerank=5)
    assert half[0] == [10, 12]
