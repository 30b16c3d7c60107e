"""Host-side logic that needs no GPU: config bag, factory dispatch, score normalisation, shard bounds, packing."""
import numpy as np
import pytest
import torch

from comorag_b200.dist import merge_partials_reference, pack_partial, shard_bounds, unpack_partials
from comorag_b200.retrieval import min_max_normalize, normalize_topk_scores
from oracle import search_oracle as so


def test_embedding_config_bag():
    from comorag_b200.embedding_model.base import EmbeddingConfig
    c = EmbeddingConfig.from_dict({"norm": True, "encode_params": {"max_length": 512}})
    assert c.norm is True and c["encode_params"]["max_length"] == 512 and "norm" in c
    c.extra = 3
    assert c.to_dict()["extra"] == 3
    with pytest.raises(AttributeError):
        c.missing
    with pytest.raises(KeyError):
        c["missing"]
    del c["extra"]
    assert "extra" not in c


def test_normalize_topk_scores_reproduces_min_max_of_all_scores():
    rng = np.random.default_rng(0)
    all_scores = rng.standard_normal((3, 1000)).astype(np.float32)
    top = np.argsort(-all_scores, axis=1)[:, :10]
    raw = np.take_along_axis(all_scores, top, axis=1)
    minmax = np.stack([all_scores.min(1), all_scores.max(1)], 1)
    want = np.stack([so.min_max_normalize(all_scores[i])[top[i]] for i in range(3)])
    np.testing.assert_allclose(normalize_topk_scores(raw, minmax), want, atol=1e-6)
    # zero range -> all ones, as the reference (misc_utils.py:147-148)
    np.testing.assert_array_equal(normalize_topk_scores(np.full((1, 4), 0.5, np.float32), np.array([[0.5, 0.5]], np.float32)),
                                  np.ones((1, 4), np.float32))
    np.testing.assert_array_equal(min_max_normalize(np.array([2.0, 2.0])), np.ones(2))


def test_shard_bounds_cover_rows_exactly():
    for n, w in [(10_000_000, 8), (7, 3), (5, 8), (0, 2), (1024, 1)]:
        o = shard_bounds(n, w)
        assert o[0] == 0 and o[-1] == n and len(o) == w + 1
        sizes = [o[i + 1] - o[i] for i in range(w)]
        assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


def test_pack_unpack_roundtrip_and_merge_rule():
    g = torch.Generator().manual_seed(0)
    world, nq, k = 3, 5, 4
    scores = torch.randn(world, nq, k, generator=g).sort(dim=2, descending=True).values
    ids = torch.arange(world * nq * k).view(world, nq, k)
    ids[2, :, 3] = -1
    scores[2, :, 3] = float("-inf")
    mm = torch.randn(world, nq, 2, generator=g)
    buf = torch.cat([pack_partial(ids[r], scores[r], mm[r]) for r in range(world)])
    i2, s2, m2 = unpack_partials(buf, world, nq, k)
    assert torch.equal(i2, ids) and torch.equal(s2, scores) and torch.equal(m2, mm)
    oi, os_, om = merge_partials_reference(ids, scores, mm, k)
    flat_s = scores.permute(1, 0, 2).reshape(nq, -1)
    assert torch.equal(os_, flat_s.sort(dim=1, descending=True).values[:, :k])
    assert torch.equal(om[:, 0], mm[..., 0].min(0).values) and torch.equal(om[:, 1], mm[..., 1].max(0).values)
    assert (oi >= 0).all()


def test_rerank_surface_keeps_reference_contract():
    """DSPyFilter(narrtiverag)(query, candidate_items, candidate_indices, len_after_rerank) -> (indices, items, dict)
    (rerank.py:97-123); ranking rule: dense similarity of "s p o" to the query, ties by candidate position."""
    import types
    from comorag_b200.rerank import DSPyFilter

    vocab = {"cinderella": 0, "prince": 1, "slipper": 2, "pumpkin": 3}

    class Model:
        def batch_encode(self, texts, **kw):
            out = np.zeros((len(texts), 4), np.float32)
            for i, t in enumerate(texts):
                for w in t.lower().split():
                    if w in vocab:
                        out[i, vocab[w]] += 1
            return out

    rag = types.SimpleNamespace(global_config=types.SimpleNamespace(), embedding_model=Model())
    f = DSPyFilter(rag)
    items = [("pumpkin", "became", "coach"), ("prince", "found", "slipper"), ("cinderella", "lost", "slipper")]
    idx, kept, meta = f("who found the slipper prince", items, [10, 11, 12], len_after_rerank=2)
    assert idx == [11, 12] and kept == [items[1], items[2]] and len(meta["confidence"]) == 2
    assert f.rerank("x", [], [], 3) == ([], [], {"confidence": None})
    idx_all, _, _ = f("slipper", items, [0, 1, 2])          # len_after_rerank=None keeps everything
    assert idx_all == [1, 2, 0]                              # tie between items 1 and 2 -> candidate order


def test_full_score_contracts_on_a_fake_index():
    """get_fact_scores / dense_passage_retrieval(top_k=None) rebuild the reference's full-array contracts from the
    ranked output of DenseIndex.search (checked against the numpy oracle with a CPU stand-in for the index)."""
    from comorag_b200 import retrieval as rt
    rng = np.random.default_rng(4)
    E = rng.standard_normal((300, 16)).astype(np.float32)
    q = rng.standard_normal((1, 16)).astype(np.float32)

    class FakeIndex:
        n_rows = 300

        def search(self, queries, k):
            sc = np.asarray(queries, np.float32) @ E.T
            ids = np.argsort(-sc, axis=1, kind="stable")[:, :k]
            return ids, np.take_along_axis(sc, ids, 1), np.stack([sc.min(1), sc.max(1)], 1)

        # CPU stand-ins for the device entry points of the full-array contracts (score-all pass + device ranking)
        def prepare_queries(self, queries):
            import torch
            return torch.as_tensor(np.asarray(queries, np.float32))

        def scores_device(self, q):
            import torch
            sc = q @ torch.from_numpy(E).T
            return sc, torch.stack([sc.min(1).values, sc.max(1).values], 1)

        def rank_device(self, row):
            import torch
            order = torch.argsort(row, descending=True, stable=True)
            return order, row[order]

    ids, scores = rt.dense_passage_retrieval(FakeIndex(), q)
    want_ids, want_scores = so.dense_passage_retrieval(E, q)
    np.testing.assert_array_equal(ids, want_ids)
    np.testing.assert_allclose(scores, want_scores, atol=1e-6)
    np.testing.assert_allclose(rt.get_fact_scores(FakeIndex(), q), so.fact_scores(E, q), atol=1e-6)
    top_ids, top_sc = rt.get_fact_scores_topk(FakeIndex(), q, 5)
    np.testing.assert_array_equal(top_ids, so.top_facts(so.fact_scores(E, q), 5))


def test_cross_encoder_reranker_host_logic():
    """Pair tokenisation, budget-cut launches over length-sorted pairs, un-permutation and the DSPyFilter return
    convention -- with a CPU stand-in for the device encoder (the arithmetic is covered by the -m gpu tests)."""
    import os
    import torch
    from transformers import AutoTokenizer
    from comorag_b200.rerank import CrossEncoderReranker, DSPyFilter
    tok = AutoTokenizer.from_pretrained(os.path.join(os.path.dirname(__file__), "golden", "bge-tiny-synth"))

    class FakeEncoder:
        n_labels = 1
        launches = []

        class config:
            max_position_embeddings, position_offset = 64, 0

        def classify_token_lists(self, seqs):
            self.launches.append([len(s) for s in seqs])
            return torch.tensor([[float(sum(s) % 97) + 0.001 * len(s)] for s in seqs])

    enc = FakeEncoder()
    rr = CrossEncoderReranker("unused", encoder=enc, tokenizer=tok, max_length=512, token_budget=40)
    assert rr.max_length == 64                                   # clamped to the position table
    pairs = [("who lost a slipper", "cinderella lost her glass slipper at the ball " * (i % 4 + 1)) for i in range(9)]
    ids = tok([q for q, _ in pairs], [p for _, p in pairs], truncation=True, max_length=64)["input_ids"]
    want = np.array([float(sum(s) % 97) + 0.001 * len(s) for s in ids], dtype=np.float32)
    got = rr.score(pairs)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    assert len(enc.launches) > 1 and all(sum(l) <= 40 or len(l) == 1 for l in enc.launches)
    assert all(max(l) <= 64 for l in enc.launches)
    assert rr.score([]).shape == (0,)
    order, sc = rr.rerank_passages("who lost a slipper", [p for _, p in pairs], top_k=3)
    assert list(order) == list(np.argsort(-want.astype(np.float64), kind="stable")[:3])

    class Host:
        global_config = type("Cfg", (), {})()
        embedding_model = None

    flt = DSPyFilter(Host())
    flt.cross_encoder = rr
    items = [("cinderella", "lost", "slipper"), ("prince", "found", "slipper"), ("fairy", "made", "coach")]
    idx, kept, meta = flt("who lost a slipper", items, [7, 8, 9], 2)
    s = rr.score([("who lost a slipper", " ".join(t)) for t in items])
    o = np.argsort(-s.astype(np.float64), kind="stable")[:2]
    assert idx == [[7, 8, 9][i] for i in o] and kept == [items[i] for i in o] and len(meta["confidence"]) == 2
    assert flt("q", [], [], 5) == ([], [], {"confidence": None})


def test_ivf_layout_is_list_major_and_tile_aligned():
    """comorag_b200.ivf.ivf_layout against the oracle's grouping: same (list, id) order, every list starting on a
    128-row tile, destinations unique, empty lists taking no tile."""
    import torch
    from comorag_b200.ivf import TILE_ROWS, ivf_layout
    from oracle import ivf_oracle as ivf
    rng = np.random.default_rng(3)
    n, nlist = 5000, 37
    a = rng.integers(0, nlist, n)
    a[a == 5] = 6                                                   # list 5 is empty
    a[:300] = 9                                                     # list 9 spans several tiles
    order, dest, tile_start, list_rows = ivf_layout(torch.from_numpy(a), nlist)
    order, dest, tile_start, list_rows = order.numpy(), dest.numpy(), tile_start.numpy(), list_rows.numpy()
    np.testing.assert_array_equal(order, np.lexsort((np.arange(n), a)))
    np.testing.assert_array_equal(list_rows, np.bincount(a, minlength=nlist))
    assert tile_start[0] == 0 and tile_start[6] == tile_start[5]    # the empty list owns no tile
    np.testing.assert_array_equal(np.diff(tile_start), (list_rows + TILE_ROWS - 1) // TILE_ROWS)
    assert len(set(dest.tolist())) == n and dest.max() < tile_start[-1] * TILE_ROWS
    for l in range(nlist):
        mine = dest[a[order] == l]
        np.testing.assert_array_equal(mine, tile_start[l] * TILE_ROWS + np.arange(list_rows[l]))
    # the oracle's back-to-back layout is the same order without the padding
    x = rng.standard_normal((n, 8)).astype(np.float32)
    c = rng.standard_normal((nlist, 8)).astype(np.float32)
    L = ivf.IVFLists(x, c, assignment=a)
    np.testing.assert_array_equal(L.ids, order)
    with pytest.raises(ValueError):
        ivf_layout(torch.tensor([0, 99]), 4)


def test_shard_bounds_and_packed_records_properties():
    """Property checks (hypothesis) of the host-side index arithmetic the multi-GPU path relies on."""
    from hypothesis import given, settings, strategies as st
    from comorag_b200.dist import pack_partial, shard_bounds, unpack_partials
    from comorag_b200.index import packed_record_bytes, packed_views

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 10**9), st.integers(1, 64))
    def bounds(n, world):
        b = shard_bounds(n, world)
        sizes = np.diff(b)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1 and np.all(np.diff(sizes) <= 0)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 128), st.integers(1, 5))
    def records(nq, k, world):
        per = packed_record_bytes(nq, k)
        assert per % 16 == 0 and per >= nq * k * 12 + nq * 8
        g = torch.Generator().manual_seed(nq * 1000 + k)
        parts = []
        for r in range(world):
            ids = torch.randint(-1, 10**12, (nq, k), generator=g)
            sc = torch.randn(nq, k, generator=g)
            mm = torch.randn(nq, 2, generator=g)
            buf = pack_partial(ids, sc, mm)
            assert buf.numel() == per
            v = packed_views(buf, nq, k)
            assert torch.equal(v[0], ids) and torch.equal(v[1], sc) and torch.equal(v[2], mm)
            parts.append((ids, sc, mm, buf))
        gi, gs, gm = unpack_partials(torch.cat([p[3] for p in parts]), world, nq, k)
        for r, (ids, sc, mm, _) in enumerate(parts):
            assert torch.equal(gi[r], ids) and torch.equal(gs[r], sc) and torch.equal(gm[r], mm)

    bounds()
    records()


def test_embedding_cache_classmethods():
    """base.py:222-260's EmbeddingCache API (get / set / contains / clear), including from several threads."""
    import threading
    from comorag_b200.embedding_model import EmbeddingCache
    EmbeddingCache.clear()
    assert EmbeddingCache.get("a") is None and not EmbeddingCache.contains("a")
    v = np.arange(4, dtype=np.float32)
    EmbeddingCache.set("a", v)
    assert EmbeddingCache.contains("a") and np.array_equal(EmbeddingCache.get("a"), v)
    ts = [threading.Thread(target=lambda i=i: EmbeddingCache.set(f"k{i}", i)) for i in range(16)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(EmbeddingCache.get(f"k{i}") == i for i in range(16))
    EmbeddingCache.clear()
    assert not EmbeddingCache.contains("a") and not EmbeddingCache.contains("k3")


def test_token_flattening_matches_the_plain_loops():
    from comorag_b200.encoder import _flatten
    rng = np.random.default_rng(1)
    seqs = [rng.integers(0, 250002, int(n)).tolist() for n in (1, 512, 37, 2, 300)]
    seqs[2] = tuple(seqs[2])                                           # any sequence type
    flat, cu, longest = _flatten(seqs)
    assert flat.dtype == np.int32 and cu.dtype == np.int32 and longest == 512
    np.testing.assert_array_equal(flat, np.array([t for s in seqs for t in s], dtype=np.int32))
    np.testing.assert_array_equal(cu, np.concatenate([[0], np.cumsum([len(s) for s in seqs])]))
    with pytest.raises(ValueError):
        _flatten([[1, 2], []])
    with pytest.raises(ValueError):
        _flatten([])


def test_retrieval_wave_shares_one_pass_per_shard_between_concurrent_tri_retrieves():
    """SURVEY 8f item 1 (host logic, CPU stand-ins for the device entry points): 12 threads run the retrieval calls of
    tri_retrieve (ComoRAG.py:470-531) for 24 queries; every caller gets exactly what an un-coalesced call returns,
    while encodes and shard passes are shared."""
    import threading
    import time
    import types
    from concurrent.futures import ThreadPoolExecutor

    import torch
    from comorag_b200 import comorag_methods as cm
    from comorag_b200 import retrieval as rt

    rng = np.random.default_rng(3)
    D = 16
    mats = {name: rng.standard_normal((n, D)).astype(np.float32) for name, n in (("fact", 40), ("chunk", 9), ("summary", 5), ("level", 4))}
    passes = {name: 0 for name in mats}
    encodes = []

    class FakeIndex:
        def __init__(self, name):
            self.name, self.E = name, torch.from_numpy(mats[name])

        def prepare_queries(self, q):
            return torch.as_tensor(np.asarray(q, np.float32))

        def scores_device(self, q):
            passes[self.name] += 1
            time.sleep(0.002)
            sc = q @ self.E.T
            return sc, torch.stack([sc.min(1).values, sc.max(1).values], 1)

        def rank_device(self, row):
            order = torch.argsort(row, descending=True, stable=True)
            return order, row[order]

    class FakeStore:
        def __init__(self, name):
            self.index = FakeIndex(name)
            self.hash_ids = [f"{name}-{i}" for i in range(mats[name].shape[0])]
            self.texts = [f"{name} text {i}" for i in range(mats[name].shape[0])]
            self._dim = D

        def get_all_ids(self):
            return list(self.hash_ids)

        def search(self, q, k):
            passes[self.index.name] += 1
            sc = np.asarray(q, np.float32) @ mats[self.index.name].T
            ids = np.argsort(-sc, axis=1, kind="stable")[:, :k]
            return ids, np.take_along_axis(sc, ids, 1), np.stack([sc.min(1), sc.max(1)], 1)

    class FakeModel:
        def batch_encode(self, texts, **kw):
            texts = [texts] if isinstance(texts, str) else list(texts)
            encodes.append(len(texts))
            time.sleep(0.001)
            return np.stack([np.random.default_rng(abs(hash(t)) % (2 ** 32)).standard_normal(D).astype(np.float32) for t in texts])

    stores = {n: FakeStore(n) for n in mats}
    rag = types.SimpleNamespace(
        global_config=types.SimpleNamespace(need_cluster=True, qa_epi_top_k=3), embedding_model=FakeModel(),
        fact_embeddings=cm.ShardMatrix(stores["fact"]), passage_embeddings=cm.ShardMatrix(stores["chunk"]),
        summary_embeddings=cm.ShardMatrix(stores["summary"]), level_store=stores["level"],
        fact_node_keys=stores["fact"].hash_ids, passage_node_keys=stores["chunk"].hash_ids,
        query_to_embedding={"triple": {}, "passage": {}})
    queries = [f"probe number {i}" for i in range(24)]

    def tri_retrieve(q):
        cm.get_query_embeddings(rag, q)
        facts = cm.get_fact_scores(rag, q)
        ver = cm.dense_passage_retrieval(rag, q)
        sem = cm.dense_passage_retrieval(rag, q, need_cluster=True)
        epi = rt.get_similar_summaries(q, rag.level_store, rag.embedding_model, top_k=3)
        return facts, ver, sem, epi

    with ThreadPoolExecutor(12) as ex:
        got = list(ex.map(tri_retrieve, queries))
    stats = rag._crag_wave.stats
    rag._crag_wave.close()
    assert stats["queries"] == 24 and stats["waves"] < 24
    assert sum(encodes) == 24 and len(encodes) == stats["waves"]             # one packed encode per wave, nothing else
    assert passes["fact"] == passes["chunk"] == passes["summary"] == passes["level"] == stats["waves"]
    for q, (facts, ver, sem, epi) in zip(queries, got):
        e = FakeModel().batch_encode(q)
        np.testing.assert_allclose(facts, so.fact_scores(mats["fact"], e), atol=1e-6)
        for (ids, sc), name in ((ver, "chunk"), (sem, "summary")):
            w_ids, w_sc = so.dense_passage_retrieval(mats[name], e)
            np.testing.assert_array_equal(ids, w_ids)
            np.testing.assert_allclose(sc, w_sc, atol=1e-6)
        raw = (e @ mats["level"].T)[0]
        want = np.argsort(-raw, kind="stable")[:3]
        assert epi[0] == [f"level text {i}" for i in want]
    # a timeline row added after the wave ran: the parked result is stale and the search runs again
    before = passes["level"]
    mats["level"] = np.concatenate([mats["level"], rng.standard_normal((1, D)).astype(np.float32)])
    stores["level"].hash_ids.append("level-4")
    stores["level"].texts.append("level text 4")
    rt.get_similar_summaries(queries[0], rag.level_store, rag.embedding_model, top_k=3)
    assert passes["level"] == before + 1


def test_prepare_retrieval_objects_from_many_threads_keeps_the_wave_and_refreshes_on_growth():
    """ComoRAG.py:467-468 is reached by up to 16 threads at once: duplicate calls must not retire the retrieval wave
    the earlier threads are already using; a call after the stores grew must rebuild and retire it."""
    import types
    from concurrent.futures import ThreadPoolExecutor

    from comorag_b200 import comorag_methods as cm

    class Store:
        def __init__(self, name, n):
            self.hash_ids = [f"{name}-{i}" for i in range(n)]
            self._dim = 8
            self.uploads = 0

        def get_all_ids(self):
            return list(self.hash_ids)

        @property
        def index(self):
            self.uploads += 1
            return object()

        def search(self, q, k):
            raise AssertionError("not used here")

    ent, chunk, fact = Store("entity", 5), Store("chunk", 3), Store("fact", 7)
    graph = types.SimpleNamespace(vs=[{"name": n} for n in ent.hash_ids + chunk.hash_ids])
    rag = types.SimpleNamespace(global_config=types.SimpleNamespace(need_cluster=False), graph=graph,
                                entity_embedding_store=ent, ver_embedding_store=chunk, fact_embedding_store=fact,
                                ready_to_retrieve=False)
    cm.prepare_retrieval_objects(rag)
    assert rag.ready_to_retrieve and rag.fact_node_keys == fact.hash_ids and rag.passage_node_idxs == [5, 6, 7]
    assert rag.fact_embeddings.shape == (7, 8) and rag.entity_embeddings.shape == (5, 8)
    closed = []
    rag._crag_wave = types.SimpleNamespace(close=lambda: closed.append(1))
    rag.query_to_embedding["triple"]["q"] = "cached"
    with ThreadPoolExecutor(8) as ex:                    # the late duplicates of the first tri_retrieve wave
        list(ex.map(lambda _: cm.prepare_retrieval_objects(rag), range(16)))
    assert not closed and rag._crag_wave is not None and rag.query_to_embedding["triple"] == {"q": "cached"}
    fact.hash_ids.append("fact-7")                       # index() added rows: the next call rebuilds
    cm.prepare_retrieval_objects(rag)
    assert closed == [1] and rag._crag_wave is None and len(rag.fact_node_keys) == 8 and rag.query_to_embedding["triple"] == {}
    with pytest.raises(TypeError):                       # a reference store (no device shard) is refused, not papered over
        rag.fact_embedding_store = types.SimpleNamespace(hash_ids=[], get_all_ids=lambda: [])
        cm.prepare_retrieval_objects(rag)
