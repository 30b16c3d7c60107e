"""Generate the search golden fixtures by RUNNING THE REFERENCE's own retrieval code in the build container.

    python tests/golden/make_golden_search.py

The reference has no tests or golden vectors for this path, so the fixtures are
outputs of its own functions on seeded inputs:
  - misc_utils.min_max_normalize                          (misc_utils.py:141-150)
  - ComoRAG.get_fact_scores / dense_passage_retrieval     (ComoRAG.py:937-967), called as unbound methods
    on a stub `self` that only carries the attributes those methods read
  - embed_utils.get_similar_summaries                     (embed_utils.py:109-161) with a stub store/model
  - embed_utils.retrieve_knn                              (embed_utils.py:8-97)
Inputs are bf16-rounded unit vectors stored as fp32.  The float64 gaps between consecutive
ranks are stored next to the reference's ranking: ranks closer than fp32 summation noise cannot be pinned
by ANY fp32 implementation (numpy's own order depends on the BLAS kernel) and are compared as sets.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_harness import import_reference  # noqa: E402


def unit_rows(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    return x.bfloat16().float().numpy()


def main():
    import_reference()
    from src.comorag.ComoRAG import ComoRAG
    from src.comorag.utils import embed_utils
    from src.comorag.utils.misc_utils import min_max_normalize

    out = {}
    # --- min_max_normalize, including the constant-vector branch
    v = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    out["mm_in"] = v
    out["mm_out"] = min_max_normalize(v)
    out["mm_const_out"] = min_max_normalize(np.full(7, 0.25, dtype=np.float32))

    # --- dense_passage_retrieval / get_fact_scores on three corpus shapes
    for tag, (n, d, seed) in {"a": (5000, 64, 11), "b": (20000, 384, 12), "c": (3000, 1024, 13)}.items():
        E = unit_rows(n, d, seed)
        Q = unit_rows(8, d, 1000 + seed)
        # float64 gaps between consecutive ranks: ranks closer than fp32 summation noise are "near ties"
        # that the parity tests compare as sets (oracle.search_oracle.assert_topk_matches)
        top = -np.sort(-(E.astype(np.float64) @ Q.astype(np.float64).T), axis=0)[:129]
        out[f"dpr_{tag}_gaps"] = (-np.diff(top, axis=0)).T.copy()
        ids_all, scores_all, facts_all, summ_ids, summ_scores = [], [], [], [], []
        for qi in range(Q.shape[0]):
            q = Q[qi:qi + 1]  # batch_encode returns [1, D]
            stub = types.SimpleNamespace(query_to_embedding={"passage": {"q": q}, "triple": {"q": q}},
                                         passage_embeddings=E, summary_embeddings=E[: n // 2], fact_embeddings=E)
            sid, ssc = ComoRAG.dense_passage_retrieval(stub, "q")
            ids_all.append(sid[:128].astype(np.int64))
            scores_all.append(ssc[:128].astype(np.float32))
            s2, sc2 = ComoRAG.dense_passage_retrieval(stub, "q", need_cluster=True)
            summ_ids.append(s2[:50].astype(np.int64))
            summ_scores.append(sc2[:50].astype(np.float32))
            fs = ComoRAG.get_fact_scores(stub, "q")
            facts_all.append(np.argsort(fs)[-5:][::-1].astype(np.int64))  # ComoRAG.py:475
        out[f"dpr_{tag}_shape"] = np.array([n, d, seed])
        out[f"dpr_{tag}_Q"] = Q
        out[f"dpr_{tag}_ids"] = np.stack(ids_all)
        out[f"dpr_{tag}_scores"] = np.stack(scores_all)
        out[f"dpr_{tag}_summary_ids"] = np.stack(summ_ids)
        out[f"dpr_{tag}_summary_scores"] = np.stack(summ_scores)
        out[f"dpr_{tag}_fact_top5"] = np.stack(facts_all)

    # --- get_similar_summaries with a stub store / model
    n, d = 400, 128
    E = unit_rows(n, d, 21)
    q = unit_rows(1, d, 22)
    ids = [f"level_0-{i:04d}" for i in range(n)]
    store = types.SimpleNamespace(get_all_ids=lambda: list(ids), hash_id_to_text={h: f"summary {i}" for i, h in enumerate(ids)},
                                  get_embeddings=lambda keys: E[[ids.index(k) for k in keys]])
    model = types.SimpleNamespace(batch_encode=lambda text, **kw: q)
    texts, scores = embed_utils.get_similar_summaries("query", store, model, top_k=50)
    out["gss_E"], out["gss_q"] = E, q
    out["gss_idx"] = np.array([int(t.split()[1]) for t in texts], dtype=np.int64)
    out["gss_scores"] = np.array(scores, dtype=np.float32)

    # --- retrieve_knn (entity x entity, k=2047 in the reference; small blocks here to exercise the 2-stage merge)
    nq, nk, d, k = 300, 2500, 64, 100
    Qv, Kv = unit_rows(nq, d, 31), unit_rows(nk, d, 32)
    res = embed_utils.retrieve_knn([f"q{i}" for i in range(nq)], [f"k{i}" for i in range(nk)], Qv, Kv, k=k,
                                   query_batch_size=128, key_batch_size=1000)
    out["knn_Q"], out["knn_K"] = Qv, Kv
    out["knn_ids"] = np.array([[int(x[1:]) for x in res[f"q{i}"][0]] for i in range(nq)], dtype=np.int64)
    out["knn_scores"] = np.array([res[f"q{i}"][1] for i in range(nq)], dtype=np.float32)

    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), **out)
    print("wrote search_golden.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
