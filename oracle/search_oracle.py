"""CPU restatement (numpy) of the reference's index-side similarity search.

TEST INFRASTRUCTURE ONLY: the product path never imports this module.

Each function cites the reference code it restates (paths relative to the
reference tree, EternityJune25/ComoRAG @ a4f84337).  The reference ships no
tests or golden vectors for this path, so the restatement
is pinned against outputs of the reference's own modules executed in the build
container: tests/golden/make_golden_search.py imports
src/comorag/utils/misc_utils.min_max_normalize, embed_utils.get_similar_summaries
and embed_utils.retrieve_knn and stores their outputs as fixtures
(tests/golden/search_*.npz), which tests/test_oracle_search.py replays against
this file.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def min_max_normalize(x: np.ndarray) -> np.ndarray:
    """misc_utils.py:141-150 (twin: embed_utils.py:99-107): (x-min)/(max-min), all ones if the range is 0."""
    min_val = np.min(x)
    max_val = np.max(x)
    range_val = max_val - min_val
    if range_val == 0:
        return np.ones_like(x)
    return (x - min_val) / range_val


def query_scores(matrix: np.ndarray, query_embedding: np.ndarray) -> np.ndarray:
    """ComoRAG.py:944-945 / :958-962 / embed_utils.py:153-154: np.dot(E, q.T) squeezed to [N]."""
    s = np.dot(matrix, query_embedding.T)
    return np.squeeze(s) if s.ndim == 2 else s


def dense_passage_retrieval(matrix: np.ndarray, query_embedding: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """ComoRAG.py:950-967: dot -> min-max -> full descending argsort; returns (sorted ids, sorted scores)."""
    scores = min_max_normalize(query_scores(matrix, query_embedding))
    sorted_doc_ids = np.argsort(scores)[::-1]
    return sorted_doc_ids, scores[sorted_doc_ids.tolist()]


def fact_scores(matrix: np.ndarray, query_embedding: np.ndarray) -> np.ndarray:
    """ComoRAG.py:937-948: min-max-normalised dot of every fact row with the query."""
    return min_max_normalize(query_scores(matrix, query_embedding))


def top_facts(scores: np.ndarray, link_top_k: int) -> np.ndarray:
    """ComoRAG.py:475 / :1073: np.argsort(scores)[-k:][::-1]."""
    return np.argsort(scores)[-link_top_k:][::-1]


def similar_summaries(matrix: np.ndarray, query_embedding: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """embed_utils.py:153-160: dot -> min-max -> argsort[::-1][:top_k]; returns (indices, normalised scores)."""
    scores = min_max_normalize(query_scores(matrix, query_embedding))
    idx = np.argsort(scores)[::-1][:top_k]
    return idx, scores[idx]


# ---------------------------------------------------------------------------
# Batched form used by the parity tests: the same arithmetic for a block of
# queries, with the deterministic tie rule the engine documents (score
# descending, then row id ascending) and float64 accumulation so that the
# ground-truth order does not depend on a BLAS summation order.
def topk_exact(matrix: np.ndarray, queries: np.ndarray, k: int, row_offset: int = 0,
               chunk: int = 262144):
    """Returns (ids int64 [nq,k], scores float64 [nq,k], minmax float64 [nq,2], gaps float64 [nq,k]).

    gaps[q, j] = score of rank j minus score of rank j+1 (inf if there is no
    rank j+1): ranks whose gap is below the fp32 accumulation noise are
    compared as sets by the tests ("near ties").  Missing entries (N < k) have
    id -1 and score -inf, as the engine writes them.
    """
    matrix = np.asarray(matrix)
    queries = np.asarray(queries, dtype=np.float64)
    n = matrix.shape[0]
    nq = queries.shape[0]
    kk = min(k + 1, n)
    best_s = np.full((nq, 0), -np.inf)
    best_i = np.zeros((nq, 0), dtype=np.int64)
    mn = np.full(nq, np.inf)
    mx = np.full(nq, -np.inf)
    for s0 in range(0, n, chunk):
        blk = matrix[s0:s0 + chunk].astype(np.float64)
        sc = queries @ blk.T  # [nq, c]
        mn = np.minimum(mn, sc.min(axis=1))
        mx = np.maximum(mx, sc.max(axis=1))
        ids = np.broadcast_to(np.arange(s0, s0 + blk.shape[0], dtype=np.int64), sc.shape)
        cs = np.concatenate([best_s, sc], axis=1)
        ci = np.concatenate([best_i, ids], axis=1)
        # order: score desc, id asc  (lexsort: last key is primary)
        order = np.stack([np.lexsort((ci[q], -cs[q]))[:kk] for q in range(nq)])
        best_s = np.take_along_axis(cs, order, axis=1)
        best_i = np.take_along_axis(ci, order, axis=1)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    out_s = np.full((nq, k), -np.inf)
    gaps = np.full((nq, k), np.inf)
    m = min(k, n)
    out_i[:, :m] = best_i[:, :m] + row_offset
    out_s[:, :m] = best_s[:, :m]
    if n > 0:
        d = best_s[:, :-1] - best_s[:, 1:]
        gaps[:, :d.shape[1]][:, :k] = d[:, :k]
    return out_i, out_s, np.stack([mn, mx], axis=1), gaps


def assert_topk_matches(got_ids: np.ndarray, got_scores: np.ndarray, want_ids: np.ndarray, want_scores: np.ndarray,
                        gaps: np.ndarray, score_tol: float = 1e-3, tie_tol: float = 2e-6) -> None:
    """Bit-exact id/rank parity, except that a run of ranks separated by gaps < tie_tol
    (indistinguishable under fp32 accumulation) may appear in any order / be swapped
    with the (k+1)-th candidate.  Scores must agree within score_tol."""
    nq, k = want_ids.shape
    assert got_ids.shape == want_ids.shape, (got_ids.shape, want_ids.shape)
    for q in range(nq):
        j = 0
        while j < k:
            e = j
            while e < k and gaps[q, e] < tie_tol:
                e += 1
            # ranks j..e form one near-tie group; if e == k the group extends past rank k
            hi = min(e, k - 1)
            if e < k:
                a, b = set(got_ids[q, j:hi + 1].tolist()), set(want_ids[q, j:hi + 1].tolist())
                assert a == b, f"query {q} ranks {j}..{hi}: got {sorted(a)} want {sorted(b)}"
            else:
                # open-ended group: every returned id must have a score within tie_tol of the group
                lo_score = want_scores[q, j] + tie_tol
                assert np.all(got_scores[q, j:k] <= lo_score + score_tol), f"query {q} ranks {j}..: scores too high"
            j = hi + 1
        valid = want_ids[q] >= 0
        assert np.allclose(got_scores[q][valid], want_scores[q][valid], atol=score_tol, rtol=0), (
            f"query {q}: scores differ by {np.abs(got_scores[q][valid] - want_scores[q][valid]).max()}")
        assert np.all(got_ids[q][~valid] == -1)
