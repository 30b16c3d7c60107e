"""`DSPyFilter` call surface (reference: src/comorag/rerank.py:15-123).

The reference's filter is an LLM few-shot prompt over <= linking_top_k fact triples -- there is no arithmetic to
port (component 5).  This class keeps the constructor and `__call__`/`rerank` signature and
return convention `(sorted_indices[:k], sorted_items[:k], {'confidence': ...})` and ranks the candidate triples by
dense similarity to the query with the engine's own encoder.  PARITY UNPINNED: the reference has no
implementation to compare with; the ordering rule is ours (score descending, candidate position ascending).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


class DSPyFilter:
    def __init__(self, narrtiverag):
        """Same single argument as the reference (rerank.py:16): the ComoRAG instance (needs .global_config and
        .embedding_model; the LLM handle the reference reads is not used)."""
        self.global_config = narrtiverag.global_config
        self.embedding_model = narrtiverag.embedding_model

    def __call__(self, *args, **kwargs):
        return self.rerank(*args, **kwargs)

    def rerank(self, query: str, candidate_items: List[Tuple], candidate_indices: List[int],
               len_after_rerank: Optional[int] = None) -> Tuple[List[int], List[Tuple], dict]:
        if len(candidate_items) == 0:
            return [], [], {"confidence": None}
        texts = [" ".join(str(x) for x in item) for item in candidate_items]
        emb = self.embedding_model.batch_encode([query] + texts)
        scores = emb[1:] @ emb[0]
        order = np.argsort(-scores, kind="stable")
        idx = [candidate_indices[i] for i in order][:len_after_rerank]
        items = [candidate_items[i] for i in order][:len_after_rerank]
        return idx, items, {"confidence": [float(scores[i]) for i in order][:len_after_rerank]}
