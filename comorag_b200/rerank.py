"""`DSPyFilter` call surface (reference: src/comorag/rerank.py:15-123) and the cross-encoder behind it.

The reference's filter is an LLM few-shot prompt over <= linking_top_k fact triples -- there is no arithmetic to
port (component 5).  `DSPyFilter` keeps the constructor and `__call__`/`rerank` signature and
return convention `(sorted_indices[:k], sorted_items[:k], {'confidence': ...})` and scores the candidates on the
GPU instead:

* with `global_config.rerank_model_name` set to a sequence-classification checkpoint directory (bge-reranker-*:
  XLM-R encoder + classification head) -> `CrossEncoderReranker`: every (query, candidate) pair runs through the
  engine's encoder kernels and the classification head (`crag_encoder_classify`), BASELINE config 5;
* otherwise -> dense similarity of the candidate text to the query with the engine's embedding model.

PARITY UNPINNED against the reference (it has no implementation to compare with); the cross-encoder arithmetic is
pinned against `transformers`' XLMRobertaForSequenceClassification instead (oracle/encoder_oracle.py
`classifier_logits`, tests/test_oracle_encoder.py).  Ordering rule (ours): score descending, candidate position
ascending on ties.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .config import cfg_get


def _order(scores: np.ndarray) -> np.ndarray:
    return np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable")


class CrossEncoderReranker:
    """(query, passage) -> relevance logit with a BERT-family sequence-classification checkpoint on one GPU."""

    def __init__(self, model_path: str, device=None, max_length: int = 512, token_budget: int = 16384,
                 encoder=None, tokenizer=None):
        if encoder is None:
            from .encoder import BertEncoderB200
            encoder = BertEncoderB200.from_pretrained(model_path, device)
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(model_path)
        if getattr(encoder, "n_labels", 0) < 1:
            raise ValueError(f"{model_path!r} has no classifier head (classifier.dense / classifier.out_proj)")
        self.encoder, self.tokenizer = encoder, tokenizer
        cfg = encoder.config
        self.max_length = min(int(max_length), cfg.max_position_embeddings - cfg.position_offset)
        self.token_budget = int(token_budget)
        self._tok_lock = threading.Lock()

    def _tokenize(self, pairs: Sequence[Tuple[str, str]]) -> List[List[int]]:
        # <s> query </s></s> passage </s> for XLM-R ([CLS] q [SEP] p [SEP] for BERT); over-long pairs lose tokens
        # from the longer side first (truncation=True, as FlagEmbedding's reranker recipe tokenises)
        with self._tok_lock:
            enc = self.tokenizer([q for q, _ in pairs], [p for _, p in pairs], padding=False, truncation=True,
                                 max_length=self.max_length)
        return enc["input_ids"]

    def score_token_lists(self, ids: Sequence[Sequence[int]]) -> np.ndarray:
        """Logits [n, n_labels] for already-tokenised pairs; launches are cut by packed-token budget over the
        length-sorted pairs (rows are independent, so batching never changes a row's value)."""
        n = len(ids)
        out = np.empty((n, self.encoder.n_labels), dtype=np.float32)
        if n == 0:
            return out
        order = sorted(range(n), key=lambda i: len(ids[i]))
        parts, start, tokens = [], 0, 0
        for pos, i in enumerate(order):
            if pos > start and tokens + len(ids[i]) > self.token_budget:
                parts.append(self.encoder.classify_token_lists([ids[j] for j in order[start:pos]]))
                start, tokens = pos, 0
            tokens += len(ids[i])
        parts.append(self.encoder.classify_token_lists([ids[j] for j in order[start:]]))
        import torch
        got = (parts[0] if len(parts) == 1 else torch.cat(parts, 0)).cpu().numpy()   # one D2H of n*n_labels floats
        out[np.asarray(order)] = got
        return out

    def score(self, pairs: Sequence[Tuple[str, str]]) -> np.ndarray:
        """Relevance logit per pair (label 0), float32 [n]."""
        if len(pairs) == 0:
            return np.empty((0,), dtype=np.float32)
        return self.score_token_lists(self._tokenize(pairs))[:, 0]

    def rerank_passages(self, query: str, passages: Sequence[str], top_k: Optional[int] = None):
        """-> (positions into `passages` best first, their logits)."""
        scores = self.score([(query, p) for p in passages])
        order = _order(scores)[:top_k]
        return order, scores[order]


class DSPyFilter:
    def __init__(self, narrtiverag):
        """Same single argument as the reference (rerank.py:16): the ComoRAG instance (needs .global_config and
        .embedding_model; the LLM handle the reference reads is not used)."""
        self.global_config = narrtiverag.global_config
        self.embedding_model = getattr(narrtiverag, "embedding_model", None)
        self.cross_encoder = None
        # The reference's LLM filter returns a SUBSET of the candidates, possibly empty, and ComoRAG falls back to plain
        # dense retrieval when nothing survives (ComoRAG.py:486-488).  A scorer only orders, so the filtering is made
        # explicit: candidates scoring below `rerank_score_threshold` (cross-encoder logit, or cosine for the dense
        # scorer) are dropped, and at most `rerank_keep_fraction` of them are kept.  The defaults keep everything --
        # with them the empty-facts fallback of the reference can only be reached through an empty candidate list.
        thr = cfg_get(self.global_config, "rerank_score_threshold", None)
        self.score_threshold = None if thr is None else float(thr)
        self.keep_fraction = float(cfg_get(self.global_config, "rerank_keep_fraction", 1.0))
        path = cfg_get(self.global_config, "rerank_model_name", None)
        if path:
            self.cross_encoder = CrossEncoderReranker(
                path, max_length=int(cfg_get(self.global_config, "rerank_max_seq_len", 512)),
                token_budget=int(cfg_get(self.global_config, "embedding_token_budget", 16384)))

    def __call__(self, *args, **kwargs):
        return self.rerank(*args, **kwargs)

    def _scores(self, query: str, texts: List[str]) -> np.ndarray:
        if self.cross_encoder is not None:
            return self.cross_encoder.score([(query, t) for t in texts])
        emb = self.embedding_model.batch_encode([query] + texts)
        return emb[1:] @ emb[0]

    def rerank(self, query: str, candidate_items: List[Tuple], candidate_indices: List[int],
               len_after_rerank: Optional[int] = None) -> Tuple[List[int], List[Tuple], dict]:
        if len(candidate_items) == 0:
            return [], [], {"confidence": None}
        texts = [" ".join(str(x) for x in item) for item in candidate_items]
        scores = self._scores(query, texts)
        order = _order(scores)
        if self.keep_fraction < 1.0:
            order = order[: max(0, int(np.ceil(len(order) * self.keep_fraction)))]
        if self.score_threshold is not None:
            order = [i for i in order if scores[i] >= self.score_threshold]     # may leave nothing: DPR fallback
        idx = [candidate_indices[i] for i in order][:len_after_rerank]
        items = [candidate_items[i] for i in order][:len_after_rerank]
        return idx, items, {"confidence": [float(scores[i]) for i in order][:len_after_rerank]}
