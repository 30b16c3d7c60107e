"""Drop the engine under an unmodified ComoRAG checkout by import-time substitution.

    import comorag_b200.install as crag
    crag.install("src.comorag")          # before or after `from src.comorag import ComoRAG`
    rag = ComoRAG(global_config=BaseConfig(embedding_model_name=".../bge-large-en-v1.5", ...))

`ComoRAG.py` binds `_get_embedding_model_class`, `EmbeddingStore`, `DSPyFilter`, `get_similar_summaries` by
`from ... import` (ComoRAG.py:21-35), so the substitution rewrites those names in every already-imported module of the
package as well as in their defining modules; class bodies resolve them from module globals at call time.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict


def install(package: str = "src.comorag", rerank: bool = False, summaries: bool = True) -> Dict[str, int]:
    """Returns {name: number of module attributes rebound}.  `rerank=True` also swaps the LLM filter for the
    dense reranker (new arithmetic, off by default so answers stay reference-identical)."""
    from . import embedding_model as em
    from . import embedding_store as es
    from . import rerank as rr
    from . import retrieval as rt

    ref_em = importlib.import_module(package + ".embedding_model")
    ref_es = importlib.import_module(package + ".embedding_store")
    swaps = {
        "_get_embedding_model_class": (ref_em._get_embedding_model_class, em._get_embedding_model_class),
        "BGEEmbeddingModel": (ref_em.BGEEmbeddingModel, em.BGEEmbeddingModel),
        "EmbeddingStore": (ref_es.EmbeddingStore, es.EmbeddingStore),
    }
    if summaries:
        ref_eu = importlib.import_module(package + ".utils.embed_utils")
        swaps["get_similar_summaries"] = (ref_eu.get_similar_summaries, rt.get_similar_summaries)
    if rerank:
        ref_rr = importlib.import_module(package + ".rerank")
        swaps["DSPyFilter"] = (ref_rr.DSPyFilter, rr.DSPyFilter)
    counts = {k: 0 for k in swaps}
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == package or name.startswith(package + ".")):
            continue
        for attr, (old, new) in swaps.items():
            if getattr(mod, attr, None) is old:
                setattr(mod, attr, new)
                counts[attr] += 1
    return counts
