"""comorag_b200.install rebinds the hot-path names inside an unmodified reference checkout (build container only:
the GPU box has no /root/reference, where this test skips)."""
import os
import sys
import types

import pytest

REF = os.environ.get("COMORAG_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "comorag")), reason="reference checkout not present")
def test_install_rebinds_names_package_wide():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    for m in ("igraph", "umap"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import src.comorag  # noqa: F401  (its __init__ imports ComoRAG.py, which binds the names with `from ... import`)
    ref_main = sys.modules["src.comorag.ComoRAG"]   # the package attribute of that name is the class, not the module

    import comorag_b200.install as crag
    from comorag_b200.embedding_model import BGEEmbeddingModel, _get_embedding_model_class
    from comorag_b200.embedding_store import EmbeddingStore
    from comorag_b200.retrieval import get_similar_summaries
    counts = crag.install("src.comorag")
    assert counts["EmbeddingStore"] >= 3 and counts["_get_embedding_model_class"] >= 2
    assert ref_main.EmbeddingStore is EmbeddingStore
    assert ref_main._get_embedding_model_class is _get_embedding_model_class
    assert ref_main.get_similar_summaries is get_similar_summaries
    import src.comorag.utils.timeline_utils as tl
    assert tl.EmbeddingStore is EmbeddingStore
    assert ref_main._get_embedding_model_class("BAAI/bge-large-en-v1.5") is BGEEmbeddingModel
    assert ref_main.DSPyFilter.__module__.startswith("src.comorag")      # LLM filter untouched unless rerank=True
    crag.install("src.comorag", rerank=True)
    assert ref_main.DSPyFilter.__module__ == "comorag_b200.rerank"
