"""Import the reference's own modules (read-only tree at /root/reference) inside the BUILD container.

Only the golden-vector generators use this; nothing that runs on the GPU box may import it.
Shims: `igraph`/`umap` are absent -> empty stub modules so the package imports;
`accelerate` is absent -> a 3-line subclass drops `device_map` from the HF init params.  The
reference code itself is untouched.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("COMORAG_REFERENCE", "/root/reference")


def import_reference():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for m in ("igraph", "umap"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    from src.comorag.embedding_model.BGEEmbedding import BGEEmbeddingModel
    from src.comorag.embedding_store import EmbeddingStore
    from src.comorag.utils.config_utils import BaseConfig

    class OracleBGE(BGEEmbeddingModel):
        def _init_embedding_config(self):
            super()._init_embedding_config()
            self.embedding_config.model_init_params.pop("device_map", None)

    return types.SimpleNamespace(BaseConfig=BaseConfig, BGEEmbeddingModel=BGEEmbeddingModel, OracleBGE=OracleBGE,
                                 EmbeddingStore=EmbeddingStore)
