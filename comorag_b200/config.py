"""Engine-side view of the reference's BaseConfig (utils/config_utils.py:17-298).

The engine only READS the embedding_* fields of whatever `global_config` object ComoRAG hands it
(config_utils.py:128-146); when none is given the reference falls back to `BaseConfig()`
(embedding_model/base.py:198-203) -- this dataclass carries the same defaults for exactly those fields so
the engine does not depend on the reference package being importable.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class EngineConfig:
    embedding_model_name: str = "nvidia/NV-Embed-v2"      # config_utils.py:128
    embedding_batch_size: int = 32                        # config_utils.py:132
    embedding_return_as_normalized: bool = True           # config_utils.py:136
    embedding_max_seq_len: int = 2048                     # config_utils.py:140
    embedding_model_dtype: str = "auto"                   # config_utils.py:144 (never read by the reference)
    # engine-only knobs; defaults keep ComoRAG.py unchanged
    embedding_device: str = "cuda"
    embedding_token_budget: int = 16384                   # max packed tokens per encoder launch
    embedding_store_append_only: bool = False             # raw append-only shards instead of per-upsert parquet rewrite
    embedding_coalesce: bool = False                      # batch concurrent callers' encodes/searches (coalescer.py)
    embedding_coalesce_wait_ms: float = 0.3
    embedding_coalesce_max_texts: int = 64
    rerank_model_name: str = ""                           # sequence-classification checkpoint dir -> cross-encoder rerank
    rerank_max_seq_len: int = 512


def cfg_get(cfg, name: str, default=None):
    return getattr(cfg, name, default) if cfg is not None else default
