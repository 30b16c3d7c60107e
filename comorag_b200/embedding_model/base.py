"""BaseEmbeddingModel / EmbeddingConfig call surface (reference: embedding_model/base.py).

Same names, argument meaning and error behaviour as the reference classes so ComoRAG.py and its helpers run
unchanged; the implementation is our own.
"""
from __future__ import annotations

import hashlib
import json
import sqlite3
import threading
from typing import Any, Dict, List, Optional

import numpy as np

from ..config import EngineConfig


class EmbeddingConfig:
    """Attribute/dict hybrid config bag (reference: base.py:21-104, a dataclass wrapping `_data`)."""

    def __init__(self) -> None:
        object.__setattr__(self, "_data", {})

    def __getattr__(self, key: str) -> Any:
        data = object.__getattribute__(self, "_data")
        if key in data:
            return data[key]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'")

    def __setattr__(self, key: str, value: Any) -> None:
        self._data[key] = value

    def __delattr__(self, key: str) -> None:
        if key not in self._data:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'")
        del self._data[key]

    def __getitem__(self, key: str) -> Any:
        if key not in self._data:
            raise KeyError(f"'{key}' not found in configuration.")
        return self._data[key]

    def __setitem__(self, key: str, value: Any) -> None:
        self._data[key] = value

    def __delitem__(self, key: str) -> None:
        if key not in self._data:
            raise KeyError(f"'{key}' not found in configuration.")
        del self._data[key]

    def __contains__(self, key: str) -> bool:
        return key in self._data

    def batch_upsert(self, updates: Dict[str, Any]) -> None:
        self._data.update(updates)

    def to_dict(self) -> Dict[str, Any]:
        return self._data

    def to_json(self) -> str:
        return json.dumps(self._data)

    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]) -> "EmbeddingConfig":
        inst = cls()
        inst.batch_upsert(config_dict)
        return inst

    @classmethod
    def from_json(cls, json_str: str) -> "EmbeddingConfig":
        return cls.from_dict(json.loads(json_str))

    def __str__(self) -> str:
        return json.dumps(self._data, indent=4, default=str)


class BaseEmbeddingModel:
    """reference: base.py:189-218."""

    global_config: Any
    embedding_model_name: str
    embedding_config: EmbeddingConfig
    embedding_dim: int

    def __init__(self, global_config: Optional[Any] = None) -> None:
        self.global_config = EngineConfig() if global_config is None else global_config
        self.embedding_model_name = self.global_config.embedding_model_name

    def batch_encode(self, texts: List[str], **kwargs) -> None:
        raise NotImplementedError

    def get_query_doc_scores(self, query_vec: np.ndarray, doc_vecs: np.ndarray):
        """base.py:212-218: np.dot(query_vec, doc_vecs.T)."""
        return np.dot(query_vec, doc_vecs.T)


def make_cache_embed(encode_func, cache_file_name: str, device):
    """Optional sqlite embedding cache (reference: base.py:112-187).

    Same contract: keyword-only call with `prompts`, rows keyed by sha256 of
    json{"instruction", "promps" (sic), "max_length"}, fp32 blobs, returns a [n, D] tensor on `device`.
    A process-wide lock replaces the reference's FileLock (filelock is an optional dependency here).
    """
    import torch

    lock = _CACHE_LOCKS.setdefault(cache_file_name, threading.Lock())

    def _keys(kwargs):
        instruction = kwargs.get("instruction", "")
        max_length = kwargs.get("max_length", "")
        return [hashlib.sha256(json.dumps({"instruction": instruction, "promps": p, "max_length": max_length},
                                          sort_keys=True, default=str).encode("utf-8")).hexdigest()
                for p in kwargs["prompts"]]

    def wrapper(**kwargs):
        keys = _keys(kwargs)
        rows: List[Optional[np.ndarray]] = []
        with lock, sqlite3.connect(cache_file_name) as conn:
            conn.execute("CREATE TABLE IF NOT EXISTS embeddings (hash TEXT PRIMARY KEY, embedding BLOB)")
            for h in keys:
                hit = conn.execute("SELECT embedding FROM embeddings WHERE hash = ?", (h,)).fetchone()
                rows.append(np.frombuffer(hit[0], dtype=np.float32).copy() if hit else None)
        missed = [i for i, r in enumerate(rows) if r is None]
        if missed:
            sub = dict(kwargs)
            sub["prompts"] = [kwargs["prompts"][i] for i in missed]
            fresh = encode_func(**sub)
            fresh_np = fresh.detach().float().cpu().numpy() if hasattr(fresh, "detach") else np.asarray(fresh, np.float32)
            with lock, sqlite3.connect(cache_file_name) as conn:
                for j, i in enumerate(missed):
                    rows[i] = fresh_np[j]
                    conn.execute("INSERT OR REPLACE INTO embeddings (hash, embedding) VALUES (?, ?)",
                                 (keys[i], fresh_np[j].astype(np.float32).tobytes()))
                conn.commit()
        return torch.from_numpy(np.stack(rows)).to(device)

    return wrapper


_CACHE_LOCKS: Dict[str, threading.Lock] = {}


class EmbeddingCache:
    """Process-wide content -> embedding cache with the reference's classmethod API (base.py:222-260: get / set /
    contains / clear).  The reference backs it with a `multiprocessing.Manager().dict()` that it never uses from a
    second process (the class has no caller in the reference tree); here it is a lock-guarded dict, and
    `share_across_processes()` swaps in a manager dict for callers that do fork workers."""

    _cache: Dict[Any, Any] = {}
    _manager = None
    _lock = threading.Lock()

    @classmethod
    def share_across_processes(cls) -> None:
        with cls._lock:
            if cls._manager is None:
                import multiprocessing
                cls._manager = multiprocessing.Manager()
                shared = cls._manager.dict()
                shared.update(cls._cache)
                cls._cache = shared

    @classmethod
    def get(cls, content):
        return cls._cache.get(content)

    @classmethod
    def set(cls, content, embedding) -> None:
        with cls._lock:
            cls._cache[content] = embedding

    @classmethod
    def contains(cls, content) -> bool:
        return content in cls._cache

    @classmethod
    def clear(cls) -> None:
        with cls._lock:
            cls._cache.clear()
