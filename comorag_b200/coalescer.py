"""Request coalescing for the <=16 concurrent `tri_retrieve` threads.

ComoRAG answers questions from a ThreadPoolExecutor (ComoRAG.py:436-441); every thread issues batch-1 encodes and
single-query searches.  The engine is fastest when those share one launch: a 32-query pass over the index costs the
same HBM traffic as a 1-query pass.  `Batcher` is a tiny dynamic batcher: callers block on a future, one worker
thread drains whatever arrived within `max_wait_s` (or `max_items`) and runs ONE call for all of them.

Rows of a batch are independent in both kernels (unpadded packing for encode, one selector per query for search),
so coalescing never changes a caller's result.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, Hashable, List, Optional, Sequence, Tuple

import numpy as np


class BatcherClosed(RuntimeError):
    """submit() after close(): the caller should run its request directly."""


class Batcher:
    """fn(group_key, [payload, ...]) -> [result, ...] (same length, same order)."""

    def __init__(self, fn: Callable[[Hashable, List[Any]], Sequence[Any]], max_items: int = 32, max_wait_s: float = 2e-4,
                 weight: Callable[[Any], int] = lambda p: 1, name: str = "crag-batcher"):
        self._fn, self._max_items, self._max_wait, self._weight = fn, max_items, max_wait_s, weight
        self._q: "queue.Queue[Tuple[Hashable, Any, Future]]" = queue.Queue()
        self._closed = False
        self._gate = threading.Lock()   # closed-check + enqueue are one step: nothing can land behind the shutdown sentinel
        self.batches = 0   # statistics: number of fn calls / payloads served
        self.items = 0
        self._t = threading.Thread(target=self._run, name=name, daemon=True)
        self._t.start()

    def submit(self, key: Hashable, payload: Any) -> Future:
        f: Future = Future()
        with self._gate:
            if self._closed:
                raise BatcherClosed("batcher is closed")
            self._q.put((key, payload, f))
        return f

    def call(self, key: Hashable, payload: Any) -> Any:
        return self.submit(key, payload).result()

    def close(self) -> None:
        """Requests already queued are still served; later submit() calls raise BatcherClosed."""
        with self._gate:
            if self._closed:
                return
            self._closed = True
            self._q.put(None)  # type: ignore[arg-type]
        if threading.current_thread() is not self._t:
            self._t.join(timeout=5)

    def _run(self) -> None:
        carry = None
        while True:
            first = carry if carry is not None else self._q.get()
            carry = None
            if first is None:
                return
            key, group, total = first[0], [first], self._weight(first[1])
            deadline = time.perf_counter() + self._max_wait
            while total < self._max_items:
                left = deadline - time.perf_counter()
                try:
                    nxt = self._q.get(timeout=max(left, 0)) if left > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None or nxt[0] != key:   # different group (or shutdown): start the next batch with it
                    carry = nxt
                    if nxt is None:
                        self._q.put(None)          # type: ignore[arg-type]
                        carry = None
                    break
                group.append(nxt)
                total += self._weight(nxt[1])
            self.batches += 1            # counted before the waiters are released, so stats read after a result are complete
            self.items += len(group)
            try:
                results = self._fn(key, [g[1] for g in group])
                if len(results) != len(group):
                    raise RuntimeError(f"batched function returned {len(results)} results for {len(group)} requests")
                for (_, _, fut), r in zip(group, results):
                    fut.set_result(r)
            except BaseException as e:  # every waiter sees the failure; nothing is swallowed
                for _, _, fut in group:
                    if not fut.done():
                        fut.set_exception(e)


class CoalescedSearch:
    """Many threads -> one fused pass.  `search(q, k)` has the contract of DenseIndex.search / EmbeddingStore.search."""

    def __init__(self, index_getter: Callable[[], Any], max_queries: int = 32, max_wait_s: float = 2e-4):
        self._index_getter = index_getter
        self._b = Batcher(self._run, max_items=max_queries, max_wait_s=max_wait_s,
                          weight=lambda p: p[0].shape[0], name="crag-search-coalescer")

    def _run(self, key, payloads):
        k = max(p[1] for p in payloads)
        q = np.concatenate([p[0] for p in payloads], axis=0)
        ids, scores, minmax = self._index_getter().search(q, k)
        out, s = [], 0
        for qq, kk in payloads:
            n = qq.shape[0]
            out.append((ids[s:s + n, :kk], scores[s:s + n, :kk], minmax[s:s + n]))
            s += n
        return out

    def search(self, queries, k: int):
        q = np.asarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        return self._b.call("search", (q, int(k)))

    @property
    def stats(self):
        return {"passes": self._b.batches, "requests": self._b.items}

    def close(self):
        self._b.close()


class CoalescedEncode:
    """Many threads' batch_encode/_encode calls -> one packed encoder forward per wave.  Requests are grouped by
    their encode parameters (instruction, max_length, normalize), exactly the things that change a row's value."""

    def __init__(self, encode_fn: Callable[..., Any], max_texts: int = 64, max_wait_s: float = 3e-4):
        self._encode_fn = encode_fn
        self._b = Batcher(self._run, max_items=max_texts, max_wait_s=max_wait_s, weight=len, name="crag-encode-coalescer")

    def _run(self, key, payloads):
        instruction, max_length, normalize = key
        texts = [t for p in payloads for t in p]
        emb = self._encode_fn(texts, instruction=instruction, max_length=max_length, normalize=normalize)
        out, s = [], 0
        for p in payloads:
            out.append(emb[s:s + len(p)])
            s += len(p)
        return out

    def encode(self, prompts, instruction: str = "", max_length: int = 512, normalize: bool = True):
        if isinstance(prompts, str):
            prompts = [prompts]
        return self._b.call((instruction, int(max_length), bool(normalize)), list(prompts))

    @property
    def stats(self):
        return {"forwards": self._b.batches, "requests": self._b.items}

    def close(self):
        self._b.close()
