"""Dynamic batching of concurrent encode / search requests (host logic, no GPU): results are routed back to the
right caller, groups with different parameters never mix, failures reach every waiter."""
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from comorag_b200.coalescer import Batcher, CoalescedEncode, CoalescedSearch


def test_search_requests_from_16_threads_share_passes():
    rng = np.random.default_rng(0)
    E = rng.standard_normal((500, 16)).astype(np.float32)

    class FakeIndex:
        def __init__(self):
            self.calls = []

        def search(self, q, k):
            self.calls.append(q.shape[0])
            time.sleep(0.002)                      # a pass takes a while; requests pile up behind it
            sc = q @ E.T
            ids = np.argsort(-sc, axis=1)[:, :k]
            return ids, np.take_along_axis(sc, ids, 1), np.stack([sc.min(1), sc.max(1)], 1)

    idx = FakeIndex()
    cs = CoalescedSearch(lambda: idx, max_queries=32, max_wait_s=5e-3)
    Q = rng.standard_normal((64, 16)).astype(np.float32)
    ks = [5 if i % 2 else 10 for i in range(64)]
    with ThreadPoolExecutor(16) as ex:
        res = list(ex.map(lambda i: cs.search(Q[i], ks[i]), range(64)))
    for i, (ids, sc, mm) in enumerate(res):
        want = np.argsort(-(Q[i] @ E.T))[:ks[i]]
        assert ids.shape == (1, ks[i]) and (ids[0] == want).all() and mm.shape == (1, 2)
    assert sum(idx.calls) == 64 and len(idx.calls) < 40 and max(idx.calls) <= 32     # coalesced, never over 32 per pass
    assert cs.stats["requests"] == 64
    cs.close()


def test_encode_groups_do_not_mix_and_order_is_kept():
    seen = []

    def enc(texts, instruction="", max_length=512, normalize=True):
        seen.append((instruction, list(texts)))
        return np.array([[len(t), len(instruction)] for t in texts], dtype=np.float32)

    ce = CoalescedEncode(enc, max_texts=64, max_wait_s=5e-3)
    jobs = [(["a" * (i + 1), "b" * (i + 2)], "Q:" if i % 3 == 0 else "") for i in range(30)]
    with ThreadPoolExecutor(12) as ex:
        outs = list(ex.map(lambda j: ce.encode(j[0], instruction=j[1]), jobs))
    for (texts, ins), out in zip(jobs, outs):
        assert out.shape == (2, 2) and out[:, 0].tolist() == [len(t) for t in texts] and (out[:, 1] == len(ins)).all()
    assert all(len({ins}) == 1 for ins, _ in seen) and len(seen) < 30
    assert ce.encode("xyz").shape == (1, 2)
    ce.close()


def test_failure_reaches_every_waiter():
    def boom(key, payloads):
        raise ValueError("kernel failed")

    b = Batcher(boom, max_items=8, max_wait_s=2e-3)
    futs = [b.submit("k", i) for i in range(5)]
    for f in futs:
        with pytest.raises(ValueError):
            f.result(timeout=2)
    b.close()


def test_close_serves_what_is_queued_and_refuses_the_rest():
    from comorag_b200.coalescer import BatcherClosed
    release = threading.Event()

    def slow(key, payloads):
        release.wait(2)
        return [p * 2 for p in payloads]

    b = Batcher(slow, max_items=1, max_wait_s=1e-3)
    futs = [b.submit("k", i) for i in range(4)]          # the worker blocks in the first batch, three wait in the queue
    closer = threading.Thread(target=b.close)
    closer.start()
    time.sleep(0.05)
    with pytest.raises(BatcherClosed):                   # nothing can be enqueued behind the shutdown sentinel
        b.submit("k", 99)
    release.set()
    assert [f.result(timeout=3) for f in futs] == [0, 2, 4, 6]
    closer.join(timeout=6)
    b.close()                                            # idempotent
