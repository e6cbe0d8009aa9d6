"""Host logic of the request micro-batcher behind GpuSearchEngine.submit with a stand-in engine: no GPU.
The reference answers one query per request on the event loop (app.py:84-111); the batcher groups concurrent submitters
into one batched launch, must propagate errors to every waiter and must not leave waiters blocked on close."""
import threading
import time
import types

import numpy as np
import pytest

from diskrag_b200.search_engine import _MicroBatcher


class FakeEngine:
    def __init__(self, fail=False, delay=0.0):
        self.calls, self.fail, self.delay = [], fail, delay

    def search_vectors(self, Q, k=10, L=100):
        self.calls.append(Q.shape[0])
        if self.delay:
            time.sleep(self.delay)
        if self.fail:
            raise ValueError("boom")
        ids = np.tile(np.arange(k, dtype=np.int32), (Q.shape[0], 1)) + Q[:, :1].astype(np.int32)   # answer depends on the query
        ids[:, -1] = -1                                                                            # a short list: -1 padding is dropped
        return types.SimpleNamespace(ids=ids, dists=np.full(ids.shape, 0.5, np.float32))


def test_concurrent_submitters_share_launches_and_get_their_own_answer():
    eng = FakeEngine()
    b = _MicroBatcher(eng, max_batch=16, max_wait_ms=50.0, k=4, L=8)
    futs = {}

    def client(i):
        futs[i] = b.submit(np.full(6, float(i), np.float32))
    th = [threading.Thread(target=client, args=(i,)) for i in range(40)]
    [t.start() for t in th]; [t.join() for t in th]
    for i, f in futs.items():
        res = f.result(timeout=5)
        assert res == [(0.5, i), (0.5, i + 1), (0.5, i + 2)]
    b.stop()
    assert sum(eng.calls) == 40 and max(eng.calls) <= 16 and len(eng.calls) < 40


def test_errors_reach_every_waiter():
    b = _MicroBatcher(FakeEngine(fail=True), max_batch=8, max_wait_ms=20.0, k=4, L=8)
    futs = [b.submit(np.zeros(6, np.float32)) for _ in range(5)]
    for f in futs:
        with pytest.raises(ValueError, match="boom"):
            f.result(timeout=5)
    b.stop()


def test_stop_fails_queued_waiters_instead_of_blocking_them():
    eng = FakeEngine(delay=0.3)
    b = _MicroBatcher(eng, max_batch=1, max_wait_ms=1.0, k=4, L=8)
    first = b.submit(np.zeros(6, np.float32))
    time.sleep(0.1)                                   # the worker is inside the first (slow) launch now
    rest = [b.submit(np.ones(6, np.float32)) for _ in range(3)]
    b.stop()
    assert first.result(timeout=5)
    for f in rest:
        try:
            f.result(timeout=5)                       # answered if the worker got to it before stopping ...
        except RuntimeError as e:
            assert "closed" in str(e)                 # ... failed loudly otherwise; never left hanging
