"""Host-side tests of the asyncio autobatcher (reference: genlm/backend/trie/async_impl.py, tests/test_trie.py:157-207).
A recording stand-in replaces the trie so the batching contract is tested without a GPU; the GPU suite runs the
same front end over the real kernels."""
import asyncio

import numpy as np
import pytest
import torch

from genlm_backend_b200 import AsyncTokenCharacterTrie, Token


class RecordingTrie:
    def __init__(self, n=4):
        self.decode = [Token(i, bytes([97 + i])) for i in range(n)]
        self.calls = []

    def batch_weight_sum(self, ws):
        self.calls.append(("sum", len(ws), type(ws)))
        return np.stack([np.asarray(w, dtype=np.float32) * 2 for w in ws])

    def batch_weight_max(self, ws):
        self.calls.append(("max", len(ws), type(ws)))
        return np.stack([np.asarray(w, dtype=np.float32) + 1 for w in ws])


def test_concurrent_requests_become_one_batch_per_op():
    async def main():
        trie = RecordingTrie()
        at = AsyncTokenCharacterTrie(trie)
        rows = [torch.full((4,), float(i)) for i in range(128)]
        sums = await asyncio.gather(*[at.weight_sum(r) for r in rows])
        assert trie.calls == [("sum", 128, tuple)]  # one 128-row batch, handed over as a tuple of rows
        for i, s in enumerate(sums):
            assert np.array_equal(s, np.full(4, 2.0 * i, dtype=np.float32))
        trie.calls.clear()
        mixed = await asyncio.gather(*[(at.weight_sum if i % 2 == 0 else at.weight_max)(rows[i]) for i in range(10)])
        assert sorted(trie.calls) == [("max", 5, tuple), ("sum", 5, tuple)]
        for i, r in enumerate(mixed):
            assert np.array_equal(r, np.full(4, 2.0 * i if i % 2 == 0 else i + 1.0, dtype=np.float32))
        await at.cleanup()
        assert at._task is None

    asyncio.run(main())


def test_cleanup_shutdown_and_restart():
    async def main():
        at = AsyncTokenCharacterTrie(RecordingTrie())
        at.start()
        task = at._task
        at.start()
        assert at._task is task  # idempotent while running
        await at.cleanup()
        assert at._task is None
        r = await at.weight_sum(torch.ones(4))  # restarts lazily
        assert np.array_equal(r, np.full(4, 2.0, dtype=np.float32))
        at.shutdown()
        assert at._task is None

    asyncio.run(main())
    at = AsyncTokenCharacterTrie(RecordingTrie())
    at.shutdown()  # never started
    del at


def test_unknown_op_fails_the_future_and_the_task_restarts():
    async def main():
        at = AsyncTokenCharacterTrie(RecordingTrie())
        at.start()
        with pytest.raises(ValueError, match="Unknown operation"):
            future = await at._queue_request(torch.ones(4), "invalid-op")
            await future
        await asyncio.sleep(0)
        assert at._task.done()  # the background task re-raised and ended (async_impl.py:132-137)
        r = await at.weight_max(torch.ones(4))  # next request restarts it
        assert np.array_equal(r, np.full(4, 2.0, dtype=np.float32))
        await at.cleanup()

    asyncio.run(main())


def test_backend_exception_reaches_every_pending_future():
    class Boom(RecordingTrie):
        def batch_weight_sum(self, ws):
            raise RuntimeError("kernel failed")

    async def main():
        at = AsyncTokenCharacterTrie(Boom())
        results = await asyncio.gather(*[at.weight_sum(torch.ones(4)) for _ in range(5)], return_exceptions=True)
        assert all(isinstance(r, RuntimeError) for r in results)
        at.shutdown()

    asyncio.run(main())


def test_from_vocab_builds_both_backends():
    vocab = [Token(0, b"a"), Token(1, b"b"), Token(2, b"ab")]
    from genlm_backend_b200 import TokenCharacterTrie, ParallelTokenCharacterTrie

    assert type(AsyncTokenCharacterTrie.from_vocab(vocab, backend="sequential").trie) is TokenCharacterTrie
    par = AsyncTokenCharacterTrie.from_vocab(vocab, backend="parallel", device="cpu").trie
    assert type(par) is ParallelTokenCharacterTrie and par.device == "cpu"
    with pytest.raises(ValueError):
        AsyncTokenCharacterTrie.from_vocab(vocab, backend="invalid")


def test_event_loop_keeps_running_while_a_batch_is_in_flight():
    """The batched call runs on a worker thread: other coroutines make progress during it, and requests that arrive
    meanwhile form the next batch (the reference blocks the loop for the whole batch, async_impl.py:96-137)."""
    import time

    class Slow(RecordingTrie):
        def batch_weight_sum(self, ws):
            time.sleep(0.15)
            return super().batch_weight_sum(ws)

    async def main():
        trie = Slow()
        at = AsyncTokenCharacterTrie(trie)
        ticks = 0

        async def ticker():
            nonlocal ticks
            while True:
                await asyncio.sleep(0.005)
                ticks += 1

        async def late():  # arrives while the first batch is on the worker thread
            await asyncio.sleep(0.03)
            return await asyncio.gather(*[at.weight_sum(torch.full((4,), 7.0)) for _ in range(3)])

        t = asyncio.create_task(ticker())
        first = asyncio.gather(*[at.weight_sum(torch.full((4,), float(i))) for i in range(5)])
        a, b = await asyncio.gather(first, late())
        t.cancel()
        assert ticks >= 10, ticks  # ~0.3 s of batches at a 5 ms tick
        assert [c[:2] for c in trie.calls] == [("sum", 5), ("sum", 3)]
        assert np.array_equal(a[4], np.full(4, 8.0, dtype=np.float32)) and np.array_equal(b[0], np.full(4, 14.0, dtype=np.float32))
        await at.cleanup()

    asyncio.run(main())


def test_sparse_readout_requests_are_batched_per_shape():
    class Sparse(RecordingTrie):
        def _preprocess_ws(self, rows):
            return torch.stack([torch.as_tensor(r, dtype=torch.float32) for r in rows])

        def batch_weight_sum_at(self, ws, ids, normalizer=None, log=False):
            self.calls.append(("sum_at", ws.shape[0], ids.shape, normalizer is not None, log))
            full = ws.numpy() * 2
            out = np.take_along_axis(full, ids.astype(np.int64), axis=1)
            if normalizer is not None:
                out = out / full[np.arange(len(full)), normalizer][:, None]
            return out

        def batch_weight_max_at(self, ws, ids, log=False):
            self.calls.append(("max_at", ws.shape[0], ids.shape, log))
            return np.take_along_axis(ws.numpy() + 1, ids.astype(np.int64), axis=1)

    async def main():
        trie = Sparse()
        at = AsyncTokenCharacterTrie(trie)
        rows = [torch.arange(4, dtype=torch.float32) + i for i in range(6)]
        res = await asyncio.gather(*[at.weight_sum_at(rows[i], [0, 3]) for i in range(4)],
                                   at.weight_sum_at(rows[4], [1, 2, 3], normalizer=3), at.weight_max_at(rows[5], [2]))
        assert sorted(c[:3] for c in trie.calls) == [("max_at", 1, (1, 1)), ("sum_at", 1, (1, 3)), ("sum_at", 4, (4, 2))]
        assert np.allclose(res[1], [2.0, 8.0]) and np.allclose(res[4], [10 / 14, 12 / 14, 1.0]) and np.allclose(res[5], [8.0])
        await at.cleanup()

    asyncio.run(main())
