"""``AsyncTokenCharacterTrie``: asyncio front end that batches concurrent mass queries.

Mirrors the reference's ``genlm/backend/trie/async_impl.py``: requests are queued as ``(weights, future, op)``;
a background task drains whatever is queued, groups by operation, issues ONE batched call per group and
resolves the futures; an exception fails every pending future of that drain and ends the task (the next
request restarts it).  Everything runs on the event-loop thread.
"""
import asyncio
import logging
from collections import defaultdict

from .base import TokenCharacterTrie
from .parallel import ParallelTokenCharacterTrie

logger = logging.getLogger(__name__)


class AsyncTokenCharacterTrie:
    """An asynchronous wrapper for trie implementations that provides automatic request batching."""

    def __init__(self, trie):
        self.trie = trie
        self._queue = None
        self._task = None

    @classmethod
    def from_vocab(cls, vocab, backend="parallel", **kwargs):
        """Build the trie for ``vocab`` with the ``"sequential"`` or ``"parallel"`` implementation."""
        if backend == "sequential":
            trie = TokenCharacterTrie(decode=vocab, **kwargs)
        elif backend == "parallel":
            trie = ParallelTokenCharacterTrie(decode=vocab, **kwargs)
        else:
            raise ValueError(f"Unknown backend: {backend}. Must be one of ['sequential', 'parallel']")
        return cls(trie)

    async def _queue_request(self, request, op):
        if not self._task or self._task.done():
            self.start()
        future = asyncio.get_running_loop().create_future()
        await self._queue.put((request, future, op))
        return future

    async def weight_sum(self, ws):
        """Queue a ``weight_sum`` request; concurrent calls are batched together."""
        future = await self._queue_request(ws, "sum")
        return await future

    async def weight_max(self, ws):
        """Queue a ``weight_max`` request; concurrent calls are batched together."""
        future = await self._queue_request(ws, "max")
        return await future

    def start(self):
        """Start the background task (binds a fresh queue to the running loop)."""
        if not self._task or self._task.done():
            self._queue = asyncio.Queue()
            self._task = asyncio.create_task(self._background_loop())

    def _do_weight_sums(self, batch_weights):
        return self.trie.batch_weight_sum(batch_weights)

    def _do_weight_maxs(self, batch_weights):
        return self.trie.batch_weight_max(batch_weights)

    async def _background_loop(self):
        while True:
            groups = defaultdict(list)
            try:
                request, future, op = await self._queue.get()
                groups[op].append((request, future))
                while not self._queue.empty():
                    request, future, op = self._queue.get_nowait()
                    groups[op].append((request, future))

                for op, group in groups.items():
                    requests, futures = zip(*group)
                    if op == "sum":
                        logger.debug(f"processing {len(requests)} sum requests")
                        results = self._do_weight_sums(requests)
                    elif op == "max":
                        logger.debug(f"processing {len(requests)} max requests")
                        results = self._do_weight_maxs(requests)
                    else:
                        raise ValueError(f"Unknown operation: {op}")
                    for future, result in zip(futures, results):
                        if not future.done():
                            future.set_result(result)
            except Exception as e:
                for group in groups.values():
                    for _, future in group:
                        if not future.done():
                            future.set_exception(e)
                raise

    async def cleanup(self):
        """Cancel the background task and wait for it to finish."""
        if self._task and not self._task.done():
            self._task.cancel()
            try:
                await self._task
            except asyncio.CancelledError:
                pass
            self._task = None

    def shutdown(self):
        """Cancel the background task without awaiting it (safe when the loop is already closed)."""
        if self._task is not None:
            try:
                self._task.cancel()
            except RuntimeError:
                pass
            self._task = None

    def __del__(self):
        self.shutdown()
