"""Request coalescing at the tool boundary.

The reference serves one query per call: every voice turn builds a RAG2Retriever and runs retrieve() on its own
(src/voice_agent/tools/crm_knowledge.py:105-124), so N concurrent calls are N scans of the corpus.  On the GPU one
scan serves a whole batch (K1 is an M = 256 tile: a second query in the batch is free), so this front end collects
the candidate-retrieval step of concurrent calls into one GpuRAG2Retriever.retrieve_batch launch: a request waits
at most `max_wait_ms` for company, or until `max_batch` requests are waiting.

Only host logic lives here (asyncio queue, futures); the scoring is retrieve_batch's, i.e. K1 + K2 + K3.
"""
from __future__ import annotations

import asyncio
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Sequence

import torch


@dataclass
class _Pending:
    query: str
    vector: torch.Tensor
    keywords: Sequence[str]
    graph_ids: Optional[Sequence[str]]
    collection: Optional[str]
    future: "asyncio.Future" = field(repr=False, default=None)


class CoalescingFrontEnd:
    """`batch_fn(queries, query_vectors [B, D], keywords, graph_ids, collections) -> list of per-query results` is
    normally `functools.partial(retriever.retrieve_batch, top_k=..., k_sem=..., k_lex=...)`.  It runs on ONE worker
    thread owned by this front end: the event loop keeps accepting (and coalescing) requests while the GPU works on
    the previous batch, and batches reach the engine strictly one after another — a libthr handle takes one thread at
    a time (include/thr.h), two overlapping batches would share its scratch, counters and status word."""

    def __init__(self, batch_fn: Callable[..., List[Any]], max_batch: int = 256, max_wait_ms: float = 2.0):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        from concurrent.futures import ThreadPoolExecutor
        self._executor = ThreadPoolExecutor(max_workers=1, thread_name_prefix="thr-batch")
        self.batch_fn = batch_fn
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._waiting: List[_Pending] = []
        self._timer: Optional[asyncio.TimerHandle] = None
        self._inflight: set = set()
        self.batches: List[int] = []          # sizes of the launched batches (observability, tests)

    async def retrieve_candidates(self, query: str, query_vector: torch.Tensor, keywords: Sequence[str],
                                  graph_ids: Optional[Sequence[str]] = None, collection: Optional[str] = None) -> Any:
        """One request; resolves with its own entry of the batch result (or raises what the batch raised)."""
        loop = asyncio.get_running_loop()
        p = _Pending(query, query_vector, keywords, graph_ids, collection, loop.create_future())
        self._waiting.append(p)
        if len(self._waiting) >= self.max_batch:
            self._flush()
        elif self._timer is None:
            self._timer = loop.call_later(self.max_wait, self._flush)
        return await p.future

    def _flush(self) -> None:
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None
        while self._waiting:
            batch, self._waiting = self._waiting[: self.max_batch], self._waiting[self.max_batch:]
            task = asyncio.get_running_loop().create_task(self._run(batch))
            self._inflight.add(task)
            task.add_done_callback(self._inflight.discard)

    async def _run(self, batch: List[_Pending]) -> None:
        self.batches.append(len(batch))
        try:
            graph = None
            if any(p.graph_ids is not None for p in batch):
                graph = [list(p.graph_ids or []) for p in batch]
            colls = [p.collection for p in batch]
            args = ([p.query for p in batch], torch.stack([p.vector.reshape(-1) for p in batch]),
                    [list(p.keywords) for p in batch], graph, colls if any(c is not None for c in colls) else None)
            res = await asyncio.get_running_loop().run_in_executor(self._executor, lambda: self.batch_fn(*args))
            if len(res) != len(batch):
                raise RuntimeError(f"batch function returned {len(res)} results for {len(batch)} requests")
            for p, r in zip(batch, res):
                if not p.future.done():
                    p.future.set_result(r)
        except Exception as e:  # every waiter of the batch sees the failure; nothing is swallowed
            for p in batch:
                if not p.future.done():
                    p.future.set_exception(e)

    async def drain(self) -> None:
        """Launch what is waiting and wait for every batch in flight (shutdown, tests)."""
        self._flush()
        if self._inflight:
            await asyncio.gather(*list(self._inflight), return_exceptions=True)

    def close(self) -> None:
        self._executor.shutdown(wait=True)
