"""Host logic of the coalescing front end (triple_hybrid_rag_b200/frontend.py) with a stub batch function:
concurrent requests share launches, a lone request leaves after max_wait, a full batch leaves at once, every
request gets its own row back, a failing batch fails all of its waiters."""
import asyncio
import time

import pytest
import torch

from triple_hybrid_rag_b200.frontend import CoalescingFrontEnd


def _stub(calls):
    def fn(queries, vectors, keywords, graph, collections):
        calls.append((list(queries), tuple(vectors.shape), graph, collections))
        time.sleep(0.01)   # the "GPU" is busy; the loop must stay responsive
        return [(q, float(v.sum()), kw, c) for q, v, kw, c in
                zip(queries, vectors, keywords, collections or [None] * len(queries))]
    return fn


def test_concurrent_requests_share_one_launch():
    calls = []
    async def go():
        fe = CoalescingFrontEnd(_stub(calls), max_batch=64, max_wait_ms=20)
        reqs = [fe.retrieve_candidates(f"q{i}", torch.full((8,), float(i)), [f"k{i}"],
                                       collection="a" if i % 2 else None) for i in range(10)]
        out = await asyncio.gather(*reqs)
        return fe, out
    fe, out = asyncio.run(go())
    assert fe.batches == [10] and len(calls) == 1 and calls[0][1] == (10, 8)
    assert [o[0] for o in out] == [f"q{i}" for i in range(10)]
    assert [o[1] for o in out] == [8.0 * i for i in range(10)]
    assert [o[3] for o in out] == ["a" if i % 2 else None for i in range(10)]
    assert calls[0][2] is None          # nobody passed a graph list


def test_full_batch_leaves_at_once_and_overflow_splits():
    calls = []
    async def go():
        fe = CoalescingFrontEnd(_stub(calls), max_batch=4, max_wait_ms=10_000)   # the timer never fires in this test
        t0 = time.perf_counter()
        out = await asyncio.gather(*[fe.retrieve_candidates(f"q{i}", torch.zeros(4), []) for i in range(8)])
        return fe, out, time.perf_counter() - t0
    fe, out, dt = asyncio.run(go())
    assert fe.batches == [4, 4] and dt < 5 and [o[0] for o in out] == [f"q{i}" for i in range(8)]


def test_lone_request_leaves_after_max_wait_and_drain():
    calls = []
    async def go():
        fe = CoalescingFrontEnd(_stub(calls), max_batch=256, max_wait_ms=30)
        t0 = time.perf_counter()
        r = await fe.retrieve_candidates("solo", torch.ones(4), ["x"], graph_ids=["c1", "c2"])
        dt = time.perf_counter() - t0
        pending = asyncio.ensure_future(fe.retrieve_candidates("late", torch.ones(4), []))
        await asyncio.sleep(0)
        await fe.drain()                  # launches the waiting request without waiting for the timer
        return fe, r, dt, await pending
    fe, r, dt, late = asyncio.run(go())
    assert r[0] == "solo" and 0.02 <= dt < 1.0 and late[0] == "late" and fe.batches == [1, 1]
    assert calls[0][2] == [["c1", "c2"]]


def test_a_failing_batch_fails_every_waiter():
    def boom(*a):
        raise RuntimeError("libthr error -2: launch failed")
    async def go():
        fe = CoalescingFrontEnd(boom, max_batch=8, max_wait_ms=5)
        return await asyncio.gather(*[fe.retrieve_candidates(f"q{i}", torch.zeros(2), []) for i in range(3)],
                                    return_exceptions=True)
    res = asyncio.run(go())
    assert len(res) == 3 and all(isinstance(r, RuntimeError) and "libthr" in str(r) for r in res)


# ---- the tool boundary's response dictionary, against the reference's own function ---------------------------------
def test_tool_response_matches_the_reference_goldens():
    """format_tool_response / search_knowledge_base_rag2 == what the reference's `_search_knowledge_base_rag2`
    (crm_knowledge.py:126-182, run unmodified by tests/golden/make_tool_golden.py) returned for the same
    RetrievalResult: refusals, empty results, missing parent text, falsy scores, tables."""
    import json
    from pathlib import Path
    from triple_hybrid_rag_b200.retriever import RetrievalCandidate, RetrievalResult
    from triple_hybrid_rag_b200.tool import format_tool_response, search_knowledge_base_rag2
    cases = json.loads((Path(__file__).parent / "golden" / "tool_golden.json").read_text(encoding="utf-8"))
    assert len(cases) == 8
    for case in cases:
        r = case["result"]
        result = RetrievalResult(success=r["success"], contexts=[RetrievalCandidate(**c) for c in r["contexts"]],
                                 max_rerank_score=r["max_rerank_score"], refused=r["refused"],
                                 refusal_reason=r["refusal_reason"], timings=dict(r["timings"]))
        assert format_tool_response(case["query"], case["category"], result) == case["response"]

        class Stub:
            async def retrieve(self, query, collection=None, top_k=None):
                assert (query, collection, top_k) == (case["query"], case["category"], case["limit"])
                return result
        assert search_knowledge_base_rag2(case["query"], case["category"], case["limit"], retriever=Stub()) == case["response"]


def test_blocking_tool_handler_refuses_a_running_loop():
    import asyncio
    from triple_hybrid_rag_b200.tool import search_knowledge_base_rag2

    class Stub:
        async def retrieve(self, query, collection=None, top_k=None):
            return None

    async def inside():
        with pytest.raises(RuntimeError, match="blocking tool handler"):
            search_knowledge_base_rag2("q", retriever=Stub())
    asyncio.run(inside())


def test_fallback_planner_matches_the_reference_goldens():
    """FallbackQueryPlanner == the plan the reference's QueryPlanner returns when its LLM call fails
    (query_planner.py:178-187; goldens: tests/golden/make_planner_golden.py), field for field, defaults included."""
    import asyncio
    import dataclasses
    import json
    from pathlib import Path
    from triple_hybrid_rag_b200.retriever import FallbackQueryPlanner
    cases = json.loads((Path(__file__).parent / "golden" / "planner_golden.json").read_text(encoding="utf-8"))
    assert len(cases) == 10
    p = FallbackQueryPlanner()
    for case in cases:
        assert dataclasses.asdict(p.plan(case["query"], case["collection"])) == case["plan"]
        assert dataclasses.asdict(asyncio.run(p.plan_async(case["query"], case["collection"]))) == case["plan"]


def test_settings_defaults_match_the_reference():
    """Rag2Settings() == the reference's default rag2_* retrieval knobs (config.py:280-314; golden written by
    tests/golden/make_planner_golden.py from the reference's Settings class)."""
    import json
    from pathlib import Path
    from triple_hybrid_rag_b200.retriever import Rag2Settings
    want = json.loads((Path(__file__).parent / "golden" / "settings_golden.json").read_text(encoding="utf-8"))
    mine = Rag2Settings()
    have = {k: getattr(mine, k) for k in dir(mine) if k.startswith("rag2_")}
    assert have == want and all(type(have[k]) is type(want[k]) for k in want)
